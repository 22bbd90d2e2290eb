"""GPU parity tests of the tcgen05 tile engine through the C-ABI against fp32 torch references of the same op.
Tolerances: operands are bf16-rounded in the reference too, accumulation is fp32 on both sides, the output is
bf16-rounded -> |diff| <= 2^-8 * |ref| + small absolute slack."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16).float()


def _act(x):
    """round to the build's activation storage format (fp16 by default)"""
    from tests import engine_helpers as H
    return x.to(H.act_dtype()).float()


def _actdt():
    from tests import engine_helpers as H
    return H.act_dtype()


def _close(a, b, rtol=1.0 / 128, atol=2e-2):
    d = (a.float() - b.float()).abs()
    lim = rtol * b.float().abs() + atol
    assert bool((d <= lim).all()), f"max diff {d.max().item()} at ref {b.float().flatten()[d.argmax()].item()}"


def test_linear_gemm_store():
    from mmnn_sts_b200 import _lib as L
    from tests import engine_helpers as H
    torch.manual_seed(0)
    M, Cin, N = 1000, 96, 128
    a = torch.randn(M, Cin, device="cuda").to(_actdt())
    w = torch.randn(N, Cin, device="cuda") * 0.2
    bp = H.pack(w, N, 128, Cin, 64, 1, Cin, 1, 0)
    out = torch.zeros(M, N, dtype=_actdt(), device="cuda")
    H.rows(M, 128, N, Cin, 64, 1, (1, 1, M), a, Cin, bp, out, N)
    torch.cuda.synchronize()
    ref = a.float() @ _act(w).t()
    _close(out, ref)


def test_linear_bnrelu_stats_ntiles_strided():
    from mmnn_sts_b200 import _lib as L
    from tests import engine_helpers as H
    torch.manual_seed(1)
    M, Ctot, Cin, N = 777, 256, 160, 224
    buf = torch.randn(M, Ctot, device="cuda").to(_actdt())
    x = buf[:, :Cin].float()
    w = torch.randn(N, Cin, device="cuda") * 0.1
    gamma = torch.rand(Cin, device="cuda") + 0.5
    beta = torch.randn(Cin, device="cuda") * 0.3
    s1 = x.double().sum(0); s2 = (x.double() ** 2).sum(0)
    bp = H.pack(w, N, 128, Cin, 64, 1, Cin, 1, 0)
    out = torch.zeros(M, N, dtype=_actdt(), device="cuda")
    st = torch.zeros(2, N, dtype=torch.float64, device="cuda")
    H.rows(M, 128, N, Cin, 64, 1, (1, 1, M), buf, Ctot, bp, out, N, trans=L.T_BNRELU, epi=L.EP_STORE_STATS,
           bnA=H.bnsrc(s1, s2, gamma, beta, count=M), st_sum=st[0], st_sq=st[1])
    torch.cuda.synchronize()
    a = _act(F.relu(F.batch_norm(x, None, None, gamma, beta, True, 0.0, 1e-5)))
    ref = a @ _act(w).t()
    _close(out, ref, rtol=1 / 64, atol=5e-2)
    o = out.double()
    assert torch.allclose(st[0], o.sum(0), rtol=1e-5, atol=1e-3)
    assert torch.allclose(st[1], (o ** 2).sum(0), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("M", [40000 + 77, 3 * 148 * 128])
def test_persistent_1x1_fprop_and_dgrad(M):
    """1x1x1 GEMMs at >= 2 tiles per SM: the one-tile-per-CTA kernel by default, the persistent warp-specialised kernel
    (rows_persist.cuh) under MMNN_ROWS_PERSIST=1 (both are run by the round's GPU logs): forward with
    BN+ReLU prologue + statistics over 2 N tiles, and the data gradient with ReLU mask + BN-backward statistics; several
    tiles per CTA (TMEM double buffering, stage ring wrap-around), ragged last tile."""
    from mmnn_sts_b200 import _lib as L
    from tests import engine_helpers as H
    torch.manual_seed(21)
    Ctot, Cin, N = 256, 160, 224
    buf = torch.randn(M, Ctot, device="cuda").to(_actdt())
    x = buf[:, :Cin].float()
    w = torch.randn(N, Cin, device="cuda") * 0.1
    gamma = torch.rand(Cin, device="cuda") + 0.5
    beta = torch.randn(Cin, device="cuda") * 0.3
    s1 = x.double().sum(0); s2 = (x.double() ** 2).sum(0)
    bp = H.pack(w, N, 128, Cin, 64, 1, Cin, 1, 0)
    out = torch.zeros(M, N, dtype=_actdt(), device="cuda")
    st = torch.zeros(2, N, dtype=torch.float64, device="cuda")
    H.rows(M, 128, N, Cin, 64, 1, (1, 1, M), buf, Ctot, bp, out, N, trans=L.T_BNRELU, epi=L.EP_STORE_STATS,
           bnA=H.bnsrc(s1, s2, gamma, beta, count=M), st_sum=st[0], st_sq=st[1])
    torch.cuda.synchronize()
    a = _act(F.relu(F.batch_norm(x, None, None, gamma, beta, True, 0.0, 1e-5)))
    ref = a @ _act(w).t()
    _close(out, ref, rtol=1 / 64, atol=5e-2)
    o = out.double()
    assert torch.allclose(st[0], o.sum(0), rtol=1e-5, atol=1e-2)
    assert torch.allclose(st[1], (o ** 2).sum(0), rtol=1e-5, atol=1e-2)
    # data gradient of conv1: dX[m][ci] = sum_co g[m][co] W[co][ci], gated by relu(bn1(x)) > 0, + BN1-backward statistics
    Co = 128
    g = (torch.randn(M, Co, device="cuda") * 0.1).to(torch.bfloat16)
    w1 = torch.randn(Co, Cin, device="cuda") * 0.1                       # conv1.weight [co][ci]
    bpd = H.pack(w1, Cin, 128, Co, 64, 1, 1, Cin, 0, fwd=False)           # operand rows n = ci, channels = co
    dr = torch.zeros(M, Cin, dtype=torch.bfloat16, device="cuda")
    std = torch.zeros(2, Cin, dtype=torch.float64, device="cuda")
    H.rows(M, 128, Cin, Co, 64, 1, (1, 1, M), g, Co, bpd, dr, Cin, epi=L.EP_MASK_STATS, grad=1, st_sum=std[0], st_sq=std[1],
           e_src=buf, e_pitch=Ctot, bnE=H.bnsrc(s1, s2, gamma, beta, count=M))
    torch.cuda.synchronize()
    mean = x.mean(0); rstd = (x.var(0, unbiased=False) + 1e-5).rsqrt()
    xhat = (x - mean) * rstd
    pre = xhat * gamma + beta
    refd = (g.float() @ _bf(w1)) * (pre > 0)
    band = pre.abs() < 1e-3
    d = ((dr.float() - refd).abs() - (refd.abs() / 64 + 5e-2)).masked_fill(band, -1)
    assert float(d.max()) <= 0
    od = dr.double()
    assert torch.allclose(std[0], od.sum(0), rtol=1e-4, atol=5e-2)
    assert torch.allclose(std[1], (od * xhat.double()).sum(0), rtol=2e-3, atol=0.5)


def test_conv3x3x3_fprop_bnrelu_dropout_slice():
    from mmnn_sts_b200 import _lib as L
    from tests import engine_helpers as H
    torch.manual_seed(2)
    B, Dz, Dy, Dx, Cin, N, Ctot = 2, 5, 6, 8, 128, 32, 96
    M = B * Dz * Dy * Dx
    bott = torch.randn(M, Cin, device="cuda").to(_actdt())
    w = torch.randn(N, Cin, 3, 3, 3, device="cuda") * 0.05
    gamma = torch.rand(Cin, device="cuda") + 0.5
    beta = torch.randn(Cin, device="cuda") * 0.3
    rmean = torch.randn(Cin, device="cuda") * 0.2
    rvar = torch.rand(Cin, device="cuda") + 0.5
    keep = (torch.rand(B, N, device="cuda") > 0.3).float() / 0.7
    bp = H.pack(w, N, 32, Cin, 64, 27, Cin * 27, 27, 1)
    buf = torch.zeros(M, Ctot, dtype=_actdt(), device="cuda")
    st = torch.zeros(2, N, dtype=torch.float64, device="cuda")
    H.rows(M, 32, N, Cin, 64, 27, (Dz, Dy, Dx), bott, Cin, bp, buf[:, 64:], Ctot, trans=L.T_BNRELU,
           epi=L.EP_STORE_STATS, bnA=H.bnsrc(None, None, gamma, beta, rmean, rvar, use_batch=0), colscale=keep,
           st_sum=st[0], st_sq=st[1])
    torch.cuda.synchronize()
    x = bott.float().view(B, Dz, Dy, Dx, Cin).permute(0, 4, 1, 2, 3)
    a = _act(F.relu(F.batch_norm(x, rmean, rvar, gamma, beta, False, 0.0, 1e-5)))
    ref = F.conv3d(a, _act(w), padding=1) * keep[:, :, None, None, None]
    ref = ref.permute(0, 2, 3, 4, 1).reshape(M, N)
    _close(buf[:, 64:], ref, rtol=1 / 64, atol=5e-2)
    assert float(buf[:, :64].abs().max()) == 0.0
    o = buf[:, 64:].double()
    assert torch.allclose(st[0], o.sum(0), rtol=1e-5, atol=1e-3)


def test_conv3x3x3_dgrad_mask_stats():
    from mmnn_sts_b200 import _lib as L
    from tests import engine_helpers as H
    torch.manual_seed(3)
    B, Dz, Dy, Dx, Cg, N = 2, 4, 5, 8, 32, 128
    M = B * Dz * Dy * Dx
    g = torch.randn(M, Cg, device="cuda").to(torch.bfloat16)
    w = torch.randn(Cg, N, 3, 3, 3, device="cuda") * 0.05      # conv2.weight [co=32][ci=128][27]
    xb = torch.randn(M, N, device="cuda").to(_actdt())          # bottleneck (BN2 input), forward activation format
    gamma = torch.rand(N, device="cuda") + 0.5
    beta = torch.randn(N, device="cuda") * 0.3
    s1 = xb.double().sum(0); s2 = (xb.double() ** 2).sum(0)
    # dgrad operand: n = ci, channel = co, tap
    bp = H.pack(w, N, 128, Cg, 32, 27, 27, N * 27, 1, fwd=False)
    out = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    st = torch.zeros(2, N, dtype=torch.float64, device="cuda")
    H.rows(M, 128, N, Cg, 32, 27, (Dz, Dy, Dx), g, Cg, bp, out, N, epi=L.EP_MASK_STATS, tap_sign=-1, grad=1,
           st_sum=st[0], st_sq=st[1], e_src=xb, e_pitch=N, bnE=H.bnsrc(s1, s2, gamma, beta, count=M))
    torch.cuda.synchronize()
    g5 = g.float().view(B, Dz, Dy, Dx, Cg).permute(0, 4, 1, 2, 3)
    dA = F.conv_transpose3d(g5, _bf(w), padding=1).permute(0, 2, 3, 4, 1).reshape(M, N)
    mean = xb.float().mean(0); var = xb.float().var(0, unbiased=False); rstd = (var + 1e-5).rsqrt()
    xhat = (xb.float() - mean) * rstd
    act = (xhat * gamma + beta) > 0
    ref = dA * act
    # elements whose pre-activation is within rounding of zero may flip: exclude a thin band
    band = (xhat * gamma + beta).abs() < 1e-3
    d = ((out.float() - ref).abs() - (ref.abs() / 64 + 5e-2)).masked_fill(band, -1)
    assert float(d.max()) <= 0
    o = out.double()
    assert torch.allclose(st[0], o.sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(st[1], (o * xhat.double()).sum(0), rtol=1e-3, atol=5e-2)


def test_stem_conv7_s2():
    from mmnn_sts_b200 import _lib as L
    from tests import engine_helpers as H
    torch.manual_seed(4)
    for cin in (1, 2):
        B, X, Y, Z = 2, 24, 20, 16
        img = torch.rand(B, cin, X, Y, Z, device="cuda")
        w = torch.randn(64, cin, 7, 7, 7, device="cuda") * 0.05
        Dz, Dy, Dx = (X - 1) // 2 + 1, (Y - 1) // 2 + 1, (Z - 1) // 2 + 1
        Sz, Sy, Sx = Dz + 3, Dy + 3, Dx + 3
        # reference-side construction of the padded space-to-depth input [B][Sz][Sy][Sx][(pz,py,px,c2)]
        pad = torch.zeros(B, 2, 2 * Sz, 2 * Sy, 2 * Sx, device="cuda")
        pad[:, :cin, 3:3 + X, 3:3 + Y, 3:3 + Z] = img
        s2d = pad.view(B, 2, Sz, 2, Sy, 2, Sx, 2).permute(0, 2, 4, 6, 3, 5, 7, 1).contiguous().to(_actdt())
        s2d = torch.cat([s2d.view(-1), torch.zeros(64, dtype=_actdt(), device="cuda")])  # tail slack for the 4-voxel rows
        M = B * Dz * Dy * Dx
        bp = H.pack(w, 64, 64, 64, 64, 16, 0, 0, 0, mode=L.PACK_STEM, cin_real=cin)
        out = torch.zeros(M, 64, dtype=_actdt(), device="cuda")
        st = torch.zeros(2, 64, dtype=torch.float64, device="cuda")
        H.rows(M, 64, 64, 64, 64, 16, (Dz, Dy, Dx), s2d, 16, bp, out, 64, amode=L.A_STEM, epi=L.EP_STORE_STATS,
               sdims=(Sz, Sy, Sx), st_sum=st[0], st_sq=st[1])
        torch.cuda.synchronize()
        ref = F.conv3d(_act(img), _act(w), stride=2, padding=3).permute(0, 2, 3, 4, 1).reshape(M, 64)
        _close(out, ref, rtol=1 / 64, atol=5e-2)
        assert torch.allclose(st[0], out.double().sum(0), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("shape", [(2, 24, 20, 16), (3, 16, 64, 32), (5, 40, 36, 24)])
def test_stem_brick_conv7_s2(shape):
    """Brick-mode stem (stem.cuh) against F.conv3d on the same rounded operands: partial tiles in y and x, more tiles
    than SMs (persistent loop, 3-deep brick ring, TMEM double buffering), register-accumulated statistics."""
    from mmnn_sts_b200 import _lib as L
    from tests import engine_helpers as H
    torch.manual_seed(14)
    for cin in (1, 2):
        B, X, Y, Z = shape
        img = torch.rand(B, cin, X, Y, Z, device="cuda")
        w = torch.randn(64, cin, 7, 7, 7, device="cuda") * 0.05
        Dz, Dy, Dx = (X - 1) // 2 + 1, (Y - 1) // 2 + 1, (Z - 1) // 2 + 1
        Sz, Sy, Sx = Dz + 3, Dy + 3, Dx + 3
        pad = torch.zeros(B, 2, 2 * Sz, 2 * Sy, 2 * Sx, device="cuda")
        pad[:, :cin, 3:3 + X, 3:3 + Y, 3:3 + Z] = img
        s2d = pad.view(B, 2, Sz, 2, Sy, 2, Sx, 2).permute(0, 2, 4, 6, 3, 5, 7, 1).contiguous().to(_actdt())
        M = B * Dz * Dy * Dx
        bp = H.pack(w, 64, 64, 64, 64, 16, 0, 0, 0, mode=L.PACK_STEM_SW32, cin_real=cin)   # the brick kernel's swizzled weight image
        out = torch.zeros(M, 64, dtype=_actdt(), device="cuda")
        st = torch.zeros(2, 64, dtype=torch.float64, device="cuda")
        H.stem_brick(B, (Dz, Dy, Dx), s2d, bp, out, 64, st_sum=st[0], st_sq=st[1])
        torch.cuda.synchronize()
        ref = F.conv3d(_act(img), _act(w), stride=2, padding=3).permute(0, 2, 3, 4, 1).reshape(M, 64)
        _close(out, ref, rtol=1 / 64, atol=5e-2)
        assert (out.float() - ref).abs().max() <= 2e-3 * ref.abs().max() + 1e-3, (out.float() - ref).abs().max()
        assert torch.allclose(st[0], out.double().sum(0), rtol=1e-5, atol=1e-3)
        assert torch.allclose(st[1], (out.double() ** 2).sum(0), rtol=1e-5, atol=1e-3)


def test_wgrad_conv1x1():
    from tests import engine_helpers as H
    torch.manual_seed(5)
    M, Ctot, Cin, Co = 1300, 256, 160, 128
    buf = torch.randn(M, Ctot, device="cuda").to(_actdt())
    dbott = (torch.randn(M, Co, device="cuda") * 0.1).to(torch.bfloat16)
    gamma = torch.rand(Cin, device="cuda") + 0.5
    beta = torch.randn(Cin, device="cuda") * 0.3
    x = buf[:, :Cin].float()
    s1 = x.double().sum(0); s2 = (x.double() ** 2).sum(0)
    dw = torch.zeros(Co, Cin, device="cuda")
    for split in (1, 5):
        dw.zero_()
        H.wgrad(0, M, 128, 1, Cin, Co, (1, 1, M), buf, Ctot, dbott, Co, dw, 1, Cin, bnA=H.bnsrc(s1, s2, gamma, beta, count=M), split=split)
        torch.cuda.synchronize()
        a = _bf(F.relu(F.batch_norm(x, None, None, gamma, beta, True, 0.0, 1e-5)))
        ref = dbott.float().t() @ a
        assert (dw - ref).abs().max() <= 2e-3 * ref.abs().max() + 1e-3, (dw - ref).abs().max()


@pytest.mark.parametrize("B,Dz,Dy,Dx", [
    (2, 4, 6, 8),       # tiles are not boxes: register path
    (3, 3, 16, 16),     # box 16 x 8 of one z slice: halo boxes (engine.cuh tma_b == 2), several tiles per slice
    (2, 5, 16, 8),      # box 8 x 16
    (2, 2, 4, 32),      # box 32 x 4 = a whole slice: every dy tap leaves the volume on one side
    (5, 4, 8, 4),       # box spans z slices: nine boxes per stage (tma_b == 1)
])
def test_wgrad_conv3x3x3(B, Dz, Dy, Dx):
    from tests import engine_helpers as H
    torch.manual_seed(6)
    Cb, Cg = 128, 32
    M = B * Dz * Dy * Dx
    bott = torch.randn(M, Cb, device="cuda").to(_actdt())
    g = (torch.randn(M, Cg, device="cuda") * 0.1).to(torch.bfloat16)
    gamma = torch.rand(Cb, device="cuda") + 0.5
    beta = torch.randn(Cb, device="cuda") * 0.3
    s1 = bott.double().sum(0); s2 = (bott.double() ** 2).sum(0)
    dwt = torch.zeros(27, Cg, Cb, device="cuda")   # scratch layout [tap][co][ci]
    H.wgrad(1, M, 32, 9, Cb, Cg, (Dz, Dy, Dx), bott, Cb, g, Cg, dwt, 1, Cb, Cg * Cb, bnA=H.bnsrc(s1, s2, gamma, beta, count=M))
    torch.cuda.synchronize()
    x5 = bott.float().view(B, Dz, Dy, Dx, Cb).permute(0, 4, 1, 2, 3)
    a = _bf(F.relu(F.batch_norm(x5, None, None, gamma, beta, True, 0.0, 1e-5))).requires_grad_(False)
    w = torch.zeros(Cg, Cb, 3, 3, 3, device="cuda", requires_grad=True)
    y = F.conv3d(a, w, padding=1)
    g5 = g.float().view(B, Dz, Dy, Dx, Cg).permute(0, 4, 1, 2, 3)
    (ref,) = torch.autograd.grad(y, w, g5)
    got = dwt.view(27, Cg, Cb).permute(1, 2, 0).reshape(Cg, Cb, 3, 3, 3)
    assert (got - ref).abs().max() <= 2e-3 * ref.abs().max() + 1e-3, (got - ref).abs().max()


def test_wgrad_raw_and_stem():
    from tests import engine_helpers as H
    torch.manual_seed(7)
    # raw x raw (transition): dW[co][ci] = sum_m g[m][co] * pooled[m][ci];  A = pooled (ci), B = g (co)
    M, C, Co = 900, 256, 128
    pooled = torch.randn(M, C, device="cuda").to(_actdt())
    g = (torch.randn(M, Co, device="cuda") * 0.1).to(torch.bfloat16)
    dw = torch.zeros(Co, C, device="cuda")
    H.wgrad(2, M, 128, 1, C, Co, (1, 1, M), pooled, C, g, Co, dw, 1, C)
    torch.cuda.synchronize()
    ref = g.float().t() @ _bf(pooled.float())
    assert (dw - ref).abs().max() <= 2e-3 * ref.abs().max() + 1e-3
    # stem
    for cin in (1, 2):
        B, X, Y, Z = 2, 24, 20, 16
        img = torch.rand(B, cin, X, Y, Z, device="cuda")
        Dz, Dy, Dx = (X - 1) // 2 + 1, (Y - 1) // 2 + 1, (Z - 1) // 2 + 1
        Sz, Sy, Sx = Dz + 3, Dy + 3, Dx + 3
        pad = torch.zeros(B, 2, 2 * Sz, 2 * Sy, 2 * Sx, device="cuda")
        pad[:, :cin, 3:3 + X, 3:3 + Y, 3:3 + Z] = img
        s2d = pad.view(B, 2, Sz, 2, Sy, 2, Sx, 2).permute(0, 2, 4, 6, 3, 5, 7, 1).contiguous().to(_actdt())
        s2d = torch.cat([s2d.view(-1), torch.zeros(64, dtype=_actdt(), device="cuda")])
        M0 = B * Dz * Dy * Dx
        dconv = (torch.randn(M0, 64, device="cuda") * 0.1).to(torch.bfloat16)
        dw0 = torch.zeros(64, cin, 7, 7, 7, device="cuda")
        H.wgrad(3, M0, 64, 1, 128, 64, (Dz, Dy, Dx), s2d, 16, dconv, 64, dw0, 0, 0, sdims=(Sz, Sy, Sx), cin_real=cin)
        torch.cuda.synchronize()
        w = torch.zeros(64, cin, 7, 7, 7, device="cuda", requires_grad=True)
        y = F.conv3d(_bf(_act(img)), w, stride=2, padding=3)
        (ref,) = torch.autograd.grad(y, w, dconv.float().view(B, Dz, Dy, Dx, 64).permute(0, 4, 1, 2, 3))
        assert (dw0 - ref).abs().max() <= 2e-3 * ref.abs().max() + 1e-3, (dw0 - ref).abs().max()


@pytest.mark.parametrize("cin,shape,B", [(2, (16, 16, 32), 3), (1, (32, 16, 64), 2), (2, (8, 8, 8), 5)])
def test_stem_wgrad_all_tma_matches_register_path(cin, shape, B):
    """Stem weight gradient with BOTH operands by TMA (bf16 space-to-depth image viewed as overlapping 64-element rows,
    engine.cuh tma_a): volumes whose 128-voxel tiles are boxes, incl. one whose box spans several samples.  Checked against
    autograd on the same rounded operands and, bit for bit, against the register path fed the same bf16 image."""
    from tests import engine_helpers as H
    torch.manual_seed(17)
    X, Y, Z = shape
    img = torch.rand(B, cin, X, Y, Z, device="cuda")
    Dz, Dy, Dx = (X - 1) // 2 + 1, (Y - 1) // 2 + 1, (Z - 1) // 2 + 1
    Sz, Sy, Sx = Dz + 3, Dy + 3, Dx + 3
    pad = torch.zeros(B, 2, 2 * Sz, 2 * Sy, 2 * Sx, device="cuda")
    pad[:, :cin, 3:3 + X, 3:3 + Y, 3:3 + Z] = img
    s2d = pad.view(B, 2, Sz, 2, Sy, 2, Sx, 2).permute(0, 2, 4, 6, 3, 5, 7, 1).contiguous().to(_actdt()).to(torch.bfloat16)
    s2d = torch.cat([s2d.view(-1), torch.zeros(64, dtype=torch.bfloat16, device="cuda")])
    M0 = B * Dz * Dy * Dx
    dconv = (torch.randn(M0, 64, device="cuda") * 0.1).to(torch.bfloat16)
    got = []
    for a_bf16 in (1, 2):      # 1: bf16 image, operand by TMA; 2: bf16 image through the register path
        dw0 = torch.zeros(64, cin, 7, 7, 7, device="cuda")
        H.wgrad(3, M0, 64, 1, 128, 64, (Dz, Dy, Dx), s2d, 16, dconv, 64, dw0, 0, 0, sdims=(Sz, Sy, Sx), cin_real=cin,
                a_bf16=a_bf16, split=1)
        torch.cuda.synchronize()
        got.append(dw0)
    w = torch.zeros(64, cin, 7, 7, 7, device="cuda", requires_grad=True)
    y = F.conv3d(_bf(_act(img)), w, stride=2, padding=3)
    (ref,) = torch.autograd.grad(y, w, dconv.float().view(B, Dz, Dy, Dx, 64).permute(0, 4, 1, 2, 3))
    assert (got[0] - ref).abs().max() <= 2e-3 * ref.abs().max() + 1e-3, (got[0] - ref).abs().max()
    assert torch.equal(got[0], got[1])


@pytest.mark.parametrize("dims,B", [((3, 20, 12), 2), ((4, 16, 8), 3), ((2, 32, 16), 40)])
def test_brick_conv3_fprop(dims, B):
    """Brick-mode 3x3x3 forward (BN+ReLU prologue, dropout scale, statistics) incl. partial tiles and a grid with
    more tiles than SMs (persistent loop, TMEM double buffering)."""
    from mmnn_sts_b200 import _lib as L
    from tests import engine_helpers as H
    torch.manual_seed(8)
    Dz, Dy, Dx = dims
    Cin, N, Ctot = 128, 32, 96
    M = B * Dz * Dy * Dx
    bott = torch.randn(M, Cin, device="cuda").to(_actdt())
    w = torch.randn(N, Cin, 3, 3, 3, device="cuda") * 0.05
    gamma = torch.rand(Cin, device="cuda") + 0.5
    beta = torch.randn(Cin, device="cuda") * 0.3
    x = bott.float()
    s1 = x.double().sum(0); s2 = (x.double() ** 2).sum(0)
    keep = (torch.rand(B, N, device="cuda") > 0.3).float() / 0.7
    bp = H.pack(w, N, 32, Cin, 64, 27, Cin * 27, 27, 1)
    buf = torch.zeros(M, Ctot, dtype=_actdt(), device="cuda")
    st = torch.zeros(2, N, dtype=torch.float64, device="cuda")
    H.brick(B, dims, Cin, N, bott, Cin, bp, buf[:, 64:], Ctot, bnA=H.bnsrc(s1, s2, gamma, beta, count=M), colscale=keep,
            st_sum=st[0], st_sq=st[1])
    torch.cuda.synchronize()
    x5 = x.view(B, Dz, Dy, Dx, Cin).permute(0, 4, 1, 2, 3)
    a = _act(F.relu(F.batch_norm(x5, None, None, gamma, beta, True, 0.0, 1e-5)))
    ref = (F.conv3d(a, _act(w), padding=1) * keep[:, :, None, None, None]).permute(0, 2, 3, 4, 1).reshape(M, N)
    _close(buf[:, 64:], ref, rtol=1 / 128, atol=2e-2)
    assert float(buf[:, :64].abs().max()) == 0.0
    o = buf[:, 64:].double()
    assert torch.allclose(st[0], o.sum(0), rtol=1e-5, atol=1e-3)
    assert torch.allclose(st[1], (o ** 2).sum(0), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("dims,B", [((3, 20, 12), 2), ((2, 32, 16), 40)])
def test_brick_conv3_dgrad(dims, B):
    from mmnn_sts_b200 import _lib as L
    from tests import engine_helpers as H
    torch.manual_seed(9)
    Dz, Dy, Dx = dims
    Cg, N = 32, 128
    M = B * Dz * Dy * Dx
    g = torch.randn(M, Cg, device="cuda").to(torch.bfloat16)
    w = torch.randn(Cg, N, 3, 3, 3, device="cuda") * 0.05
    xb = torch.randn(M, N, device="cuda").to(_actdt())
    gamma = torch.rand(N, device="cuda") + 0.5
    beta = torch.randn(N, device="cuda") * 0.3
    s1 = xb.double().sum(0); s2 = (xb.double() ** 2).sum(0)
    bp = H.pack(w, N, 128, Cg, 32, 27, 27, N * 27, 1, fwd=False)
    out = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    st = torch.zeros(2, N, dtype=torch.float64, device="cuda")
    H.brick(B, dims, Cg, N, g, Cg, bp, out, N, grad=1, tap_sign=-1, st_sum=st[0], st_sq=st[1], e_src=xb, e_pitch=N,
            bnE=H.bnsrc(s1, s2, gamma, beta, count=M))
    torch.cuda.synchronize()
    g5 = g.float().view(B, Dz, Dy, Dx, Cg).permute(0, 4, 1, 2, 3)
    dA = F.conv_transpose3d(g5, _bf(w), padding=1).permute(0, 2, 3, 4, 1).reshape(M, N)
    mean = xb.float().mean(0); var = xb.float().var(0, unbiased=False); rstd = (var + 1e-5).rsqrt()
    xhat = (xb.float() - mean) * rstd
    pre = xhat * gamma + beta
    ref = dA * (pre > 0)
    band = pre.abs() < 1e-3
    d = ((out.float() - ref).abs() - (ref.abs() / 64 + 5e-2)).masked_fill(band, -1)
    assert float(d.max()) <= 0
    o = out.double()
    assert torch.allclose(st[0], o.sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(st[1], (o * xhat.double()).sum(0), rtol=1e-3, atol=5e-2)
