"""Test-side helpers that drive the fine-grained C-ABI entry points (mmnn_conv_rows / mmnn_pack_weights) on torch
CUDA tensors.  The product path calls the same kernels from the C++ encoder orchestrator."""
import ctypes as C

import torch

from mmnn_sts_b200 import _lib as L


def act_dtype():
    """Storage dtype of forward activations / forward weight images in this build (fp16 by default)."""
    return torch.float16 if L.lib().mmnn_act_is_fp16() else torch.bfloat16


GRD = torch.bfloat16


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def pack(src_f32, N, NT, Cin, kbw, ntaps, sn, sc, st, mode=L.PACK_GENERIC, cin_real=0, fwd=True):
    dst = torch.empty(L.packed_elems(N, NT, Cin, kbw, ntaps), dtype=torch.bfloat16, device="cuda")
    f16 = 1 if (fwd and act_dtype() == torch.float16) else 0
    d = L.PackDesc(ptr(src_f32), ptr(dst), N, NT, Cin, kbw, ntaps, mode, cin_real, f16, sn, sc, st)
    scratch = torch.empty(C.sizeof(L.PackDesc), dtype=torch.uint8, device="cuda")
    L.check(L.lib().mmnn_pack_weights(C.byref(d), 1, ptr(scratch), stream_ptr()), "pack")
    torch.cuda.synchronize()
    return dst


def bnsrc(sum_=None, sumsq=None, gamma=None, beta=None, rmean=None, rvar=None, count=1, eps=1e-5, use_batch=1):
    return L.BnSrc(ptr(sum_), ptr(sumsq), ptr(gamma), ptr(beta), ptr(rmean), ptr(rvar), 1.0 / count, eps, use_batch)


def rows(M, NT, Ncols, Cin, kbw, ntaps, dims, a_src, a_pitch, b_packed, out, out_pitch, amode=L.A_LINEAR_CONV,
         trans=L.T_NONE, epi=L.EP_STORE, tap_sign=1, sdims=(0, 0, 0), bnA=None, colscale=None, st_sum=None, st_sq=None,
         e_src=None, e_pitch=0, bnE=None, stages=0, grad=0):
    p = L.RowsParams()
    p.M, p.NT, p.Ncols, p.Cin, p.kbw, p.ntaps, p.tap_sign = M, NT, Ncols, Cin, kbw, ntaps, tap_sign
    p.Dz, p.Dy, p.Dx = dims
    p.Sz, p.Sy, p.Sx = sdims
    p.a_src, p.a_pitch = a_src.data_ptr(), a_pitch
    p.bnA = bnA if bnA is not None else L.BnSrc()
    p.b_packed = b_packed.data_ptr()
    p.out, p.out_pitch = out.data_ptr(), out_pitch
    p.colscale = colscale.data_ptr() if colscale is not None else None
    p.st_sum = st_sum.data_ptr() if st_sum is not None else None
    p.st_sq = st_sq.data_ptr() if st_sq is not None else None
    p.e_src = e_src.data_ptr() if e_src is not None else None
    p.e_pitch = e_pitch
    p.bnE = bnE if bnE is not None else L.BnSrc()
    p.stages = stages
    L.check(L.lib().mmnn_conv_rows(C.byref(p), amode, trans, epi, grad, stream_ptr()), "conv_rows")


def wgrad(kind, M, CB, NB, na_total, nb_total, dims, a_src, a_pitch, b_src, b_pitch, dw, so_a, so_b, so_j=0,
          bnA=None, bnB=None, sdims=(0, 0, 0), cin_real=0, split=0, stages=0, a_bf16=0):
    p = L.WgradParams()
    p.M, p.CB, p.NB, p.na_total, p.nb_total = M, CB, NB, na_total, nb_total
    p.Dz, p.Dy, p.Dx = dims
    p.Sz, p.Sy, p.Sx = sdims
    p.a_src, p.a_pitch = a_src.data_ptr(), a_pitch
    p.bnA = bnA if bnA is not None else L.BnSrc()
    p.b_src, p.b_pitch = b_src.data_ptr(), b_pitch
    p.bnB = bnB if bnB is not None else L.BnSrc()
    p.dw, p.so_a, p.so_b, p.so_j = dw.data_ptr(), so_a, so_b, so_j
    p.cin_real, p.stages = cin_real, stages
    p.a_bf16 = a_bf16
    L.check(L.lib().mmnn_conv_wgrad(C.byref(p), kind, split, stream_ptr()), "conv_wgrad")


def brick(B, dims, CH, NT, a_src, a_pitch, b_packed, out, out_pitch, grad=0, tap_sign=1, bnA=None, colscale=None,
          st_sum=None, st_sq=None, e_src=None, e_pitch=0, bnE=None):
    p = L.BrickParams()
    p.B = B
    p.Dz, p.Dy, p.Dx = dims
    p.CH, p.NT, p.tap_sign = CH, NT, tap_sign
    p.a_src, p.a_pitch = a_src.data_ptr(), a_pitch
    p.bnA = bnA if bnA is not None else L.BnSrc()
    p.b_packed = b_packed.data_ptr()
    p.out, p.out_pitch = out.data_ptr(), out_pitch
    p.colscale = colscale.data_ptr() if colscale is not None else None
    p.st_sum = st_sum.data_ptr() if st_sum is not None else None
    p.st_sq = st_sq.data_ptr() if st_sq is not None else None
    p.e_src = e_src.data_ptr() if e_src is not None else None
    p.e_pitch = e_pitch
    p.bnE = bnE if bnE is not None else L.BnSrc()
    L.check(L.lib().mmnn_conv3_brick(C.byref(p), grad, stream_ptr()), "conv3_brick")


def stem_brick(B, dims, xs2d, w_packed, out, out_pitch, st_sum=None, st_sq=None):
    p = L.StemBrickParams()
    p.B = B
    p.D0, p.H0, p.W0 = dims
    p.Sz, p.Sy, p.Sx = dims[0] + 3, dims[1] + 3, dims[2] + 3
    p.xs2d, p.w_packed = xs2d.data_ptr(), w_packed.data_ptr()
    p.out, p.out_pitch = out.data_ptr(), out_pitch
    p.st_sum = st_sum.data_ptr() if st_sum is not None else None
    p.st_sq = st_sq.data_ptr() if st_sq is not None else None
    L.check(L.lib().mmnn_stem_brick(C.byref(p), stream_ptr()), "stem_brick")
