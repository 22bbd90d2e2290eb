"""torch.autograd wrappers around the small C-ABI kernels (feature head, clinical MLP + fusion heads, Cox loss).
Thin by design: tensors in, raw pointers + the current CUDA stream out.  No CPU path."""
import ctypes as C

import torch

from . import _lib as L


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return t.data_ptr() if t is not None else None


def _require_cuda(t, what):
    if not t.is_cuda:
        raise L.MMNNLibraryError(f"{what}: mmnn_sts_b200 has no CPU path (got a {t.device} tensor)")


class GapLinearDropout(torch.autograd.Function):
    """ReLU -> AdaptiveAvgPool3d(1) -> Flatten -> Linear -> Dropout  (/root/reference/models/densenet.py:234-247)."""

    @staticmethod
    def run(x, weight, bias, p, injected_mask=None):
        mask = None
        if injected_mask is not None:
            mask = injected_mask.to(device=x.device, dtype=torch.float32).contiguous()
        elif p > 0:
            mask = torch.bernoulli(torch.full((x.shape[0], weight.shape[0]), 1.0 - p, device=x.device)) / (1.0 - p)
        return GapLinearDropout.apply(x, weight, bias, mask)

    @staticmethod
    def forward(ctx, x, weight, bias, mask):
        _require_cuda(x, "features")
        B, Cc = x.shape[0], x.shape[1]
        y = x.permute(0, 2, 3, 4, 1).contiguous().float()       # NDHWC; a no-op for the backbone's own output
        V = y.numel() // (B * Cc)
        F = weight.shape[0]
        w = weight.contiguous().float()
        pooled = torch.empty((B, Cc), dtype=torch.float32, device=x.device)
        out = torch.empty((B, F), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            L.check(L.lib().mmnn_gap_linear_fwd(_p(y), B, V, Cc, _p(w), _p(bias), _p(mask), F, _p(pooled), _p(out), _stream()),
                    "mmnn_gap_linear_fwd")
        ctx.save_for_backward(y, pooled, w, mask)
        ctx.dims = (B, V, Cc, F, tuple(x.shape))
        return out

    @staticmethod
    def backward(ctx, dout):
        y, pooled, w, mask = ctx.saved_tensors
        B, V, Cc, F, xshape = ctx.dims
        dout = dout.contiguous().float()
        dy = torch.empty_like(y)
        dW = torch.empty_like(w)
        db = torch.empty((F,), dtype=torch.float32, device=y.device)
        with torch.cuda.device(y.device):
            L.check(L.lib().mmnn_gap_linear_bwd(_p(y), _p(pooled), B, V, Cc, _p(w), _p(dout), _p(mask), F, _p(dy), _p(dW),
                                                _p(db), _stream()), "mmnn_gap_linear_bwd")
        return dy.permute(0, 4, 1, 2, 3), dW, db, None
