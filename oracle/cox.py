"""Cox partial-likelihood loss -- restatement of pycox.models.loss (third-party; NOT vendored in the reference).

TEST INFRASTRUCTURE.  The reference's call site is /root/reference/losses/losses.py:3,6-9
(`CoxPHLoss()(log_h, events, duration)`); pycox is unpinned in /root/reference/requirements.txt:16.
Restated from the published pycox algorithm (pycox.models.loss.cox_ph_loss / cox_ph_loss_sorted),
SURVEY.md appendix B.1:

    idx  = durations.sort(descending=True)[1]
    e, h = events[idx], log_h[idx]
    g    = h.max()
    lcs  = log(cumsum(exp(h - g)) + eps) + g          # eps inside the gamma-shifted log
    loss = -sum((h - lcs) * e) / sum(e)

Quirk Q1 (kept on purpose): the reference passes (log_h, events, duration) into pycox's
(log_h, durations, events) slots, so as shipped the sort key is the 0/1 event flag and the per-row
weight is the duration.  `CoxPH` below reproduces that call order verbatim.
Quirk Q2: torch.sort is not stable; `stable=True` here fixes the tie order to original index order, which is
what the CUDA kernel implements.  Tie-free inputs are identical under both.
"""
import numpy as np
import torch


def cox_ph_loss_sorted(log_h, events, eps=1e-7):
    if events.dtype is torch.bool:
        events = events.float()
    events = events.view(-1)
    log_h = log_h.view(-1)
    gamma = log_h.max()
    log_cumsum_h = log_h.sub(gamma).exp().cumsum(0).add(eps).log().add(gamma)
    return -log_h.sub(log_cumsum_h).mul(events).sum().div(events.sum())


def cox_ph_loss(log_h, durations, events, eps=1e-7, stable=True, perm=None):
    if perm is None:
        perm = durations.sort(descending=True, stable=stable)[1]
    return cox_ph_loss_sorted(log_h[perm], events[perm], eps)


class CoxPHLoss(torch.nn.Module):
    """pycox.models.loss.CoxPHLoss: forward(log_h, durations, events)."""

    def forward(self, log_h, durations, events):
        return cox_ph_loss(log_h, durations, events)


def CoxPH(log_h, events, duration):
    """As written at /root/reference/losses/losses.py:6-9 (argument order swapped into pycox, quirk Q1)."""
    return CoxPHLoss()(log_h, events, duration)


def surv_criterion(loss_func, preds, events, durations, device="cpu"):
    """/root/reference/utils/utils.py:24-29 -- SUM over classes."""
    losses = 0
    for i in range(preds.shape[1]):
        losses = losses + loss_func(preds[:, i], events[:, i], durations[:, i]).to(device)
    return losses


def cox_np(log_h, sort_key, weight, eps=1e-7, dtype=np.float64):
    """NumPy restatement with explicit roles: rows sorted by `sort_key` descending (stable), weighted by `weight`.
    Returns (loss, dloss/dlog_h) treating gamma as a constant (SURVEY.md appendix B.1 gradient formula)."""
    h = np.asarray(log_h, dtype=dtype)
    key = np.asarray(sort_key)
    w = np.asarray(weight, dtype=dtype)
    n = h.shape[0]
    order = np.argsort(-key.astype(np.float64), kind="stable")
    hs, ws = h[order], w[order]
    g = hs.max()
    p = np.exp(hs - g)
    S = np.cumsum(p) + dtype(eps)
    lcs = np.log(S) + g
    W = ws.sum()
    loss = -np.sum((hs - lcs) * ws) / W
    # d/dh_j = -(1/W) [ w_j - p_j * sum_{i>=j} w_i / S_i ]
    tail = np.cumsum((ws / S)[::-1])[::-1]
    gs = -(ws - p * tail) / W
    grad = np.zeros(n, dtype=dtype)
    grad[order] = gs
    return loss, grad
