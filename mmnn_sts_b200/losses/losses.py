"""Cox partial-likelihood loss with the reference's name, argument order and (swapped) semantics
(/root/reference/losses/losses.py:6-9 -> pycox CoxPHLoss, restated in oracle/cox.py)."""
import torch

from ..ops import cox_ph_segments


def CoxPH(log_h, events, duration):
    """As shipped, the reference calls pycox's CoxPHLoss()(log_h, events, duration) whose signature is
    (log_h, durations, events): the sort key is `events` and the row weight is `duration` (quirk Q1, kept).
    Ties of the (binary) sort key keep their original order (stable), see DESIGN.md."""
    return cox_ph_segments(log_h.reshape(1, -1), events.reshape(1, -1), duration.reshape(1, -1))[0]


def CoxPH_intended(log_h, events, duration):
    """The evidently intended call (sort by duration, weight by event) -- not what the reference runs."""
    return cox_ph_segments(log_h.reshape(1, -1), duration.reshape(1, -1), events.reshape(1, -1))[0]


def _coxph_columns(preds, events, durations):
    """CoxPH of every class column (and every stacked head) in one launch. preds [..., B, C] -> losses [..., C]."""
    lead = preds.shape[:-2]
    B, Cc = preds.shape[-2:]
    h = preds.reshape(-1, B, Cc).transpose(1, 2).reshape(-1, B)          # [H*C, B]
    H = h.shape[0] // Cc
    key = events.t().repeat(H, 1)
    w = durations.t().repeat(H, 1)
    return cox_ph_segments(h, key, w).reshape(*lead, Cc)
