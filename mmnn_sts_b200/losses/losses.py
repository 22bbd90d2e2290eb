"""Cox partial-likelihood loss with the reference's name, argument order and (swapped) semantics
(/root/reference/losses/losses.py:6-9 -> pycox CoxPHLoss, restated in oracle/cox.py)."""
import torch

from ..ops import bce_with_logits, cox_ph_segments


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


class BCEWithLogitsLoss(torch.nn.Module):
    """`nn.BCEWithLogitsLoss(pos_weight=..., reduction='sum' | 'none' | 'mean')` as the reference's classification path
    builds it (/root/reference/main.py:148-153) on the fused kernel.  Targets may have fewer leading dimensions than the
    logits (GradientBlender stacks the same [N, C] targets for each of the k+1 heads,
    /root/reference/losses/GradientBlender.py:166-168): they are reused per head without materialising the stack.
    `last_counts` holds the tp / fp / fn counters of the FIRST head (main.py:226-229) when `count_threshold` is set."""

    def __init__(self, pos_weight=None, reduction="mean", count_threshold=None):
        super().__init__()
        self.register_buffer("pos_weight", pos_weight if pos_weight is None else torch.as_tensor(pos_weight, dtype=torch.float32))
        self.reduction = reduction
        self.count_threshold = count_threshold
        self.last_counts = None

    def forward(self, logits, targets):
        t = targets
        while t.dim() > 2 and t.shape[0] == logits.shape[0] and t.dim() == logits.dim():
            # an explicitly stacked target tensor [H, N, C] with identical slices is what the reference passes; accept it
            if logits.dim() == 2:
                break
            t = t[0]
        counts = None
        if self.count_threshold is not None:
            counts = torch.zeros((3, logits.shape[-1]), dtype=torch.int32, device=logits.device)
        loss = bce_with_logits(logits, t, self.pos_weight, self.count_threshold if self.count_threshold is not None else 0.5, counts)
        self.last_counts = counts
        if self.reduction == "sum":
            return loss.sum()
        if self.reduction == "mean":
            return loss.mean()
        return loss
