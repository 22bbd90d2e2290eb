"""GradientBlender with the reference's constructor, method names and update rule
(/root/reference/losses/GradientBlender.py:9-256; survival branch :48-103,181-205, classification branch :105-179).  Differences from the reference,
all on purpose: the per-head / per-class Cox losses of one step are ONE kernel launch instead of six, and the weights
live on the predictions' device so a step has no device->host synchronisation (quirk Q5)."""
import numpy as np
import torch
import torch.nn.functional as F


class GradientBlender:
    def __init__(self, loss_function, survival=False, reduction="sum", device="cpu", surv_criterion=None):
        self.loss_function = loss_function
        self.weights = None
        self.reduction = reduction.lower()
        self.survival = survival
        self.lvn = None
        self.ltn = None
        self.device = device
        self.surv_criterion = surv_criterion
        self.history = []

    def _set_weights(self, new):
        """New blending weights.  Once the weights live on the device they are updated IN PLACE: a training step captured
        in a CUDA graph (mmnn_sts_b200.graph) keeps reading this tensor's storage, so rebinding it would freeze the
        replayed step at the weights of capture time."""
        new = new.detach()
        w = self.weights
        if w is not None and w.is_cuda and w.shape == new.shape:
            w.copy_(new.to(device=w.device, dtype=w.dtype))
        else:
            self.weights = new

    # ---- survival
    def _head_losses(self, preds, events, durations):
        from .losses import CoxPH, _coxph_columns
        from ..utils.utils import surv_criterion as fused
        if self.loss_function is CoxPH and self.surv_criterion is fused and preds.is_cuda:
            return _coxph_columns(preds, events, durations).sum(dim=-1)          # [heads]
        return torch.stack([self.surv_criterion(self.loss_function, preds[i, ...], events, durations, preds.device)
                            for i in range(preds.shape[0])], dim=0)

    def computeLossSurv(self, preds, events, durations, reduceToHeads=False):
        head_losses = self._head_losses(preds, events, durations)
        if self.weights is None:
            self.weights = self.normalize(torch.ones(preds.shape[0]))
        if reduceToHeads:
            return head_losses
        if self.weights.device != head_losses.device or self.weights.dtype != head_losses.dtype:
            # one-time move: a host->device copy of a pageable tensor every step would synchronise the stream (quirk Q5)
            self.weights = self.weights.to(device=head_losses.device, dtype=head_losses.dtype)
        return self.reduce(self.weights * head_losses), head_losses[0]

    def updateWeightsSurv(self, train_preds, train_events, train_durations, val_preds, val_events, val_durations):
        train_loss = self.computeLossSurv(train_preds, train_events, train_durations, reduceToHeads=True).detach()
        val_loss = self.computeLossSurv(val_preds, val_events, val_durations, reduceToHeads=True).detach()
        if self.lvn is None or self.ltn is None:
            self._set_weights(self.normalize(torch.ones(train_preds.shape[0], device=train_loss.device)))
        else:
            o_n = self.lvn - self.ltn
            o_npn = val_loss - train_loss
            delta_g = self.lvn - val_loss
            delta_o = o_npn - o_n
            self._set_weights(self.normalize(delta_g / torch.pow(delta_o, 2)))
        self.lvn, self.ltn = val_loss, train_loss
        self.history.append(self.weights.detach().cpu().numpy())

    # ---- classification (/root/reference/losses/GradientBlender.py:105-179,207-226)
    def computeLossClassification(self, preds, targets, reduceToHeads=False, no_reduce=False):
        """preds [k+1, N, C] stacked head logits, targets [N, C]; loss_function has reduction='none'.  The reference stacks
        the targets k+1 times (:166-168); the fused loss reuses one copy per head instead."""
        loss = self.loss_function(preds, targets)
        if self.weights is None:
            self.weights = self.normalize(torch.ones(preds.shape[0]))
            self.history.append(self.weights.detach().cpu().numpy())
        if no_reduce:
            return loss
        loss = self.reduceToHeads(loss)
        if reduceToHeads:
            return loss
        if self.weights.device != loss.device or self.weights.dtype != loss.dtype:
            self.weights = self.weights.to(device=loss.device, dtype=loss.dtype)
        return self.reduce(self.weights * loss)

    def updateWeightsClass(self, train_preds, train_targs, val_preds, val_targs):
        train_loss = self.computeLossClassification(train_preds, train_targs, reduceToHeads=True).detach()
        val_loss = self.computeLossClassification(val_preds, val_targs, reduceToHeads=True).detach()
        if self.lvn is None or self.ltn is None:
            self._set_weights(self.normalize(torch.ones(train_preds.shape[0], device=train_loss.device)))
        else:
            o_n = self.lvn - self.ltn
            o_npn = val_loss - train_loss
            delta_g = val_loss - self.lvn            # sign as in the reference's classification branch (:128)
            delta_o = o_npn - o_n
            self._set_weights(self.normalize(delta_g / torch.pow(delta_o, 2)))
        self.lvn, self.ltn = val_loss, train_loss

    def reduceToHeads(self, loss):
        if self.reduction.startswith("sum"):
            return torch.sum(loss, dim=(1, 2))
        if self.reduction.startswith("mean"):
            return torch.mean(loss, dim=(1, 2))
        if self.reduction.startswith("none"):
            return loss
        raise ValueError("Unable to reduce loss, unrecognized reduction: {}".format(self.reduction))

    # ---- dispatch
    def updateWeights(self, *args, **kwargs):
        if self.survival:
            self.updateWeightsSurv(*args, **kwargs)
        else:
            self.updateWeightsClass(*args, **kwargs)

    def computeLoss(self, *args, **kwargs):
        if self.survival:
            return self.computeLossSurv(*args, **kwargs)
        return self.computeLossClassification(*args, **kwargs)

    def reduce(self, loss):
        if self.reduction.startswith("sum"):
            return torch.sum(loss)
        if self.reduction.startswith("mean"):
            return torch.mean(loss)
        if self.reduction.startswith("none"):
            return loss
        raise ValueError("Unable to reduce loss, unrecognized reduction: {}".format(self.reduction))

    def normalize(self, weights):
        return F.softmax(weights, dim=0)

    def saveHistory(self):
        np.savetxt("gblend_weights_history.csv", np.array(self.history), delimiter=",")
