"""GradientBlender restatement (survival branch) -- /root/reference/losses/GradientBlender.py:37-103,181-205,228-253.

TEST INFRASTRUCTURE.  head k=0 is the multimodal head.  First computeLoss sets weights = softmax(ones) = 1/3 each
(:199-200,253; F.softmax without dim on a 1-D tensor);  loss = sum_k w_k * head_loss_k  (reduction 'sum');
updateWeights: first call -> uniform; later  w = softmax( (lvn - val) / ((val-train) - (lvn-ltn))^2 ).
"""
import torch

from .cox import CoxPH, surv_criterion


class GradientBlenderOracle:
    def __init__(self, loss_function=CoxPH):
        self.loss_function = loss_function
        self.weights = None
        self.lvn = None
        self.ltn = None
        self.history = []

    def head_losses(self, preds, events, durations):
        return torch.stack([surv_criterion(self.loss_function, preds[k], events, durations)
                            for k in range(preds.shape[0])], dim=0)

    def computeLoss(self, preds, events, durations, reduceToHeads=False):
        hl = self.head_losses(preds, events, durations)
        if self.weights is None:
            self.weights = torch.softmax(torch.ones(preds.shape[0]), dim=0)
        if reduceToHeads:
            return hl
        return torch.sum(self.weights.to(hl.dtype) * hl), hl[0]

    def updateWeights(self, tp, te, td, vp, ve, vd):
        train_loss = self.computeLoss(tp, te, td, reduceToHeads=True)
        val_loss = self.computeLoss(vp, ve, vd, reduceToHeads=True)
        if self.lvn is None or self.ltn is None:
            self.weights = torch.softmax(torch.ones(tp.shape[0]), dim=0)
        else:
            o_n = self.lvn - self.ltn
            o_npn = val_loss - train_loss
            delta_g = self.lvn - val_loss
            delta_o = o_npn - o_n
            self.weights = torch.softmax(delta_g / torch.pow(delta_o, 2), dim=0)
        self.lvn, self.ltn = val_loss, train_loss
        self.history.append(self.weights.detach().cpu().numpy())
