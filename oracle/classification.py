"""CPU restatement of the reference's classification loss path.  TEST INFRASTRUCTURE.

  bce_with_logits      torch.nn.BCEWithLogitsLoss(pos_weight, reduction) as constructed at /root/reference/main.py:148-153
                       (torch itself is the third-party dependency; the restatement is checked against it in
                       tests/test_classification.py)
  GradientBlenderClassOracle   the UNCHANGED control flow of /root/reference/losses/GradientBlender.py:105-179,207-253
                       (computeLossClassification / updateWeightsClass / reduceToHeads / reduce / normalize)
  f1_counts, f1_score  /root/reference/main.py:98-104,224-229
"""
import torch
import torch.nn.functional as F


def bce_with_logits(x, y, pos_weight=None):
    x = x.double(); y = y.double()
    w = 1.0 + ((pos_weight.double() if pos_weight is not None else 1.0) - 1.0) * y
    return ((1.0 - y) * x + w * (torch.log1p(torch.exp(-x.abs())) + torch.clamp(-x, min=0.0)))


class GradientBlenderClassOracle:
    def __init__(self, loss_function, reduction="sum"):
        self.loss_function, self.reduction = loss_function, reduction
        self.weights = self.lvn = self.ltn = None

    def normalize(self, w):
        return F.softmax(w, dim=0)

    def reduceToHeads(self, loss):
        return torch.sum(loss, dim=(1, 2)) if self.reduction == "sum" else torch.mean(loss, dim=(1, 2))

    def reduce(self, loss):
        return torch.sum(loss) if self.reduction == "sum" else torch.mean(loss)

    def computeLoss(self, preds, targets, reduceToHeads=False, no_reduce=False):
        targets = torch.stack([targets for _ in range(preds.shape[0])], dim=0)
        loss = self.loss_function(preds, targets)
        if self.weights is None:
            self.weights = self.normalize(torch.ones(preds.shape[0]))
        if no_reduce:
            return loss
        loss = self.reduceToHeads(loss)
        if reduceToHeads:
            return loss
        return self.reduce(self.weights * loss)

    def updateWeights(self, train_preds, train_targs, val_preds, val_targs):
        train_loss = self.computeLoss(train_preds, train_targs, reduceToHeads=True)
        val_loss = self.computeLoss(val_preds, val_targs, reduceToHeads=True)
        if self.lvn is None or self.ltn is None:
            self.weights = self.normalize(torch.ones(train_preds.shape[0]))
        else:
            o_n = self.lvn - self.ltn
            o_npn = val_loss - train_loss
            delta_g = val_loss - self.lvn
            delta_o = o_npn - o_n
            self.weights = self.normalize(delta_g / torch.pow(delta_o, 2))
        self.lvn, self.ltn = val_loss, train_loss


def f1_counts(logits, labels, threshold=0.5):
    preds = torch.sigmoid(logits) > threshold
    tps = torch.sum((preds == 1) * (labels == 1), 0)
    fps = torch.sum((preds == 1) * (labels == 0), 0)
    fns = torch.sum((preds == 0) * (labels == 1), 0)
    return tps, fps, fns


def f1_score(tps, fps, fns):
    return [(tps[i] / (tps[i] + 0.5 * (fns[i] + fps[i]))).item() for i in range(len(tps))]
