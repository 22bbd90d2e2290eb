"""CPU restatement of the reference's survival TRAINING LOOP (the caller of the hot path).

TEST INFRASTRUCTURE.  Follows /root/reference/main.py:385-601 (`train_survival`), which cannot itself be imported
(/root/reference/main.py:1 imports a name that does not exist):
  * :402-407  super_batch_interval = 64 / batch_size;  steps_per_epoch = ceil(N / 64)
  * :410-414  SGD(lr, momentum, nesterov=True, weight_decay) + OneCycleLR(max_lr=lr, steps_per_epoch, epochs)
  * :445-469  forward -> GradientBlender.computeLoss()[0] (blend) or surv_criterion(CoxPH, ...) -> loss.backward() on EVERY micro-batch
  * :478-481  optimizer.step(); scheduler.step(); optimizer.zero_grad()  when (i + 1) % super_batch_interval == 0 or on the last batch
  * :484-498  running concatenation of predictions; per-epoch C-index of head 0 through getCIndices
  * :509-569  eval-mode validation under no_grad: summed loss / len(val), `selection_loss` = head-0 loss of the LAST validation batch
  * :584-588  every blend_update_interval epochs: gradient_blender.updateWeights(all train preds, all val preds)
The loop is generic over the model / loss / blender objects so that the SAME code drives (a) the unchanged reference classes under
oracle/shim.py (tests/golden/make_trajectory_golden.py -> tests/golden/trajectory.npz) and (b) the oracle restatement
(tests/test_oracle.py pins it to that file).  The CUDA build's mmnn_sts_b200.main.train_survival is compared with the same file."""
import math
from types import SimpleNamespace

import numpy as np
import torch

SUPER_BATCH_SIZE = 64   # /root/reference/main.py:62


def train_survival_loop(model, train_batches, val_batches, args, blender, surv_criterion, loss_function, get_c_indices):
    n_train = args.num_train
    super_batch_interval = SUPER_BATCH_SIZE / args.batch_size
    steps_per_epoch = n_train // SUPER_BATCH_SIZE if n_train % SUPER_BATCH_SIZE == 0 else 1 + n_train // SUPER_BATCH_SIZE
    optimizer = torch.optim.SGD(model.parameters(), args.lr, momentum=args.momentum, nesterov=True, weight_decay=args.weight_decay)
    scheduler = torch.optim.lr_scheduler.OneCycleLR(optimizer, max_lr=args.lr, steps_per_epoch=steps_per_epoch, epochs=args.epochs)
    hist = SimpleNamespace(train_loss=[], val_loss=[], train_c=[], val_c=[], lr_trace=[], momentum_trace=[], step_at=[],
                           best_loss=math.inf, selection=[], blender_weights=[])
    n_val = sum(b[1].shape[0] for b in val_batches)
    for epoch in range(args.epochs):
        model.train()
        epoch_loss = 0.0
        c_pred, c_events, c_durations = [], [], []
        for i, (inputs, events, durations) in enumerate(train_batches):
            outputs = model(inputs)
            if args.blend:
                loss, _ = blender.computeLoss(outputs, events, durations)
            else:
                loss = surv_criterion(loss_function, outputs, events, durations, "cpu")
            loss.backward()
            epoch_loss += loss.item()
            if (i + 1) % super_batch_interval == 0 or i == len(train_batches) - 1:
                optimizer.step()
                scheduler.step()
                optimizer.zero_grad()
                hist.step_at.append((epoch, i))
                hist.lr_trace.append(optimizer.param_groups[0]["lr"])
                hist.momentum_trace.append(optimizer.param_groups[0]["momentum"])
            c_pred.append(outputs.detach()); c_events.append(events); c_durations.append(durations)
        c_pred = torch.cat(c_pred, dim=1 if args.blend else 0)
        c_events, c_durations = torch.cat(c_events), torch.cat(c_durations)
        hist.train_c.append(get_c_indices((c_pred[0] if args.blend else c_pred).numpy(), c_events.numpy(), c_durations.numpy()))
        hist.train_loss.append(epoch_loss / n_train)
        model.eval()
        with torch.no_grad():
            y_pred, y_events, y_durations, test_loss, selection_loss = [], [], [], 0.0, None
            for inputs, events, durations in val_batches:
                preds = model(inputs)
                if args.blend:
                    loss, selection_loss = blender.computeLoss(preds, events, durations)
                else:
                    loss = selection_loss = surv_criterion(loss_function, preds, events, durations, "cpu")
                test_loss += float(loss)
                y_pred.append(preds); y_events.append(events); y_durations.append(durations)
            y_pred = torch.cat(y_pred, dim=1 if args.blend else 0)
            y_events, y_durations = torch.cat(y_events), torch.cat(y_durations)
            hist.val_c.append(get_c_indices((y_pred[0] if args.blend else y_pred).numpy(), y_events.numpy(), y_durations.numpy()))
            hist.val_loss.append(test_loss / n_val)
            hist.selection.append(float(selection_loss))
            hist.best_loss = min(hist.best_loss, float(selection_loss))
        if args.blend and (epoch + 1) % args.blend_update_interval == 0:
            blender.updateWeights(c_pred, c_events, c_durations, y_pred, y_events, y_durations)
            hist.blender_weights.append(np.asarray(blender.weights.detach().cpu().numpy(), dtype=np.float64))
    return hist


class OracleMultiModal(torch.nn.Module):
    """The oracle's functional network (oracle/model.py) behind the nn.Module surface the loop needs."""

    def __init__(self, sd, blend):
        super().__init__()
        from . import model as om
        self._om, self.blend = om, blend
        self.sd = {k: v.clone() for k, v in sd.items()}
        self._params = []
        for k, v in self.sd.items():
            if v.dtype.is_floating_point and "running_" not in k:
                v.requires_grad_(True)
                self._params.append(v)

    def parameters(self, recurse=True):
        return iter(self._params)

    def forward(self, x):
        return self._om.multimodal_forward(self.sd, x["image"], x["clinical"], self.training, self.blend, None)


def trajectory_case(mini=False):
    """The committed trajectory case: 72 patients in micro-batches of 8 (9 per epoch: the optimiser steps after the 8th --
    64 patients -- and after the last one), 2 epochs = 4 optimiser steps = the whole OneCycle schedule, blending weights updated
    every epoch, 16 validation patients; 1x64x64x32 volumes, no dropout.  mini: 24 patients (3 micro-batches, one optimiser step
    per epoch) -- the cheap case that pins the oracle's own restatement on the CPU."""
    from . import synth
    nb = 3 if mini else 9
    args = SimpleNamespace(lr=2e-3, momentum=0.9, weight_decay=1e-4, epochs=2, batch_size=8, blend=True, blend_update_interval=1, num_train=8 * nb)
    train = []
    for i in range(nb):
        im, cl, ev, du = synth.make_batch(300 + i, 8, 1, (64, 64, 32))
        train.append(({"image": im, "clinical": cl}, ev, du))
    val = []
    for i in range(2):
        im, cl, ev, du = synth.make_batch(400 + i, 8, 1, (64, 64, 32))
        val.append(({"image": im, "clinical": cl}, ev, du))
    sd = synth.make_state_dict(42, in_channels=1)
    return args, train, val, sd


TRACKED = ["image_model.model.backbone.conv0.weight", "image_model.model.backbone.denseblock1.denselayer1.layers.conv1.weight",
           "image_model.model.backbone.denseblock2.denselayer5.layers.conv2.weight", "image_model.model.backbone.transition2.conv.weight",
           "image_model.model.backbone.denseblock4.denselayer16.layers.conv2.weight", "image_model.model.backbone.norm5.weight",
           "image_model.model.features.feature_layer.weight", "clinical_model.model.backbone.dense0.weight",
           "clinical_model.model.features.bn5.weight", "output_head.weight", "image_output_head.weight", "clinical_output_head.weight"]
