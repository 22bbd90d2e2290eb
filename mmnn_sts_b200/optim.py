"""SGD with the reference's hyper-parameters (`torch.optim.SGD(model.parameters(), lr, momentum=0.9, nesterov=True,
weight_decay=1e-4)`, /root/reference/main.py:410-414) whose `step()` is ONE launch of `mmnn_sgd_step` over every
parameter tensor instead of torch's 3 foreach passes x 11 launches.  Subclass of `torch.optim.SGD`: `param_groups`,
`state_dict()`, LR schedulers (OneCycleLR cycles `lr` and `momentum`, :402-409) work unchanged; one-line swap in main.py.
`capturable=True`: lr / momentum / weight decay are read by the kernel from a small DEVICE tensor per parameter group
(`mmnn_sgd_step_dev`) that `refresh_hyper()` rewrites in place, so a step captured in a CUDA graph (mmnn_sts_b200.graph)
keeps following the scheduler; without it the hyper-parameters are launch-time scalars and a captured step would replay the
values of capture time.  No CPU path: CPU parameters raise."""
import ctypes as C

import torch

from . import _lib as L


class SGD(torch.optim.SGD):
    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, capturable=False, **kw):
        if dampening != 0.0:
            raise NotImplementedError("mmnn_sgd_step implements dampening = 0 (the reference's setting)")
        if kw.get("maximize", False):
            raise NotImplementedError("maximize is not supported")
        super().__init__(params, lr=lr, momentum=momentum, dampening=0.0, weight_decay=weight_decay, nesterov=nesterov)
        self._tables = {}
        self.capturable = bool(capturable)
        self._hyper_dev = {}      # group index -> device float[3] {lr, momentum, weight_decay}
        self._hyper_host = {}     # group index -> pinned staging of the same

    def refresh_hyper(self):
        """capturable mode: copy the CURRENT lr / momentum / weight_decay of every group into its device tensor (in place,
        asynchronously from pinned memory).  Call it after scheduler.step() / before each CUDA-graph replay; eager steps
        call it themselves.  Not capturable itself (it is the host's side of the schedule)."""
        for gi, g in enumerate(self.param_groups):
            dev = self._hyper_dev.get(gi)
            if dev is None:
                continue
            host = self._hyper_host[gi]
            host[0], host[1], host[2] = float(g["lr"]), float(g["momentum"]), float(g["weight_decay"])
            dev.copy_(host, non_blocking=True)

    def _hyper(self, gi, device):
        if gi not in self._hyper_dev:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("mmnn_sts_b200.optim.SGD(capturable=True): run one eager step before capturing")
            self._hyper_host[gi] = torch.zeros(3, dtype=torch.float32).pin_memory()
            self._hyper_dev[gi] = torch.zeros(3, dtype=torch.float32, device=device)
        return self._hyper_dev[gi]

    def _static(self, gi, ps):
        """Per-group arrays that do not change between steps: parameter / momentum pointers and sizes."""
        for p in ps:
            st = self.state[p]
            if st.get("momentum_buffer") is None:
                st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        # the cached pointer tables are valid only while every parameter AND every momentum buffer still lives at the
        # same address: load_state_dict() replaces the momentum tensors, `p.data = ...` / model.to() move a parameter
        # without changing id(p) -- so the addresses themselves are the key (cheap to compare each step)
        key = tuple((p.data_ptr(), self.state[p]["momentum_buffer"].data_ptr(), p.numel()) for p in ps)
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1:]
        for p in ps:
            st = self.state[p]
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and st["momentum_buffer"].is_contiguous()):
                raise L.MMNNLibraryError("mmnn_sts_b200.optim.SGD needs contiguous fp32 CUDA parameters (no CPU path)")
        n = len(ps)
        pp = (C.c_void_p * n)(*[p.data_ptr() for p in ps])
        mm = (C.c_void_p * n)(*[self.state[p]["momentum_buffer"].data_ptr() for p in ps])
        nn = (C.c_longlong * n)(*[p.numel() for p in ps])
        self._tables[gi] = (key, pp, mm, nn)
        return pp, mm, nn

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, g in enumerate(self.param_groups):
            ps = [p for p in g["params"] if p.grad is not None]
            if not ps:
                continue
            pp, mm, nn = self._static(gi, ps)
            grads = [p.grad for p in ps]
            for p, gr in zip(ps, grads):
                if not (gr.is_cuda and gr.dtype == torch.float32 and gr.is_contiguous() and gr.numel() == p.numel()):
                    raise L.MMNNLibraryError("mmnn_sts_b200.optim.SGD needs contiguous fp32 CUDA gradients (no CPU path)")
            gg = (C.c_void_p * len(ps))(*[gr.data_ptr() for gr in grads])
            with torch.cuda.device(ps[0].device):
                if self.capturable:
                    hyper = self._hyper(gi, ps[0].device)
                    if not torch.cuda.is_current_stream_capturing():
                        self.refresh_hyper()
                    rc = L.lib().mmnn_sgd_step_dev(pp, gg, mm, nn, len(ps), hyper.data_ptr(), int(bool(g["nesterov"])),
                                                   torch.cuda.current_stream().cuda_stream)
                else:
                    if torch.cuda.is_current_stream_capturing():
                        raise RuntimeError("capturing an SGD step whose lr / momentum are launch-time scalars would freeze the "
                                           "schedule at its capture-time values: build the optimizer with capturable=True")
                    rc = L.lib().mmnn_sgd_step(pp, gg, mm, nn, len(ps), float(g["lr"]), float(g["momentum"]),
                                               float(g["weight_decay"]), int(bool(g["nesterov"])),
                                               torch.cuda.current_stream().cuda_stream)
            L.check(rc, "mmnn_sgd_step")
        return loss
