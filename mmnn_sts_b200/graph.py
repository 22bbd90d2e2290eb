"""CUDA-graph capture of a whole training step (forward + loss + backward + optimizer) for static shapes.

The late dense blocks of the network are launch-latency bound (SURVEY.md section 7 hard part 5): at configs[0] the
Python / launch overhead of the ~600 kernels of a step exceeds their GPU time.  Everything the step enqueues is
capturable -- the C-ABI entry points never allocate or synchronise, device tables are uploaded once during warm-up,
the weight-gradient side stream is forked / joined with events -- so the step can be replayed as one graph launch."""
import torch


class GraphedTrainStep:
    """step_fn(image, clinical, events, durations) -> loss tensor; must do backward + optimizer.step + zero_grad itself.
    Inputs are copied into static buffers before each replay; the returned loss is a static tensor.

    What a replay re-reads and what it does not: kernel ARGUMENTS (scalars, pointers) are frozen at capture time, device
    MEMORY is read afresh.  Hence
      * `optimizer` (optional): must be `mmnn_sts_b200.optim.SGD(capturable=True)`, whose lr / momentum / weight decay live in
        device memory; `__call__` refreshes them from `param_groups` before every replay, so a scheduler stepped between
        replays (OneCycleLR after each optimiser step, /root/reference/main.py:414,480) takes effect.  A torch optimizer or a
        non-capturable SGD is refused here rather than replayed with stale hyper-parameters;
      * `GradientBlender.updateWeights` updates its device weight tensor in place, so captured steps see the new weights."""

    def __init__(self, step_fn, example_batch, warmup=3, optimizer=None):
        if optimizer is not None and not getattr(optimizer, "capturable", False):
            raise ValueError("GraphedTrainStep needs an optimizer whose hyper-parameters live in device memory "
                             "(mmnn_sts_b200.optim.SGD(..., capturable=True)); scalar lr / momentum would be frozen at capture time")
        self.optimizer = optimizer
        self.static = [t.clone() for t in example_batch]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):           # allocator warm-up + one-time table uploads happen outside the capture
                step_fn(*self.static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = step_fn(*self.static)

    def __call__(self, image, clinical, events, durations):
        for dst, src in zip(self.static, (image, clinical, events, durations)):
            dst.copy_(src, non_blocking=True)
        if self.optimizer is not None:
            self.optimizer.refresh_hyper()
        self.graph.replay()
        return self.loss
