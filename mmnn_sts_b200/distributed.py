"""Data-parallel plumbing: one process per GPU, patients sharded by rank, gradients all-reduced with NCCL.

The reference has no distributed code.  Semantics chosen so that "B patients per GPU on G GPUs"
equals the reference accumulating G micro-batches of B patients before one optimizer step (/root/reference/main.py:
469,478-481): BatchNorm statistics and the Cox risk set stay rank-local, gradients are SUMMED (not averaged).
There is no data-path collective: the only exchange is the gradient all-reduce (payload 11.3 M fp32 = 45 MB)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun). Returns (rank, world, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    device = torch.device("cuda", local) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world)
    return rank, world, device


def shard_patients(num_patients, rank, world, permutation=None):
    """Rank r takes patients perm[r::world]."""
    perm = permutation if permutation is not None else torch.arange(num_patients)
    return perm[rank::world]


class GradientAllReducer:
    """Flat-bucket gradient all-reduce (SUM by default).  Parameters that never receive a gradient (the reference's
    unused class_layers.out / dense6 / modality heads, SURVEY.md section 7 hard part 9) are skipped.
    Buckets are launched in reverse parameter order (the order backward produces them) on NCCL's own stream via
    async_op, so the reduction of the late layers overlaps whatever the caller still has queued."""

    def __init__(self, params, bucket_bytes=64 << 20, average=False, group=None, model=None):
        self.params = [p for p in params if p.requires_grad]
        self.bucket_bytes = bucket_bytes
        self.average = average
        self.group = group
        # modules that keep their gradients in one flat buffer (the DenseNet trunk) are reduced in place, without packing
        self.flat_modules = [m for m in model.modules() if hasattr(m, "flat_grad_buffer")] if model is not None else []

    def __call__(self, *_):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        world = dist.get_world_size(self.group)
        pending_flat, covered = [], set()
        for m in self.flat_modules:
            flat = m.flat_grad_buffer()
            if flat is not None:
                pending_flat.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True), flat))
                covered.update(id(p) for p in m.parameters())
        grads = [p.grad for p in reversed(self.params) if p.grad is not None and id(p) not in covered]
        buckets, cur, size = [], [], 0
        for g in grads:
            cur.append(g); size += g.numel() * g.element_size()
            if size >= self.bucket_bytes:
                buckets.append(cur); cur, size = [], 0
        if cur:
            buckets.append(cur)
        pending = []
        for b in buckets:
            flat = torch.cat([g.reshape(-1) for g in b])
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            pending.append((work, flat, b))
        for work, flat in pending_flat:
            work.wait()
            if self.average:
                flat.div_(world)
        for work, flat, b in pending:
            work.wait()
            if self.average:
                flat.div_(world)
            off = 0
            for g in b:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n


def allgather_rows(t, group=None):
    """All-gather of per-rank row blocks (predictions / targets for the per-epoch C-index and updateWeights)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t.contiguous(), group=group)
    return torch.cat(out, dim=0)
