"""Data-parallel replicas for inference: observations are independent, so a global batch is split
contiguously across ranks, every rank holds a full bf16 replica and there is NO collective on the data path.
(ZoeDepth's metric-head router votes over the *local* batch, so parity is defined per shard -- exactly what the
reference would compute if it were handed that shard.)"""
from __future__ import annotations


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of n items for `rank`; the first n % world ranks get one extra item."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: dict, rank: int, world: int):
    """Slices every per-observation tensor of a model-input dict; a shared (3,3) intrinsic is passed through."""
    n = batch["input_ids"].shape[0]
    lo, hi = shard_bounds(n, rank, world)
    out = {}
    for k, v in batch.items():
        if hasattr(v, "shape") and v.dim() >= 1 and v.shape[0] == n and not (k == "intrinsic" and v.dim() == 2):
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def reduce_loss(outputs, group=None):
    """The one exchange step of the labelled forward over a sharded batch (evaluation / the loss a data-parallel training step
    logs): every rank holds the per-row losses of ITS observations, the global values are
    loss = sum_r sum(row_loss_r) / sum_r R_r and token accuracy = sum_r hits_r / sum_r R_r -- one all-reduce (SUM) of three fp32
    numbers over NCCL (gloo on CPU).  Equals the mean of the per-rank mean losses the reference's DDP / ZeRO-1 step effectively
    optimises whenever every rank labels the same number of tokens (13 per sample at config #5).
    outputs: the result of forward(labels=...) on this rank's shard.  Returns (loss, token_accuracy, labelled_rows) as floats/int."""
    import torch
    import torch.distributed as dist
    rl = outputs.row_loss
    if rl is None or rl.numel() == 0:
        part = torch.zeros(3, dtype=torch.float32, device=outputs.loss.device)
    else:
        hits = (outputs.row_argmax == outputs.row_labels).sum().to(torch.float32)
        part = torch.stack([rl.sum(), torch.tensor(float(rl.numel()), device=rl.device), hits]).to(torch.float32)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    s, n, h = (float(v) for v in part.cpu())
    return (s / n if n else float("nan")), (h / n if n else float("nan")), int(n)


def allreduce_gradients(arena, group=None, average: bool = True):
    """The gradient exchange of the data-parallel fine-tune step (BASELINE.json config #5): ONE all-reduce over the flat gradient
    buffer of the LoRA arena (spatialvla_b200/lora.py; 59.2 M elements at the 4B-224 size), NCCL over NVLink on GPUs, gloo on CPU;
    `average` divides by the world size (the mean DDP / ZeRO-1 applies, scripts/zero1.json).  Returns the number of elements sent."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(arena.grad, op=dist.ReduceOp.SUM, group=group)
        if average:
            arena.grad.div_(dist.get_world_size(group))
    return arena.grad.numel()


def world_size(group=None) -> int:
    import torch.distributed as dist
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


class GradientReducer:
    """Gradient exchange of the data-parallel LoRA step, overlapped with the backward pass: the arena is laid out in backward order
    of completion (Gemma2 adapters first, then SigLIP / projector / Ego3D), so the Gemma2 segment (`n_first` elements, ~82 % of the
    arena) is all-reduced on NCCL's stream as soon as the language-model backward has finished, while the SigLIP backward is still
    running; the remaining segment follows at the end of the backward.  Two collectives per step over disjoint slices of the ONE
    flat buffer (DeepSpeed ZeRO-1 buckets the same way: scripts/zero1.json reduce_bucket_size / overlap_comm).  gloo on CPU."""

    def __init__(self, arena, n_first: int, group=None):
        self.arena, self.n_first, self.group = arena, int(n_first), group
        self.work = []
        self.collectives = 0

    def _active(self):
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def first_segment_ready(self):
        if self._active():
            import torch.distributed as dist
            self.work.append(dist.all_reduce(self.arena.grad[: self.n_first], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self.collectives += 1

    def finish(self):
        """Launch the all-reduce of the rest and make the compute stream wait for both.  Returns the world size."""
        if not self._active():
            return 1
        import torch.distributed as dist
        if not self.work:                                  # no overlap hook fired: one collective over the whole arena
            self.work.append(dist.all_reduce(self.arena.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self.work.append(dist.all_reduce(self.arena.grad[self.n_first:], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.collectives += 1
        for w in self.work:
            w.wait()
        self.work = []
        return dist.get_world_size(self.group)
