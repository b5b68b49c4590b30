"""Data parallelism over subjects: one process per GPU, ONE all-reduce over a flat gradient buffer (NCCL over NVLink).

The reference is single-device (``cuda:0`` hard-coded, kfold_train_adversarial.py:24); data parallelism is added by
this build.  Every rank holds a full replica (4.17 M parameters, 16.7 MB of fp32 gradients),
runs the unchanged model on its shard of the global batch with per-rank BatchNorm statistics (DDP semantics), and
averages gradients before the optimizer step.

Design (measured on 2 x B200, profiles/r2_dp_notes.md).  The payload is tiny, so the exchange is latency-bound; what
costs time is everything AROUND the collective:
  * round 1 packed gradients into 8 MB buckets (pack -> ncclAllReduce -> unpack per bucket, serial): +0.15 ms per step;
  * overlapping bucketed all-reduces with the backward pass was tried in round 2 and is SLOWER (+0.37 ms): the conv
    kernels are persistent with one CTA per SM, and an NCCL kernel that holds even a few SMs forces their last CTAs into
    a second wave, doubling those kernels.
So the gradients are born in place: every parameter owns a 256-byte aligned slot of one flat fp32 buffer, the backward
kernels of this package write their weight gradients straight into the slots (``functional.grad_out``; autograd then
adopts the slot view as ``.grad`` without a copy), and ``finish()`` is a single ``all_reduce(flat, AVG)`` between the
backward pass and the optimizer -- no pack, no unpack, no bucket bookkeeping.  The few gradients produced by torch's
own autograd (BatchNorm1d affine parameters of the heads) are copied into / out of their slots by one fused copy each.
Everything is stream-ordered and capturable: ``train.GraphedTrainStep`` records the step, collective included, into ONE
CUDA graph.  The same code runs on CPU tensors with the ``gloo`` backend (tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

SLOT_ALIGN = 64          # floats: 256-byte aligned slots (vectorised kernels, NCCL-friendly)


def shard_slice(global_batch: int, rank: int, world: int) -> slice:
    """Contiguous, equal shards of the global batch (global_batch must divide by world)."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return slice(rank * per, (rank + 1) * per)


class FlatGradReducer:
    """``params``: iterable of parameters (or ``model=`` an ``nn.Module``).  ``install()`` makes the slots visible to
    the backward kernels; ``finish()`` averages the gradients over the ranks.  World size 1: ``finish()`` is a no-op
    (the slots are still used, so single- and multi-GPU runs execute the same kernels on the same addresses)."""

    def __init__(self, params=None, process_group=None, model=None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        plist = list(model.parameters()) if model is not None else list(params)
        self.params = [p for p in plist if p.requires_grad]
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + SLOT_ALIGN - 1) // SLOT_ALIGN * SLOT_ALIGN
        self.numel = n
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(max(n, 1), dtype=torch.float32, device=dev)
        self.slots = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]
        backend = dist.get_backend(process_group) if self.world > 1 else None
        self._native_avg = backend == "nccl"
        self.allreduce_launches = 0
        self.packed_last = 0            # parameters whose gradient had to be copied into its slot in the last finish()
        self._installed = False

    def layout(self):
        """(total floats incl. alignment padding, number of parameters) -- for tests / DESIGN.md."""
        return self.numel, len(self.params)

    def install(self):
        """Register the slots with ``functional.grad_out`` (one active reducer per process)."""
        from . import functional as TF
        TF.set_grad_slots({p.data_ptr(): (s, p) for p, s in zip(self.params, self.slots)})
        self._installed = True
        return self

    def remove(self):
        from . import functional as TF
        if self._installed:
            TF.set_grad_slots(None)
            self._installed = False

    def finish(self):
        """Call after ``loss.backward()`` and before ``optimizer.step()``."""
        if self.world == 1:
            return
        dst, src = [], []
        for p, s in zip(self.params, self.slots):
            g = p.grad
            if g is None:
                s.zero_()                                   # (a parameter without a gradient contributes zeros)
            elif g.data_ptr() != s.data_ptr():
                dst.append(s)
                src.append(g)
        self.packed_last = len(dst)
        if dst:
            torch._foreach_copy_(dst, src)
        op = dist.ReduceOp.AVG if self._native_avg else dist.ReduceOp.SUM
        dist.all_reduce(self.flat, op=op, group=self.group)
        if not self._native_avg:
            self.flat.div_(self.world)
        self.allreduce_launches += 1
        if dst:
            torch._foreach_copy_(src, dst)
        for p, s in zip(self.params, self.slots):
            if p.grad is None:
                p.grad = s.clone()


# Round-1 name, kept for callers: the bucketed reducer is gone (see the module docstring for why).
GradBucketReducer = FlatGradReducer
