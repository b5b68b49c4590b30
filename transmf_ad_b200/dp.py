"""Data parallelism over subjects: one process per GPU, bucketed gradient all-reduce (NCCL over NVLink).

The reference is single-device (``cuda:0`` hard-coded, kfold_train_adversarial.py:24); data parallelism is added by
this build.  Every rank holds a full replica (4.17 M parameters, 16.7 MB of fp32 gradients),
runs the unchanged model on its shard of the global batch with per-rank BatchNorm statistics (DDP semantics), and
averages gradients before the optimizer step.  Gradients are packed into a few flat fp32 buckets in reverse
registration order (heads and fusion transformer first -- they finish first in backward -- conv towers last); a
bucket's all-reduce is launched asynchronously from the autograd hook of its last parameter so it overlaps with
the remaining backward kernels.  The same code runs on CPU tensors with the ``gloo`` backend (tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_slice(global_batch: int, rank: int, world: int) -> slice:
    """Contiguous, equal shards of the global batch (global_batch must divide by world)."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return slice(rank * per, (rank + 1) * per)


class GradBucketReducer:
    def __init__(self, params, process_group=None, bucket_bytes: int = 8 << 20):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad][::-1]
        self.buckets = []                   # list of dicts: params, offsets, numel, flat, views, pending, work
        cur, cur_bytes = [], 0
        for p in self.params:
            nbytes = p.numel() * 4
            if cur and cur_bytes + nbytes > bucket_bytes:
                self._close(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._close(cur)
        self._where = {}
        for bi, b in enumerate(self.buckets):
            for p in b["params"]:
                self._where[p] = bi
        self._hooks = []
        if self.world > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        backend = dist.get_backend(process_group) if self.world > 1 else None
        self._native_avg = backend == "nccl"
        self.allreduce_launches = 0

    def _close(self, plist):
        offs, n = [], 0
        for p in plist:
            offs.append(n)
            n += p.numel()
        self.buckets.append({"params": list(plist), "offsets": offs, "numel": n, "flat": None, "views": None,
                             "pending": len(plist), "work": None, "launched": False})

    def bucket_layout(self):
        """[(numel, n_params)] per bucket, in launch order (for tests / DESIGN.md)."""
        return [(b["numel"], len(b["params"])) for b in self.buckets]

    def _ensure_flat(self, b):
        if b["flat"] is None:
            ref = b["params"][0]
            b["flat"] = torch.zeros(b["numel"], dtype=torch.float32, device=ref.device)
            b["views"] = [b["flat"][o:o + p.numel()].view_as(p) for o, p in zip(b["offsets"], b["params"])]

    def _on_grad(self, p):
        b = self.buckets[self._where[p]]
        b["pending"] -= 1
        if b["pending"] == 0 and not b["launched"]:
            self._launch(b)

    def _launch(self, b):
        self._ensure_flat(b)
        have = [(v, p.grad) for v, p in zip(b["views"], b["params"]) if p.grad is not None]
        if len(have) != len(b["params"]):
            b["flat"].zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        op = dist.ReduceOp.AVG if self._native_avg else dist.ReduceOp.SUM
        b["work"] = dist.all_reduce(b["flat"], op=op, group=self.group, async_op=True)
        b["launched"] = True
        self.allreduce_launches += 1

    def finish(self):
        """Call after ``loss.backward()`` and before ``optimizer.step()``: waits for the in-flight buckets, launches
        any bucket whose parameters did not all receive a gradient, and writes the averaged gradients back."""
        if self.world == 1:
            return
        for b in self.buckets:
            if not b["launched"]:
                self._launch(b)
        for b in self.buckets:
            b["work"].wait()
            if not self._native_avg:
                b["flat"].div_(self.world)
            dst, src = [], []
            for v, p in zip(b["views"], b["params"]):
                if p.grad is None:
                    p.grad = v.clone()
                else:
                    dst.append(p.grad)
                    src.append(v)
            if dst:
                torch._foreach_copy_(dst, src)
            b["pending"], b["work"], b["launched"] = len(b["params"]), None, False

    # ---- hook-free form for CUDA-graphed steps (train.GraphedTrainStep) ---------------------------------------
    def bind_static_grads(self):
        """After the backward pass has been captured into a CUDA graph the ``.grad`` tensors are static: remember them
        (graph replays do not run autograd hooks, so the hooks are dropped) and reduce with ``reduce_now()``."""
        self.remove()
        for b in self.buckets:
            self._ensure_flat(b)
            b["grads"] = [p.grad for p in b["params"]]
            if any(g is None for g in b["grads"]):
                raise RuntimeError("bind_static_grads: every parameter must have received a gradient in the captured step")

    def reduce_now(self):
        """Pack -> all-reduce(AVG) -> unpack for every bucket, on the current stream (world size 1: no-op)."""
        if self.world == 1:
            return
        op = dist.ReduceOp.AVG if self._native_avg else dist.ReduceOp.SUM
        for b in self.buckets:
            self._ensure_flat(b)
            grads = b.get("grads") or [p.grad for p in b["params"]]
            torch._foreach_copy_(b["views"], grads)
            dist.all_reduce(b["flat"], op=op, group=self.group)
            if not self._native_avg:
                b["flat"].div_(self.world)
            torch._foreach_copy_(grads, b["views"])
            self.allreduce_launches += 1
            b["pending"], b["work"], b["launched"] = len(b["params"]), None, False

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
