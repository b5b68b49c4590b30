"""Data parallelism over subjects: one process per GPU, bucketed gradient all-reduce (NCCL over NVLink).

The reference is single-device (``cuda:0`` hard-coded, kfold_train_adversarial.py:24); data parallelism is added by
this build.  Every rank holds a full replica (4.17 M parameters, 16.7 MB of fp32 gradients),
runs the unchanged model on its shard of the global batch with per-rank BatchNorm statistics (DDP semantics), and
averages gradients before the optimizer step.

Overlap.  Gradients are packed into flat fp32 buckets in the order in which the backward pass PRODUCES them:
  1. heads and fusion transformer, in reverse registration order (their autograd nodes run first);
  2. the two sNet towers layer by layer, conv4.3 first and conv1.0 last -- the conv stack is ONE autograd node
     (``functional.SNetFunction``) whose parameter gradients would all surface at its end, so it hands every layer's
     gradients to the reducer the moment they exist (``functional.set_grad_sink``): 7 MB of the 10.7 MB of tower
     gradients (conv4.x) are on the wire while conv3 .. conv1 still compute.
A bucket's all-reduce is launched asynchronously as soon as its last gradient is packed; only the small last bucket
(blocks 1-2, 0.7 MB) is exposed.  Everything -- pack, ncclAllReduce, unpack -- is stream-ordered and capturable, so
``train.GraphedTrainStep`` records the whole step, collectives included, into ONE CUDA graph.
The same code runs on CPU tensors with the ``gloo`` backend (tests).
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


def readiness_order(model):
    """Parameters of ``model`` in the order the backward pass produces their gradients, as a list of GROUPS (a bucket
    boundary is allowed only between groups): [heads + fusion, reverse registration] then per conv layer 6..0 the
    parameters of that layer in every sNet tower.  Models without sNet towers degenerate to reverse registration."""
    from .models.networks import sNet
    towers = [m for m in model.modules() if isinstance(m, sNet)]
    tower_params = set()
    layers = [[] for _ in range(7)]
    for t in towers:
        for l, (conv, bn) in enumerate(t._units()):
            for p in (conv.weight, conv.bias, bn.weight, bn.bias):
                if p.requires_grad:
                    layers[l].append(p)
                    tower_params.add(p)
    rest = [p for p in model.parameters() if p.requires_grad and p not in tower_params][::-1]
    groups = [[p] for p in rest]
    for l in range(6, -1, -1):
        if layers[l]:
            groups.append(layers[l])
    return groups


class GradBucketReducer:
    """``params``: an iterable of parameters (bucketed in reverse order) or, better, ``model=`` an ``nn.Module`` (bucketed
    in gradient-readiness order, see the module docstring).  ``tail_bytes``: the buckets are cut so that the LAST one --
    the only one whose all-reduce cannot hide behind remaining backward work -- holds at most this many bytes."""

    def __init__(self, params=None, process_group=None, bucket_bytes: int = 4 << 20, model=None, tail_bytes: int = 1 << 20):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        if model is not None:
            groups = readiness_order(model)
        else:
            groups = [[p] for p in list(params) if p.requires_grad][::-1]
        self.params = [p for g in groups for p in g]
        self.buckets = []                   # list of dicts: params, offsets, numel, flat, views, pending, work
        # the tail: trailing groups that together fit tail_bytes form the last bucket
        tail_start, acc = len(groups), 0
        while tail_start > 1:
            nb = sum(p.numel() for p in groups[tail_start - 1]) * 4
            if acc + nb > tail_bytes:
                break
            acc += nb
            tail_start -= 1
        cur, cur_bytes = [], 0
        for gi, g in enumerate(groups):
            nbytes = sum(p.numel() for p in g) * 4
            if cur and (cur_bytes + nbytes > bucket_bytes or gi == tail_start):
                self._close(cur)
                cur, cur_bytes = [], 0
            cur.extend(g)
            cur_bytes += nbytes
        if cur:
            self._close(cur)
        self._where = {}
        for bi, b in enumerate(self.buckets):
            for i, p in enumerate(b["params"]):
                self._where[p] = (bi, i)
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
                             "pending": len(plist), "packed": [False] * len(plist), "work": None, "launched": False})

    def bucket_layout(self):
        """[(numel, n_params)] per bucket, in launch order (for tests / DESIGN.md)."""
        return [(b["numel"], len(b["params"])) for b in self.buckets]

    def _ensure_flat(self, b):
        if b["flat"] is None:
            ref = b["params"][0]
            b["flat"] = torch.zeros(b["numel"], dtype=torch.float32, device=ref.device)
            b["views"] = [b["flat"][o:o + p.numel()].view_as(p) for o, p in zip(b["offsets"], b["params"])]

    # ---- gradient arrival ------------------------------------------------------------------------------------------
    def early_grads(self, params, grads):
        """Called by ``functional.SNetFunction.backward`` (through ``functional.set_grad_sink``) with the finished
        gradients of one conv layer, long before autograd assigns them to ``.grad``: pack now, launch when full."""
        if self.world == 1:
            return
        per_bucket = {}
        for p, g in zip(params, grads):
            loc = self._where.get(p)
            if loc is None or g is None:
                continue
            per_bucket.setdefault(loc[0], []).append((loc[1], g))
        for bi, items in per_bucket.items():
            b = self.buckets[bi]
            self._ensure_flat(b)
            torch._foreach_copy_([b["views"][i] for i, _ in items], [g for _, g in items])
            for i, _ in items:
                b["packed"][i] = True
            b["pending"] -= len(items)
            if b["pending"] == 0 and not b["launched"]:
                self._launch(b)

    def _on_grad(self, p):
        bi, i = self._where[p]
        b = self.buckets[bi]
        if b["packed"][i]:                   # arrived early through the sink
            return
        b["pending"] -= 1
        if b["pending"] == 0 and not b["launched"]:
            self._launch(b)

    def _launch(self, b):
        self._ensure_flat(b)
        todo = [(v, p.grad) for v, p, done in zip(b["views"], b["params"], b["packed"]) if not done]
        have = [(v, g) for v, g in todo if g is not None]
        if len(have) != len(todo):
            for v, g in todo:
                if g is None:
                    v.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        op = dist.ReduceOp.AVG if self._native_avg else dist.ReduceOp.SUM
        b["work"] = dist.all_reduce(b["flat"], op=op, group=self.group, async_op=True)
        b["launched"] = True
        self.allreduce_launches += 1

    def finish(self):
        """Call after ``loss.backward()`` and before ``optimizer.step()``: launches any bucket whose parameters did not
        all receive a gradient, waits for the buckets (stream-ordered: no host sync on CUDA) and writes the averaged
        gradients back into ``.grad``."""
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
            b["packed"] = [False] * len(b["params"])

    # ---- hook-free form (kept for callers that drive the reduction themselves) ----------------------------------
    def reduce_now(self):
        """Pack -> all-reduce(AVG) -> unpack for every bucket, on the current stream (world size 1: no-op)."""
        if self.world == 1:
            return
        op = dist.ReduceOp.AVG if self._native_avg else dist.ReduceOp.SUM
        for b in self.buckets:
            self._ensure_flat(b)
            grads = [p.grad for p in b["params"]]
            torch._foreach_copy_(b["views"], grads)
            dist.all_reduce(b["flat"], op=op, group=self.group)
            if not self._native_avg:
                b["flat"].div_(self.world)
            torch._foreach_copy_(grads, b["views"])
            self.allreduce_launches += 1
            b["pending"], b["work"], b["launched"] = len(b["params"]), None, False
            b["packed"] = [False] * len(b["params"])

    def install(self):
        """Route the conv stack's per-layer gradients to this reducer (one active reducer per process)."""
        from . import functional as TF
        TF.set_grad_sink(self.early_grads if self.world > 1 else None)
        return self

    def remove(self):
        from . import functional as TF
        for h in self._hooks:
            h.remove()
        self._hooks = []
        if TF.get_grad_sink() == self.early_grads:
            TF.set_grad_sink(None)
