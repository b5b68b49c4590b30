"""CUDA-graphed train step and host->device prefetch for the TransMF_AD hot path.

The reference's train step (kfold_train_adversarial.py:101-136 / kfold_train_single.py:95-112) is

    optimizer.zero_grad(); outs = net_model(mri, pet); loss = criterion(...); loss.backward(); optimizer.step()

i.e. ~500 kernel launches of a few microseconds each from Python.  On a B200 the kernels of one step take about as
long as the host needs to enqueue them, so the eager loop is launch-bound.  ``GraphedTrainStep`` captures the same
sequence -- the unchanged ``nn.Module`` forward, the caller's loss function, autograd backward and the optimizer
step -- once into CUDA graphs with static input buffers and replays them; nothing in the model code changes, and the
eager path stays available (it is what the parity tests drive).

``DevicePrefetcher`` is the matching input side: batch i+1 is copied from pinned host memory on a side stream while
batch i computes (the reference loads synchronously with ``num_workers=0``, datasets/ADNI.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def _as_tuple(x):
    return x if isinstance(x, (tuple, list)) else (x,)


class GraphedTrainStep:
    """One train step as CUDA graph replays.

    ``loss_fn(outputs, *targets) -> (total_loss, *extras)`` must be made of capturable torch ops (no host syncs);
    ``total_loss.backward()`` is captured with the forward.  ``inputs`` / ``targets`` passed to ``__call__`` are
    copied device-to-device into the static buffers (tens of microseconds at HBM speed) and must keep the shapes and
    dtypes of the examples.  Returns the static ``(total_loss, *extras)`` tensors of the captured step: read them
    (``.item()``) before the next call.

    world size 1 : zero_grad + forward + loss + backward + optimizer.step in ONE graph.
    world size N : the same ONE graph with the data-parallel exchange inside it: ``reducer.finish()`` -- one ncclAllReduce
                   over the flat gradient buffer the backward kernels wrote into (``dp.FlatGradReducer``) -- is captured
                   between the backward pass and the optimizer.  ``TMF_DP_GRAPH=split`` keeps the collective outside
                   (graph A = forward + backward; eager all-reduce; graph B = optimizer).

    The optimizer must be graph-capturable (e.g. ``torch.optim.Adam(..., capturable=True)``).

    Warm-up.  CUDA-graph capture needs ``warmup`` real steps on the example batch first (allocator, lazy inits).  With
    ``restore_state=True`` (default) the model's parameters and buffers (BatchNorm running statistics,
    ``num_batches_tracked``) and the optimizer's moments / step counts are put back afterwards, in place, so that the
    first replay is step 1 of training exactly as in the eager loop (tests/test_gpu_train_path.py).
    """

    def __init__(self, model, optimizer, loss_fn, example_inputs, example_targets, reducer=None, warmup=3,
                 restore_state=True):
        self.model, self.optimizer, self.loss_fn, self.reducer = model, optimizer, loss_fn, reducer
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if self.world > 1 and reducer is None:
            raise ValueError("world size > 1 needs a GradBucketReducer")
        import os
        self.split = self.world > 1 and os.environ.get("TMF_DP_GRAPH", "fused") == "split"
        if reducer is not None:
            reducer.install()                              # weight gradients are written straight into the flat buffer
        self.launches_per_step = 0                         # libtmf kernel launches captured per step
        self.static_inputs = [t.clone() for t in _as_tuple(example_inputs)]
        self.static_targets = [t.clone() for t in _as_tuple(example_targets)]
        self.graph_fb, self.graph_opt = torch.cuda.CUDAGraph(), None
        self.outputs, self.losses = None, None
        self._capture(warmup, restore_state)

    # ---- the step, in eager form (used for warm-up and as the thing that is captured) ---------------------------
    def _fwd_bwd(self):
        self.optimizer.zero_grad(set_to_none=True)
        outs = self.model(*self.static_inputs)
        losses = _as_tuple(self.loss_fn(_as_tuple(outs), *self.static_targets))
        losses[0].backward()
        return outs, losses

    def _reduce(self):
        if self.world > 1:
            self.reducer.finish()

    # ---- state snapshot around the warm-up steps ------------------------------------------------------------------
    def _snapshot(self):
        model = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        opt = {}
        for p, st in self.optimizer.state.items():
            opt[p] = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
        return model, opt

    @torch.no_grad()
    def _restore(self, snap):
        model, opt = snap
        for k, v in self.model.state_dict().items():
            v.copy_(model[k])                              # in place: the captured graph keeps these addresses
        for p, st in self.optimizer.state.items():
            old = opt.get(p)
            for k, v in st.items():
                if torch.is_tensor(v):
                    if old is not None and k in old:
                        v.copy_(old[k])
                    else:
                        v.zero_()                          # state created by the warm-up: back to "no step taken"
                elif old is not None and k in old:
                    st[k] = old[k]
                elif isinstance(v, (int, float)):
                    st[k] = type(v)(0)

    def _capture(self, warmup, restore_state=True):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        snap = self._snapshot() if restore_state else None
        with torch.cuda.stream(side):                      # warm-up on a side stream (allocator + lazy inits + autotune)
            for _ in range(max(1, warmup)):
                self._fwd_bwd()
                self._reduce()
                self.optimizer.step()
            if snap is not None:
                self._restore(snap)
            # Cached bf16 conv-weight packs (functional.ConvPack).  FusedAdam refreshes them inside its kernel, so they are
            # packed once here and the captured step has no pack launches; any other optimizer updates the weights behind
            # the cache's back during replays, so the cache is invalidated and the pack launches are captured with the forward.
            emits = bool(getattr(self.optimizer, "emits_conv_packs", False))
            for m in self.model.modules():
                if hasattr(m, "pack_now"):
                    if emits:
                        m.pack_now()
                    else:
                        m.invalidate_packs()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        n0 = _lib.launch_count()
        if self.world == 1:
            with torch.cuda.graph(self.graph_fb):
                self.outputs, self.losses = self._fwd_bwd()
                self.optimizer.step()
        elif not self.split:
            # NCCL's watchdog thread polls CUDA events while we capture: only this thread's calls are policed
            with torch.cuda.graph(self.graph_fb, capture_error_mode="thread_local"):
                self.outputs, self.losses = self._fwd_bwd()
                self.reducer.finish()
                self.optimizer.step()
        else:
            with torch.cuda.graph(self.graph_fb, capture_error_mode="thread_local"):
                self.outputs, self.losses = self._fwd_bwd()
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt, pool=self.graph_fb.pool(), capture_error_mode="thread_local"):
                self.optimizer.step()
        self.launches_per_step = _lib.launch_count() - n0
        torch.cuda.synchronize()

    def __call__(self, inputs, targets):
        for dst, src in zip(self.static_inputs, _as_tuple(inputs)):
            dst.copy_(src, non_blocking=True)
        for dst, src in zip(self.static_targets, _as_tuple(targets)):
            dst.copy_(src, non_blocking=True)
        refresh = getattr(self.optimizer, "refresh_hyperparams", None)
        if refresh is not None:
            refresh()                                      # e.g. FusedAdam: a scheduler-changed lr reaches the device scalar
        from . import functional as TF
        TF.bump_param_epoch()                              # replays change parameters / running statistics without Python
        self.graph_fb.replay()
        if self.split:
            self.reducer.finish()
            self.graph_opt.replay()
        return self.losses


class DevicePrefetcher:
    """Iterates over host batches (tuples of pinned CPU tensors) and yields device batches; the copy of batch i+1 runs
    on a side stream while the caller works on batch i.  Two device buffer sets are recycled: a yielded batch stays
    valid until the next ``next()`` call has returned, and everything that reads it must have been enqueued on the
    current stream by then (single-threaded use)."""

    def __init__(self, host_batches, device, reuse=None, transform=None):
        """``reuse``: an exhausted prefetcher of the same batch shapes whose side stream and device buffers are taken over
        (one per epoch: no cudaMalloc / stream creation after the first).  ``transform``: a callable applied to the device
        batch ON THE COPY STREAM right after the copy (e.g. ``data.GpuBatchTransform``: scaling + augmentation), so that
        it overlaps the previous step like the copy itself; it must return the tuple the consumer gets."""
        self.transform = transform
        self.it = iter(host_batches)
        self.device = torch.device(device)
        if reuse is not None:
            self.stream, self.slots, self.consumed = reuse.stream, reuse.slots, reuse.consumed
            if reuse._last is not None:                     # readers of its last batch are already enqueued
                reuse.consumed[reuse._last].record(torch.cuda.current_stream(self.device))
                reuse._last = None
        else:
            self.stream = torch.cuda.Stream(device=self.device)
            self.slots = [None, None]
            self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.k = 0
        self._last = None
        self._next = None
        self._preload()

    def _preload(self):
        try:
            hb = next(self.it)
        except StopIteration:
            self._next = None
            return
        k = self.k
        self.k ^= 1
        with torch.cuda.stream(self.stream):
            if self.slots[k] is None:
                self.slots[k] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in hb)
            self.stream.wait_event(self.consumed[k])        # no-op until the slot has been handed out once
            for d, h in zip(self.slots[k], hb):
                d.copy_(h, non_blocking=True)
            out = self.slots[k] if self.transform is None else tuple(self.transform(self.slots[k]))
            ready = self.stream.record_event()
        self._next = (k, ready, out)

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        if self._last is not None:                          # all readers of the previous batch are enqueued by now
            self.consumed[self._last].record(cur)
        if self._next is None:
            raise StopIteration
        k, ready, batch = self._next
        cur.wait_event(ready)
        self._last = k
        for t in batch:                                     # transform outputs were allocated on the copy stream
            if torch.is_tensor(t):
                t.record_stream(cur)
        self._preload()                                     # batch i+1: H2D overlaps with the caller's work on batch i
        return batch


class LossReader:
    """Host-side reads of per-step scalars (the reference's ``loss.item()`` pair, kfold_train_adversarial.py:127-128)
    without stalling the launch queue: ``push(tensors)`` enqueues a device->pinned-host copy of this step's scalars and
    returns the PREVIOUS step's values (``None`` on the first call); ``flush()`` returns the last step's.  One step of
    lag keeps the next graph launch and H2D copy in flight while the current step computes."""

    def __init__(self, n, device):
        self.n = n
        self.host = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.stage = [torch.empty(n, dtype=torch.float32, device=device) for _ in range(2)]
        self.done = [torch.cuda.Event(), torch.cuda.Event()]
        self.k = 0
        self.pending = None

    def _wait(self):
        if self.pending is None:
            return None
        self.done[self.pending].synchronize()
        return tuple(float(v) for v in self.host[self.pending])

    def push(self, tensors):
        k = self.k
        self.k ^= 1
        for i, t in enumerate(tensors[: self.n]):
            self.stage[k][i:i + 1].copy_(t.detach().reshape(1), non_blocking=True)
        self.host[k].copy_(self.stage[k], non_blocking=True)
        self.done[k].record()
        prev = self._wait()
        self.pending = k
        return prev

    def flush(self):
        prev = self._wait()
        self.pending = None
        return prev
