"""Fused Adam for the TransMF_AD train step.

The reference builds ``torch.optim.Adam(params, lr, weight_decay)`` (utils/utils.py:38-41).  On a B200 the ~100
parameter tensors of these models make that optimizer ~200 tiny launches per step -- as long as a convolution layer.
``FusedAdam`` has the same constructor arguments, ``param_groups`` / ``state_dict`` layout (``step``, ``exp_avg``,
``exp_avg_sq`` per parameter) and arithmetic, but one ``tmf_adam_step`` launch per parameter group; the step count and the
learning rate are device scalars, so the step can be captured in a CUDA graph and LR schedulers keep working.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L

CHUNK = 16384          # elements per block
TILE_CO, TILE_CI = 16, 32      # (co, ci) tile of a packed conv weight per block (csrc/adam.cu: ADAM_TCO, ADAM_TCI)


class FusedAdam(torch.optim.Optimizer):
    emits_conv_packs = True         # the step kernel rewrites the cached bf16 conv-weight packs (functional.ConvPack)

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        if amsgrad:
            raise ValueError("FusedAdam: amsgrad is not supported (the reference does not use it)")
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("FusedAdam: invalid hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))
        self._tables = {}

    # ---- per-group device state: step counter, lr, ticket, chunk table ------------------------------------------------
    def _group_state(self, gi, group):
        st = self._tables.get(gi)
        params = [p for p in group["params"] if p.grad is not None]
        if not params:
            return None
        for p in params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters (there is no CPU fallback)")
            if p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                raise RuntimeError("FusedAdam needs contiguous fp32 gradients")
        dev = params[0].device
        if st is None:
            st = {"step": torch.zeros(1, dtype=torch.float32, device=dev), "lr": torch.empty(1, dtype=torch.float32, device=dev),
                  "lr_host": None, "ticket": torch.zeros(1, dtype=torch.int32, device=dev), "key": None, "table": None, "n": 0}
            self._tables[gi] = st
        for p in params:
            s = self.state[p]
            if "exp_avg" not in s:
                s["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                s["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if s.get("step") is not st["step"]:
                # state that came from load_state_dict (ours or torch.optim.Adam's) carries its own step value: the
                # group's device counter -- the only one the kernel reads -- is seeded from it, then shared again
                old = s.get("step")
                if old is not None:
                    val = float(old)
                    if st.get("seeded") is not None and st["seeded"] != val:
                        raise RuntimeError("FusedAdam: parameters of one group carry different step counts "
                                           f"({st['seeded']} vs {val}); one device counter per group is supported")
                    if st.get("seeded") is None:
                        st["step"].fill_(val)
                        st["seeded"] = val
                s["step"] = st["step"]                      # one device counter shared by the group (same value for all)
        def pack_of(p):
            # conv weights whose bf16 operand packs are cached by the module (functional.ConvPack): the kernel refreshes them
            ep = getattr(p, "_tmf_encpack", None)
            if ep is not None and ep[0] is not None and ep[0].device == p.device:
                return (ep[0].data_ptr(), ep[1].data_ptr(), 0, 0, -1)       # taps = -1: encoder hi / lo pack (functional.EncPack)
            pk = getattr(p, "_tmf_pack", None)
            if pk is None or pk[0] is None or pk[0].device != p.device:
                return (0, 0, 0, 0, 0)
            wf, wd, cout, cin, taps = pk
            return (wf.data_ptr(), 0 if wd is None else wd.data_ptr(), cout, cin, taps)

        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr(), p.numel()) + pack_of(p) for p in params)
        if key != st["key"]:
            rec = int(L.load().tmf_adam_chunk_bytes())
            rows = []
            for pp, gp, mp, vp, n, wf, wd, cout, cin, taps in key:
                if wf and taps > 0 and cout > 0 and cin > 0 and cout % TILE_CO == 0 and cin % TILE_CI == 0 and taps <= 27:
                    # conv weight with operand packs: one block per (16 co x 32 ci) tile with all its taps -- the kernel
                    # transposes the tile in shared memory so that the bf16 pack stores are whole 32 / 64-byte runs
                    for co0 in range(0, cout, TILE_CO):
                        for ci0 in range(0, cin, TILE_CI):
                            rows.append((pp, gp, mp, vp, wf, wd, TILE_CO * TILE_CI * taps, co0 * cin + ci0, cout, cin, taps, 1))
                    continue
                enc = taps == -1                    # wf / wd = the hi / lo arrays of an encoder weight (same element order)
                for off in range(0, n, CHUNK):
                    rows.append((pp + 4 * off, gp + 4 * off, mp + 4 * off, vp + 4 * off, wf, wd, min(CHUNK, n - off), off,
                                 cout, cin, 0 if enc else taps, 2 if enc else 0))
            arr = np.array(rows, dtype=np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("wf", "<u8"),
                                                 ("wd", "<u8"), ("n", "<i4"), ("off", "<i4"), ("cout", "<i4"), ("cin", "<i4"),
                                                 ("taps", "<i4"), ("tile", "<i4")]))
            assert arr.dtype.itemsize == rec
            # pinned staging buffer + async copy: legal inside CUDA-graph capture (the captured backward hands out
            # new, then static, gradient buffers, so the table is rebuilt once while capturing)
            host = torch.from_numpy(arr.view(np.uint8).copy()).pin_memory()
            st["table_host"] = host
            st["table"] = torch.empty(host.numel(), dtype=torch.uint8, device=dev)
            st["table"].copy_(host, non_blocking=True)
            st["n"] = len(rows)
            st["key"] = key
        self._sync_lr(st, group)
        return st

    @staticmethod
    def _sync_lr(st, group):
        if st["lr_host"] != group["lr"]:                    # LR schedulers change group["lr"] between steps
            if st.get("lr_pinned") is None:
                st["lr_pinned"] = torch.empty(1, dtype=torch.float32).pin_memory()
            st["lr_pinned"][0] = float(group["lr"])
            st["lr"].copy_(st["lr_pinned"], non_blocking=True)
            st["lr_host"] = group["lr"]

    def load_state_dict(self, state_dict):
        """torch.optim.Adam-compatible: moments and the step count are restored (the per-parameter ``step`` entries are
        re-bound to the group's device counter on the next ``step()``); chunk tables are rebuilt."""
        super().load_state_dict(state_dict)
        self._tables = {}

    def refresh_hyperparams(self):
        """Push a changed learning rate to the device without running ``step()`` (CUDA-graph replays skip the Python
        side of ``step``; ``GraphedTrainStep`` calls this before every replay)."""
        for gi, group in enumerate(self.param_groups):
            st = self._tables.get(gi)
            if st is not None:
                self._sync_lr(st, group)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            st = self._group_state(gi, group)
            if st is None:
                continue
            b1, b2 = group["betas"]
            L.call("tmf_adam_step", L.ptr(st["table"]), st["n"], L.ptr(st["lr"]), float(b1), float(b2), float(group["eps"]),
                   float(group["weight_decay"]), L.ptr(st["step"]), L.ptr(st["ticket"]))
            from . import functional as TF
            TF.bump_param_epoch()                           # parameters changed behind torch's version counters
            for p in group["params"]:                       # the kernel rewrote these parameters' packs from the new values
                cache = getattr(p, "_tmf_pack_cache", None)
                if cache is not None and p.grad is not None:
                    cache.mark(p)
        return loss
