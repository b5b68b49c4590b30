"""Host-side operators: thin wrappers over the C ABI (include/tmf.h) plus the ``torch.autograd.Function``s that
give the reference's ``nn.Module``s their forward and backward on the B200 kernels.

Nothing here computes on the CPU or through a torch library kernel on the hot path: tensors are allocated with
torch (caching allocator), every arithmetic op is a ``tmf_*`` launch on the current CUDA stream.
"""
from __future__ import annotations

import os

import torch

from . import _lib as L

BN_EPS, BN_MOMENTUM, LRELU_SLOPE, LN_EPS = 1e-5, 0.1, 0.01, 1e-5


# Data parallelism (dp.FlatGradReducer.install): parameter data_ptr -> (slot view of the flat gradient buffer, parameter).
# The backward kernels write weight gradients straight into the slots, autograd adopts the view as ``.grad`` without a
# copy, and the gradient exchange is one all-reduce over the flat buffer.
_GRAD_SLOTS = None


def set_grad_slots(table):
    global _GRAD_SLOTS
    _GRAD_SLOTS = table
    _SLOT_HANDED.clear()


_SLOT_HANDED = {}        # parameter address -> id of the backward pass (autograd graph task) that last received its slot


def grad_out(param, shape=None, dtype=torch.float32):
    """Output tensor for the gradient of ``param``: its slot of the flat buffer when one is registered (and the parameter
    is not accumulating into an existing ``.grad``), else a fresh tensor.

    A module that runs TWICE in one forward pass (the discriminator ``D`` of model_ad / model_CNN_ad, reference
    mymodel.py:210-211) asks twice during one backward pass: only the first request gets the slot -- a second kernel
    writing into the same memory would overwrite the contribution autograd still holds for the sum (found by the 2-rank
    NCCL parity test: D.0.weight was twice the PET contribution).  Autograd then adds slot + fresh tensor out of place and
    ``FlatGradReducer.finish()`` copies the sum into the slot."""
    if _GRAD_SLOTS is not None and param is not None:
        hit = _GRAD_SLOTS.get(param.data_ptr())
        if hit is not None:
            slot, owner = hit
            if owner.grad is None and (shape is None or tuple(slot.shape) == tuple(shape)) and slot.dtype == dtype:
                task = torch._C._current_graph_task_id()
                key = param.data_ptr()
                if task < 0 or _SLOT_HANDED.get(key) != task:
                    _SLOT_HANDED[key] = task
                    return slot.detach()    # a fresh alias: autograd adopts it as .grad only if nobody else holds it
    return torch.empty(tuple(param.shape) if shape is None else tuple(shape), dtype=dtype, device=param.device)


def conv_impl():
    """TMF_CONV_IMPL = auto | direct | umma  (bring-up / cross-check switch; default auto)."""
    return {"auto": L.CONV_AUTO, "direct": L.CONV_DIRECT, "umma": L.CONV_UMMA}[os.environ.get("TMF_CONV_IMPL", "auto")]


def keep_ymax():
    """TMF_KEEP_YMAX=0: max-pool layers re-read y in the backward reduction (cross-check switch; default keeps ymax)."""
    return os.environ.get("TMF_KEEP_YMAX", "1") != "0"


def _f32c(t):
    """fp32, contiguous, plain torch.Tensor (MONAI MetaTensor inputs are unwrapped)."""
    if hasattr(t, "as_tensor"):
        t = t.as_tensor()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ==============================================================================================================
# sNet conv stack  (reference models/networks.py:18-61)
# ==============================================================================================================
class SNetSpec:
    """Static description of one sNet: per layer (Cin, Cout, ksize, pool)."""

    def __init__(self, dim):
        if dim < 32 or dim % 32 != 0:
            raise ValueError(f"sNet(dim={dim}): the B200 kernels work on 8-channel vectors of dim/4 channels; "
                             f"dim must be a positive multiple of 32")
        q, h = dim // 4, dim // 2
        self.layers = [(1, q, 3, L.POOL_MAX), (q, q, 3, L.POOL_NONE), (q, h, 3, L.POOL_MAX), (h, h, 3, L.POOL_NONE),
                       (h, dim, 3, L.POOL_MAX), (dim, 2 * dim, 3, L.POOL_NONE), (2 * dim, dim, 1, L.POOL_AVG)]
        self.dim = dim


class ConvPack:
    """bf16 operand packs of one Conv3d weight -- wf[tap][Cout][Cin] (forward) and wd[taps-1-tap][Cin][Cout] (dgrad) -- kept
    across steps.  They are fresh while ``weight._version`` is the one they were packed from: torch in-place updates
    (``torch.optim``, ``load_state_dict``) bump it and force a repack; ``optim.FusedAdam`` rewrites the packs from its own
    kernel (the weight's version does not move), so a train step on FusedAdam has no pack launches at all."""

    def __init__(self):
        self.wf = self.wd = None
        self.version, self.ptr = None, None

    def stale(self, w):
        cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
        taps = k ** 3
        if self.wf is None or self.ptr != w.data_ptr() or self.wf.device != w.device:
            self.wf = torch.empty((taps, cout, cin), dtype=torch.bfloat16, device=w.device)
            self.wd = torch.empty((taps, cin, cout), dtype=torch.bfloat16, device=w.device)
            self.version, self.ptr = None, w.data_ptr()
        if getattr(w, "_tmf_pack_cache", None) is not self:
            w._tmf_pack = (self.wf, self.wd, cout, cin, taps)      # optim.FusedAdam reads these
            w._tmf_pack_cache = self
        return self.version != w._version

    def mark(self, w):
        self.version = w._version


class EncPack:
    """bf16 hi / lo copies of one encoder's five weight matrices (the operand pack of csrc/enc_fused.cu: per matrix
    [hi | lo], in the order Wq, Wkv, Wo, W1, W2), kept across steps like ``ConvPack``: fresh while every weight's ``_version`` is
    the one it was packed from; ``optim.FusedAdam`` rewrites the hi / lo values from its own kernel (``_tmf_encpack``), so a
    train step on FusedAdam has no ``tmf_encoder_pack_weights`` launches (6 per ``model_ad`` step before)."""

    def __init__(self):
        self.pack, self.key, self.versions = None, None, {}

    def stale(self, ws, mlp):
        key = tuple(w.data_ptr() for w in ws) + (str(ws[0].device),)
        if self.pack is None or key != self.key:
            self.pack = torch.empty(int(L.load().tmf_encoder_pack_bytes(int(mlp))) // 2, dtype=torch.bfloat16, device=ws[0].device)
            self.key, self.versions = key, {}
        off = 0
        for w in ws:
            n = w.numel()
            if getattr(w, "_tmf_pack_cache", None) is not self or getattr(w, "_tmf_encpack", (None,))[0] is None \
                    or w._tmf_encpack[0].data_ptr() != self.pack.data_ptr() + 2 * off:
                w._tmf_encpack = (self.pack[off:off + n], self.pack[off + n:off + 2 * n])     # optim.FusedAdam reads these
                w._tmf_pack_cache = self
            off += 2 * n
        return any(self.versions.get(w.data_ptr()) != w._version for w in ws)

    def mark(self, w):
        self.versions[w.data_ptr()] = w._version


# Bumped whenever parameters or BatchNorm running statistics change behind torch's version counters (FusedAdam's kernel,
# train-mode forwards, CUDA-graph replays): the eval-mode BatchNorm folds below are rebuilt when it moves.
_PARAM_EPOCH = [0]


def bump_param_epoch():
    _PARAM_EPOCH[0] += 1


class EvalFold:
    """Eval-mode BatchNorm3d folded into one conv layer's operands (csrc/eval_ops.cu): w*scale as the bf16 pack (or fp32 for
    conv1.0) and b' = (b - running_mean)*scale + beta; plus the identity coefficients the LeakyReLU/pool pass then takes."""

    def __init__(self):
        self.sig = None
        self.wf = self.w32 = self.bias = self.ident = None

    def stale(self, l, tensors):
        w = tensors[0]
        sig = (_PARAM_EPOCH[0],) + tuple((t.data_ptr(), t._version) for t in tensors)
        if self.bias is None or self.bias.device != w.device:
            cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
            if l == 0:
                self.w32 = torch.empty_like(w)
            else:
                self.wf = torch.empty((k ** 3, cout, cin), dtype=torch.bfloat16, device=w.device)
            self.bias = torch.empty(cout, dtype=torch.float32, device=w.device)
            self.ident = torch.zeros(4 * cout, dtype=torch.float32, device=w.device)
            self.ident[:cout] = 1.0
            self.ident[3 * cout:] = 1.0
            self.sig = None
        return sig != self.sig, sig


class SNetRun:
    """Per-call settings of the conv stack: train / eval, whether autograd is recording (``torch.is_grad_enabled()`` at
    the call site: inside ``Function.forward`` grad mode is always off), and the hyper-parameters read from the
    ``nn.BatchNorm3d`` / ``nn.LeakyReLU`` children -- per layer (eps, momentum, negative_slope)."""

    def __init__(self, training, grad_enabled, hyper=None, packs=None, folds=None, sync_group=False):
        self.training, self.grad_enabled = bool(training), bool(grad_enabled)
        self.hyper = hyper if hyper is not None else [(BN_EPS, BN_MOMENTUM, LRELU_SLOPE)] * 7
        self.packs = packs              # per tower: 7 ConvPack (index 0 unused: conv1.0 consumes the fp32 weight)
        self.folds = folds              # per tower: 7 EvalFold (inference path: BatchNorm folded into the conv operands)
        # False: per-rank BatchNorm statistics (DDP semantics, the default).  None / a process group: the BatchNorm3d children
        # are nn.SyncBatchNorm (torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)): statistics over the GLOBAL batch
        self.sync_group = sync_group


def _sync_world(group):
    """World size of the SyncBatchNorm exchange (1: nothing to exchange)."""
    import torch.distributed as dist
    if group is False or not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


def _allreduce_stat_rows(bufs, group):
    """SyncBatchNorm: replace the per-CTA partial rows of every tower's statistics buffer by the sum over all ranks (row 0 =
    the global totals, the other rows zero), so that the finalize kernels see global-batch statistics."""
    import torch.distributed as dist
    tot = torch.stack([b.sum(0) for b in bufs])
    dist.all_reduce(tot, group=group)
    for b, t in zip(bufs, tot):
        b.zero_()
        b[0].copy_(t)


def wgrad_workspace(ng, impl, B, D, H, W, cin, cout, ks, dev):
    """Caller-owned scratch for the split-K partials of the tcgen05 weight-gradient kernel (None if not needed)."""
    n = int(L.load().tmf_conv3d_wgrad_workspace_bytes(ng, impl, B, D, H, W, cin, cout, ks))
    return torch.empty(n, dtype=torch.uint8, device=dev) if n > 0 else None


def _pooled(D, H, W, pool):
    return (D, H, W) if pool == L.POOL_NONE else (D // 2, H // 2, W // 2)


class SNetFunction(torch.autograd.Function):
    """Both towers (or one) of the 3D-CNN encoder as ONE autograd node.

    apply(spec, run, buffers, ng, x_0[, x_1], *params) with run an ``SNetRun`` and params = for each tower, for each of the 7 layers:
    conv.weight, conv.bias, bn.weight, bn.bias;  buffers[t][l] = (running_mean, running_var, num_batches_tracked).
    Returns one fp32 tensor per tower with logical shape (B, dim, d, h, w) in channels-last-3d memory, i.e. the
    token matrix (B, d*h*w, dim) is a free view of it.
    """

    @staticmethod
    def forward(ctx, spec, run, buffers, ng, *args):
        if not isinstance(run, SNetRun):
            run = SNetRun(bool(run), True)
        training = run.training
        if run.grad_enabled and any(ctx.needs_input_grad[4:4 + ng]):
            raise RuntimeError("sNet: gradients with respect to the input volumes are not implemented (conv1.0 has no "
                               "dgrad kernel); detach the inputs or compute saliency with the reference modules")
        xs = [_f32c(a) for a in args[:ng]]
        params = args[ng:]
        assert len(params) == ng * 7 * 4
        B, cin0, D, H, W = xs[0].shape
        if cin0 != 1:
            raise RuntimeError(f"sNet expects single-channel volumes (B,1,D,H,W); got {tuple(xs[0].shape)}")
        for x in xs:
            if tuple(x.shape) != tuple(xs[0].shape):
                raise RuntimeError("MRI and PET volumes must have the same shape")
        dev = xs[0].device
        need_grad = run.grad_enabled and any(ctx.needs_input_grad[4 + ng:])
        impl = conv_impl()
        P = lambda t, l, k: params[(t * 7 + l) * 4 + k]
        if not training and not need_grad and run.folds is not None and os.environ.get("TMF_EVAL_FOLD", "1") != "0":
            return SNetFunction._eval_forward(ctx, spec, run, buffers, ng, xs, P, impl)
        if training:
            bump_param_epoch()                              # running statistics are about to change
        saved = []
        act = xs
        dims = (D, H, W)
        sync_world = _sync_world(run.sync_group) if training else 1
        for l, (cin, cout, ks, pool) in enumerate(spec.layers):
            Dl, Hl, Wl = dims
            count = B * Dl * Hl * Wl
            bn_eps, bn_momentum, slope = run.hyper[l]
            y = [torch.empty((B, Dl, Hl, Wl, cout), dtype=torch.bfloat16, device=dev) for _ in range(ng)]
            # per-CTA partial rows, summed in a fixed order by tmf_bn_finalize (deterministic; never zeroed)
            stats = L.stat_buffers(ng, cout, dev) if training else None
            w = [P(t, l, 0) for t in range(ng)]
            b = [P(t, l, 1) for t in range(ng)]
            wd = None
            if l == 0:
                L.call("tmf_conv1_fwd", ng, L.ptrs(act), L.ptrs(w), L.ptrs(b), L.ptrs(y), L.ptrs(stats),
                       B, Dl, Hl, Wl, cout, impl)
            else:
                taps = ks ** 3
                if run.packs is not None:
                    caches = [run.packs[t][l] for t in range(ng)]
                    stale = [c.stale(wt) for c, wt in zip(caches, w)]
                    wf, wd = [c.wf for c in caches], [c.wd for c in caches]
                    if any(stale):
                        L.call("tmf_pack_conv_weights", ng, L.ptrs(w), L.ptrs(wf), L.ptrs(wd), cout, cin, ks)
                        for c, wt in zip(caches, w):
                            c.mark(wt)
                else:
                    wf = [torch.empty((taps, cout, cin), dtype=torch.bfloat16, device=dev) for _ in range(ng)]
                    if need_grad:
                        wd = [torch.empty((taps, cin, cout), dtype=torch.bfloat16, device=dev) for _ in range(ng)]
                    L.call("tmf_pack_conv_weights", ng, L.ptrs(w), L.ptrs(wf), L.ptrs(wd), cout, cin, ks)
                L.call("tmf_conv3d_fwd", ng, L.ptrs(act), L.ptrs(wf), L.ptrs(b), L.ptrs(y), L.ptrs(stats),
                       B, Dl, Hl, Wl, cin, cout, ks, impl, tag=f"tmf_conv3d_fwd@L{l}")
            if training and sync_world > 1:
                _allreduce_stat_rows(stats, run.sync_group)           # SyncBatchNorm: global-batch mean / variance
                count *= sync_world
            coef = list(torch.empty((ng, 4 * cout), dtype=torch.float32, device=dev).unbind(0))
            L.call("tmf_bn_finalize", ng, L.ptrs(stats), L.ptrs([P(t, l, 2) for t in range(ng)]),
                   L.ptrs([P(t, l, 3) for t in range(ng)]), L.ptrs([buffers[t][l][0] for t in range(ng)]),
                   L.ptrs([buffers[t][l][1] for t in range(ng)]), L.ptrs([buffers[t][l][2] for t in range(ng)]),
                   L.ptrs(coef), cout, count, bn_momentum, bn_eps, int(training))
            last = l == len(spec.layers) - 1
            Do, Ho, Wo = _pooled(Dl, Hl, Wl, pool)
            out = [torch.empty((B, Do, Ho, Wo, cout), dtype=torch.float32 if last else torch.bfloat16, device=dev)
                   for _ in range(ng)]
            ymax = None
            if need_grad and pool == L.POOL_MAX and keep_ymax():
                # max-pool layers keep the pre-BN value behind every window maximum: the backward reduction then reads
                # 1/8 of the voxels instead of all of y
                ymax = [torch.empty((B, Do, Ho, Wo, cout), dtype=torch.bfloat16, device=dev) for _ in range(ng)]
                L.call("tmf_bn_act_pool_fwd_keepmax", ng, L.ptrs(y), L.ptrs(coef), L.ptrs(out), L.ptrs(ymax), int(last),
                       B, Dl, Hl, Wl, cout, slope, tag=f"tmf_bn_act_pool_fwd@L{l}")
            else:
                L.call("tmf_bn_act_pool_fwd", ng, L.ptrs(y), L.ptrs(coef), L.ptrs(out), int(last), B, Dl, Hl, Wl, cout,
                       pool, slope, tag=f"tmf_bn_act_pool_fwd@L{l}")
            if need_grad:
                saved.append((act, y, coef, wd, dims, ymax))
            act = out
            dims = (Do, Ho, Wo)
        # The fused block-1 backward (conv1_bwd_fused.cu) reads the input image as six bf16 hi/lo planes.  That split depends
        # on x alone: it is launched HERE, behind the last forward launch of the towers, on a side stream (a parallel branch of a
        # captured graph), so that it runs beside the fusion transformer / heads / the small deep layers of the backward pass,
        # which leave most of the GPU idle -- instead of in front of the backward kernel (145 MB written, 42 us at B = 8).
        # (Launched at the START of the forward pass it ran beside conv1.0 forward and cost that kernel what it saved.)
        # OFF by default (TMF_C1B_PRESPLIT=1): measured on B200 over 60 graph replays the step is 4.304 ms with the early split
        # and 4.295 ms without -- the parallel branch costs the kernels it overlaps what it saves the backward pass.
        ctx.c1split = None
        cout0, pool0 = spec.layers[0][1], spec.layers[0][3]
        if (need_grad and pool0 == L.POOL_MAX and impl != L.CONV_DIRECT and len(spec.layers) > 1
                and os.environ.get("TMF_C1B_PRESPLIT", "0") != "0"):
            nws0 = int(L.load().tmf_conv1_bwd_fused_workspace_bytes(ng, B, D, H, W, cout0))
            if nws0 > 0:
                ws0 = torch.empty(nws0, dtype=torch.uint8, device=dev)
                cur = torch.cuda.current_stream(dev)
                side = _side_stream(dev)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    L.call("tmf_conv1_bwd_split_x", ng, L.ptrs(xs), B, D, H, W, cout0, L.ptr(ws0), nws0)
                    ev = torch.cuda.Event()
                    ev.record(side)
                ws0.record_stream(side)
                for x in xs:
                    x.record_stream(side)
                ctx.c1split = (ws0, nws0, ev)
        ctx.spec, ctx.training, ctx.ng, ctx.saved, ctx.B = spec, training, ng, saved, B
        ctx.hyper = run.hyper
        ctx.sync_group, ctx.sync_world = run.sync_group, sync_world
        ctx.params = params if need_grad else None          # references only (gradient slots of the flat DP buffer)
        ctx.impl = impl
        outs = tuple(o.permute(0, 4, 1, 2, 3) for o in act)      # logical (B,C,d,h,w), channels-last memory
        return outs if ng > 1 else outs[0]

    @staticmethod
    def _eval_forward(ctx, spec, run, buffers, ng, xs, P, impl):
        """Inference (val_step, reference kfold_train_adversarial.py:144-161): no statistics, no saved activations; the
        BatchNorm of every layer is folded into the conv operands once per weight / running-statistics state."""
        B, _, D, H, W = xs[0].shape
        dev = xs[0].device
        act, dims = xs, (D, H, W)
        for l, (cin, cout, ks, pool) in enumerate(spec.layers):
            Dl, Hl, Wl = dims
            bn_eps, _, slope = run.hyper[l]
            folds = [run.folds[t][l] for t in range(ng)]
            state = [fo.stale(l, [P(t, l, 0), P(t, l, 1), P(t, l, 2), P(t, l, 3), buffers[t][l][0], buffers[t][l][1]])
                     for t, fo in enumerate(folds)]
            if any(st for st, _ in state):
                L.call("tmf_fold_bn_pack", ng, L.ptrs([P(t, l, 0) for t in range(ng)]), L.ptrs([P(t, l, 1) for t in range(ng)]),
                       L.ptrs([P(t, l, 2) for t in range(ng)]), L.ptrs([P(t, l, 3) for t in range(ng)]),
                       L.ptrs([buffers[t][l][0] for t in range(ng)]), L.ptrs([buffers[t][l][1] for t in range(ng)]),
                       L.ptrs([fo.wf for fo in folds]), L.ptrs([fo.w32 for fo in folds]), L.ptrs([fo.bias for fo in folds]),
                       cout, cin, ks, float(bn_eps))
                for fo, (_, sig) in zip(folds, state):
                    fo.sig = sig
            y = [torch.empty((B, Dl, Hl, Wl, cout), dtype=torch.bfloat16, device=dev) for _ in range(ng)]
            bias = [fo.bias for fo in folds]
            if l == 0:
                L.call("tmf_conv1_fwd", ng, L.ptrs(act), L.ptrs([fo.w32 for fo in folds]), L.ptrs(bias), L.ptrs(y), L.ptrs(None),
                       B, Dl, Hl, Wl, cout, impl)
            else:
                L.call("tmf_conv3d_fwd", ng, L.ptrs(act), L.ptrs([fo.wf for fo in folds]), L.ptrs(bias), L.ptrs(y), L.ptrs(None),
                       B, Dl, Hl, Wl, cin, cout, ks, impl, tag=f"tmf_conv3d_fwd@L{l}")
            last = l == len(spec.layers) - 1
            Do, Ho, Wo = _pooled(Dl, Hl, Wl, pool)
            out = [torch.empty((B, Do, Ho, Wo, cout), dtype=torch.float32 if last else torch.bfloat16, device=dev)
                   for _ in range(ng)]
            L.call("tmf_bn_act_pool_fwd", ng, L.ptrs(y), L.ptrs([fo.ident for fo in folds]), L.ptrs(out), int(last), B, Dl, Hl,
                   Wl, cout, pool, slope, tag=f"tmf_bn_act_pool_fwd@L{l}")
            act, dims = out, (Do, Ho, Wo)
        ctx.saved = None
        outs = tuple(o.permute(0, 4, 1, 2, 3) for o in act)
        return outs if ng > 1 else outs[0]

    @staticmethod
    def backward(ctx, *grads):
        spec, ng, B, training = ctx.spec, ctx.ng, ctx.B, ctx.training
        saved = ctx.saved
        if saved is None or len(saved) != len(spec.layers):
            raise RuntimeError("sNet backward: the saved activations are gone -- backward through this forward ran "
                               "already (they are freed layer by layer; retain_graph / double backward are not "
                               "supported), or the forward ran without autograd recording")
        dev = saved[0][1][0].device
        dout = []
        for t in range(ng):
            g = grads[t]
            cout = spec.layers[-1][1]
            if g is None:
                Dl, Hl, Wl = _pooled(*saved[-1][4], spec.layers[-1][3])
                g = torch.zeros((B, Dl, Hl, Wl, cout), dtype=torch.float32, device=dev)
            else:
                g = _f32c(g.permute(0, 2, 3, 4, 1))
            dout.append(g)
        dout_fp32 = 1
        pgrads = [None] * (ng * 7 * 4)
        for l in range(len(spec.layers) - 1, -1, -1):
            cin, cout, ks, pool = spec.layers[l]
            act, y, coef, wd, (Dl, Hl, Wl), ymax = saved[l]
            count = B * Dl * Hl * Wl
            slope = ctx.hyper[l][2]
            sums = L.stat_buffers(ng, cout, dev)
            if ymax is not None:
                L.call("tmf_bn_maxpool_bwd_reduce_kept", ng, L.ptrs(dout), dout_fp32, L.ptrs(ymax), L.ptrs(coef),
                       L.ptrs(sums), B, Dl // 2, Hl // 2, Wl // 2, cout, slope,
                       tag=f"tmf_bn_act_pool_bwd_reduce@L{l}")
            else:
                L.call("tmf_bn_act_pool_bwd_reduce", ng, L.ptrs(dout), dout_fp32, L.ptrs(y), L.ptrs(coef), L.ptrs(sums),
                       B, Dl, Hl, Wl, cout, pool, slope, tag=f"tmf_bn_act_pool_bwd_reduce@L{l}")
            PG = lambda t, k: ctx.params[(t * 7 + l) * 4 + k]
            dbias = [grad_out(PG(t, 1)) for t in range(ng)]
            dgamma = [grad_out(PG(t, 2)) for t in range(ng)]
            dbeta = [grad_out(PG(t, 3)) for t in range(ng)]
            bcoef = list(torch.empty((ng, 2 * cout), dtype=torch.float32, device=dev).unbind(0))
            L.call("tmf_bn_bwd_finalize", ng, L.ptrs(sums), L.ptrs(coef), L.ptrs(dgamma), L.ptrs(dbeta),
                   L.ptrs(dbias), L.ptrs(bcoef), cout, count, int(training))
            if training and ctx.sync_world > 1:
                # SyncBatchNorm: the parameter gradients above are this rank's own (the gradient exchange averages them);
                # the input gradient needs the GLOBAL means of dz and dz * xhat
                _allreduce_stat_rows(sums, ctx.sync_group)
                scratch_g = list(torch.empty((3 * ng, cout), dtype=torch.float32, device=dev).unbind(0))
                L.call("tmf_bn_bwd_finalize", ng, L.ptrs(sums), L.ptrs(coef), L.ptrs(scratch_g[:ng]), L.ptrs(scratch_g[ng:2 * ng]),
                       L.ptrs(scratch_g[2 * ng:]), L.ptrs(bcoef), cout, count * ctx.sync_world, int(training))
            dw = [grad_out(PG(t, 0)) for t in range(ng)]
            fused_ws = 0
            if l == 0 and pool == L.POOL_MAX and not dout_fp32 and ctx.impl != L.CONV_DIRECT:
                fused_ws = int(L.load().tmf_conv1_bwd_fused_workspace_bytes(ng, B, Dl, Hl, Wl, cout))
            if fused_ws > 0:
                # block 1: BN/LeakyReLU/MaxPool backward "apply" + conv1.0 weight gradient in one pass; dy stays on chip
                if ctx.c1split is not None and ctx.c1split[1] == fused_ws:
                    ws, _, ev = ctx.c1split
                    torch.cuda.current_stream(dev).wait_event(ev)        # the image split launched by the forward pass
                    L.call("tmf_conv1_bwd_fused_presplit", ng, L.ptrs(dout), L.ptrs(y), L.ptrs(coef), L.ptrs(bcoef),
                           L.ptrs(dw), B, Dl, Hl, Wl, cout, slope, L.ptr(ws), fused_ws, tag="tmf_conv1_bwd_fused")
                    ctx.c1split = None
                else:
                    ws = torch.empty(fused_ws, dtype=torch.uint8, device=dev)
                    L.call("tmf_conv1_bwd_fused", ng, L.ptrs(dout), L.ptrs(y), L.ptrs(coef), L.ptrs(bcoef), L.ptrs(act),
                           L.ptrs(dw), B, Dl, Hl, Wl, cout, slope, L.ptr(ws), fused_ws)
                SNetFunction._layer_grads(ctx, pgrads, l, dw, dbias, dgamma, dbeta)
                saved[l] = None
                continue
            dy = [torch.empty((B, Dl, Hl, Wl, cout), dtype=torch.bfloat16, device=dev) for _ in range(ng)]
            L.call("tmf_bn_act_pool_bwd_apply", ng, L.ptrs(dout), dout_fp32, L.ptrs(y), L.ptrs(coef), L.ptrs(bcoef),
                   L.ptrs(dy), B, Dl, Hl, Wl, cout, pool, slope, tag=f"tmf_bn_act_pool_bwd_apply@L{l}")
            if l == 0:
                nws = int(L.load().tmf_conv1_wgrad_workspace_bytes(ng, ctx.impl, Wl, cout))
                ws = torch.empty(nws, dtype=torch.uint8, device=dev) if nws > 0 else None
                L.call("tmf_conv1_wgrad", ng, L.ptrs(dy), L.ptrs(act), L.ptrs(dw), B, Dl, Hl, Wl, cout, ctx.impl,
                       L.ptr(ws), nws)
            else:
                ws = wgrad_workspace(ng, ctx.impl, B, Dl, Hl, Wl, cin, cout, ks, dev)
                L.call("tmf_conv3d_wgrad", ng, L.ptrs(dy), L.ptrs(act), L.ptrs(dw), B, Dl, Hl, Wl, cin, cout, ks,
                       ctx.impl, L.ptr(ws), 0 if ws is None else ws.numel(), tag=f"tmf_conv3d_wgrad@L{l}")
                da = [torch.empty((B, Dl, Hl, Wl, cin), dtype=torch.bfloat16, device=dev) for _ in range(ng)]
                L.call("tmf_conv3d_fwd", ng, L.ptrs(dy), L.ptrs(wd), L.ptrs(None), L.ptrs(da), L.ptrs(None),
                       B, Dl, Hl, Wl, cout, cin, ks, ctx.impl, tag=f"tmf_conv3d_dgrad@L{l}")
                dout, dout_fp32 = da, 0
            SNetFunction._layer_grads(ctx, pgrads, l, dw, dbias, dgamma, dbeta)
            saved[l] = None
        ctx.saved = None
        ctx.params = None
        return (None, None, None, None) + (None,) * ng + tuple(pgrads)

    @staticmethod
    def _layer_grads(ctx, pgrads, l, dw, dbias, dgamma, dbeta):
        for t in range(ctx.ng):
            base = (t * 7 + l) * 4
            pgrads[base + 0], pgrads[base + 1], pgrads[base + 2], pgrads[base + 3] = dw[t], dbias[t], dgamma[t], dbeta[t]


# ==============================================================================================================
# Linear / LayerNorm / attention / pooling / GRL
# ==============================================================================================================
class LinearFunction(torch.autograd.Function):
    """y = act(x W^T + b) + residual   (nn.Linear; act = exact GELU when gelu=True)."""

    @staticmethod
    def forward(ctx, x, w, b, residual, gelu):
        x2 = _f32c(x).reshape(-1, x.shape[-1])
        M, K = x2.shape
        N = w.shape[0]
        w = _f32c(w)
        y = torch.empty((M, N), dtype=torch.float32, device=x2.device)
        pre = torch.empty_like(y) if gelu else None
        res2 = _f32c(residual).reshape(M, N) if residual is not None else None
        L.call("tmf_linear_fwd", L.ptr(x2), L.ptr(w), L.ptr(None if b is None else _f32c(b)), L.ptr(res2), L.ptr(y),
               L.ptr(pre), M, K, N, int(gelu))
        ctx.save_for_backward(x2, w, pre)
        ctx.bias_ref = b                                     # reference only: its gradient slot (flat DP buffer)
        ctx.has_bias, ctx.has_res, ctx.gelu = b is not None, residual is not None, gelu
        ctx.xshape = x.shape
        return y.reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w, pre = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[0]
        dy2 = _f32c(dy).reshape(M, N)
        dres = dy2.reshape(dy.shape) if ctx.has_res else None
        if ctx.gelu:
            dpre = torch.empty_like(dy2)
            L.call("tmf_gelu_bwd", L.ptr(dy2), L.ptr(pre), L.ptr(dpre), M * N)
            dy2 = dpre
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, K), dtype=torch.float32, device=dy2.device)
            L.call("tmf_linear_dgrad", L.ptr(dy2), L.ptr(w), L.ptr(dx), M, K, N, 0)
            dx = dx.reshape(ctx.xshape)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw = grad_out(w, (N, K))
            db = grad_out(ctx.bias_ref, (N,)) if ctx.has_bias else None
            ws, nws = L.scratch(dy2.device)
            L.call("tmf_linear_wgrad", L.ptr(dy2), L.ptr(x2), L.ptr(dw), L.ptr(db), M, K, N, L.ptr(ws), nws)
        return dx, dw, db, dres, None


def linear(x, w, b=None, residual=None, gelu=False):
    return LinearFunction.apply(x, w, b, residual, gelu)


class LayerNormFunction(torch.autograd.Function):
    """y = LayerNorm(x) * gamma + beta (+ residual)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, residual, eps):
        x2 = _f32c(x).reshape(-1, x.shape[-1])
        rows, dim = x2.shape
        y = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x2.device)
        rstd = torch.empty_like(mean)
        res2 = _f32c(residual).reshape(rows, dim) if residual is not None else None
        g = _f32c(gamma)
        L.call("tmf_layernorm_fwd", L.ptr(x2), L.ptr(g), L.ptr(_f32c(beta)), L.ptr(res2), L.ptr(y), L.ptr(mean),
               L.ptr(rstd), rows, dim, eps)
        ctx.save_for_backward(x2, g, mean, rstd)
        ctx.beta_ref = beta
        ctx.has_res = residual is not None
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, g, mean, rstd = ctx.saved_tensors
        rows, dim = x2.shape
        dy2 = _f32c(dy).reshape(rows, dim)
        dx = torch.empty_like(x2)
        dgamma, dbeta = grad_out(g, (dim,)), grad_out(ctx.beta_ref, (dim,))
        ws, nws = L.scratch(x2.device)
        L.call("tmf_layernorm_bwd", L.ptr(dy2), L.ptr(x2), L.ptr(g), L.ptr(mean), L.ptr(rstd), L.ptr(dx), L.ptr(dgamma),
               L.ptr(dbeta), rows, dim, 0, L.ptr(ws), nws)
        return dx.reshape(dy.shape), dgamma, dbeta, (dy if ctx.has_res else None), None


def layer_norm(x, gamma, beta, residual=None, eps=LN_EPS):
    return LayerNormFunction.apply(x, gamma, beta, residual, eps)


class AttentionCoreFunction(torch.autograd.Function):
    """softmax(q k^T * scale) v per head on short token sequences; q (B,Nq,h*dh), kv (B,Nk,2*h*dh)."""

    @staticmethod
    def forward(ctx, q, kv, heads, scale):
        q, kv = _f32c(q), _f32c(kv)
        B, Nq, inner = q.shape
        Nk = kv.shape[1]
        dh = inner // heads
        out = torch.empty_like(q)
        lse = torch.empty((B, heads, Nq), dtype=torch.float32, device=q.device)
        L.call("tmf_attn_fwd", L.ptr(q), L.ptr(kv), L.ptr(out), L.ptr(lse), B, Nq, Nk, heads, dh, float(scale))
        ctx.save_for_backward(q, kv, out, lse)
        ctx.heads, ctx.scale = heads, float(scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, kv, out, lse = ctx.saved_tensors
        B, Nq, inner = q.shape
        Nk = kv.shape[1]
        dq, dkv = torch.empty_like(q), torch.empty_like(kv)
        L.call("tmf_attn_bwd", L.ptr(_f32c(dout)), L.ptr(q), L.ptr(kv), L.ptr(out), L.ptr(lse), L.ptr(dq), L.ptr(dkv),
               B, Nq, Nk, ctx.heads, inner // ctx.heads, ctx.scale)
        return dq, dkv, None, None


def attention_core(q, kv, heads, scale):
    return AttentionCoreFunction.apply(q, kv, heads, scale)


# Side stream for work that is off the backward pass's critical path (the encoders' weight gradients): launched there, it
# overlaps the next encoder's dgrad chain; the launching stream waits for it when autograd finishes the backward pass
# (engine callback), i.e. before anything can consume the gradients.  Under CUDA-graph capture this becomes a parallel branch.
_SIDE = {}
_SIDE_PENDING = []


def _side_stream(device):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    st = _SIDE.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _SIDE[key] = st
    return st


def join_side_streams():
    """Make the current stream wait for everything launched on the side stream (idempotent)."""
    while _SIDE_PENDING:
        torch.cuda.current_stream().wait_event(_SIDE_PENDING.pop())


class EncoderFunction(torch.autograd.Function):
    """One ``Transformer(depth=1)`` encoder (reference models/networks.py:215-230) with cross-attention context, as three
    forward and five backward launches (csrc/enc_fused.cu + the attention core):
        a = Attn(LN1(x), ctx) + x;   y = LNf(FF(LN2(a)) + a) (+ x when ``add_input``: the caller's outer residual).
    params = (ln1_w, ln1_b, wq, wkv, wo, bo, ln2_w, ln2_b, w1, b1, w2, b2, lnf_w, lnf_b)."""

    @staticmethod
    def forward(ctx, x, context, heads, scale, add_input, eps1, eps2, epsf, pack_cache, *params):
        ln1_w, ln1_b, wq, wkv, wo, bo, ln2_w, ln2_b, w1, b1, w2, b2, lnf_w, lnf_b = [_f32c(t) for t in params]
        x3, c3 = _f32c(x), _f32c(context)
        B, Nq, dim = x3.shape
        Nk = c3.shape[1]
        Mx, Mc, mlp = B * Nq, B * Nk, w1.shape[0]
        dev = x3.device
        E = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        h1, mean1, rstd1, q, kv = E(Mx, dim), E(Mx), E(Mx), E(B, Nq, dim), E(B, Nk, 2 * dim)
        # bf16 hi / lo copies of the five weight matrices for the split-weight GEMMs (one launch; the backward pass reuses
        # them).  TMF_ENC_BF16=0: three-pass TF32 straight from the fp32 weights (pack = NULL).
        pack = None
        if os.environ.get("TMF_ENC_BF16", "1") != "0":
            ws5 = (wq, wkv, wo, w1, w2)
            cached = pack_cache is not None and all(w is t for w, t in zip(ws5, (params[2], params[3], params[4], params[8], params[10])))
            if cached:                                      # the module's pack, refreshed by FusedAdam between steps
                if pack_cache.stale(ws5, mlp):
                    L.call("tmf_encoder_pack_weights", L.ptr(wq), L.ptr(wkv), L.ptr(wo), L.ptr(w1), L.ptr(w2), mlp, L.ptr(pack_cache.pack))
                    for w in ws5:
                        pack_cache.mark(w)
                pack = pack_cache.pack
            else:
                pack = torch.empty(int(L.load().tmf_encoder_pack_bytes(int(mlp))) // 2, dtype=torch.bfloat16, device=dev)
                L.call("tmf_encoder_pack_weights", L.ptr(wq), L.ptr(wkv), L.ptr(wo), L.ptr(w1), L.ptr(w2), mlp, L.ptr(pack))
        L.call("tmf_encoder_proj_fwd", L.ptr(x3), L.ptr(c3), L.ptr(ln1_w), L.ptr(ln1_b), L.ptr(wq), L.ptr(wkv), L.ptr(h1),
               L.ptr(mean1), L.ptr(rstd1), L.ptr(q), L.ptr(kv), Mx, Mc, float(eps1), L.ptr(pack))
        o, lse = E(B, Nq, dim), E(B, heads, Nq)
        L.call("tmf_attn_fwd", L.ptr(q), L.ptr(kv), L.ptr(o), L.ptr(lse), B, Nq, Nk, heads, dim // heads, float(scale))
        a, h2, mean2, rstd2, pre, f, g = E(Mx, dim), E(Mx, dim), E(Mx), E(Mx), E(Mx, mlp), E(Mx, mlp), E(Mx, dim)
        meanf, rstdf, y = E(Mx), E(Mx), E(B, Nq, dim)
        L.call("tmf_encoder_chain_fwd", L.ptr(o), L.ptr(x3), L.ptr(wo), L.ptr(bo), L.ptr(ln2_w), L.ptr(ln2_b), L.ptr(w1),
               L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(lnf_w), L.ptr(lnf_b), L.ptr(a), L.ptr(h2), L.ptr(mean2), L.ptr(rstd2),
               L.ptr(pre), L.ptr(f), L.ptr(g), L.ptr(meanf), L.ptr(rstdf), L.ptr(y), Mx, mlp, int(add_input), float(eps2),
               float(epsf), L.ptr(pack))
        ctx.save_for_backward(x3, c3, h1, mean1, rstd1, q, kv, o, lse, a, h2, mean2, rstd2, pre, f, g, meanf, rstdf,
                              ln1_w, wq, wkv, wo, ln2_w, w1, w2, lnf_w)
        ctx.pack = pack
        ctx.param_refs = params                              # references only: gradient slots of the flat DP buffer
        ctx.cfg = (heads, float(scale), bool(add_input), B, Nq, Nk, dim, mlp)
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        (x3, c3, h1, mean1, rstd1, q, kv, o, lse, a, h2, mean2, rstd2, pre, f, g, meanf, rstdf,
         ln1_w, wq, wkv, wo, ln2_w, w1, w2, lnf_w) = ctx.saved_tensors
        heads, scale, add_input, B, Nq, Nk, dim, mlp = ctx.cfg
        Mx, Mc = B * Nq, B * Nk
        dev = x3.device
        E = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        P = ctx.param_refs
        G = lambda i: grad_out(P[i])
        d_ln1_w, d_ln1_b, d_wq, d_wkv, d_wo, d_bo, d_ln2_w, d_ln2_b, d_w1, d_b1, d_w2, d_b2, d_lnf_w, d_lnf_b = [G(i) for i in range(14)]
        ws, nws = L.scratch(dev)
        dy2 = _f32c(dy).reshape(Mx, dim)
        dg, dp, da, dout, dxp = E(Mx, dim), E(Mx, mlp), E(Mx, dim), E(B, Nq, dim), E(Mx, dim)
        L.call("tmf_encoder_chain_bwd", L.ptr(dy2), L.ptr(g), L.ptr(a), L.ptr(pre), L.ptr(wo), L.ptr(w1), L.ptr(w2), L.ptr(ln2_w),
               L.ptr(lnf_w), L.ptr(mean2), L.ptr(rstd2), L.ptr(meanf), L.ptr(rstdf), L.ptr(dg), L.ptr(dp), L.ptr(da),
               L.ptr(dout), L.ptr(dxp), L.ptr(d_lnf_w), L.ptr(d_lnf_b), L.ptr(d_ln2_w), L.ptr(d_ln2_b), Mx, mlp,
               int(add_input), L.ptr(ctx.pack), L.ptr(ws), nws)
        dq, dkv = E(B, Nq, dim), E(B, Nk, 2 * dim)
        L.call("tmf_attn_bwd", L.ptr(dout), L.ptr(q), L.ptr(kv), L.ptr(o), L.ptr(lse), L.ptr(dq), L.ptr(dkv), B, Nq, Nk, heads,
               dim // heads, scale)
        dx, dctx = E(B, Nq, dim), E(B, Nk, dim)
        L.call("tmf_encoder_proj_bwd", L.ptr(dq), L.ptr(dkv), L.ptr(dxp), L.ptr(x3), L.ptr(ln1_w), L.ptr(mean1), L.ptr(rstd1),
               L.ptr(wq), L.ptr(wkv), L.ptr(dx), L.ptr(dctx), L.ptr(d_ln1_w), L.ptr(d_ln1_b), Mx, Mc, L.ptr(ctx.pack),
               L.ptr(ws), nws)
        table = (ctypes_ptr_array([dq, h1, d_wq, dkv, c3, d_wkv, da, o, d_wo, d_bo, dp, h2, d_w1, d_b1, dg]))
        if os.environ.get("TMF_ENC_SIDE", "1") != "0":
            # all five weight gradients feed nothing downstream in this backward pass: off the critical path
            cur = torch.cuda.current_stream(dev)
            side = _side_stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                ws2, nws2 = L.scratch(dev)                   # the side stream's own scratch buffer
                L.call("tmf_encoder_wgrad", table, L.ptr(f), L.ptr(d_w2), L.ptr(d_b2), Mx, Mc, mlp, L.ptr(ws2), nws2)
                ev = torch.cuda.Event()
                ev.record(side)
            for t in (dq, h1, dkv, c3, da, o, dp, h2, dg, f, d_wq, d_wkv, d_wo, d_bo, d_w1, d_b1, d_w2, d_b2):
                t.record_stream(side)
            if not _SIDE_PENDING:
                torch.autograd.Variable._execution_engine.queue_callback(join_side_streams)
            _SIDE_PENDING.append(ev)
        else:
            L.call("tmf_encoder_wgrad", table, L.ptr(f), L.ptr(d_w2), L.ptr(d_b2), Mx, Mc, mlp, L.ptr(ws), nws)
        return (dx.reshape(x3.shape), dctx.reshape(c3.shape), None, None, None, None, None, None, None,
                d_ln1_w, d_ln1_b, d_wq, d_wkv, d_wo, d_bo, d_ln2_w, d_ln2_b, d_w1, d_b1, d_w2, d_b2, d_lnf_w, d_lnf_b)


def ctypes_ptr_array(tensors):
    import ctypes as C
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def encoder_supported(dim, inner, mlp):
    return os.environ.get("TMF_ENC_FUSED", "1") != "0" and bool(L.load().tmf_encoder_supported(int(dim), int(inner), int(mlp)))


def encoder(x, context, heads, scale, add_input, eps1, eps2, epsf, params, pack_cache=None):
    return EncoderFunction.apply(x, context, heads, scale, add_input, eps1, eps2, epsf, pack_cache, *params)


class TokenPoolFunction(torch.autograd.Function):
    """x (B,N,C) -> mean over N (B,C) and/or max over N (B,C) (first-maximum gradient routing)."""

    @staticmethod
    def forward(ctx, x, want_mean, want_max):
        x = _f32c(x)
        B, N, C = x.shape
        mean = torch.empty((B, C), dtype=torch.float32, device=x.device) if want_mean else None
        mx = torch.empty((B, C), dtype=torch.float32, device=x.device) if want_max else None
        amax = torch.empty((B, C), dtype=torch.int32, device=x.device) if want_max else None
        L.call("tmf_token_pool_fwd", L.ptr(x), L.ptr(mean), L.ptr(mx), L.ptr(amax), B, N, C)
        ctx.shape, ctx.amax = (B, N, C), amax
        ctx.want = (want_mean, want_max)
        if want_mean and want_max:
            return mean, mx
        return mean if want_mean else mx

    @staticmethod
    def backward(ctx, *g):
        B, N, C = ctx.shape
        want_mean, want_max = ctx.want
        if want_mean and want_max:
            dmean, dmax = g
        elif want_mean:
            dmean, dmax = g[0], None
        else:
            dmean, dmax = None, g[0]
        dmean = _f32c(dmean) if dmean is not None else None
        dmax = _f32c(dmax) if dmax is not None else None
        dx = torch.empty((B, N, C), dtype=torch.float32, device=(dmean if dmean is not None else dmax).device)
        L.call("tmf_token_pool_bwd", L.ptr(dmean), L.ptr(dmax), L.ptr(ctx.amax), L.ptr(dx), B, N, C, 0)
        return dx, None, None


def token_pool(x, want_mean=True, want_max=True):
    return TokenPoolFunction.apply(x, want_mean, want_max)


class GradientReversalFunction(torch.autograd.Function):
    """reference models/gradient_reversal/functional.py:4-19: identity forward, -alpha * grad backward."""

    @staticmethod
    def forward(ctx, x, alpha):
        if torch.is_tensor(alpha):
            if alpha.is_cuda:
                ctx.alpha_dev, ctx.alpha = alpha.detach().reshape(-1).float(), 1.0
            else:
                ctx.alpha_dev, ctx.alpha = None, float(alpha.reshape(-1)[0])
        else:
            ctx.alpha_dev, ctx.alpha = None, float(alpha)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        if not ctx.needs_input_grad[0]:
            return None, None
        g = _f32c(g)
        out = torch.empty_like(g)
        L.call("tmf_scale", L.ptr(g), L.ptr(out), -ctx.alpha, L.ptr(ctx.alpha_dev), g.numel())
        return out, None


def revgrad(x, alpha):
    return GradientReversalFunction.apply(x, alpha)
