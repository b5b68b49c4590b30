"""Building blocks of TransMF_AD on the B200 kernels -- drop-in for the reference ``models/networks.py``.

Every class keeps the reference's name, constructor signature, sub-module names (hence ``state_dict`` keys, shapes and
fp32 dtype) and forward signature; the torch ``nn.Conv3d`` / ``nn.BatchNorm3d`` / ``nn.Linear`` / ``nn.LayerNorm``
children are *parameter containers only* -- their forward is never called.  All arithmetic goes through
``transmf_ad_b200.functional`` (C ABI -> sm_100a kernels).  Reference line numbers are cited per class.
"""
from __future__ import annotations

import torch
from torch import nn

from transmf_ad_b200 import functional as TF


def exists(val):
    return val is not None


def default(val, d):
    return val if exists(val) else d


# ---------------------------------------------------------------------------------------------------------------
# 3D-CNN encoder   (reference models/networks.py:18-61)
# ---------------------------------------------------------------------------------------------------------------
def _conv_unit(cin, cout, k):
    return [nn.Conv3d(cin, cout, kernel_size=(k, k, k), padding=k // 2), nn.BatchNorm3d(cout), nn.LeakyReLU()]


class sNet(nn.Module):
    """7x(Conv3d + BatchNorm3d + LeakyReLU) with three MaxPool3d(2,2) and a final AvgPool3d(2,2)."""

    def __init__(self, dim) -> None:
        super().__init__()
        q, h = dim // 4, dim // 2
        self.conv1 = nn.Sequential(*_conv_unit(1, q, 3), nn.MaxPool3d(2, stride=2))
        self.conv2 = nn.Sequential(*_conv_unit(q, q, 3), *_conv_unit(q, h, 3), nn.MaxPool3d(2, stride=2))
        self.conv3 = nn.Sequential(*_conv_unit(h, h, 3), *_conv_unit(h, dim, 3), nn.MaxPool3d(2, stride=2))
        self.conv4 = nn.Sequential(*_conv_unit(dim, dim * 2, 3), *_conv_unit(dim * 2, dim, 1), nn.AvgPool3d(2, stride=2))
        self._spec = TF.SNetSpec(dim)
        self._packs = [TF.ConvPack() for _ in range(7)]      # cached bf16 operand packs of the conv weights
        self._folds = [TF.EvalFold() for _ in range(7)]      # eval mode: BatchNorm folded into the conv operands

    def _units(self):
        """[(conv, bn)] in execution order."""
        return [(self.conv1[0], self.conv1[1]), (self.conv2[0], self.conv2[1]), (self.conv2[3], self.conv2[4]),
                (self.conv3[0], self.conv3[1]), (self.conv3[3], self.conv3[4]), (self.conv4[0], self.conv4[1]),
                (self.conv4[3], self.conv4[4])]

    def _hyper(self):
        """Per layer (eps, momentum, negative_slope) read from the BatchNorm3d / LeakyReLU children."""
        seqs = [(self.conv1, 0), (self.conv2, 0), (self.conv2, 3), (self.conv3, 0), (self.conv3, 3), (self.conv4, 0),
                (self.conv4, 3)]
        out = []
        for seq, i in seqs:
            bn, act = seq[i + 1], seq[i + 2]
            if bn.momentum is None or not bn.track_running_stats or not bn.affine:
                raise NotImplementedError("sNet on the B200 kernels needs affine BatchNorm3d with running statistics and "
                                          "a numeric momentum (the reference's nn.BatchNorm3d defaults)")
            out.append((float(bn.eps), float(bn.momentum), float(act.negative_slope)))
        return out

    def _run(self, other=None):
        packs = [self._packs] if other is None else [self._packs, other._packs]
        folds = [self._folds] if other is None else [self._folds, other._folds]
        # torch.nn.SyncBatchNorm.convert_sync_batchnorm(model) turns the BatchNorm3d children into nn.SyncBatchNorm: the kernels
        # then normalise with statistics over the global batch (optional row)
        sync = [isinstance(bn, nn.SyncBatchNorm) for _, bn in self._units()]
        if any(sync) and not all(sync):
            raise NotImplementedError("sNet: either every BatchNorm3d is an nn.SyncBatchNorm or none")
        group = self._units()[0][1].process_group if all(sync) else False
        return TF.SNetRun(self.training, torch.is_grad_enabled(), self._hyper(), packs, folds, sync_group=group)

    def invalidate_packs(self):
        for pk in self._packs:
            pk.version = None

    def pack_now(self):
        """(Re-)pack every conv weight now, e.g. after weights were restored outside an optimizer step, so that a CUDA-graph
        capture that follows contains no pack launches (``optim.FusedAdam`` keeps the packs fresh from then on)."""
        from transmf_ad_b200 import _lib as L
        for l, (conv, _) in enumerate(self._units()):
            if l == 0 or not conv.weight.is_cuda:
                continue
            pk, w = self._packs[l], conv.weight
            pk.stale(w)
            L.call("tmf_pack_conv_weights", 1, L.ptrs([w.detach()]), L.ptrs([pk.wf]), L.ptrs([pk.wd]), w.shape[0], w.shape[1],
                   w.shape[2])
            pk.mark(w)

    def _params_and_buffers(self):
        params, bufs = [], []
        for conv, bn in self._units():
            params += [conv.weight, conv.bias, bn.weight, bn.bias]
            bufs.append((bn.running_mean, bn.running_var, bn.num_batches_tracked))
        return params, bufs

    def forward(self, mri):
        params, bufs = self._params_and_buffers()
        return TF.SNetFunction.apply(self._spec, self._run(), [bufs], 1, mri, *params)


def snet_pair_forward(net_a: sNet, net_b: sNet, xa, xb):
    """Run two towers of identical shape as one grouped launch sequence (grid.z = tower)."""
    if (net_a._spec.dim != net_b._spec.dim or net_a.training != net_b.training or tuple(xa.shape) != tuple(xb.shape)
            or net_a._hyper() != net_b._hyper()):
        return net_a(xa), net_b(xb)
    pa, ba = net_a._params_and_buffers()
    pb, bb = net_b._params_and_buffers()
    return TF.SNetFunction.apply(net_a._spec, net_a._run(net_b), [ba, bb], 2, xa, xb, *pa, *pb)


def tokens_of(feat):
    """'b d x y z -> b (x y z) d' (reference models/mymodel.py:218-219); a free view of the channels-last output."""
    B, C = feat.shape[:2]
    return feat.permute(0, 2, 3, 4, 1).reshape(B, -1, C)


# ---------------------------------------------------------------------------------------------------------------
# transformer blocks   (reference models/networks.py:114-175, 215-230)
# ---------------------------------------------------------------------------------------------------------------
class Linear(nn.Linear):
    """nn.Linear whose forward is the tmf GEMM kernel (same parameters / state_dict keys)."""

    def forward(self, x):
        return TF.linear(x, self.weight, self.bias)


def _dropout_active(mod):
    return mod.training and mod.p > 0


class PreNorm(nn.Module):
    """LayerNorm on the first argument only; ``context`` passes through raw (reference :114-121)."""

    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn

    def forward(self, x, **kwargs):
        return self.fn(TF.layer_norm(x, self.norm.weight, self.norm.bias, eps=self.norm.eps), **kwargs)


class FeedForward(nn.Module):
    """Linear - exact GELU - Dropout - Linear - Dropout (reference :125-137)."""

    def __init__(self, dim, hidden_dim, dropout=0.):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))

    def forward(self, x, residual=None):
        h = TF.linear(x, self.net[0].weight, self.net[0].bias, gelu=True)
        if _dropout_active(self.net[2]):
            h = self.net[2](h)
        if _dropout_active(self.net[4]):
            y = self.net[4](TF.linear(h, self.net[3].weight, self.net[3].bias))
            return y if residual is None else y + residual
        return TF.linear(h, self.net[3].weight, self.net[3].bias, residual=residual)


class Attention(nn.Module):
    """Multi-head (cross-)attention, bias-free q / kv projections (reference :141-175)."""

    def __init__(self, dim, heads=4, dim_head=64, dropout=0.):
        super().__init__()
        inner_dim = dim_head * heads
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.attend = nn.Softmax(dim=-1)
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(dim, inner_dim * 2, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))

    def forward(self, x, context=None, kv_include_self=False, residual=None):
        context = default(context, x)
        if kv_include_self:
            context = torch.cat((x, context), dim=1)
        q = TF.linear(x, self.to_q.weight)
        kv = TF.linear(context, self.to_kv.weight)
        out = TF.attention_core(q, kv, self.heads, self.scale)
        proj, drop = self.to_out[0], self.to_out[1]
        if _dropout_active(drop):
            y = drop(TF.linear(out, proj.weight, proj.bias))
            return y if residual is None else y + residual
        return TF.linear(out, proj.weight, proj.bias, residual=residual)


class Transformer(nn.Module):
    """depth x (pre-LN attention + pre-LN FFN, each with a residual) and a trailing LayerNorm (reference :215-230)."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.):
        super().__init__()
        self.layers = nn.ModuleList([])
        self.norm = nn.LayerNorm(dim)
        self._enc_pack = TF.EncPack()          # cached bf16 hi / lo operand pack of the fused encoder (depth 1)
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                PreNorm(dim, Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout)),
                PreNorm(dim, FeedForward(dim, mlp_dim, dropout=dropout))]))

    def _enc_weights(self):
        attn, ff = self.layers[0]
        return (attn.fn.to_q.weight, attn.fn.to_kv.weight, attn.fn.to_out[0].weight, ff.fn.net[0].weight, ff.fn.net[3].weight)

    def invalidate_packs(self):
        self._enc_pack.versions = {}

    def pack_now(self):
        """(Re-)pack the five weight matrices now (see ``sNet.pack_now``): a CUDA-graph capture that follows then contains no
        ``tmf_encoder_pack_weights`` launch; ``optim.FusedAdam`` keeps the pack fresh from its own kernel."""
        from transmf_ad_b200 import _lib as L
        ws = self._enc_weights()
        if len(self.layers) != 1 or not ws[0].is_cuda or not TF.encoder_supported(ws[0].shape[1], ws[0].shape[0], ws[3].shape[0]):
            return
        pk = self._enc_pack
        pk.stale(ws, ws[3].shape[0])
        L.call("tmf_encoder_pack_weights", *[L.ptr(w.detach()) for w in ws], int(ws[3].shape[0]), L.ptr(pk.pack))
        for w in ws:
            pk.mark(w)

    def _fused_params(self):
        attn, ff = self.layers[0]
        A, F = attn.fn, ff.fn
        return (attn.norm.weight, attn.norm.bias, A.to_q.weight, A.to_kv.weight, A.to_out[0].weight, A.to_out[0].bias,
                ff.norm.weight, ff.norm.bias, F.net[0].weight, F.net[0].bias, F.net[3].weight, F.net[3].bias,
                self.norm.weight, self.norm.bias)

    def _can_fuse(self, x, context):
        if len(self.layers) != 1 or context is None or x.dim() != 3 or context.dim() != 3 or not x.is_cuda:
            return False
        attn, ff = self.layers[0]
        A, F = attn.fn, ff.fn
        if _dropout_active(A.to_out[1]) or _dropout_active(F.net[2]) or _dropout_active(F.net[4]):
            return False
        dim = x.shape[-1]
        return (A.to_out[0].bias is not None and F.net[0].bias is not None and F.net[3].bias is not None and
                TF.encoder_supported(dim, A.to_q.weight.shape[0], F.net[0].weight.shape[0]) and context.shape[-1] == dim)

    def forward(self, x, context=None, add_input=False):
        """``add_input`` fuses the caller's outer residual ``enc(x) + x`` into the final LayerNorm kernel."""
        if self._can_fuse(x, context):
            attn, ff = self.layers[0]          # the whole encoder in 3 launches (csrc/enc_fused.cu)
            return TF.encoder(x, context, attn.fn.heads, attn.fn.scale, add_input, attn.norm.eps, ff.norm.eps,
                              self.norm.eps, self._fused_params(), pack_cache=self._enc_pack)
        x_in = x
        for attn, ff in self.layers:
            x = attn(x, context=context, residual=x)
            x = ff(x, residual=x)
        return TF.layer_norm(x, self.norm.weight, self.norm.bias, residual=x_in if add_input else None,
                             eps=self.norm.eps)


class _TokenGAP(nn.Module):
    """'b n d -> b d' mean over tokens (reference :264-266)."""

    def forward(self, x):
        return TF.token_pool(x, True, False)


class _TokenGMP(nn.Module):
    """'b n d -> b d' max over tokens (reference :267-269)."""

    def forward(self, x):
        return TF.token_pool(x, False, True)


def _encoder_pairs(dim, depth, heads, dim_head, mlp_dim, dropout):
    return nn.ModuleList([nn.ModuleList([Transformer(dim, 1, heads, dim_head, mlp_dim, dropout=dropout),
                                         Transformer(dim, 1, heads, dim_head, mlp_dim, dropout=dropout)])
                          for _ in range(depth)])


class CrossTransformer(nn.Module):
    """Both encoders attend to cat([mri, pet]) (reference :233-252; ``share=True`` is broken upstream and unused)."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout, share=False):
        super().__init__()
        self.share = share
        if share:
            self.layers = nn.ModuleList([Transformer(dim, 1, heads, dim_head, mlp_dim, dropout=dropout)
                                         for _ in range(depth)])
        else:
            self.layers = _encoder_pairs(dim, depth, heads, dim_head, mlp_dim, dropout)

    def forward(self, mri_tokens, pet_tokens):
        for mri_enc, pet_enc in self.layers:
            mri_tokens = mri_enc(mri_tokens, context=torch.cat([mri_tokens, pet_tokens], dim=1), add_input=True)
            pet_tokens = pet_enc(pet_tokens, context=torch.cat([mri_tokens, pet_tokens], dim=1), add_input=True)
        return mri_tokens, pet_tokens


class CrossTransformer_MOD_AVG(nn.Module):
    """MRI attends to PET, then PET to the UPDATED MRI, ``depth`` times; cat[GAP mri, GAP pet, GMP mri, GMP pet]
    (reference :255-281)."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout):
        super().__init__()
        self.layers = _encoder_pairs(dim, depth, heads, dim_head, mlp_dim, dropout)
        self.gap = _TokenGAP()
        self.gmp = _TokenGMP()

    def forward(self, mri_tokens, pet_tokens):
        for mri_enc, pet_enc in self.layers:
            mri_tokens = mri_enc(mri_tokens, context=pet_tokens, add_input=True)
            pet_tokens = pet_enc(pet_tokens, context=mri_tokens, add_input=True)
        mri_avg, mri_max = TF.token_pool(mri_tokens, True, True)
        pet_avg, pet_max = TF.token_pool(pet_tokens, True, True)
        return torch.cat([mri_avg, pet_avg, mri_max, pet_max], dim=1)
