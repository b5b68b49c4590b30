"""CPU restatement of the TransMF_AD training hot path  --  TEST INFRASTRUCTURE ONLY.

This file is the *oracle*: a plain, functional (state-dict driven) fp32 PyTorch-CPU restatement of the
reference algorithm.  It is never imported by the product package ``transmf_ad_b200``; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it.

Where the arithmetic lives: the reference (Kateridge/TransMF_AD) contains no kernels of its own; every
op is a ``torch.nn`` call (pinned ``torch==2.4.0``, ``einops==0.8.0`` in the reference's requirements.txt:4,6).
The functions below restate the reference's *composition* of those ops, citing the reference file:line each
follows.  Parity pin: ``oracle/make_golden.py`` imports the real reference modules from ``/root/reference``
in the authoring container and (a) asserts this restatement equals them bit-for-bit on CPU, (b) writes the
golden fixtures under ``tests/golden/`` which travel to the GPU box (``/root/reference`` does not).

All functions take ``sd`` -- a dict with exactly the reference ``state_dict()`` keys -- and a key prefix.
Parameters that need gradients must be leaf tensors with ``requires_grad=True`` inside ``sd``.

``rnd`` (optional) is a callable applied at the points where the CUDA path rounds to bf16 (conv operands and
stored pre-BN conv outputs); ``None`` = pure fp32 reference semantics ("Oracle-B"); ``bf16_round`` =
"Oracle-A" (SURVEY.md section 8c).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # nn.BatchNorm3d / BatchNorm1d default, reference models/networks.py:23
BN_MOMENTUM = 0.1
LRELU_SLOPE = 0.01   # nn.LeakyReLU() default, reference models/networks.py:24
LN_EPS = 1e-5        # nn.LayerNorm default, reference models/networks.py:117


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to bf16 and back, with a straight-through gradient."""
    return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()


def _id(t):
    return t


# --------------------------------------------------------------------------------------------------------------
# sNet  (reference models/networks.py:18-61)
# --------------------------------------------------------------------------------------------------------------
# (sequential name, conv index, bn index, pool after activation)
SNET_LAYERS = (
    ("conv1", 0, 1, "max"),   # networks.py:21-26  Conv3d(1, dim/4, 3, p=1) BN LeakyReLU MaxPool(2,2)
    ("conv2", 0, 1, None),    # networks.py:28-30
    ("conv2", 3, 4, "max"),   # networks.py:31-34
    ("conv3", 0, 1, None),    # networks.py:37-39
    ("conv3", 3, 4, "max"),   # networks.py:40-43
    ("conv4", 0, 1, None),    # networks.py:46-48
    ("conv4", 3, 4, "avg"),   # networks.py:49-52  Conv3d(2dim, dim, 1) BN LeakyReLU AvgPool(2,2)
)


def batchnorm(sd, pfx, x, training):
    """train-mode batch statistics (biased var to normalise, unbiased into running_var, momentum 0.1),
    eval-mode running statistics; increments num_batches_tracked like nn.BatchNorm*d."""
    rm, rv = sd[pfx + ".running_mean"], sd[pfx + ".running_var"]
    out = F.batch_norm(x, rm, rv, sd[pfx + ".weight"], sd[pfx + ".bias"], training, BN_MOMENTUM, BN_EPS)
    if training:
        sd[pfx + ".num_batches_tracked"] += 1
    return out


def snet_forward(sd, pfx, x, training=True, rnd=None, taps=None):
    """reference models/networks.py:55-61.  x: (B,1,D,H,W) fp32 -> (B,dim,d,h,w)."""
    rnd = rnd or _id
    a = x
    for li, (blk, ci, bi, pool) in enumerate(SNET_LAYERS):
        w = sd[f"{pfx}.{blk}.{ci}.weight"]
        b = sd[f"{pfx}.{blk}.{ci}.bias"]
        pad = 1 if w.shape[-1] == 3 else 0
        if li == 0:
            y = F.conv3d(a, w, b, padding=pad)          # conv1 keeps fp32 operands on the CUDA path too
        else:
            y = F.conv3d(rnd(a), rnd(w), b, padding=pad)
        y = rnd(y)                                       # CUDA path stores the pre-BN conv output in bf16
        z = batchnorm(sd, f"{pfx}.{blk}.{bi}", y, training)
        a = F.leaky_relu(z, LRELU_SLOPE)
        if pool == "max":
            a = F.max_pool3d(a, 2, 2)
        elif pool == "avg":
            a = F.avg_pool3d(a, 2, 2)
        if taps is not None:
            taps[f"{pfx}.{blk}.{ci}"] = (y, a)
    return a


# --------------------------------------------------------------------------------------------------------------
# Transformer blocks (reference models/networks.py:114-175, 215-230)
# --------------------------------------------------------------------------------------------------------------
def layernorm(sd, pfx, x):
    return F.layer_norm(x, (x.shape[-1],), sd[pfx + ".weight"], sd[pfx + ".bias"], LN_EPS)


def attention(sd, pfx, x, context, heads):
    """reference models/networks.py:157-175.  x is already layer-normed by PreNorm; context is RAW
    (PreNorm normalises only its first argument, networks.py:120-121).  to_q / to_kv have no bias."""
    B, n, _ = x.shape
    q = F.linear(x, sd[pfx + ".to_q.weight"])
    kv = F.linear(context, sd[pfx + ".to_kv.weight"])
    k, v = kv.chunk(2, dim=-1)
    inner = q.shape[-1]
    dh = inner // heads

    def split(t):
        return t.reshape(B, t.shape[1], heads, dh).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    dots = torch.matmul(q, k.transpose(-1, -2)) * (dh ** -0.5)
    attn = torch.softmax(dots, dim=-1)
    out = torch.matmul(attn, v)
    out = out.permute(0, 2, 1, 3).reshape(B, n, inner)
    return F.linear(out, sd[pfx + ".to_out.0.weight"], sd[pfx + ".to_out.0.bias"])   # Dropout(p=0) = identity


def feedforward(sd, pfx, x):
    """reference models/networks.py:125-137: Linear - exact-erf GELU - Linear (dropout p = 0)."""
    h = F.linear(x, sd[pfx + ".net.0.weight"], sd[pfx + ".net.0.bias"])
    h = F.gelu(h)
    return F.linear(h, sd[pfx + ".net.3.weight"], sd[pfx + ".net.3.bias"])


def transformer_encoder(sd, pfx, x, context, heads):
    """reference models/networks.py:215-230 with depth = len(layers) (always 1 in the fusion stacks)."""
    depth = 0
    while f"{pfx}.layers.{depth}.0.norm.weight" in sd:
        depth += 1
    for l in range(depth):
        p = f"{pfx}.layers.{l}"
        xn = layernorm(sd, p + ".0.norm", x)
        ctx = xn if context is None else context      # default(context, x) is applied to the NORMED x (:160)
        x = attention(sd, p + ".0.fn", xn, ctx, heads) + x
        x = feedforward(sd, p + ".1.fn", layernorm(sd, p + ".1.norm", x)) + x
    return layernorm(sd, pfx + ".norm", x)


def cross_transformer_mod_avg(sd, pfx, mri, pet, heads):
    """reference models/networks.py:272-281: MRI attends to PET, then PET attends to the UPDATED MRI;
    outer residual added after each encoder's trailing LayerNorm; cat[GAP mri, GAP pet, GMP mri, GMP pet]."""
    depth = 0
    while f"{pfx}.layers.{depth}.0.norm.weight" in sd:
        depth += 1
    for l in range(depth):
        mri = transformer_encoder(sd, f"{pfx}.layers.{l}.0", mri, pet, heads) + mri
        pet = transformer_encoder(sd, f"{pfx}.layers.{l}.1", pet, mri, heads) + pet
    return torch.cat([mri.mean(1), pet.mean(1), mri.amax(1), pet.amax(1)], dim=1)


def cross_transformer(sd, pfx, mri, pet, heads):
    """reference models/networks.py:248-252 (share=False): context = cat([mri, pet]) for both encoders."""
    depth = 0
    while f"{pfx}.layers.{depth}.0.norm.weight" in sd:
        depth += 1
    for l in range(depth):
        mri = transformer_encoder(sd, f"{pfx}.layers.{l}.0", mri, torch.cat([mri, pet], 1), heads) + mri
        pet = transformer_encoder(sd, f"{pfx}.layers.{l}.1", pet, torch.cat([mri, pet], 1), heads) + pet
    return mri, pet


# --------------------------------------------------------------------------------------------------------------
# Gradient reversal (reference models/gradient_reversal/functional.py:4-19)
# --------------------------------------------------------------------------------------------------------------
class _RevGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = float(alpha)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return -ctx.alpha * g, None


def revgrad(x, alpha):
    return _RevGrad.apply(x, alpha)


# --------------------------------------------------------------------------------------------------------------
# Heads and model assemblies (reference models/mymodel.py)
# --------------------------------------------------------------------------------------------------------------
def _tokens(feat):
    """rearrange 'b d x y z -> b (x y z) d'  (mymodel.py:218-219)."""
    B, C = feat.shape[:2]
    return feat.reshape(B, C, -1).transpose(1, 2)


def _gap(feat):
    """AdaptiveAvgPool3d(1) + flatten (mymodel.py:193)."""
    return feat.mean(dim=(2, 3, 4))


def discriminator(sd, pfx, v, training):
    """D = Linear(dim,128) BN1d ReLU Linear(128,2)  (mymodel.py:194)."""
    h = F.linear(v, sd[pfx + ".0.weight"], sd[pfx + ".0.bias"])
    h = F.relu(batchnorm(sd, pfx + ".1", h, training))
    return F.linear(h, sd[pfx + ".3.weight"], sd[pfx + ".3.bias"])


def fc_cls_transformer(sd, pfx, c, training, p_drop=0.5):
    """Linear(4dim,512) BN1d ReLU Dropout(.5) Linear(512,64) BN1d ReLU Dropout(.5) Linear(64,2)  (mymodel.py:190-192)."""
    h = F.linear(c, sd[pfx + ".0.weight"], sd[pfx + ".0.bias"])
    h = F.dropout(F.relu(batchnorm(sd, pfx + ".1", h, training)), p_drop, training)
    h = F.linear(h, sd[pfx + ".4.weight"], sd[pfx + ".4.bias"])
    h = F.dropout(F.relu(batchnorm(sd, pfx + ".5", h, training)), p_drop, training)
    return F.linear(h, sd[pfx + ".8.weight"], sd[pfx + ".8.bias"])


def model_ad_forward(sd, mri, pet, heads=4, training=True, rnd=None, p_drop=0.5, alpha=2.0, taps=None):
    """reference models/mymodel.py:204-222."""
    fm = snet_forward(sd, "mri_cnn", mri, training, rnd, taps)
    fp = snet_forward(sd, "pet_cnn", pet, training, rnd, taps)
    d_mri = discriminator(sd, "D", revgrad(_gap(fm), alpha), training)     # MRI first, then PET (BN1d stats order)
    d_pet = discriminator(sd, "D", revgrad(_gap(fp), alpha), training)
    cls = cross_transformer_mod_avg(sd, "fuse_transformer", _tokens(fm), _tokens(fp), heads)
    if taps is not None:
        taps["cls"] = cls
    logits = fc_cls_transformer(sd, "fc_cls", cls, training, p_drop)
    return logits, d_mri, d_pet


def model_cnn_ad_forward(sd, mri, pet, training=True, rnd=None, alpha=2.0):
    """reference models/mymodel.py:162-179."""
    fm = snet_forward(sd, "mri_cnn", mri, training, rnd)
    fp = snet_forward(sd, "pet_cnn", pet, training, rnd)
    d_mri = discriminator(sd, "D", revgrad(_gap(fm), alpha), training)
    d_pet = discriminator(sd, "D", revgrad(_gap(fp), alpha), training)
    h = torch.cat([_gap(fm), _gap(fp)], dim=1)
    h = F.relu(F.linear(h, sd["fc_cls.0.weight"], sd["fc_cls.0.bias"]))
    logits = F.linear(h, sd["fc_cls.2.weight"], sd["fc_cls.2.bias"])
    return logits, d_mri, d_pet


def model_single_forward(sd, img, training=True, rnd=None):
    """reference models/mymodel.py:30-37."""
    f = _gap(snet_forward(sd, "cnn", img, training, rnd))
    h = F.relu(F.linear(f, sd["fc.0.weight"], sd["fc.0.bias"]))
    return F.linear(h, sd["fc.2.weight"], sd["fc.2.bias"])


def model_cnn_forward(sd, mri, pet, training=True, rnd=None):
    """reference models/mymodel.py:59-66."""
    h = torch.cat([_gap(snet_forward(sd, "mri_cnn", mri, training, rnd)),
                   _gap(snet_forward(sd, "pet_cnn", pet, training, rnd))], dim=1)
    h = F.relu(F.linear(h, sd["fc.0.weight"], sd["fc.0.bias"]))
    return F.linear(h, sd["fc.2.weight"], sd["fc.2.bias"])


def model_transformer_forward(sd, mri, pet, heads=4, training=True, rnd=None, p_drop=0.5):
    """reference models/mymodel.py:87-98."""
    fm = snet_forward(sd, "mri_cnn", mri, training, rnd)
    fp = snet_forward(sd, "pet_cnn", pet, training, rnd)
    cls = cross_transformer_mod_avg(sd, "fuse_transformer", _tokens(fm), _tokens(fp), heads)
    return fc_cls_transformer(sd, "fc_cls", cls, training, p_drop)


def model_transformer_res_forward(sd, mri, pet, heads=4, training=True, rnd=None, p_drop=0.5):
    """reference models/mymodel.py:125-141."""
    tm = _tokens(snet_forward(sd, "mri_cnn", mri, training, rnd))
    tp = _tokens(snet_forward(sd, "pet_cnn", pet, training, rnd))
    fm, fp = cross_transformer(sd, "fuse_transformer", tm, tp, heads)
    c = torch.cat([(fm + tm).mean(1), (fp + tp).mean(1)], dim=1)
    h = F.dropout(F.relu(F.linear(c, sd["fc_cls.0.weight"], sd["fc_cls.0.bias"])), p_drop, training)
    h = F.dropout(F.relu(F.linear(h, sd["fc_cls.3.weight"], sd["fc_cls.3.bias"])), p_drop, training)
    return F.linear(h, sd["fc_cls.6.weight"], sd["fc_cls.6.bias"])


# --------------------------------------------------------------------------------------------------------------
# MiSePyNet / Mnet baseline (reference models/MiSePyNet.py:5-163), pure fp32
# --------------------------------------------------------------------------------------------------------------
def _conv_bn_relu(sd, pfx, ci, x, training, stride=1):
    y = F.conv3d(x, sd[f"{pfx}.{ci}.weight"], sd[f"{pfx}.{ci}.bias"], stride=stride)
    return F.relu(batchnorm(sd, f"{pfx}.{ci + 1}", y, training))


def slice_cnn_forward(sd, pfx, img, training):
    """reference MiSePyNet.py:32-38: three stacks of (1,1,k) convolutions that collapse the last axis."""
    c1 = _conv_bn_relu(sd, pfx + ".conv1", 0, img, training)
    c2 = _conv_bn_relu(sd, pfx + ".conv2", 3, _conv_bn_relu(sd, pfx + ".conv2", 0, img, training), training)
    c3 = _conv_bn_relu(sd, pfx + ".conv3", 0, img, training)
    c3 = _conv_bn_relu(sd, pfx + ".conv3", 6, _conv_bn_relu(sd, pfx + ".conv3", 3, c3, training), training)
    return c1, c2, c3


def spatial_cnn_forward(sd, pfx, s1, s2, s3, training):
    """reference MiSePyNet.py:89-94: ``self.conv1`` is applied to ALL three inputs (conv2 / conv3 are dead weights)."""
    def conv1(x):
        x = F.max_pool3d(_conv_bn_relu(sd, pfx + ".conv1", 0, x, training, stride=2), (3, 3, 1))
        x = F.max_pool3d(_conv_bn_relu(sd, pfx + ".conv1", 4, x, training), (3, 3, 1))
        return _conv_bn_relu(sd, pfx + ".conv1", 8, x, training)
    return conv1(s1) + conv1(s2) + conv1(s3)


def misepynet_forward(sd, pfx, img, training):
    """reference MiSePyNet.py:117-136."""
    B = img.shape[0]
    views = (("axial", img), ("col", img.permute(0, 1, 2, 4, 3)), ("sag", img.permute(0, 1, 4, 3, 2)))
    feats = []
    for name, v in views:
        s = slice_cnn_forward(sd, f"{pfx}.slice_cnn_{name}", v, training)
        feats.append(spatial_cnn_forward(sd, f"{pfx}.spatial_cnn_{name}", *s, training).reshape(B, -1))
    return torch.cat(feats, dim=1)


def mnet_forward(sd, mri, pet, training=True, p_drop=0.5):
    """reference MiSePyNet.py:155-163; fc = Linear BN1d ReLU Dropout Linear BN1d ReLU Dropout Linear (:144-153)."""
    c = torch.cat([misepynet_forward(sd, "mri", mri, training), misepynet_forward(sd, "pet", pet, training)], dim=-1)
    return fc_cls_transformer(sd, "fc", c, training, p_drop)


# --------------------------------------------------------------------------------------------------------------
# Training step of the caller (reference kfold_train_adversarial.py:101-136, kfold_train_single.py:91-113)
# --------------------------------------------------------------------------------------------------------------
def adversarial_losses(logits, d_mri, d_pet, label):
    """ce = CE(logits,label); ad = (CE(D_MRI, ones) + CE(D_PET, zeros)) / 2; total = ad + ce."""
    ce = F.cross_entropy(logits, label)
    ones = torch.ones(d_mri.shape[0], dtype=torch.int64, device=d_mri.device)
    zeros = torch.zeros(d_pet.shape[0], dtype=torch.int64, device=d_pet.device)
    ad = (F.cross_entropy(d_mri, ones) + F.cross_entropy(d_pet, zeros)) / 2
    return ce, ad, ad + ce


def clone_state(sd, requires_grad=True, dtype=None):
    """Deep-copy a state dict into fresh leaves (float tensors get requires_grad unless they are buffers)."""
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if t.is_floating_point():
            if dtype is not None:
                t = t.to(dtype)
            if requires_grad and not (k.endswith("running_mean") or k.endswith("running_var")):
                t.requires_grad_(True)
        out[k] = t
    return out
