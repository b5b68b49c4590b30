"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import json
import os

import torch

from oracle import restatement as R
from transmf_ad_b200.synthetic import make_labels, make_volumes, procedural_state

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def manifest():
    with open(os.path.join(GOLDEN_DIR, "keys.json")) as f:
        return json.load(f)


def manifest_key(kind, kwargs):
    return kind + json.dumps(kwargs, sort_keys=True)


def template_from_manifest(kind, kwargs):
    """Ordered state-dict template (zeros of the right shape / dtype) from the golden key manifest."""
    man = manifest()[manifest_key(kind, kwargs)]
    dt = {"torch.float32": torch.float32, "torch.int64": torch.int64}
    return {k: torch.zeros(shape, dtype=dt[d]) for k, shape, d in man}


def case_inputs(gold):
    B, shape = gold["batch"], tuple(gold["shape"])
    label = make_labels(B)
    s_mri, s_pet = gold["input_seeds"]
    mri = make_volumes(B, shape, seed=s_mri, labels=label)
    pet = make_volumes(B, shape, seed=s_pet, labels=label)
    return mri, pet, label


def case_state(gold):
    return procedural_state(template_from_manifest(gold["kind"], gold["kwargs"]), seed=gold["weight_seed"])


def oracle_forward(kind, sd, inputs, kwargs, training, p_drop=0.0, rnd=None):
    heads = kwargs.get("heads", 4)
    if kind == "model_ad":
        return R.model_ad_forward(sd, *inputs, heads=heads, training=training, p_drop=p_drop, rnd=rnd)
    if kind == "model_CNN_ad":
        return R.model_cnn_ad_forward(sd, *inputs, training=training, rnd=rnd)
    if kind == "model_single":
        return (R.model_single_forward(sd, inputs[0], training=training, rnd=rnd),)
    if kind == "model_CNN":
        return (R.model_cnn_forward(sd, *inputs, training=training, rnd=rnd),)
    if kind == "model_transformer":
        return (R.model_transformer_forward(sd, *inputs, heads=heads, training=training, p_drop=p_drop, rnd=rnd),)
    if kind == "model_transformer_res":
        return (R.model_transformer_res_forward(sd, *inputs, heads=heads, training=training, p_drop=p_drop, rnd=rnd),)
    raise KeyError(kind)


def losses(outs, label):
    if len(outs) == 3:
        return R.adversarial_losses(outs[0], outs[1], outs[2], label)
    ce = torch.nn.functional.cross_entropy(outs[0], label)
    return ce, torch.zeros((), device=ce.device), ce


def sample(t, n=48):
    f = t.detach().flatten()
    stride = max(1, f.numel() // n)
    return f[::stride][:n].clone()


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosine(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def set_head_dropout(module, p):
    """Only the Dropout(0.5) layers of the classifier heads (reference mymodel.py:190-191)."""
    for m in module.modules():
        if isinstance(m, torch.nn.Dropout):
            if not hasattr(m, "_p0"):
                m._p0 = m.p
            if m._p0 == 0.5:
                m.p = p
