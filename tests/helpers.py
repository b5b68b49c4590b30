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
    if kind == "Mnet":
        return (R.mnet_forward(sd, *inputs, training=training, p_drop=p_drop),)
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


def is_conv_bias(k):
    """Conv3d biases inside the sNet towers: train-mode BatchNorm cancels them exactly (true gradient 0)."""
    parts = k.split(".")
    return k.endswith(".bias") and ".conv" in k and parts[-2] in ("0", "3")


def parity_report(name, device="cuda"):
    """Run one train step + eval forward of golden case ``name`` on the CUDA path and on the CPU oracle (with and
    without bf16 rounding emulation); return a dict of error metrics (no assertions)."""
    import copy
    from transmf_ad_b200.models import mymodel as M
    gold = load_golden(name)
    mri, pet, label = case_inputs(gold)
    kind, kwargs = gold["kind"], gold["kwargs"]
    inputs = (mri,) if kind == "model_single" else (mri, pet)
    state = case_state(gold)
    model = getattr(M, kind)(**kwargs)
    model.load_state_dict(state, strict=True)
    model = model.to(device).train()
    set_head_dropout(model, 0.0)
    rep = {"name": name}
    # ---- tower features (kernel exactness, upstream of the ill-conditioned BatchNorm1d heads)
    towers = [("cnn", 0)] if kind == "model_single" else [("mri_cnn", 0), ("pet_cnn", 1)]
    sd_feat = R.clone_state(state, requires_grad=False)
    feat_err = {}
    with torch.no_grad():
        for pfx, idx in towers:
            net = copy.deepcopy(getattr(model, pfx))
            got = net(inputs[idx].to(device)).float().cpu()
            want = R.snet_forward(sd_feat, pfx, inputs[idx], True, R.bf16_round)
            want32 = R.snet_forward(R.clone_state(state, requires_grad=False), pfx, inputs[idx], True, None)
            feat_err[pfx] = (rel_err(got, want), rel_err(got, want32), rel_err(want, want32))
    rep["feat_rel(ours:A, ours:B, A:B)"] = feat_err
    # ---- full train step
    model.zero_grad(set_to_none=True)
    outs = model(*[t.to(device) for t in inputs])
    outs = outs if isinstance(outs, tuple) else (outs,)
    ce, ad, total = losses(outs, label.to(device))
    total.backward()
    sd = R.clone_state(state)
    o_outs = oracle_forward(kind, sd, inputs, kwargs, True, 0.0, rnd=R.bf16_round)
    o_total = losses(o_outs, label)[2]
    o_total.backward()
    rep["logit_err_A"] = [float((o.detach().cpu() - a.detach()).abs().max()) for o, a in zip(outs, o_outs)]
    rep["logit_err_B"] = [float((o.detach().cpu() - b).abs().max()) for o, b in zip(outs, gold["train_outs"])]
    rep["logit_err_A_vs_B"] = [float((a.detach() - b).abs().max()) for a, b in zip(o_outs, gold["train_outs"])]
    rep["loss"] = (float(total.detach()), float(o_total.detach()), gold["train_losses"][2])
    # train-mode labels (argmax of the classification logits) against the REAL reference's
    t_ref = gold["train_outs"][0]
    t_margin = (t_ref[:, 0] - t_ref[:, 1]).abs()
    rep["train_margin_min"] = float(t_margin.min())
    rep["train_argmax_equal"] = bool(torch.equal(outs[0].detach().cpu().argmax(1), t_ref.argmax(1)))
    rep["train_argmax_equal_sure"] = lambda tol: bool(torch.equal(outs[0].detach().cpu().argmax(1)[t_margin > 2 * tol],
                                                                  t_ref.argmax(1)[t_margin > 2 * tol]))
    grads = {}
    for k, p in model.named_parameters():
        g = None if p.grad is None else p.grad.detach().cpu()
        ga = sd[k].grad
        gs = gold["grad_sample"][k]
        entry = {"missing": g is None}
        if g is not None:
            entry.update(finite=bool(torch.isfinite(g).all()), norm=float(g.norm()), norm_A=float(ga.norm()),
                         norm_B=gold["grad_norm"][k], rel_A=rel_err(g, ga), cos_A=cosine(g, ga),
                         cos_B=cosine(sample(g), gs), cos_A_vs_B=cosine(sample(ga), gs), conv_bias=is_conv_bias(k),
                         absmax=float(g.abs().max()))
        grads[k] = entry
    rep["grads"] = grads
    # whole-model gradient direction (all non-trivial tensors concatenated)
    ours = torch.cat([p.grad.detach().cpu().flatten() for k, p in model.named_parameters()
                      if not is_conv_bias(k) and float(sd[k].grad.norm()) >= 1e-5])
    ref = torch.cat([sd[k].grad.flatten() for k, p in model.named_parameters()
                     if not is_conv_bias(k) and float(sd[k].grad.norm()) >= 1e-5])
    rep["global_grad_cos_A"] = cosine(ours, ref)
    rep["global_grad_rel_A"] = rel_err(ours, ref)
    # the same direction for the fp32 oracle (no rounding emulation): how far bf16 rounding ALONE moves the gradient
    sd32 = R.clone_state(state)
    losses(oracle_forward(kind, sd32, inputs, kwargs, True, 0.0, rnd=None), label)[2].backward()
    ref32 = torch.cat([sd32[k].grad.flatten() for k, p in model.named_parameters()
                       if not is_conv_bias(k) and float(sd[k].grad.norm()) >= 1e-5])
    rep["global_grad_cos_B"] = cosine(ours, ref32)
    rep["global_grad_cos_A_vs_B"] = cosine(ref, ref32)
    for k, p in model.named_parameters():          # per tensor: how far bf16 rounding alone moves the gradient (full tensors)
        if not grads[k]["missing"]:
            grads[k]["cos_AB_full"] = cosine(sd[k].grad, sd32[k].grad)
            grads[k]["cos_B_full"] = cosine(p.grad.detach().cpu(), sd32[k].grad)
    msd = model.state_dict()
    rep["buffers"] = {k: (int(msd[k]) == int(v)) if k.endswith("num_batches_tracked")
                      else float((msd[k].cpu() - v).abs().max() / v.abs().max().clamp_min(1e-6))
                      for k, v in gold["buffers_after"].items()}
    # ---- eval forward with the updated statistics
    model.eval()
    with torch.no_grad():
        e = model(*[t.to(device) for t in inputs])
    e = e if isinstance(e, tuple) else (e,)
    sd_eval = R.clone_state({k: v.cpu() for k, v in model.state_dict().items()}, requires_grad=False)
    with torch.no_grad():
        oe = oracle_forward(kind, sd_eval, inputs, kwargs, False, rnd=R.bf16_round)
        oe32 = oracle_forward(kind, R.clone_state(sd_eval, requires_grad=False), inputs, kwargs, False, rnd=None)
    rep["eval_err_A"] = [float((a.cpu() - b).abs().max()) for a, b in zip(e, oe)]
    rep["eval_err_B"] = [float((a.cpu() - b).abs().max()) for a, b in zip(e, oe32)]
    margin = (oe32[0][:, 0] - oe32[0][:, 1]).abs()
    rep["eval_margin_min"] = float(margin.min())
    rep["eval_argmax_equal"] = bool(torch.equal(e[0].cpu().argmax(1), oe32[0].argmax(1)))
    rep["eval_argmax_equal_sure"] = lambda tol: bool(torch.equal(e[0].cpu().argmax(1)[margin > 2 * tol],
                                                                 oe32[0].argmax(1)[margin > 2 * tol]))
    return rep
