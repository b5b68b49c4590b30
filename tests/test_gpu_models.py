"""Model-level parity (GPU): the drop-in modules against (a) the golden fixtures written from the REAL reference
and (b) the CPU oracle with bf16 rounding emulated at the CUDA path's rounding points ("Oracle-A").

Tolerances (the network is ill-conditioned under train-mode BatchNorm with tiny batches):
  vs Oracle-A : logits <= 2e-2 abs, losses <= 1e-2, per-tensor gradient relative L2 <= 0.15 / cosine >= 0.98
  vs fp32 golden (Oracle-B): logits <= 6e-2 abs, losses <= 3e-2, gradient cosine >= 0.8, eval argmax identical
    wherever the reference margin exceeds twice the logit tolerance.
Conv-bias gradients are excluded from relative comparisons (train-mode BN cancels them; the reference value is
rounding noise ~1e-6) but must be tiny in absolute terms.
"""
import pytest
import torch

from oracle import restatement as R
from tests import helpers as H
from transmf_ad_b200.models import mymodel as M

pytestmark = pytest.mark.gpu
DEV = "cuda"

LOGIT_A, LOSS_A, GRAD_REL_A, GRAD_COS_A = 2e-2, 1e-2, 0.15, 0.98
LOGIT_B, LOSS_B, GRAD_COS_B = 6e-2, 3e-2, 0.8


def is_conv_bias(k):
    parts = k.split(".")
    return k.endswith(".bias") and ".conv" in k and parts[-2] in ("0", "3")


def build(gold):
    model = getattr(M, gold["kind"])(**gold["kwargs"])
    model.load_state_dict(H.case_state(gold), strict=True)
    return model.to(DEV)


def run_train_step(model, inputs, label):
    model.train()
    H.set_head_dropout(model, 0.0)
    model.zero_grad(set_to_none=True)
    outs = model(*[t.to(DEV) for t in inputs])
    outs = outs if isinstance(outs, tuple) else (outs,)
    ce, ad, total = H.losses(outs, label.to(DEV))
    total.backward()
    return outs, (float(ce), float(ad), float(total))


CASES = ["model_ad_h4", "model_ad_h8", "model_cnn_ad", "model_single", "model_transformer", "model_transformer_res",
         "model_cnn", "model_ad_dim64"]


@pytest.mark.parametrize("name", CASES)
def test_train_step_against_reference_golden_and_oracle_a(name):
    gold = H.load_golden(name)
    mri, pet, label = H.case_inputs(gold)
    inputs = (mri,) if gold["kind"] == "model_single" else (mri, pet)
    model = build(gold)
    outs, (ce, ad, total) = run_train_step(model, inputs, label)
    # ---- Oracle-A on the CPU
    sd = R.clone_state(H.case_state(gold))
    o_outs = H.oracle_forward(gold["kind"], sd, inputs, gold["kwargs"], True, 0.0, rnd=R.bf16_round)
    o_ce, o_ad, o_total = H.losses(o_outs, label)
    o_total.backward()
    report = []
    for o, a, b in zip(outs, o_outs, gold["train_outs"]):
        da = float((o.detach().cpu() - a.detach()).abs().max())
        db = float((o.detach().cpu() - b).abs().max())
        report.append((da, db))
        assert da <= LOGIT_A, f"logits vs Oracle-A {da}"
        assert db <= LOGIT_B, f"logits vs fp32 reference {db}"
    assert abs(total - float(o_total)) <= LOSS_A
    assert abs(total - gold["train_losses"][2]) <= LOSS_B
    worst_rel, worst_cos_b = 0.0, 1.0
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        assert torch.isfinite(p.grad).all(), k
        g = p.grad.detach().cpu()
        ga = sd[k].grad
        if is_conv_bias(k):
            assert float(g.abs().max()) <= 1e-3, k
            continue
        if float(ga.norm()) < 1e-7:
            continue
        rel, cos = H.rel_err(g, ga), H.cosine(g, ga)
        worst_rel = max(worst_rel, rel)
        assert rel <= GRAD_REL_A or cos >= GRAD_COS_A, f"{k}: rel {rel:.3g} cos {cos:.4f} vs Oracle-A"
        # fp32 reference: norm and strided sample
        gs = gold["grad_sample"][k]
        if gold["grad_norm"][k] > 1e-7 and gs.numel() >= 8:
            cb = H.cosine(H.sample(g), gs)
            worst_cos_b = min(worst_cos_b, cb)
            assert cb >= GRAD_COS_B, f"{k}: cosine {cb:.3f} vs fp32 reference sample"
    print(f"[parity] {name}: logits(A,B)={report} loss={total:.5f}/{float(o_total):.5f}/{gold['train_losses'][2]:.5f} "
          f"worst grad rel(A)={worst_rel:.3g} worst cos(B)={worst_cos_b:.3f}")
    # ---- BatchNorm buffers after the step (running stats are fp32 statistics of bf16-rounded activations)
    msd = model.state_dict()
    for k, v in gold["buffers_after"].items():
        if k.endswith("num_batches_tracked"):
            assert int(msd[k]) == int(v), k
        else:
            assert torch.allclose(msd[k].cpu(), v, atol=2e-2, rtol=5e-2), k
    # ---- eval path (val_step): argmax labels identical where the reference margin is meaningful
    model.eval()
    with torch.no_grad():
        e = model(*[t.to(DEV) for t in inputs])
    e = e if isinstance(e, tuple) else (e,)
    sd_eval = R.clone_state({k: v.cpu() for k, v in model.state_dict().items()}, requires_grad=False)
    with torch.no_grad():
        oe = H.oracle_forward(gold["kind"], sd_eval, inputs, gold["kwargs"], False, rnd=R.bf16_round)
    for a, b in zip(e, oe):
        assert float((a.cpu() - b).abs().max()) <= LOGIT_A
    ref_logits = oe[0]
    margin = (ref_logits[:, 0] - ref_logits[:, 1]).abs()
    sure = margin > 2 * LOGIT_A
    assert torch.equal(e[0].cpu().argmax(1)[sure], ref_logits.argmax(1)[sure])


def test_ragged_and_single_sample_eval_batches():
    """val/test loaders have no drop_last (kfold_train_adversarial.py:65-66): B = 1 and odd B must work in eval."""
    gold = H.load_golden("model_ad_h4")
    model = build(gold).eval()
    mri, pet, _ = H.case_inputs(gold)
    with torch.no_grad():
        full = model(mri.to(DEV), pet.to(DEV))[0]
        one = model(mri[:1].to(DEV), pet[:1].to(DEV))[0]
    assert torch.allclose(full[:1], one, atol=1e-5)


def test_reference_call_forms():
    """revgrad with the reference's 1-element device tensor, sNet standalone, PreNorm/Attention standalone."""
    from transmf_ad_b200.models import networks as N
    from transmf_ad_b200.models.gradient_reversal import GradientReversal
    x = torch.rand(2, 8, 128, device=DEV, requires_grad=True)
    GradientReversal(2.0)(x).sum().backward()
    assert torch.equal(x.grad, torch.full_like(x, -2.0))
    net = N.sNet(128).to(DEV).train()
    out = net(torch.rand(2, 1, 32, 32, 32, device=DEV))
    assert out.shape == (2, 128, 2, 2, 2) and out.dtype == torch.float32
    enc = N.Transformer(128, 2, 4, 32, 256).to(DEV)
    t = torch.rand(2, 8, 128, device=DEV)
    ref_sd = {k: v.cpu() for k, v in enc.state_dict().items()}
    y = enc(t)               # self-attention form (context=None)
    sd = {"enc." + k: v for k, v in ref_sd.items()}
    with torch.no_grad():
        want = R.transformer_encoder(sd, "enc", t.cpu(), None, 4)
    assert torch.allclose(y.detach().cpu(), want, atol=1e-4)


def test_grouped_towers_equal_separate_towers():
    gold = H.load_golden("model_cnn_ad")
    model = build(gold).eval()
    mri, pet, _ = H.case_inputs(gold)
    from transmf_ad_b200.models.networks import snet_pair_forward
    with torch.no_grad():
        fm, fp = snet_pair_forward(model.mri_cnn, model.pet_cnn, mri.to(DEV), pet.to(DEV))
        fm1, fp1 = model.mri_cnn(mri.to(DEV)), model.pet_cnn(pet.to(DEV))
    assert torch.equal(fm, fm1) and torch.equal(fp, fp1)


def test_full_size_volume_properties():
    """BASELINE-sized input (91x109x91): shapes, finiteness, BN statistics self-consistency and determinism of the
    forward; linearity of the gradient-reversal branch in lambda."""
    from transmf_ad_b200.synthetic import make_labels, make_volumes
    torch.manual_seed(0)
    model = M.model_ad(128, 3, 4, 32, 512, 0.).to(DEV).train()
    H.set_head_dropout(model, 0.0)
    label = make_labels(2)
    mri = make_volumes(2, seed=1, labels=label).to(DEV)
    pet = make_volumes(2, seed=2, labels=label).to(DEV)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    o1 = model(mri, pet)
    model.load_state_dict(sd0)
    o2 = model(mri, pet)
    for a, b in zip(o1, o2):
        assert a.shape == (2, 2) and torch.isfinite(a).all()
        assert torch.allclose(a, b, atol=5e-3)      # fp32 atomics in the BN statistics are order-dependent
    H.losses(o2, label.to(DEV))[2].backward()
    for k, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
    feat = model.mri_cnn(mri)
    assert feat.shape == (2, 128, 5, 6, 5)
