"""Secondary volume shapes of the hot path (SURVEY.md section 8 preamble): 79x95x79 (SPM 2 mm bounding box) and 128x128x79 (the
ADVIT pad, reference datasets/ADNI.py:122).  One train step of ``model_ad`` on the CUDA path against the CPU oracle with bf16
rounding emulated (Oracle-A) and against the fp32 oracle: sNet features, logits (same noise-floor rule as
tests/test_gpu_fullsize.py), whole-model gradient direction, finiteness, and bitwise run-to-run reproducibility."""
import pytest
import torch

from oracle import restatement as R
from tests import helpers as H
from transmf_ad_b200.models import mymodel as M
from transmf_ad_b200.synthetic import make_labels, make_volumes, procedural_state

pytestmark = pytest.mark.gpu
DEV = "cuda"
KW = dict(dim=128, depth=3, heads=4, dim_head=32, mlp_dim=512, dropout=0.)


@pytest.mark.parametrize("shape,tokens", [((79, 95, 79), 80), ((128, 128, 79), 256)])
def test_secondary_shape_train_step(shape, tokens):
    B = 4
    model = M.model_ad(**KW)
    state = procedural_state(model.state_dict(), seed=21)
    model.load_state_dict(state)
    model = model.to(DEV).train()
    H.set_head_dropout(model, 0.0)
    label = make_labels(B)
    mri, pet = make_volumes(B, shape, seed=31, labels=label), make_volumes(B, shape, seed=32, labels=label)
    feat = model.mri_cnn(mri.to(DEV))
    assert feat.shape[0] == B and feat.shape[2] * feat.shape[3] * feat.shape[4] == tokens
    outs = model(mri.to(DEV), pet.to(DEV))
    H.losses(outs, label.to(DEV))[2].backward()
    torch.cuda.synchronize()
    # oracles
    sd_a, sd_b = R.clone_state(state), R.clone_state(state)
    oa = R.model_ad_forward(sd_a, mri, pet, heads=4, training=True, rnd=R.bf16_round, p_drop=0.0)
    ob = R.model_ad_forward(sd_b, mri, pet, heads=4, training=True, rnd=None, p_drop=0.0)
    H.losses(oa, label)[2].backward()
    H.losses(ob, label)[2].backward()
    with torch.no_grad():
        want = R.snet_forward(R.clone_state(state, requires_grad=False), "mri_cnn", mri, True, R.bf16_round)
    assert H.rel_err(feat.detach().float().cpu(), want) <= 2e-2
    ab = max(float((a.detach() - b.detach()).abs().max()) for a, b in zip(oa, ob))
    err_a = max(float((o.detach().cpu() - a.detach()).abs().max()) for o, a in zip(outs, oa))
    err_b = max(float((o.detach().cpu() - b.detach()).abs().max()) for o, b in zip(outs, ob))
    print(f"[shape {shape}] logits vs A {err_a:.3e} vs fp32 {err_b:.3e} (A vs fp32 {ab:.3e})")
    assert err_a <= max(3e-2, ab) and err_b <= max(3e-2, ab) + ab          # vs fp32: triangle inequality through Oracle-A
    keep = [k for k, _ in model.named_parameters() if not H.is_conv_bias(k) and float(sd_a[k].grad.norm()) >= 1e-5]
    ours = torch.cat([dict(model.named_parameters())[k].grad.detach().cpu().flatten() for k in keep])
    ga = torch.cat([sd_a[k].grad.flatten() for k in keep])
    gb = torch.cat([sd_b[k].grad.flatten() for k in keep])
    assert torch.isfinite(ours).all()
    cos_a, cos_ab = H.cosine(ours, ga), H.cosine(ga, gb)
    print(f"[shape {shape}] whole-model gradient cosine vs A {cos_a:.4f} (A vs fp32 {cos_ab:.4f})")
    assert cos_a >= min(0.95, cos_ab - 0.03)
    # bitwise reproducibility at this shape
    g1 = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.load_state_dict(state)
    model.zero_grad(set_to_none=True)
    outs2 = model(mri.to(DEV), pet.to(DEV))
    H.losses(outs2, label.to(DEV))[2].backward()
    for a, b in zip(outs, outs2):
        assert torch.equal(a, b)
    for k, p in model.named_parameters():
        assert torch.equal(p.grad, g1[k]), k
