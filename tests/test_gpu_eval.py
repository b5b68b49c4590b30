"""Inference path: BatchNorm folded into the conv operands + metric glue, against the golden eval outputs
of the REAL reference (tests/golden: ``eval_outs`` / ``eval_argmax``, ALL samples) and against torch / numpy metrics."""
import pytest
import torch

from tests import helpers as H
from transmf_ad_b200 import _lib as L
from transmf_ad_b200 import evaluate as E
from transmf_ad_b200.models import mymodel as M

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("name", ["model_ad_h4", "model_cnn_ad", "model_single", "model_ad_full_b8", "model_cnn_ad_full_b8"])
def test_eval_path_matches_reference_golden(name, monkeypatch):
    gold = H.load_golden(name)
    model = getattr(M, gold["kind"])(**gold["kwargs"])
    model.load_state_dict(H.case_state(gold))
    for k, v in gold["buffers_after"].items():                  # the reference evaluated with the post-step statistics
        model.state_dict()[k].copy_(v)
    model = model.to(DEV).eval()
    mri, pet, label = H.case_inputs(gold)
    batch = {"MRI": mri, "label": label}
    if gold["kind"] != "model_single":
        batch["PET"] = pet
    n0 = L.launch_count()
    out = E.val_step(model, batch, DEV)
    n1 = L.launch_count()
    want = gold["eval_outs"][0]
    err = float((out["logits"].cpu() - want).abs().max())
    print(f"[eval] {name}: max |logit - reference| {err:.3e}; launches {n1 - n0}")
    assert err <= 5e-3
    assert torch.equal(out["pred"].cpu(), gold["eval_argmax"])          # every sample
    assert torch.allclose(out["prob"].cpu(), torch.softmax(want, 1)[:, -1], atol=2e-3)
    # second batch with the same weights: the folds are reused (no fold launches), result bit-identical
    out2 = E.val_step(model, batch, DEV)
    n2 = L.launch_count()
    assert (n2 - n1) < (n1 - n0), (n0, n1, n2)
    assert torch.equal(out2["logits"], out["logits"]), float((out2["logits"] - out["logits"]).abs().max())
    # the unfolded path (statistics-free BN applied after the conv) agrees
    monkeypatch.setenv("TMF_EVAL_FOLD", "0")
    out3 = E.val_step(model, batch, DEV)
    assert float((out3["logits"] - out["logits"]).abs().max()) <= 5e-3


def test_fold_follows_training_and_weight_changes():
    gold = H.load_golden("model_cnn_ad")
    model = getattr(M, gold["kind"])(**gold["kwargs"])
    model.load_state_dict(H.case_state(gold))
    model = model.to(DEV)
    mri, pet, label = H.case_inputs(gold)
    batch = {"MRI": mri, "PET": pet, "label": label}
    a = E.val_step(model, batch, DEV)["logits"].clone()
    model.train()
    model(mri.to(DEV), pet.to(DEV))                              # running statistics move
    b = E.val_step(model, batch, DEV)["logits"].clone()
    assert not torch.equal(a, b)
    with torch.no_grad():
        model.mri_cnn.conv3[0].weight.mul_(1.5)
    c = E.val_step(model, batch, DEV)["logits"]
    assert not torch.equal(b, c)


def test_metrics_match_reference_formulas():
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(200, 2, generator=g)
    logits[::7] = logits[3]                                      # ties in the scores
    label = (torch.rand(200, generator=g) > 0.4).long()
    acc = E.EvalAccumulator(DEV)
    for i in range(0, 200, 64):
        lg, lb = logits[i:i + 64].to(DEV), label[i:i + 64].to(DEV)
        pred, prob = E.eval_head(lg)
        acc.update({"logits": lg, "label": lb, "pred": pred, "prob": prob, "loss": torch.nn.functional.cross_entropy(lg, lb)})
    res = acc.compute()
    pred = logits.argmax(1)
    assert res["accuracy"] == pytest.approx(float((pred == label).float().mean()))
    cm = [[int(((label == i) & (pred == j)).sum()) for j in range(2)] for i in range(2)]
    assert res["confusion"] == cm
    assert res["loss"] == pytest.approx(float(torch.nn.functional.cross_entropy(logits, label)), rel=1e-5)
    # AUC against the O(n^2) definition
    p = torch.softmax(logits, 1)[:, -1].double()
    pos, neg = p[label == 1], p[label == 0]
    auc = float(((pos[:, None] > neg[None, :]).double() + 0.5 * (pos[:, None] == neg[None, :]).double()).mean())
    assert res["auc"] == pytest.approx(auc, abs=1e-9)
    sen, spe, f1 = E.cal_confusion_metrics(cm)
    assert (res["sen"], res["spe"], res["f1"]) == pytest.approx((sen, spe, f1))
