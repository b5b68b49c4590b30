"""Model-level parity at the BASELINE configurations: full 91x109x91 volumes, ``model_ad`` at
batch 8 (C3) and batch 2 (C1) and ``model_CNN_ad`` at batch 8 (C2).  One train step (forward, the three losses of
kfold_train_adversarial.py:119-131, backward) and one eval forward of the CUDA path against

  * the golden fixtures written from the REAL reference at these sizes (oracle/make_golden.py, "Oracle-B"), and
  * the CPU oracle run here with bf16 rounding emulated where the CUDA path rounds ("Oracle-A").

Tolerances.  The conv operands and the stored conv outputs are bf16; how far THAT moves the
reference's own numbers is measured inside the test as the distance between Oracle-A and the fp32 reference ("A:B"),
because the BatchNorm1d heads of ``model_ad`` amplify it (measured at this size: train-mode logits A:B = 4.8e-2).
  * sNet features: <= 2e-2 rel-L2 vs Oracle-A (kernel exactness upstream of the heads; measured 6e-3);
  * train-mode logits: vs Oracle-A <= max(3e-2, A:B), vs the fp32 reference <= that + A:B (triangle inequality); heads without
    BatchNorm1d (``model_CNN_ad`` classifier): <= 2e-3;  losses <= 2e-2;  eval-mode logits <= 5e-3 (measured 7e-4);
  * gradients: whole-model cosine vs Oracle-A >= min(0.95, cos(Oracle-A, fp32)) (measured 0.945-0.953 with cos(A, fp32) = 0.904); per tensor >= min(0.9, cos(A, fp32) - 0.05) vs Oracle-A, i.e. at least
    as aligned with Oracle-A as Oracle-A is with the fp32 reference; BatchNorm buffers <= 2e-2 relative;
  * LABELS: arg-max identical to the REAL reference in train AND eval mode on every sample; the fixture seeds were
    scanned so that every reference margin exceeds twice the logit tolerance that applies (asserted).
Batch 2 (the reference's default ``--batch_size``): BatchNorm1d over two samples maps every head feature to +-1, so the
train-mode logits and the gradients behind them are sign patterns of tiny differences; there the step is checked on
the sNet features, the BatchNorm3d buffers, finiteness, and the eval-mode logits / labels.
"""
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"

FEAT_A, LOGIT_TOL, LOGIT_NO_BN1D, LOSS_TOL, EVAL_A = 2e-2, 3e-2, 2e-3, 2e-2, 5e-3
GRAD_COS_GLOBAL, GRAD_COS_TENSOR, BUF_REL = 0.95, 0.9, 2e-2


def _summary(r):
    g = r["grads"]
    worst = sorted(((e["cos_A"], k) for k, e in g.items() if not e["missing"] and not e["conv_bias"] and e["norm_A"] >= 1e-5))[:4]
    return (f"feat={r['feat_rel(ours:A, ours:B, A:B)']} logitA={r['logit_err_A']} logitB={r['logit_err_B']} A:B={r['logit_err_A_vs_B']} "
            f"loss={r['loss']} gcos A/B/A:B={r['global_grad_cos_A']:.4f}/{r['global_grad_cos_B']:.4f}/{r['global_grad_cos_A_vs_B']:.4f} "
            f"worst={worst} eval A/B={r['eval_err_A']}/{r['eval_err_B']} margins train/eval={r['train_margin_min']:.3f}/{r['eval_margin_min']:.3f}")


@pytest.mark.parametrize("name", ["model_ad_full_b8", "model_cnn_ad_full_b8"])
def test_full_size_train_step_batch8(name):
    r = H.parity_report(name, DEV)
    print(f"[parity-full] {name}: {_summary(r)}")
    for pfx, (ea, eb, ab) in r["feat_rel(ours:A, ours:B, A:B)"].items():
        assert ea <= FEAT_A, f"{pfx} features vs Oracle-A: {ea}"
        assert eb <= 2 * max(ab, FEAT_A), f"{pfx} features vs fp32 reference: {eb} (Oracle-A itself: {ab})"
    ab = max(r["logit_err_A_vs_B"])                    # what bf16 operand rounding alone does to the reference's logits
    tol_a = max(LOGIT_TOL, ab)
    tol_b = tol_a + ab                                 # triangle inequality through Oracle-A
    assert max(r["logit_err_A"]) <= tol_a, (r["logit_err_A"], ab)
    assert max(r["logit_err_B"]) <= tol_b, (r["logit_err_B"], ab)
    cls_tol = tol_b
    if name.startswith("model_cnn_ad"):                # classifier head without BatchNorm1d: no amplification
        assert r["logit_err_A"][0] <= LOGIT_NO_BN1D and r["logit_err_B"][0] <= LOGIT_NO_BN1D
        cls_tol = LOGIT_NO_BN1D
    assert abs(r["loss"][0] - r["loss"][1]) <= LOSS_TOL and abs(r["loss"][0] - r["loss"][2]) <= LOSS_TOL
    # at least as aligned with Oracle-A as the fp32 reference itself is (bf16 noise floor), and >= 0.95 where that allows
    assert r["global_grad_cos_A"] >= min(GRAD_COS_GLOBAL, r["global_grad_cos_A_vs_B"]), (r["global_grad_cos_A"], r["global_grad_cos_A_vs_B"])
    assert r["global_grad_cos_B"] >= min(GRAD_COS_GLOBAL, r["global_grad_cos_A_vs_B"]) - 0.03
    for k, e in r["grads"].items():
        assert not e["missing"] and e["finite"], k
        if e["conv_bias"]:
            assert e["absmax"] <= 1e-3, k
        elif e["norm_A"] < 1e-5:
            assert e["norm"] < 1e-4, k
        else:
            floor = min(GRAD_COS_TENSOR, e["cos_AB_full"] - 0.05)
            assert e["cos_A"] >= floor, f"{k}: cos {e['cos_A']:.4f} (Oracle-A vs fp32: {e['cos_AB_full']:.4f}) rel {e['rel_A']:.3g}"
    for k, v in r["buffers"].items():
        assert v is True or v <= BUF_REL, (k, v)
    assert max(r["eval_err_A"]) <= EVAL_A
    # labels: the fixture seeds were chosen so that every reference margin clears twice the logit tolerance
    assert r["train_margin_min"] > 2 * cls_tol and r["eval_margin_min"] > 2 * EVAL_A, (r["train_margin_min"], cls_tol)
    assert r["train_argmax_equal"], "train-mode arg-max labels differ from the reference"
    assert r["eval_argmax_equal"], "eval-mode arg-max labels differ from the reference"


def test_full_size_train_step_batch2_c1():
    r = H.parity_report("model_ad_full_b2", DEV)
    print(f"[parity-full] model_ad_full_b2: {_summary(r)}")
    for pfx, (ea, eb, ab) in r["feat_rel(ours:A, ours:B, A:B)"].items():
        assert ea <= FEAT_A, f"{pfx} features vs Oracle-A: {ea}"
        assert eb <= 2 * max(ab, FEAT_A)
    for k, e in r["grads"].items():
        assert not e["missing"] and e["finite"], k
    for k, v in r["buffers"].items():
        if "_cnn." in k:                      # BatchNorm3d statistics (the BatchNorm1d ones sit behind the +-1 sign patterns)
            assert v is True or v <= BUF_REL, (k, v)
    assert max(r["eval_err_A"]) <= EVAL_A
    assert r["eval_argmax_equal_sure"](EVAL_A)
