"""Model-level parity at the BASELINE configurations: full 91x109x91 volumes, ``model_ad`` at
batch 8 (C3) and batch 2 (C1) and ``model_CNN_ad`` at batch 8 (C2).  One train step (forward, the three losses of
kfold_train_adversarial.py:119-131, backward) and one eval forward of the CUDA path against

  * the golden fixtures written from the REAL reference at these sizes (oracle/make_golden.py, "Oracle-B"), and
  * the CPU oracle run here with bf16 rounding emulated where the CUDA path rounds ("Oracle-A").

Tolerances: sNet features <= 2e-2 rel-L2 (A); train-mode logits <= 3e-2 abs vs Oracle-A and vs the fp32
reference; losses <= 2e-2; eval-mode logits <= 1e-2; whole-model gradient cosine >= 0.95 (A); per-tensor gradient cosine
>= 0.9 (A); BatchNorm buffers <= 2e-2 relative; arg-max labels identical in train AND eval mode on every sample whose
reference margin exceeds twice the logit tolerance.
Batch 2 (the reference's default ``--batch_size``): BatchNorm1d over two samples maps every head feature to +-1, so the
train-mode logits and the gradients behind them are sign patterns of tiny differences; there the step is checked on
the sNet features, the BatchNorm3d buffers, finiteness, and the eval-mode logits / labels.
"""
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"

FEAT_A, LOGIT_A, LOGIT_B, LOSS_TOL, EVAL_A = 2e-2, 3e-2, 3e-2, 2e-2, 1e-2
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
    assert max(r["logit_err_A"]) <= LOGIT_A, r["logit_err_A"]
    assert max(r["logit_err_B"]) <= LOGIT_B, r["logit_err_B"]
    assert abs(r["loss"][0] - r["loss"][1]) <= LOSS_TOL and abs(r["loss"][0] - r["loss"][2]) <= LOSS_TOL
    assert r["global_grad_cos_A"] >= GRAD_COS_GLOBAL
    assert r["global_grad_cos_B"] >= min(GRAD_COS_GLOBAL, r["global_grad_cos_A_vs_B"]) - 0.03
    for k, e in r["grads"].items():
        assert not e["missing"] and e["finite"], k
        if e["conv_bias"]:
            assert e["absmax"] <= 1e-3, k
        elif e["norm_A"] < 1e-5:
            assert e["norm"] < 1e-4, k
        else:
            assert e["cos_A"] >= GRAD_COS_TENSOR, f"{k}: cos {e['cos_A']:.4f} rel {e['rel_A']:.3g} vs Oracle-A"
    for k, v in r["buffers"].items():
        assert v is True or v <= BUF_REL, (k, v)
    assert max(r["eval_err_A"]) <= EVAL_A
    # labels: the fixture seeds were chosen so that every reference margin clears twice the logit tolerance
    assert r["train_margin_min"] > 2 * LOGIT_B and r["eval_margin_min"] > 2 * EVAL_A
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
