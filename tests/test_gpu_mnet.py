"""MiSePyNet / Mnet baseline (BASELINE configs[4]) on the CUDA path against the golden fixture written from
the REAL reference (oracle/make_golden.py case ``mnet_b4``: 91x109x91, batch 4) and against the CPU oracle run here.  The whole
network is fp32 on both sides, so the tolerances are fp32-level: logits <= 2e-3 abs (BatchNorm1d over 4 samples amplifies the
~1e-6 summation-order noise), per-tensor gradient cosine >= 0.999 for every live parameter, BatchNorm buffers <= 1e-4 relative;
dead parameters (spatial_cnn.conv2 / conv3, never called by the reference) get no gradient on either side."""
import pytest
import torch

from oracle import restatement as R
from tests import helpers as H
from transmf_ad_b200 import _lib as L
from transmf_ad_b200.models.MiSePyNet import Mnet

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_mnet_train_step_and_eval_against_reference_golden():
    gold = H.load_golden("mnet_b4")
    state = H.case_state(gold)
    model = Mnet()
    model.load_state_dict(state, strict=True)
    model = model.to(DEV).train()
    H.set_head_dropout(model, 0.0)
    mri, pet, label = H.case_inputs(gold)
    n0 = L.launch_count()
    logits = model(mri.to(DEV), pet.to(DEV))
    loss = torch.nn.functional.cross_entropy(logits, label.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    assert L.launch_count() - n0 > 300                      # the conv / BN / pool stack ran on libtmf kernels
    want = gold["train_outs"][0]
    err = float((logits.detach().cpu() - want).abs().max())
    print(f"[mnet] train logits max err {err:.3e}; loss {float(loss):.6f} vs {gold['train_losses'][2]:.6f}")
    assert err <= 2e-3 and abs(float(loss) - gold["train_losses"][2]) <= 1e-3
    assert torch.equal(logits.detach().cpu().argmax(1), want.argmax(1))
    # gradients: full tensors against the oracle (bit-equal to the reference on CPU), samples against the golden fixture
    sd = R.clone_state(state)
    o = R.mnet_forward(sd, mri, pet, training=True, p_drop=0.0)
    torch.nn.functional.cross_entropy(o, label).backward()
    worst = (1.0, None)
    for k, p in model.named_parameters():
        if k not in gold["grad_norm"]:
            assert p.grad is None and sd[k].grad is None, k     # dead parameters
            continue
        g, gr = p.grad.detach().cpu(), sd[k].grad
        assert torch.isfinite(g).all(), k
        parts = k.split(".")
        if k.endswith(".bias") and parts[-2].isdigit() and isinstance(dict(model.named_modules())[".".join(parts[:-1])], torch.nn.Conv3d):
            assert float(g.abs().max()) <= 1e-4, k             # conv bias under train-mode BatchNorm: analytically zero
            continue
        if float(gr.norm()) < 1e-6:
            assert float(g.norm()) < 1e-4, k
            continue
        c = H.cosine(g, gr)
        worst = min(worst, (c, k))
        assert c >= 0.999, f"{k}: cosine {c:.6f}"
        assert H.cosine(H.sample(g), gold["grad_sample"][k]) >= 0.99, k
    print(f"[mnet] worst per-tensor gradient cosine {worst}")
    msd = model.state_dict()
    for k, v in gold["buffers_after"].items():
        if k.endswith("num_batches_tracked"):
            assert int(msd[k]) == int(v), k
        else:
            assert float((msd[k].cpu() - v).abs().max()) <= 1e-4 * float(v.abs().max().clamp_min(1e-3)), k
    # eval forward with the updated statistics (the golden eval used the original weights: no optimizer step here either)
    model.eval()
    with torch.no_grad():
        e = model(mri.to(DEV), pet.to(DEV))
    assert float((e.cpu() - gold["eval_outs"][0]).abs().max()) <= 1e-3
    assert torch.equal(e.cpu().argmax(1), gold["eval_argmax"])


def test_mnet_step_is_bitwise_reproducible():
    torch.manual_seed(0)
    model = Mnet().to(DEV).train()
    H.set_head_dropout(model, 0.0)
    x = torch.rand(2, 1, 91, 109, 91, device=DEV)
    y = torch.rand(2, 1, 91, 109, 91, device=DEV)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    runs = []
    for _ in range(2):
        model.load_state_dict(sd0)
        model.zero_grad(set_to_none=True)
        out = model(x, y)
        out.sum().backward()
        runs.append((out.detach().clone(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}))
    assert torch.equal(runs[0][0], runs[1][0])
    for k in runs[0][1]:
        assert torch.equal(runs[0][1][k], runs[1][1][k]), k
