"""The path bench.py times -- ``GraphedTrainStep`` (CUDA-graph replay) + ``FusedAdam`` -- against the reference's eager loop
(kfold_train_adversarial.py:101-136) and ``torch.optim.Adam`` (utils/utils.py:38-41), and run-to-run reproducibility.

Every kernel on the path is deterministic (include/tmf.h, DETERMINISM), which makes three exact statements possible:
  1. k graph replays == k eager steps, BITWISE, with the same optimizer (FusedAdam, and torch.optim.Adam(capturable));
     the warm-up steps inside ``GraphedTrainStep`` leave no trace (state snapshot / restore).
  2. One step of FusedAdam == one step of torch.optim.Adam as the reference builds it, to fp32 rounding (<= 2e-7 + 2e-6|p|),
     from the same state on the real model; long gradient sequences are compared in tests/test_gpu_ops.py.
  3. Two runs of a full-size step give bit-identical logits, buffers and gradients.
Why (2) is stated for one step: Adam normalises every element's update to ~lr * sign(g), and this network stores its
activations in bf16, so a 1-ulp difference in a weight (fused vs foreach arithmetic: bias corrections in fp32 vs
double) flips single bf16 roundings in the next forward and with them the sign of near-zero gradient elements -- two
CORRECT Adam implementations drift apart by ~lr per element per step (measured: identical after step 1, <= 1.7e-3 at
lr 1e-3 after step 2; torch's own capturable and non-capturable Adam differ by exactly as much).
"""
import copy

import pytest
import torch

from tests import helpers as H
from transmf_ad_b200.models import mymodel as M
from transmf_ad_b200.optim import FusedAdam
from transmf_ad_b200.synthetic import make_labels, make_volumes, procedural_state
from transmf_ad_b200.train import GraphedTrainStep

pytestmark = pytest.mark.gpu
DEV = "cuda"
STEPS = 3
LR = 1e-3                 # larger than the reference's 1e-4 so that three steps move the weights visibly


def _model(kind, kwargs, seed):
    m = getattr(M, kind)(**kwargs)
    m.load_state_dict(procedural_state(m.state_dict(), seed=seed))
    m = m.to(DEV).train()
    H.set_head_dropout(m, 0.0)          # Dropout(0.5) draws from torch's Philox stream: order-dependent, so off here
    return m


def _batches(B, shape, n):
    label = make_labels(B).to(DEV)
    return [(make_volumes(B, shape, seed=100 + i, labels=label.cpu()).to(DEV),
             make_volumes(B, shape, seed=200 + i, labels=label.cpu()).to(DEV), label) for i in range(n)]


def _loss(outs, label):
    ce, ad, total = H.losses(outs, label)
    return total, ce, ad


def _eager_steps(model, opt, batches):
    out = []
    for mri, pet, label in batches:
        opt.zero_grad()
        outs = model(mri, pet)
        total, ce, ad = _loss(outs, label)
        total.backward()
        opt.step()
        out.append(float(total))
    return out


CASES = [("model_CNN_ad", dict(dim=128), 4, (33, 35, 34)),
         ("model_ad", dict(dim=128, depth=3, heads=4, dim_head=32, mlp_dim=512, dropout=0.), 4, (32, 36, 33))]


def _graph_steps(model, opt, batches):
    step = GraphedTrainStep(model, opt, _loss, batches[0][:2], batches[0][2], warmup=3)
    assert step.launches_per_step > 50
    out = []
    for mri, pet, label in batches:
        out.append(float(step((mri, pet), label)[0]))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("optname", ["fused", "torch"])
@pytest.mark.parametrize("kind,kwargs,B,shape", CASES)
def test_graph_replay_equals_eager_loop_bitwise(kind, kwargs, B, shape, optname):
    batches = _batches(B, shape, STEPS)
    ref = _model(kind, kwargs, seed=3)
    ours = copy.deepcopy(ref)
    make = (lambda ps: FusedAdam(ps, lr=LR, weight_decay=0.0)) if optname == "fused" else \
           (lambda ps: torch.optim.Adam(ps, lr=LR, weight_decay=0.0, capturable=True))
    ref_losses = _eager_steps(ref, make(ref.parameters()), batches)
    got_losses = _graph_steps(ours, make(ours.parameters()), batches)
    assert got_losses == ref_losses
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    for k, v in sd_ref.items():
        assert torch.equal(sd_ours[k], v), f"{k}: max diff {float((sd_ours[k].double() - v.double()).abs().max()):.3e}"
    assert int(sd_ours["mri_cnn.conv1.1.num_batches_tracked"]) == STEPS      # warm-up steps left no trace
    start = procedural_state(sd_ours, seed=3)
    moved = max(float((sd_ours[k].cpu() - start[k]).abs().max()) for k in sd_ours if k.endswith("conv2.0.weight"))
    assert moved > 0.5 * LR               # the weights did move (Adam's first steps are ~lr per element)


@pytest.mark.parametrize("kind,kwargs,B,shape", CASES)
def test_one_step_of_fused_adam_equals_torch_adam_on_the_model(kind, kwargs, B, shape):
    batches = _batches(B, shape, 1)
    ref = _model(kind, kwargs, seed=3)
    ours = copy.deepcopy(ref)
    _eager_steps(ref, torch.optim.Adam(ref.parameters(), lr=LR, weight_decay=0.0), batches)      # utils/utils.py:38-41
    _graph_steps(ours, FusedAdam(ours.parameters(), lr=LR, weight_decay=0.0), batches)           # the timed path
    sd_ours = ours.state_dict()
    for k, v in ref.state_dict().items():
        w = sd_ours[k]
        if not v.dtype.is_floating_point:
            assert int(v) == int(w), k
        else:
            err = (w - v).abs()
            assert bool((err <= 2e-7 + 2e-6 * v.abs()).all()), f"{k}: max |dp| {float(err.max()):.3e}"


def test_fused_adam_keeps_the_conv_weight_packs_fresh():
    """FusedAdam rewrites the bf16 operand packs of the conv weights from inside its kernel, so the
    captured step contains no tmf_pack_conv_weights launch; the packs must equal an explicit pack of the updated weights."""
    from transmf_ad_b200 import _lib as L
    kind, kwargs, B, shape = CASES[0]
    batches = _batches(B, shape, 2)
    model = _model(kind, kwargs, seed=3)
    opt = FusedAdam(model.parameters(), lr=LR, weight_decay=0.0)
    step = GraphedTrainStep(model, opt, _loss, batches[0][:2], batches[0][2], warmup=3)
    for mri, pet, label in batches:
        step((mri, pet), label)
    torch.cuda.synchronize()
    for net in (model.mri_cnn, model.pet_cnn):
        for l, (conv, _) in enumerate(net._units()):
            if l == 0:
                continue
            w, pk = conv.weight.detach(), net._packs[l]
            wf, wd = torch.empty_like(pk.wf), torch.empty_like(pk.wd)
            L.call("tmf_pack_conv_weights", 1, L.ptrs([w]), L.ptrs([wf]), L.ptrs([wd]), w.shape[0], w.shape[1], w.shape[2])
            assert torch.equal(wf, pk.wf) and torch.equal(wd, pk.wd), f"layer {l}"
            assert not pk.stale(conv.weight)
    # 99 launches per model_CNN_ad step in round 1; the 6 pack launches (and the per-layer memsets) are gone
    assert step.launches_per_step <= 93, step.launches_per_step


def test_fused_adam_keeps_the_encoder_weight_packs_fresh():
    """The bf16 hi / lo operand packs of the fused transformer encoders (functional.EncPack) are rewritten by FusedAdam's kernel:
    after graph replays they equal an explicit tmf_encoder_pack_weights of the updated weights, bitwise, and the captured
    model_ad step contains no pack launch."""
    from transmf_ad_b200 import _lib as L
    from transmf_ad_b200.models.networks import Transformer
    kind, kwargs = "model_ad", dict(dim=128, depth=2, heads=4, dim_head=32, mlp_dim=512, dropout=0.)
    batches = _batches(4, (33, 36, 34), 2)
    model = _model(kind, kwargs, seed=3)
    before = {k: v.clone() for k, v in model.state_dict().items() if "fuse_transformer" in k and k.endswith("weight")}
    opt = FusedAdam(model.parameters(), lr=LR, weight_decay=0.0)
    step = GraphedTrainStep(model, opt, _loss, batches[0][:2], batches[0][2], warmup=3)
    for mri, pet, label in batches:
        step((mri, pet), label)
    torch.cuda.synchronize()
    after = model.state_dict()
    assert all(not torch.equal(v, after[k]) for k, v in before.items())       # (every encoder weight was updated)
    encs = [m for m in model.modules() if isinstance(m, Transformer)]
    assert len(encs) == 4
    for e in encs:
        ws = e._enc_weights()
        pk = e._enc_pack
        assert pk.pack is not None and not pk.stale(ws, ws[3].shape[0])
        want = torch.empty_like(pk.pack)
        L.call("tmf_encoder_pack_weights", *[L.ptr(w.detach()) for w in ws], int(ws[3].shape[0]), L.ptr(want))
        torch.cuda.synchronize()
        assert torch.equal(want.view(torch.int16), pk.pack.view(torch.int16))
    # an eager step with a torch optimizer moves the weights behind the cache: the next forward repacks
    topt = torch.optim.SGD(model.parameters(), lr=1e-3)
    mri, pet, label = batches[0]
    topt.zero_grad()
    _loss(model(mri, pet), label)[0].backward()
    topt.step()
    assert all(e._enc_pack.stale(e._enc_weights(), 512) for e in encs)
    with torch.no_grad():
        model(mri, pet)
    assert not any(e._enc_pack.stale(e._enc_weights(), 512) for e in encs)


def test_fused_adam_state_dict_round_trip_matches_torch_adam():
    """a resumed FusedAdam must continue with the loaded step count (bias correction) and moments."""
    torch.manual_seed(0)
    p0 = [torch.randn(257, device=DEV), torch.randn(33, 7, device=DEV)]
    grads = [[torch.randn_like(p) for p in p0] for _ in range(5)]

    def run(make_opt, resume_at=None, make_resumed=None):
        ps = [p.clone().requires_grad_(True) for p in p0]
        opt = make_opt(ps)
        for i, gs in enumerate(grads):
            if resume_at is not None and i == resume_at:
                sd = copy.deepcopy(opt.state_dict())
                opt = make_resumed(ps)
                opt.load_state_dict(sd)
            for p, g in zip(ps, gs):
                p.grad = g.clone()
            opt.step()
        return [p.detach() for p in ps]

    want = run(lambda ps: torch.optim.Adam(ps, lr=1e-2))
    fused = lambda ps: FusedAdam(ps, lr=1e-2)
    for got in (run(fused), run(fused, 2, fused),
                run(lambda ps: torch.optim.Adam(ps, lr=1e-2), 3, fused)):        # torch checkpoint -> FusedAdam
        for a, b in zip(got, want):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-7)


def test_full_size_step_is_bitwise_reproducible():
    """BASELINE-sized volumes (91x109x91, batch 2): two runs from the same state give bit-identical logits, BatchNorm
    buffers and conv-tower gradients (per-CTA partial statistics, fixed-order sums; no floating-point atomics)."""
    model = _model("model_ad", CASES[1][1], seed=0)
    label = make_labels(2).to(DEV)
    mri = make_volumes(2, seed=1, labels=label.cpu()).to(DEV)
    pet = make_volumes(2, seed=2, labels=label.cpu()).to(DEV)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    runs = []
    for _ in range(2):
        model.load_state_dict(sd0)
        model.zero_grad(set_to_none=True)
        outs = model(mri, pet)
        H.losses(outs, label)[2].backward()
        torch.cuda.synchronize()
        runs.append(([o.detach().clone() for o in outs], {k: p.grad.clone() for k, p in model.named_parameters()},
                     {k: v.clone() for k, v in model.state_dict().items() if "running" in k}))
    (o1, g1, b1), (o2, g2, b2) = runs
    for a, b in zip(o1, o2):
        assert a.shape == (2, 2) and torch.isfinite(a).all()
        assert torch.equal(a, b)
    for k in b1:
        assert torch.equal(b1[k], b2[k]), k
    for k in g1:
        assert torch.isfinite(g1[k]).all(), k
        assert torch.equal(g1[k], g2[k]), k
    assert model.mri_cnn(mri).shape == (2, 128, 5, 6, 5)


@pytest.mark.parametrize("kind,kwargs", [("model_CNN_ad", dict(dim=128)),
                                         ("model_ad", dict(dim=128, depth=1, heads=4, dim_head=32, mlp_dim=512, dropout=0.))])
def test_flat_gradient_slots_leave_every_gradient_unchanged(kind, kwargs):
    """With a FlatGradReducer installed (what bench.py and the data-parallel path run, world size 1 included) the backward
    kernels write weight gradients straight into the slots of the flat buffer.  Every gradient must equal, BITWISE, the one
    computed without slots -- in particular those of the discriminator ``D``, which runs twice per forward pass (reference
    mymodel.py:210-211): its second backward call must not write into the slot that still holds the first contribution."""
    from transmf_ad_b200.dp import FlatGradReducer
    (mri, pet, label), = _batches(4, (33, 36, 34), 1)
    plain = _model(kind, kwargs, seed=7)
    _loss(plain(mri, pet), label)[0].backward()
    slotted = _model(kind, kwargs, seed=7)
    red = FlatGradReducer(model=slotted).install()
    try:
        for _ in range(2):                                  # twice: the hand-out bookkeeping must reset between backward passes
            slotted.zero_grad(set_to_none=True)
            _loss(slotted(mri, pet), label)[0].backward()
            red.finish()
            torch.cuda.synchronize()
            for (k, p), q in zip(slotted.named_parameters(), plain.parameters()):
                assert torch.equal(p.grad, q.grad), f"{k}: max |diff| {float((p.grad - q.grad).abs().max())}"
        slot_of = {id(p): s for p, s in zip(red.params, red.slots)}
        born = sum(1 for p in slotted.parameters() if p.grad.data_ptr() == slot_of[id(p)].data_ptr())
        assert born >= 0.8 * len(red.params), (born, len(red.params))       # the slots are actually used
    finally:
        red.remove()
