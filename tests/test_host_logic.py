"""Host-side logic that needs no GPU: drop-in API surface, state_dict parity, synthetic data, loud failure."""
import inspect
import json

import pytest
import torch

from tests import helpers as H
from transmf_ad_b200.models import mymodel as M
from transmf_ad_b200.models import networks as N
from transmf_ad_b200.synthetic import make_volumes, procedural_state

CTOR = {
    "model_ad": dict(dim=128, depth=3, heads=4, dim_head=32, mlp_dim=512, dropout=0.),
    "model_CNN_ad": dict(dim=128),
    "model_single": dict(dim=128),
    "model_transformer": dict(dim=128, depth=2, heads=4, dim_head=32, mlp_dim=256, dropout=0.),
    "model_transformer_res": dict(dim=128, depth=2, heads=4, dim_head=32, mlp_dim=256, dropout=0.),
    "model_CNN": dict(dim=128),
}


@pytest.mark.parametrize("kind", sorted(CTOR))
def test_state_dict_equals_reference_manifest(kind):
    """keys, ORDER, shapes and dtypes equal the reference's state_dict (fixture written from the real reference)."""
    m = getattr(M, kind)(**CTOR[kind])
    man = H.manifest()[H.manifest_key(kind, CTOR[kind])]
    ours = [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()]
    assert ours == man


def test_reference_checkpoint_roundtrip():
    """A reference-shaped state_dict loads strictly and round-trips (ignite Checkpoint / load_objects path)."""
    m = M.model_ad(**CTOR["model_ad"])
    state = procedural_state(H.template_from_manifest("model_ad", CTOR["model_ad"]), seed=3)
    m.load_state_dict(state, strict=True)
    for k, v in m.state_dict().items():
        assert torch.equal(v, state[k]), k


def test_constructor_and_forward_signatures():
    assert list(inspect.signature(M.model_ad.__init__).parameters) == ["self", "dim", "depth", "heads", "dim_head", "mlp_dim", "dropout"]
    assert list(inspect.signature(M.model_ad.forward).parameters) == ["self", "mri", "pet"]
    assert list(inspect.signature(M.model_CNN_ad.__init__).parameters) == ["self", "dim"]
    assert list(inspect.signature(M.model_single.forward).parameters) == ["self", "img"]
    assert list(inspect.signature(N.sNet.__init__).parameters) == ["self", "dim"]
    assert list(inspect.signature(N.Attention.__init__).parameters) == ["self", "dim", "heads", "dim_head", "dropout"]
    assert list(inspect.signature(N.Transformer.__init__).parameters) == ["self", "dim", "depth", "heads", "dim_head", "mlp_dim", "dropout"]
    for name in ("sNet", "PreNorm", "FeedForward", "Attention", "Transformer", "CrossTransformer", "CrossTransformer_MOD_AVG"):
        assert hasattr(N, name)
    from transmf_ad_b200.models.gradient_reversal import GradientReversal, revgrad  # noqa: F401


def test_parameter_counts_match_survey():
    assert sum(p.numel() for p in M.model_ad(**CTOR["model_ad"]).parameters()) == 4173060
    assert sum(p.numel() for p in M.model_CNN_ad(128).parameters()) == 2720580
    assert sum(p.numel() for p in M.model_single(128).parameters()) == 1343586


def test_cnn_init_matches_reference_scheme():
    m = M.model_single(128)
    bn = m.cnn.conv2[1]
    assert torch.all(bn.weight == 1) and torch.all(bn.bias == 0)
    w = m.cnn.conv3[0].weight
    std_expected = (2.0 / (w.shape[0] * 27)) ** 0.5           # kaiming fan_out, relu
    assert abs(float(w.std()) - std_expected) < 0.1 * std_expected


def test_synthetic_volumes_are_deterministic_and_scaled():
    a = make_volumes(2, (20, 24, 18), seed=7)
    b = make_volumes(2, (20, 24, 18), seed=7)
    assert torch.equal(a, b) and a.shape == (2, 1, 20, 24, 18) and a.dtype == torch.float32
    assert float(a.min()) == 0.0 and float(a.max()) == 1.0
    assert not torch.equal(a, make_volumes(2, (20, 24, 18), seed=8))


def test_procedural_state_is_deterministic():
    t = H.template_from_manifest("model_single", CTOR["model_single"])
    a, b = procedural_state(t, 5), procedural_state(t, 5)
    assert all(torch.equal(a[k], b[k]) for k in a)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    """Without a CUDA device the product path raises instead of silently computing somewhere else."""
    m = M.model_single(128)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        m(torch.rand(1, 1, 16, 16, 16))


def test_snet_spec_matches_reference_layer_table():
    from transmf_ad_b200.functional import SNetSpec
    spec = SNetSpec(128)
    assert [(a, b, k) for a, b, k, _ in spec.layers] == [(1, 32, 3), (32, 32, 3), (32, 64, 3), (64, 64, 3), (64, 128, 3),
                                                         (128, 256, 3), (256, 128, 1)]
    assert [p for *_, p in spec.layers] == [1, 0, 1, 0, 1, 0, 2]


def test_fused_adam_host_logic():
    """FusedAdam mirrors torch.optim.Adam's constructor / param_groups and refuses anything it cannot run on the GPU
    (there is no CPU fallback)."""
    import pytest
    import torch
    from transmf_ad_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.zeros(4))
    opt = FusedAdam([p], lr=1e-4, weight_decay=0.0)
    g = opt.param_groups[0]
    assert g["lr"] == 1e-4 and g["betas"] == (0.9, 0.999) and g["eps"] == 1e-8 and g["weight_decay"] == 0.0
    opt.step()                                   # no gradients yet: nothing to do, like torch.optim.Adam
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        opt.step()                               # CPU parameter: must fail loudly
    with pytest.raises(ValueError):
        FusedAdam([p], amsgrad=True)


def test_encoder_pack_cache_follows_the_weight_versions():
    """functional.EncPack (host logic, no kernels): the cached bf16 hi / lo operand pack of a fused encoder is fresh exactly while
    every one of the five weight matrices still has the version it was packed from; the slices handed to FusedAdam
    (``_tmf_encpack``) tile the pack in the order Wq, Wkv, Wo, W1, W2, each as [hi | lo]."""
    import torch
    from transmf_ad_b200.models.networks import Transformer
    enc = Transformer(128, 1, 4, 32, 512)
    ws = enc._enc_weights()
    pk = enc._enc_pack
    assert pk.stale(ws, 512)                                   # never packed
    assert pk.pack.numel() == 2 * sum(w.numel() for w in ws) and pk.pack.dtype == torch.bfloat16
    off = 0
    for w in ws:
        hi, lo = w._tmf_encpack
        assert hi.data_ptr() == pk.pack.data_ptr() + 2 * off and lo.data_ptr() == hi.data_ptr() + 2 * w.numel()
        assert hi.numel() == lo.numel() == w.numel() and w._tmf_pack_cache is pk
        off += 2 * w.numel()
    for w in ws:
        pk.mark(w)
    assert not pk.stale(ws, 512)
    with torch.no_grad():
        ws[3].mul_(1.5)                                        # a torch optimizer / load_state_dict moves W1 in place
    assert pk.stale(ws, 512)
    pk.mark(ws[3])
    assert not pk.stale(ws, 512)
    enc.invalidate_packs()
    assert pk.stale(ws, 512)
