"""Device input pipeline (csrc/augment.cu, transmf_ad_b200/data.py) against the torch restatement oracle/augment.py."""
import numpy as np
import pytest
import torch

from oracle import augment as OA
from transmf_ad_b200.data import GpuBatchTransform, draw_params

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _raw(B, shape, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand((B, 1) + shape, generator=g) * 3000.0 - 200.0).float()      # raw scanner intensities, not [0,1]


@pytest.mark.parametrize("shape", [(19, 23, 17), (91, 109, 91)])
def test_scaling_and_affine_augmentation_match_the_restatement(shape):
    B = 4
    mri, pet = _raw(B, shape, 1), _raw(B, shape, 2)
    params = np.array([[0, 1, 0, 1],                                   # identity: scaling only
                       [1, 1, 0, 1],                                   # flip
                       [0, np.cos(0.05), np.sin(0.05), 1 / 0.95],      # rotation + zoom
                       [1, np.cos(-0.03), np.sin(-0.03), 1 / 0.97]], dtype=np.float32)
    tf = GpuBatchTransform(aug=True, seed=0)
    om, op = tf(mri.to(DEV), pet.to(DEV), params=params)
    assert om.shape == mri.shape and om.dtype == torch.float32
    for b in range(B):
        for got, src in ((om, mri), (op, pet)):
            want = OA.transform_volume(src[b, 0], bool(params[b, 0]), float(params[b, 1]), float(params[b, 2]), float(params[b, 3]))
            err = float((got[b, 0].cpu() - want).abs().max())
            assert err <= 2e-5, (b, err)
    # identity rows are exact min-max scaling
    x = mri[0, 0]
    assert torch.equal(om[0, 0].cpu(), ((x.to(DEV) - x.min()) * (1.0 / (x.max() - x.min())).to(DEV)).cpu())
    assert float(om.min()) >= 0.0 and float(om.max()) <= 1.0 + 1e-6


def test_draws_follow_the_reference_probabilities_and_ranges():
    rng = np.random.default_rng(0)
    p = draw_params(20000, rng)
    assert abs(p[:, 0].mean() - 0.3) < 0.02
    rot = p[:, 2] != 0
    assert abs(rot.mean() - 0.3) < 0.02 and np.all(np.abs(np.arcsin(p[rot, 2])) <= 0.05 + 1e-6)
    zoom = p[:, 3] != 1
    assert abs(zoom.mean() - 0.3) < 0.02 and np.all((1 / p[zoom, 3] >= 0.95 - 1e-6) & (1 / p[zoom, 3] <= 1 + 1e-6))
    assert np.array_equal(draw_params(3, rng, aug=False), np.tile(np.array([0, 1, 0, 1], np.float32), (3, 1)))


def test_prefetcher_applies_the_transform_on_the_copy_stream():
    from transmf_ad_b200.train import DevicePrefetcher
    shape = (16, 18, 17)
    host = [(_raw(2, shape, 10 + i).pin_memory(), _raw(2, shape, 20 + i).pin_memory(), torch.tensor([0, 1]).pin_memory())
            for i in range(3)]
    tf = GpuBatchTransform(aug=False)
    got = []
    for mri, pet, label in DevicePrefetcher(iter(host), DEV, transform=lambda b: tf(b[0], b[1]) + (b[2],)):
        got.append((mri.clone(), pet.clone(), label.clone()))
    assert len(got) == 3
    for (m, p, l), (hm, hp, hl) in zip(got, host):
        want = OA.transform_volume(hm[1, 0], False, 1.0, 0.0, 1.0)
        assert float((m[1, 0].cpu() - want).abs().max()) <= 2e-6 and torch.equal(l.cpu(), hl)
