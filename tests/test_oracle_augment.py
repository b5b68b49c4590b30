"""oracle/augment.py (the restatement the device input pipeline is tested against) pinned against an independent
implementation: scipy.ndimage.affine_transform (order 1, mode 'nearest') of the same output->input map, plus exact
identities.  CPU only.  (MONAI, which the reference uses, is not installed here: this row's parity is unpinned against it.)"""
import math

import numpy as np
import torch
from scipy import ndimage

from oracle import augment as OA


def test_identity_and_flip_are_exact():
    v = torch.rand(9, 11, 7, generator=torch.Generator().manual_seed(0)) * 100 - 3
    s = (v - v.min()) / (v.max() - v.min())
    assert torch.equal(OA.transform_volume(v, False, 1.0, 0.0, 1.0), s)
    assert torch.equal(OA.transform_volume(v, True, 1.0, 0.0, 1.0), torch.flip(s, [0]))
    assert float(OA.transform_volume(torch.full((4, 4, 4), 7.0), False, 1.0, 0.0, 1.0).abs().max()) == 0.0


def test_rotation_zoom_flip_against_scipy():
    v = torch.rand(13, 17, 15, generator=torch.Generator().manual_seed(1))
    th, zoom, flip = 0.05, 0.95, True
    got = OA.transform_volume(v, flip, math.cos(th), math.sin(th), 1 / zoom).numpy()
    s = ((v - v.min()) / (v.max() - v.min())).numpy().astype(np.float64)
    D, H, W = s.shape
    c = np.array([(D - 1) / 2, (H - 1) / 2, (W - 1) / 2])
    iz = 1 / zoom
    A = iz * np.array([[1, 0, 0], [0, math.cos(th), math.sin(th)], [0, -math.sin(th), math.cos(th)]])
    if flip:                      # q_d = (D-1) - (iz*p_d + c_d)
        A[0, 0] = -A[0, 0]
    offset = c - A @ c
    if flip:
        offset[0] = (D - 1) - (iz * (-c[0]) + c[0])
    want = ndimage.affine_transform(s, A, offset=offset, order=1, mode="nearest")
    assert float(np.abs(got - want).max()) < 1e-5
