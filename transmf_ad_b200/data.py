"""Input pipeline on the device -- the tensor contract of the reference's MONAI transforms
(datasets/ADNI.py:59-84) produced on the GPU instead of on the training thread.

The reference loads and transforms every sample synchronously on the host (``DataLoader(num_workers=0)``,
kfold_train_adversarial.py:63-66): at B200 speeds (5 ms per 8-subject step) that loop is the bottleneck.  Here the host
hands over RAW volumes; ``GpuBatchTransform`` computes per-volume min / max and applies scaling + flip + rotation + zoom as
one trilinear resampling (csrc/augment.cu), and ``train.DevicePrefetcher(transform=...)`` runs it on the copy stream while the
previous step computes.

Random draws follow the reference's transform parameters: each of RandFlipd / RandRotated / RandZoomd fires with
probability 0.3; theta ~ U(-0.05, 0.05) rad; zoom ~ U(0.95, 1); one draw per SUBJECT shared by MRI and PET (MONAI dictionary
transforms).  Deviations from MONAI, stated: the three geometric transforms are composed into ONE interpolation (MONAI
interpolates after each), and the zoom uses the same trilinear resampling (MONAI: area interpolation + edge padding); MONAI is
not installed here, so this row's parity is pinned to the torch restatement in oracle/augment.py, not to MONAI itself.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib as L


def draw_params(batch, rng, aug=True, prob=0.3, range_x=0.05, min_zoom=0.95, max_zoom=1.0):
    """-> float32 (batch, 4): flip, cos(theta), sin(theta), 1/zoom per subject (identity rows when ``aug`` is False)."""
    out = np.zeros((batch, 4), dtype=np.float32)
    out[:, 1] = 1.0
    out[:, 3] = 1.0
    if aug:
        for b in range(batch):
            if rng.random() < prob:                      # RandFlipd(prob=0.3, spatial_axis=0)
                out[b, 0] = 1.0
            if rng.random() < prob:                      # RandRotated(prob=0.3, range_x=0.05)
                th = rng.uniform(-range_x, range_x)
                out[b, 1], out[b, 2] = math.cos(th), math.sin(th)
            if rng.random() < prob:                      # RandZoomd(prob=0.3, min_zoom=0.95, max_zoom=1)
                out[b, 3] = 1.0 / rng.uniform(min_zoom, max_zoom)
    return out


class GpuBatchTransform:
    """``__call__(mri_raw, pet_raw[, params])`` with device tensors (B,1,D,H,W) fp32 (any intensity range) -> scaled /
    augmented (B,1,D,H,W) fp32 in [0,1].  ``aug=False`` is the test-time transform (scaling only)."""

    def __init__(self, aug=True, seed=0):
        self.aug = aug
        self.rng = np.random.default_rng(seed)
        self._pinned = None

    def __call__(self, *volumes, params=None):
        vols = [v if v.dtype == torch.float32 else v.float() for v in volumes]
        B, _, D, H, W = vols[0].shape
        dev = vols[0].device
        nv = len(vols)
        if params is None:
            params = draw_params(B, self.rng, self.aug)
        if self._pinned is None or self._pinned.shape[0] != B:
            self._pinned = torch.empty((B, 4), dtype=torch.float32).pin_memory()
        self._pinned.copy_(torch.from_numpy(np.ascontiguousarray(params, dtype=np.float32)))
        p_dev = self._pinned.to(dev, non_blocking=True)
        # volumes of one subject must be consecutive: (B, nv, D, H, W)
        src = torch.stack([v.reshape(B, D, H, W) for v in vols], dim=1).contiguous() if nv > 1 else vols[0].reshape(B, 1, D, H, W).contiguous()
        dst = torch.empty_like(src)
        mm = torch.empty(B * nv * 2, dtype=torch.float32, device=dev)
        L.call("tmf_volume_minmax", L.ptr(src), L.ptr(mm), B * nv, D * H * W)
        L.call("tmf_augment_volumes", L.ptr(src), L.ptr(dst), L.ptr(mm), L.ptr(p_dev), B * nv, nv, D, H, W)
        outs = tuple(dst[:, i].reshape(B, 1, D, H, W) for i in range(nv))
        return outs if nv > 1 else outs[0]
