"""CPU restatement of the device input pipeline (csrc/augment.cu)  --  TEST INFRASTRUCTURE ONLY.

Follows the reference's transform list (datasets/ADNI.py:59-84): ScaleIntensityd (:64), RandFlipd(spatial_axis=0) (:66),
RandRotated(range_x) (:67), RandZoomd (:68), with the parameters drawn by the caller.  MONAI itself is not installed in this
environment (no wheel, no network): PARITY UNPINNED against MONAI; this file restates the documented semantics (min-max
scaling; flip of the first spatial axis; rotation of the plane of the two other axes about the volume centre; isotropic
zoom about the centre with the size kept and border values outside) as ONE affine map sampled with torch's own trilinear
``grid_sample`` (align_corners=True, padding_mode='border')."""
import torch
import torch.nn.functional as F


def transform_volume(vol, flip, cos_t, sin_t, inv_zoom):
    """vol (D,H,W) fp32 -> (D,H,W) fp32 in [0,1]."""
    D, H, W = vol.shape
    lo, hi = vol.min(), vol.max()
    x = (vol - lo) / (hi - lo) if float(hi) > float(lo) else torch.zeros_like(vol)
    cd, ch, cw = (D - 1) / 2.0, (H - 1) / 2.0, (W - 1) / 2.0
    d, h, w = torch.meshgrid(torch.arange(D, dtype=torch.float64), torch.arange(H, dtype=torch.float64),
                             torch.arange(W, dtype=torch.float64), indexing="ij")
    pd, ph, pw = d - cd, h - ch, w - cw
    qd = inv_zoom * pd + cd
    qh = inv_zoom * (cos_t * ph + sin_t * pw) + ch
    qw = inv_zoom * (-sin_t * ph + cos_t * pw) + cw
    if flip:
        qd = (D - 1) - qd
    # grid_sample wants (x=w, y=h, z=d) in [-1, 1] with align_corners=True
    norm = lambda q, n: (2.0 * q / (n - 1) - 1.0) if n > 1 else torch.zeros_like(q)
    grid = torch.stack([norm(qw, W), norm(qh, H), norm(qd, D)], dim=-1).unsqueeze(0)
    out = F.grid_sample(x.double()[None, None], grid, mode="bilinear", padding_mode="border", align_corners=True)
    return out[0, 0].float()
