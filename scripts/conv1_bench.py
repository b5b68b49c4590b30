#!/usr/bin/env python
"""Block-1 microbenchmark (B200): tmf_conv1_fwd at the bench shape, both towers, CUDA events over `--iters` launches.
TMF_C1U_DEBUG (timing only, wrong results): 1 = one MMA pair per K step, 2 = no stores, 4 = no image loads.

    python scripts/conv1_bench.py [--iters 20] [--batch 8]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transmf_ad_b200 import _lib as L          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=8)
    a = ap.parse_args()
    dev, B, ng, cout = "cuda", a.batch, 2, 32
    D, H, W = 91, 109, 91
    x = [torch.rand((B, 1, D, H, W), device=dev) for _ in range(ng)]
    w = [torch.randn((cout, 1, 3, 3, 3), device=dev) * 0.2 for _ in range(ng)]
    b = [torch.zeros(cout, device=dev) for _ in range(ng)]
    y = [torch.empty((B, D, H, W, cout), dtype=torch.bfloat16, device=dev) for _ in range(ng)]
    stats = L.stat_buffers(ng, cout, dev)

    def fn():
        L.call("tmf_conv1_fwd", ng, L.ptrs(x), L.ptrs(w), L.ptrs(b), L.ptrs(y), L.ptrs(stats), B, D, H, W, cout, L.CONV_AUTO)

    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    gb = ng * B * D * H * W * (4 + 2 * cout) / 1e9
    print(f"conv1 fwd B={B} x{ng}: {ms * 1e3:7.1f} us  {gb / ms:7.2f} TB/s algorithmic (TMF_C1U_DEBUG={os.environ.get('TMF_C1U_DEBUG', '0')})",
          flush=True)


if __name__ == "__main__":
    main()
