#!/usr/bin/env python
"""Per-layer conv microbenchmark (B200): times tmf_conv3d_fwd (forward and dgrad operands) and tmf_conv3d_wgrad for
the sNet layers at the bench shape, both towers in one launch, CUDA events over `--iters` launches.

    python scripts/conv_bench.py [--layers 1,2] [--ops fwd,dgrad,wgrad] [--iters 20] [--batch 8] [--check]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transmf_ad_b200 import _lib as L          # noqa: E402
from transmf_ad_b200 import functional as TF   # noqa: E402

LAYERS = {1: (32, 32, 3, (45, 54, 45)), 2: (32, 64, 3, (45, 54, 45)), 3: (64, 64, 3, (22, 27, 22)),
          4: (64, 128, 3, (22, 27, 22)), 5: (128, 256, 3, (11, 13, 11)), 6: (256, 128, 1, (11, 13, 11))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", default="1,2,3,4,5,6")
    ap.add_argument("--ops", default="fwd,dgrad,wgrad")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--ng", type=int, default=2)
    args = ap.parse_args()
    dev = "cuda"
    B, ng = args.batch, args.ng
    peak = 2250.0
    total = {}
    for l in [int(x) for x in args.layers.split(",")]:
        cin, cout, ks, (D, H, W) = LAYERS[l]
        taps = ks ** 3
        g = torch.Generator(device=dev).manual_seed(l)
        a = [torch.randn((B, D, H, W, cin), device=dev, generator=g).to(torch.bfloat16) for _ in range(ng)]
        dy = [torch.randn((B, D, H, W, cout), device=dev, generator=g).to(torch.bfloat16) for _ in range(ng)]
        w = [torch.randn((cout, cin, ks, ks, ks), device=dev, generator=g) * 0.05 for _ in range(ng)]
        bias = [torch.zeros(cout, device=dev) for _ in range(ng)]
        wf = [torch.empty((taps, cout, cin), dtype=torch.bfloat16, device=dev) for _ in range(ng)]
        wd = [torch.empty((taps, cin, cout), dtype=torch.bfloat16, device=dev) for _ in range(ng)]
        L.call("tmf_pack_conv_weights", ng, L.ptrs(w), L.ptrs(wf), L.ptrs(wd), cout, cin, ks)
        y = [torch.empty((B, D, H, W, cout), dtype=torch.bfloat16, device=dev) for _ in range(ng)]
        da = [torch.empty((B, D, H, W, cin), dtype=torch.bfloat16, device=dev) for _ in range(ng)]
        dw = [torch.empty((cout, cin, ks, ks, ks), dtype=torch.float32, device=dev) for _ in range(ng)]
        stats = L.stat_buffers(ng, cout, dev)
        ws = TF.wgrad_workspace(ng, L.CONV_AUTO, B, D, H, W, cin, cout, ks, dev)
        flops = 2.0 * B * D * H * W * cout * cin * taps * ng

        def fwd():
            L.call("tmf_conv3d_fwd", ng, L.ptrs(a), L.ptrs(wf), L.ptrs(bias), L.ptrs(y), L.ptrs(stats), B, D, H, W, cin, cout,
                   ks, L.CONV_AUTO)

        def dgrad():
            L.call("tmf_conv3d_fwd", ng, L.ptrs(dy), L.ptrs(wd), L.ptrs(None), L.ptrs(da), L.ptrs(None), B, D, H, W, cout,
                   cin, ks, L.CONV_AUTO)

        def wgrad():
            L.call("tmf_conv3d_wgrad", ng, L.ptrs(dy), L.ptrs(a), L.ptrs(dw), B, D, H, W, cin, cout, ks, L.CONV_AUTO,
                   L.ptr(ws), 0 if ws is None else ws.numel())

        for name, fn in (("fwd", fwd), ("dgrad", dgrad), ("wgrad", wgrad)):
            if name not in args.ops.split(","):
                continue
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            total[name] = total.get(name, 0.0) + ms
            print(f"L{l} {name:5s} {cin:3d}->{cout:3d} k{ks} {D}x{H}x{W}: {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s "
                  f"({100 * flops / ms / 1e9 / peak:4.1f}% of {peak:.0f})", flush=True)
    print("sum ms:", {k: round(v, 4) for k, v in total.items()}, "all", round(sum(total.values()), 4))


if __name__ == "__main__":
    main()
