#!/usr/bin/env python
"""Attention-core microbenchmark (B200): tmf_attn_fwd / tmf_attn_bwd at the fusion transformer's shape (B=8, heads=4,
N=150, dim_head=32), graph replays.  TMF_ATTN_IMPL=0 selects the row-per-warp kernels for comparison.

    python scripts/attn_bench.py [--iters 50] [--batch 8] [--nk 150]
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
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--nq", type=int, default=150)
    ap.add_argument("--nk", type=int, default=150)
    ap.add_argument("--heads", type=int, default=4)
    ap.add_argument("--dh", type=int, default=32)
    a = ap.parse_args()
    dev = "cuda"
    B, Nq, Nk, H, dh = a.batch, a.nq, a.nk, a.heads, a.dh
    inner = H * dh
    q = torch.randn(B * Nq, inner, device=dev)
    kv = torch.randn(B * Nk, 2 * inner, device=dev)
    out, lse = torch.empty_like(q), torch.empty(B * H * Nq, device=dev)
    dout, dq, dkv = torch.randn_like(q), torch.empty_like(q), torch.empty_like(kv)
    scale = dh ** -0.5
    ops = {
        "fwd": lambda: L.call("tmf_attn_fwd", L.ptr(q), L.ptr(kv), L.ptr(out), L.ptr(lse), B, Nq, Nk, H, dh, scale),
        "bwd": lambda: L.call("tmf_attn_bwd", L.ptr(dout), L.ptr(q), L.ptr(kv), L.ptr(out), L.ptr(lse), L.ptr(dq), L.ptr(dkv),
                              B, Nq, Nk, H, dh, scale),
    }
    for name, fn in ops.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(a.iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"attn {name}: B={B} heads={H} Nq={Nq} Nk={Nk} dh={dh}: {e0.elapsed_time(e1) / a.iters * 1e3:7.2f} us per call "
              f"(TMF_ATTN_IMPL={os.environ.get('TMF_ATTN_IMPL', '1')})", flush=True)


if __name__ == "__main__":
    main()
