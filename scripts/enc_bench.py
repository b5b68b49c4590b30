#!/usr/bin/env python
"""Per-kernel timing of one fused transformer encoder (csrc/enc_fused.cu + attention core) at the bench shape
(B=8, 150 tokens, dim 128, mlp 512): CUDA events over --iters launches of each C-ABI entry point, fwd+bwd totals."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transmf_ad_b200 import _lib as L
from transmf_ad_b200 import functional as TF
from transmf_ad_b200.models import networks as N

ap = argparse.ArgumentParser(); ap.add_argument("--iters", type=int, default=50); ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--tokens", type=int, default=150); ap.add_argument("--mlp", type=int, default=512)
args = ap.parse_args()
dev = "cuda"
B, Nt, mlp = args.batch, args.tokens, args.mlp
enc = N.Transformer(128, 1, 4, 32, mlp).to(dev)
x = torch.randn(B, Nt, 128, device=dev, requires_grad=True); c = torch.randn(B, Nt, 128, device=dev, requires_grad=True)
w = torch.randn(B, Nt, 128, device=dev)
def step():
    enc.zero_grad(set_to_none=True); x.grad = None; c.grad = None
    y = enc(x, context=c, add_input=True); (y * w).sum().backward()
for fused in ("1", "0"):
    os.environ["TMF_ENC_FUSED"] = fused
    for _ in range(3): step()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(1e8)); L.TIMER.start()
    for _ in range(args.iters): step()
    rec = L.TIMER.stop()
    tot = sum(v[0] for v in rec.values()) / args.iters
    print(f"TMF_ENC_FUSED={fused}: {len(rec)} entry points, {sum(v[1] for v in rec.values()) // args.iters} calls, {tot * 1e3:.1f} us of kernels per encoder fwd+bwd")
    for k, (ms, n) in sorted(rec.items(), key=lambda kv: -kv[1][0]):
        print(f"   {k:28s} {ms / n * 1e3:7.1f} us x {n // args.iters}")
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g): step()
    for _ in range(3): g.replay()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"   graph replay of the encoder fwd+bwd: {e0.elapsed_time(e1) / args.iters * 1e3:.1f} us")
