"""Bring-up of the tcgen05 weight-gradient kernel against the CUDA-core implementation (one subprocess per case)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (B, D, H, W, cin, cout, ks)
    "k1_256_128": (2, 5, 6, 5, 256, 128, 1),
    "k3_32_32": (1, 4, 10, 14, 32, 32, 3),
    "k3_32_64": (2, 5, 12, 45, 32, 64, 3),
    "k3_64_64": (1, 6, 27, 22, 64, 64, 3),
    "k3_64_128": (1, 6, 27, 22, 64, 128, 3),
    "k3_128_256": (2, 5, 13, 11, 128, 256, 3),
    "k3_64_32": (1, 4, 10, 14, 64, 32, 3),
    "k1_64_64": (1, 2, 8, 16, 64, 64, 1),
}


def run_case(name):
    import torch
    from transmf_ad_b200 import _lib as L
    from transmf_ad_b200 import functional as TF
    B, D, H, W, cin, cout, ks = CASES[name]
    g = torch.Generator().manual_seed(0)
    a = torch.randn(B, D, H, W, cin, generator=g).to(torch.bfloat16).cuda()
    dy = torch.randn(B, D, H, W, cout, generator=g).to(torch.bfloat16).cuda()
    dw0 = torch.empty(cout, cin, ks, ks, ks, dtype=torch.float32, device="cuda")
    dw1 = torch.full_like(dw0, float("nan"))
    L.call("tmf_conv3d_wgrad", 1, L.ptrs([dy]), L.ptrs([a]), L.ptrs([dw0]), B, D, H, W, cin, cout, ks, L.CONV_DIRECT, L.ptr(None), 0)
    ws = TF.wgrad_workspace(1, L.CONV_UMMA, B, D, H, W, cin, cout, ks, "cuda")
    L.call("tmf_conv3d_wgrad", 1, L.ptrs([dy]), L.ptrs([a]), L.ptrs([dw1]), B, D, H, W, cin, cout, ks, L.CONV_UMMA, L.ptr(ws), ws.numel())
    torch.cuda.synchronize()
    nan = int(torch.isnan(dw1).sum())
    d = (dw1 - dw0).abs()
    scale = float(dw0.abs().max())
    rel = float((dw1 - dw0).norm() / dw0.norm()) if nan == 0 else -1
    bad = (d > 0.01 * scale) | torch.isnan(dw1)
    msg = f"rel_l2={rel:.3g} bad={int(bad.sum())}/{d.numel()} nan={nan}"
    if int(bad.sum()):
        bt = bad.reshape(cout, cin, -1)
        msg += f" bad by tap: {bt.sum(dim=(0, 1)).tolist()} by ci/16: {bt.sum(dim=(0, 2)).reshape(-1, 16).sum(1).tolist()} by co/16: {bt.sum(dim=(1, 2)).reshape(-1, 16).sum(1).tolist()}"
    print(f"RESULT {name}: {msg}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_case(sys.argv[2])
        sys.exit(0)
    for n in (sys.argv[1:] or list(CASES)):
        try:
            out = subprocess.run([sys.executable, __file__, "--one", n], capture_output=True, text=True, timeout=120)
            lines = [l for l in (out.stdout + out.stderr).splitlines() if l.startswith("RESULT") or "rror" in l or "tmf:" in l]
            print("\n".join(lines[-4:]) if lines else f"RESULT {n}: no output rc={out.returncode}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"RESULT {n}: TIMEOUT", flush=True)
