"""Bring-up of the tcgen05 conv kernel: compare against the CUDA-core implementation, one subprocess per case so that a
trap / hang in one variant does not poison the others.  usage: python scripts/umma_bringup.py [case ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (B, D, H, W, cin, cout, ks)
    "k1_aligned": (1, 2, 8, 16, 64, 64, 1),
    "k1_c256": (2, 5, 6, 5, 256, 128, 1),
    "k3_c64": (1, 4, 10, 14, 64, 64, 3),
    "k3_c32": (1, 4, 10, 14, 32, 32, 3),
    "k3_c32_64": (2, 5, 12, 45, 32, 64, 3),
    "k3_c64_32": (2, 5, 12, 45, 64, 32, 3),
    "k3_c128_256": (2, 5, 13, 11, 128, 256, 3),
    "k3_c256_128": (2, 5, 13, 11, 256, 128, 3),
    "k3_c64_128": (1, 6, 27, 22, 64, 128, 3),
}


def run_case(name):
    import torch
    from transmf_ad_b200 import _lib as L
    B, D, H, W, cin, cout, ks = CASES[name]
    g = torch.Generator().manual_seed(0)
    a = torch.randn(B, D, H, W, cin, generator=g).to(torch.bfloat16).cuda()
    taps = ks ** 3
    wf = (torch.randn(taps, cout, cin, generator=g) * (2.0 / (cin * taps)) ** 0.5).to(torch.bfloat16).cuda()
    bias = (0.1 * torch.randn(cout, generator=g)).cuda()
    y0 = torch.empty(B, D, H, W, cout, dtype=torch.bfloat16, device="cuda")
    y1 = torch.full_like(y0, float("nan"))
    s0 = torch.empty(2 * cout, dtype=torch.float64, device="cuda")
    s1 = torch.empty_like(s0)
    L.call("tmf_conv3d_fwd", 1, L.ptrs([a]), L.ptrs([wf]), L.ptrs([bias]), L.ptrs([y0]), L.ptrs([s0]), B, D, H, W, cin, cout, ks, L.CONV_DIRECT)
    L.call("tmf_conv3d_fwd", 1, L.ptrs([a]), L.ptrs([wf]), L.ptrs([bias]), L.ptrs([y1]), L.ptrs([s1]), B, D, H, W, cin, cout, ks, L.CONV_UMMA)
    torch.cuda.synchronize()
    d = (y1.float() - y0.float()).abs()
    scale = float(y0.float().abs().max())
    bad = d > 0.02 * scale
    nan = int(torch.isnan(y1.float()).sum())
    srel = float(((s1 - s0).abs() / s0.abs().clamp_min(1e-3)).max())
    msg = f"max|d|/scale={float(d[~torch.isnan(d)].max()) / scale if nan < d.numel() else -1:.4g} bad={int(bad.sum())}/{d.numel()} nan={nan} stats_rel={srel:.3g}"
    if int(bad.sum()) or nan:
        bw = bad.any(dim=-1) | torch.isnan(y1.float()).any(dim=-1)      # (B,D,H,W)
        msg += f" bad voxels by w: {bw.sum(dim=(0, 1, 2)).tolist()[:48]} by h: {bw.sum(dim=(0, 1, 3)).tolist()} by d: {bw.sum(dim=(0, 2, 3)).tolist()}"
    print(f"RESULT {name} bo={os.environ.get('TMF_UMMA_BO', '1')}: {msg}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_case(sys.argv[2])
        sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    modes = os.environ.get("BO_MODES", "1,0,2").split(",")
    for n in names:
        for bo in modes:
            env = dict(os.environ, TMF_UMMA_BO=bo)
            try:
                out = subprocess.run([sys.executable, __file__, "--one", n], env=env, capture_output=True, text=True, timeout=120)
                lines = [l for l in (out.stdout + out.stderr).splitlines() if l.startswith("RESULT") or "rror" in l or "tmf:" in l]
                print("\n".join(lines[-4:]) if lines else f"RESULT {n} bo={bo}: no output rc={out.returncode}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"RESULT {n} bo={bo}: TIMEOUT", flush=True)
