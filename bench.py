#!/usr/bin/env python
"""bench.py -- train subjects/s (MRI+PET pairs) of the TransMF_AD hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--workload cnn_ad|ad|single] [--batch B] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = the caller's train_step (reference kfold_train_adversarial.py:101-136): forward, CE + adversarial
losses, backward, (gradient all-reduce for N > 1) and the Adam step of utils/utils.py:38-41, on one batch of
synthetic paired volumes of shape (B,1,91,109,91).  Weak scaling: B per GPU is fixed.  Rank 0 prints ONE JSON line.

``value``   device-timed (CUDA events, max over ranks) with the batch already resident in HBM.
``e2e``     same step driven from PINNED HOST buffers: H2D copy of MRI/PET/labels and the two ``.item()`` loss
            reads (D2H) inside the timed region.
``roofline``     conv3d kernels (fwd + dgrad + wgrad, algorithmic FLOPs of SURVEY.md section 8d) timed live with CUDA
                 events around each launch in a separate profiled pass, against the measured bf16 peak.
``cpu_baseline`` the CPU oracle port (oracle/restatement.py; the reference is pure PyTorch) timed on the host cores
                 on a bounded sample (B=2) of the same workload.
``--impl reference`` times that CPU port alone and prints the same line shape with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPE = (91, 109, 91)
METRIC = "train subjects/s (MRI+PET pairs)"
WORKLOADS = {
    # name: (model class, ctor kwargs, towers, description)
    "cnn_ad": ("model_CNN_ad", dict(dim=128), 2, "model_CNN_ad (--model CNN) adversarial train step"),
    "ad": ("model_ad", dict(dim=128, depth=3, heads=4, dim_head=32, mlp_dim=512, dropout=0.), 2,
           "model_ad (--model Transformer) adversarial train step"),
    "single": ("model_single", dict(dim=128), 1, "model_single (kfold_train_single) train step"),
}


def conv_flops_per_subject(shape, dim=128, towers=2, first_layer=True, other_layers=True):
    """Algorithmic conv FLOPs: 2*M*Cout*Cin*k^3 for fwd and wgrad of every layer and dgrad of all but conv1.0.
    ``first_layer`` / ``other_layers`` select conv1.0 (Cin = 1, HBM-bound) and conv2.0 .. conv4.3 (tensor-bound)."""
    from transmf_ad_b200.functional import SNetSpec
    D, H, W = shape
    total = 0.0
    for l, (cin, cout, ks, pool) in enumerate(SNetSpec(dim).layers):
        f = 2.0 * D * H * W * cout * cin * ks ** 3
        if (l == 0 and first_layer) or (l > 0 and other_layers):
            total += f * (3 if l > 0 else 2)
        if pool:
            D, H, W = D // 2, H // 2, W // 2
    return total * towers


def block1_bytes_per_subject(shape, dim=128, towers=2):
    """Algorithmic HBM bytes of the four block-1 passes: x fp32, y bf16 (dim/4 channels),
    pooled activation / its gradient bf16."""
    D, H, W = shape
    c = dim // 4
    x = 4.0 * D * H * W
    y = 2.0 * D * H * W * c
    pooled = 2.0 * (D // 2) * (H // 2) * (W // 2) * c
    from transmf_ad_b200.functional import keep_ymax
    keep = keep_ymax()           # forward also writes ymax (pooled extent); the backward reduction reads ymax + dout only
    return {"tmf_conv1_fwd": towers * (x + y), "tmf_bn_act_pool_fwd@L0": towers * (y + pooled * (2 if keep else 1)),
            "tmf_bn_act_pool_bwd_reduce@L0": towers * ((2 * pooled) if keep else (y + pooled)),
            "tmf_conv1_bwd_fused": towers * (y + pooled + x)}


def committed_traffic():
    """DRAM bytes per step of the tensor-bound conv launches from the committed `ncu --set full` capture
    (profiles/conv_traffic.json, written by scripts/ncu_traffic.py); None when absent."""
    path = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------------
def cpu_port_step_fn(workload, batch):
    """The CPU oracle port of the same train step (fwd + losses + bwd + Adam) on a bounded sample."""
    from oracle import restatement as R
    from transmf_ad_b200.models import mymodel as M
    from transmf_ad_b200.synthetic import make_labels, make_volumes, procedural_state
    kind, kwargs, towers, _ = WORKLOADS[workload]
    template = getattr(M, kind)(**kwargs).state_dict()
    sd = R.clone_state(procedural_state(template, seed=0))
    label = make_labels(batch)
    mri = make_volumes(batch, SHAPE, seed=1, labels=label)
    pet = make_volumes(batch, SHAPE, seed=2, labels=label)
    params = [v for k, v in sd.items() if v.is_floating_point() and v.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4, weight_decay=0.0)
    heads = kwargs.get("heads", 4)

    def step():
        opt.zero_grad()
        if kind == "model_ad":
            outs = R.model_ad_forward(sd, mri, pet, heads=heads, training=True)
        elif kind == "model_CNN_ad":
            outs = R.model_cnn_ad_forward(sd, mri, pet, training=True)
        else:
            outs = (R.model_single_forward(sd, mri, training=True),)
        if len(outs) == 3:
            ce, ad, total = R.adversarial_losses(*outs, label)
        else:
            total = torch.nn.functional.cross_entropy(outs[0], label)
        total.backward()
        opt.step()
        return float(total)

    return step


def time_cpu_port(workload, batch, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_port_step_fn(workload, batch)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return batch / min(ts), sum(ts) / len(ts)


def run_reference_arm(args, rank):
    if rank != 0:
        return
    sample_b = 2
    kind, kwargs, towers, desc = WORKLOADS[args.workload]
    steps = max(1, min(args.steps, 5))
    warm = 1
    sps, mean_s = time_cpu_port(args.workload, sample_b, steps, warm)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "subjects/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{desc}, batch {args.batch}/GPU, volumes 91x109x91", "timed_on": "host CPU"},
        "cpu_baseline": {"value": sps, "unit": "subjects/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps of batch {sample_b} (best step), fwd+bwd+Adam, fp32, torch CPU "
                                   f"{torch.__version__}; oracle/restatement.py (the reference is pure PyTorch)"},
        "e2e": {"value": sps, "unit": "subjects/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cnn_ad", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=8, help="subjects per GPU per step")
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"],
                    help="graph: the step replayed as CUDA graphs (transmf_ad_b200.train.GraphedTrainStep); "
                         "eager: the reference's Python loop, one launch at a time")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"],
                    help="fused: transmf_ad_b200.optim.FusedAdam (one launch); torch: torch.optim.Adam as the reference builds it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch.distributed as dist
    from transmf_ad_b200 import _lib
    from transmf_ad_b200.dp import GradBucketReducer
    from transmf_ad_b200.models import mymodel as M
    from transmf_ad_b200.synthetic import make_labels, make_volumes, procedural_state

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"[bench] warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    kind, kwargs, towers, desc = WORKLOADS[args.workload]
    B = args.batch
    model = getattr(M, kind)(**kwargs)
    model.load_state_dict(procedural_state(model.state_dict(), seed=0))     # identical replicas on every rank
    model = model.to(dev).train()
    graph_mode = args.mode == "graph"
    if args.optimizer == "fused":                                            # utils/utils.py:38-41 (Adam branch), one launch
        from transmf_ad_b200.optim import FusedAdam
        opt = FusedAdam(model.parameters(), lr=1e-4, weight_decay=0.0)
    else:
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=0.0, capturable=graph_mode)
    reducer = GradBucketReducer(model.parameters())
    ce_fn = torch.nn.CrossEntropyLoss()
    ones = torch.ones(B, dtype=torch.int64, device=dev)
    zeros = torch.zeros(B, dtype=torch.int64, device=dev)

    # synthetic batches: a small pool, distinct per rank (seed offsets as in SURVEY.md section 8d)
    npool = 2
    label_h = make_labels(B)
    pool_h = []
    for i in range(npool):
        mri = make_volumes(B, SHAPE, seed=1 + 1000 * rank + i, labels=label_h).pin_memory()
        pet = make_volumes(B, SHAPE, seed=2 + 1000 * rank + i, labels=label_h).pin_memory()
        pool_h.append((mri, pet, label_h.pin_memory()))
    pool_d = [tuple(t.to(dev) for t in b) for b in pool_h]

    def losses(outs, label):
        if len(outs) == 3:                                    # kfold_train_adversarial.py:119-131
            ce = ce_fn(outs[0], label)
            ad = (ce_fn(outs[1], ones) + ce_fn(outs[2], zeros)) / 2
            return ce, ad, ad + ce
        ce = ce_fn(outs[0], label)                           # kfold_train_single.py:105
        return ce, None, ce

    def train_step(batch, read_losses):                      # the reference's eager loop
        mri, pet, label = batch
        opt.zero_grad()
        outs = model(mri) if towers == 1 else model(mri, pet)
        outs = outs if isinstance(outs, tuple) else (outs,)
        ce, ad, total = losses(outs, label)
        vals = None
        if read_losses:                                      # the reference's two .item() syncs (:127-128)
            vals = (ce.item(), ad.item() if ad is not None else 0.0)
        total.backward()
        reducer.finish()
        opt.step()
        return vals

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def inputs_of(batch):
        return batch[:1] if towers == 1 else batch[:2]

    graphed = None
    if graph_mode:
        from transmf_ad_b200.train import DevicePrefetcher, GraphedTrainStep, LossReader

        def graph_loss(outs, label):                         # (total, ce, ad): total.backward() is captured
            ce, ad, total = losses(outs, label)
            return (total, ce) if ad is None else (total, ce, ad)

        for i in range(args.warmup):                         # eager warm-up (lazy inits, allocator, attribute calls)
            train_step(pool_d[i % npool], False)
        barrier()
        graphed = GraphedTrainStep(model, opt, graph_loss, inputs_of(pool_d[0]), pool_d[0][2], reducer=reducer,
                                   warmup=args.warmup)

        def dev_step(i):
            b = pool_d[i % npool]
            graphed(inputs_of(b), b[2])
    else:
        def dev_step(i):
            train_step(pool_d[i % npool], False)

    # ---- warm-up, then the device-resident timed region ---------------------------------------------------------
    for i in range(args.warmup):
        dev_step(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    total_ms = timed(dev_step, args.steps)
    launches = (graphed.launches_per_step * args.steps) if graph_mode else (_lib.launch_count() - n0)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- end-to-end: pinned host buffers, H2D inside the timed region, loss .item() reads ----------------------------
    if graph_mode:
        # batch i+1 is copied from pinned host memory on a side stream while step i runs (DevicePrefetcher); every
        # step's inputs cross PCIe inside the timed region and both loss values are read back each step.
        reader = LossReader(2 if towers > 1 else 1, dev)
        e2e_state = {}

        def e2e_run(steps):
            def host_batches():
                for i in range(steps):
                    hb = pool_h[i % npool]
                    yield (hb[0], hb[2]) if towers == 1 else hb
            # every step's two loss values cross to the host inside the timed region, but the host blocks on step
            # i's values only after step i+1 has been enqueued (LossReader), so the graph launch, the H2D of the next
            # batch and the device work of the current one overlap instead of serialising on .item().
            vals = None
            pf = DevicePrefetcher(host_batches(), dev, reuse=e2e_state.get("pf"))
            e2e_state["pf"] = pf                             # the next "epoch" reuses its stream and device buffers
            for db in pf:
                out = graphed(db[:-1], db[-1])
                vals = reader.push(out[1:])
            vals = reader.flush()
            return vals

        e2e_run(2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_run(args.steps)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        e2e_ms = float(ms) / args.steps
    else:
        def e2e_step(i):
            hb = pool_h[i % npool]
            db = tuple(t.to(dev, non_blocking=True) for t in hb)
            train_step(db, True)

        for i in range(2):
            e2e_step(i)
        e2e_ms = timed(e2e_step, args.steps) / args.steps
    h2d = sum(t.numel() * t.element_size() for t in pool_h[0][: (1 if towers == 1 else 2)]) + pool_h[0][2].numel() * 8
    # the PCIe leg on its own (pinned host -> device, nothing else running): when e2e ~= this, the step is link-bound
    probe_dst = [torch.empty_like(t, device=dev) for t in pool_h[0][: (1 if towers == 1 else 2)]]
    h2d_ms_alone = []
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for d, hsrc in zip(probe_dst, pool_h[0]):
            d.copy_(hsrc, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        h2d_ms_alone.append(e0.elapsed_time(e1))
    del probe_dst
    e2e = {"value": world * B / (e2e_ms * 1e-3), "unit": "subjects/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": (8 if towers > 1 else 4) * world,   # whole job
           "h2d_ms_alone": round(min(h2d_ms_alone), 4), "h2d_gbs_alone": round(h2d / min(h2d_ms_alone) / 1e6, 2),      # one rank's link
           "h2d_bytes_per_rank": int(h2d)}

    # ---- roofline leg: CUDA events around every C-ABI launch (separate pass; not part of `value`) --------------------
    roofline, breakdown = None, None
    if not args.no_roofline:
        peaks = measured_peaks()
        _lib.TIMER.start()
        nprof = 2
        for i in range(nprof):
            train_step(pool_d[i % npool], False)
        rec = _lib.TIMER.stop()
        # dominant kernel family: the tcgen05 implicit-GEMM convolutions conv2.0 .. conv4.3 (fwd, dgrad, wgrad), the
        # layers SURVEY.md section 8d puts under the tensor roofline.  conv1.0 (Cin = 1, 26 FLOP/B) is HBM-bound and
        # is reported against the copy bandwidth below, together with the other block-1 passes.
        gemm_tags = ("tmf_conv3d_fwd", "tmf_conv3d_dgrad", "tmf_conv3d_wgrad")
        conv1_tags = ("tmf_conv1_fwd", "tmf_conv1_wgrad", "tmf_conv1_bwd_fused")
        gemm_ms = sum(v[0] for t, v in rec.items() if t.split("@")[0] in gemm_tags) / nprof
        conv1_ms = sum(v[0] for t, v in rec.items() if t.split("@")[0] in conv1_tags) / nprof
        flops = conv_flops_per_subject(SHAPE, kwargs["dim"], towers, first_layer=False) * B
        flops_all = conv_flops_per_subject(SHAPE, kwargs["dim"], towers) * B
        achieved = flops / (gemm_ms * 1e-3) / 1e12
        achieved_all = flops_all / ((gemm_ms + conv1_ms) * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        traffic = committed_traffic()
        hbm = {}
        for tag, nbytes in block1_bytes_per_subject(SHAPE, kwargs["dim"], towers).items():
            if tag in rec:
                ms = rec[tag][0] / nprof
                gbs = nbytes * B / (ms * 1e-3) / 1e9
                hbm[tag] = {"ms": round(ms, 4), "algorithmic_mb": round(nbytes * B / 1e6, 1), "achieved_gbs": round(gbs, 1),
                            "frac_of_hbm_peak": round(gbs / peaks["hbm_gbs"], 3)}
        roofline = {"bound": "tensor",
                    "kernel": "conv3d implicit GEMM on tcgen05: fwd + dgrad + wgrad of conv2.0 .. conv4.3, both towers "
                              "(18 launches per step)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic["dram_bytes_per_step"] if traffic else None,
                    "traffic_source": traffic["source"] if traffic else None,
                    "peak_source": f"{peaks['source']} bf16 sustained (kernels timed inside the step)",
                    "frac_of_burst_peak": achieved / peaks["bf16_tflops"],
                    "algorithmic_gflop_per_step": flops / 1e9, "conv_ms_per_step": gemm_ms,
                    "all_conv_incl_conv1": {"achieved": achieved_all, "frac": achieved_all / peak,
                                            "algorithmic_gflop_per_step": flops_all / 1e9,
                                            "ms_per_step": gemm_ms + conv1_ms,
                                            "note": "conv1.0 fwd and the fused block-1 backward (BN/LeakyReLU/MaxPool "
                                                    "backward + conv1.0 wgrad) are HBM-bound passes; their whole time is "
                                                    "counted here"},
                    "hbm_bound_block1": {"peak_gbs": peaks["hbm_gbs"], "kernels": hbm}}
        breakdown = {t: round(rec[t][0] / nprof, 4) for t in sorted(rec)}          # ms per step per entry point (@L = layer)

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sps, mean_s = time_cpu_port(args.workload, 2, 3, 1)
        cpu_baseline = {"value": sps, "unit": "subjects/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": "3 steps of batch 2 (best step) of the same workload, fwd+bwd+Adam, fp32 torch CPU; "
                                  "oracle/restatement.py (the reference is pure PyTorch, its kernels are ATen's)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "subjects/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{desc}, batch {B}/GPU, volumes 91x109x91 fp32, Adam lr 1e-4",
                       "global_batch": B * world, "parallelism": f"dp{world}",
                       "conv_impl": os.environ.get("TMF_CONV_IMPL", "auto"),
                       "optimizer": "FusedAdam (tmf_adam_step)" if args.optimizer == "fused" else "torch.optim.Adam",
                       "mode": ("CUDA-graph replay of the whole step (transmf_ad_b200.train.GraphedTrainStep); e2e adds "
                                "DevicePrefetcher (H2D of batch i+1 on a side stream) and LossReader (losses of step i "
                                "read on the host after step i+1 is enqueued)") if graph_mode else "eager launches",
                       "l2": "per-step working set (~0.2 GB/subject of activations) >> 126 MB L2; inputs rotate over a pool"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "breakdown": breakdown,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
