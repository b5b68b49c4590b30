"""2+-rank NCCL data-parallel parity worker (launched by tests/test_gpu_dp_nccl.py through torch.distributed.run).

Each rank runs ``model_CNN_ad`` / ``model_ad`` on its shard of a global batch; checks on rank 0:
  1. eager loop: gradients after ``GradBucketReducer.finish()`` == mean over ranks of the per-rank gradients (all-gathered;
     exact up to the all-reduce's summation order, 1e-6) and are aligned with the mean of the per-shard Oracle-A
     gradients (CPU oracle with bf16 rounding emulated; per-rank BatchNorm statistics = DDP semantics);
  2. timed path: k replays of the single-graph ``GraphedTrainStep`` (all-reduce captured inside, overlapped) + ``FusedAdam``
     leave the same parameters as k eager data-parallel steps + ``torch.optim.Adam``, on every rank, and all ranks agree.
Prints ``DP_NCCL_OK`` on success.
"""
import copy
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import restatement as R                                     # noqa: E402  (checker only)
from tests import helpers as H                                           # noqa: E402
from transmf_ad_b200.dp import FlatGradReducer, shard_slice            # noqa: E402
from transmf_ad_b200.models import mymodel as M                          # noqa: E402
from transmf_ad_b200.optim import FusedAdam                              # noqa: E402
from transmf_ad_b200.synthetic import make_labels, make_volumes, procedural_state   # noqa: E402
from transmf_ad_b200.train import GraphedTrainStep                       # noqa: E402


def say(msg):
    print(f"[dp-progress rank {os.environ.get('RANK')}] {msg}", flush=True)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    kind = os.environ.get("DP_MODEL", "model_ad")
    kwargs = dict(dim=128, depth=3, heads=4, dim_head=32, mlp_dim=512, dropout=0.) if kind == "model_ad" else dict(dim=128)
    per, shape = 4, (33, 36, 34)
    GB = per * world
    label = make_labels(GB)
    mri = make_volumes(GB, shape, seed=41, labels=label)
    pet = make_volumes(GB, shape, seed=42, labels=label)
    sl = shard_slice(GB, rank, world)
    state = procedural_state(getattr(M, kind)(**kwargs).state_dict(), seed=5)

    def fresh():
        m = getattr(M, kind)(**kwargs)
        m.load_state_dict(state)
        m = m.to(dev).train()
        H.set_head_dropout(m, 0.0)
        return m

    def loss_fn(outs, lab):
        ce, ad, total = H.losses(outs, lab)
        return total, ce, ad

    batch = (mri[sl].to(dev), pet[sl].to(dev), label[sl].to(dev))
    if os.environ.get("DP_MODE") == "syncbn":
        # ---- optional SyncBatchNorm: torch.nn.SyncBatchNorm.convert_sync_batchnorm(model) makes the towers
        # (our kernels) and the BatchNorm1d heads (torch) normalise over the GLOBAL batch, so `world` ranks with `per` subjects each
        # reproduce ONE device running the global batch: logits of the shard's rows, running statistics, averaged gradients.
        whole = fresh()
        outs = whole(mri.to(dev), pet.to(dev))
        loss_fn(outs, label.to(dev))[0].backward()
        synced = torch.nn.SyncBatchNorm.convert_sync_batchnorm(fresh())
        red = FlatGradReducer(model=synced).install()
        outs_s = synced(*batch[:2])
        loss_fn(outs_s, batch[2])[0].backward()
        red.finish()
        torch.cuda.synchronize()
        err_logit = max(float((a - b[sl]).abs().max()) for a, b in zip(outs_s, outs))
        sd_w, sd_s = whole.state_dict(), synced.state_dict()
        stat = [(float(((sd_s[k] - v).abs() / (v.abs() + 1e-2)).max()), k) for k, v in sd_w.items() if "running" in k]
        err_stat, worst_stat = max(stat)
        g_w = torch.cat([p.grad.flatten() for p in whole.parameters()])
        g_s = torch.cat([p.grad.flatten() for p in synced.parameters()])
        cos = float(torch.nn.functional.cosine_similarity(g_w, g_s, dim=0))
        # the same comparison WITHOUT SyncBatchNorm (per-rank statistics): how far plain data parallelism is from the global batch
        plain = fresh()
        red.remove()
        red2 = FlatGradReducer(model=plain).install()
        outs_p = plain(*batch[:2])
        loss_fn(outs_p, batch[2])[0].backward()
        red2.finish()
        torch.cuda.synchronize()
        err_plain = max(float((a - b[sl]).abs().max()) for a, b in zip(outs_p, outs))
        cos_plain = float(torch.nn.functional.cosine_similarity(g_w, torch.cat([p.grad.flatten() for p in plain.parameters()]), dim=0))
        red2.remove()
        if rank == 0:
            print(f"[dp] syncbn {kind} world {world}: vs one device on the global batch: logits {err_logit:.2e} (per-rank BN: "
                  f"{err_plain:.2e}), running statistics {err_stat:.2e} rel (worst: {worst_stat}), whole-model gradient cosine "
                  f"{cos:.4f} (per-rank BN: {cos_plain:.4f})", flush=True)
        assert err_stat <= 2e-2, (err_stat, worst_stat)
        assert err_logit <= 5e-2 and err_logit < err_plain, (err_logit, err_plain)
        assert cos >= 0.97 and cos > cos_plain, (cos, cos_plain)
        dist.barrier()
        if rank == 0:
            print("DP_NCCL_OK", flush=True)
        dist.destroy_process_group()
        return
    # ---- 1. eager: reduced gradients == mean of the ranks' own gradients, and ~ mean of per-shard Oracle-A gradients
    model = fresh()
    red = FlatGradReducer(model=model).install()
    own = fresh()                                                        # same step without any reducer
    loss_fn(own(*batch[:2]), batch[2])[0].backward()
    loss_fn(model(*batch[:2]), batch[2])[0].backward()
    red.finish()
    torch.cuda.synchronize()
    n_torch_native = sum(1 for k in (n for n, _ in model.named_parameters()) if k.split(".")[-2] in ("1", "5") and
                         (k.startswith("D.") or k.startswith("fc_cls.")))            # BatchNorm1d affine parameters
    assert red.allreduce_launches == 1
    # zero-copy: only the gradients torch's own autograd produced (BatchNorm1d) or summed (the discriminator D runs twice per
    # step, so autograd adds its two weight-gradient contributions in a buffer of its own) had to be copied into their slots
    assert red.packed_last == n_torch_native + 4, (red.packed_last, n_torch_native)
    slot_of = {id(p): s for p, s in zip(red.params, red.slots)}
    for k, p in model.named_parameters():
        if "_cnn." in k or "fuse_transformer" in k:
            assert p.grad.data_ptr() == slot_of[id(p)].data_ptr(), f"{k}: gradient was not born in its flat-buffer slot"
    names = [k for k, _ in model.named_parameters()]
    for k, p, q in zip(names, model.parameters(), own.parameters()):
        gathered = [torch.empty_like(q.grad) for _ in range(world)]
        dist.all_gather(gathered, q.grad.contiguous())
        mean = torch.stack(gathered).mean(0)
        err = float((p.grad - mean).abs().max())
        assert err <= 1e-6 + 1e-5 * float(mean.abs().max()), f"rank {rank} {k}: reduced gradient differs from the mean ({err})"
    if rank == 0:
        acc = None
        for r in range(world):                                           # Oracle-A on every shard, same weights
            s = shard_slice(GB, r, world)
            sd = R.clone_state(state)
            o = H.oracle_forward(kind, sd, (mri[s], pet[s]), kwargs, True, 0.0, rnd=R.bf16_round)
            H.losses(o, label[s])[2].backward()
            g = {k: sd[k].grad for k in names}
            acc = g if acc is None else {k: acc[k] + g[k] for k in names}
        ours = torch.cat([p.grad.detach().cpu().flatten() for k, p in zip(names, model.parameters())
                          if not H.is_conv_bias(k) and float(acc[k].norm()) >= 1e-5 * world])
        ref = torch.cat([(acc[k] / world).flatten() for k in names
                         if not H.is_conv_bias(k) and float(acc[k].norm()) >= 1e-5 * world])
        cos = H.cosine(ours, ref)
        # the same mean for the fp32 oracle: how far bf16 rounding alone moves it (BatchNorm1d over 4-sample shards amplifies)
        acc32 = None
        for r in range(world):
            s = shard_slice(GB, r, world)
            sd = R.clone_state(state)
            o = H.oracle_forward(kind, sd, (mri[s], pet[s]), kwargs, True, 0.0, rnd=None)
            H.losses(o, label[s])[2].backward()
            g = {k: sd[k].grad for k in names}
            acc32 = g if acc32 is None else {k: acc32[k] + g[k] for k in names}
        ref32 = torch.cat([(acc32[k] / world).flatten() for k in names
                           if not H.is_conv_bias(k) and float(acc[k].norm()) >= 1e-5 * world])
        cos_ab = H.cosine(ref, ref32)
        print(f"[dp] {kind} world {world}: flat buffer {red.layout()}, packed {red.packed_last}; whole-model gradient cosine vs "
              f"mean of per-shard Oracle-A gradients {cos:.4f} (Oracle-A vs fp32: {cos_ab:.4f})", flush=True)
        assert cos >= min(0.95, cos_ab - 0.03), (cos, cos_ab)
    red.remove()
    say("part 1 done")
    # ---- 2. timed path.  (a) k single-graph replays + FusedAdam == k eager data-parallel steps + FusedAdam, BITWISE (every
    # kernel and the all-reduce are deterministic);  (b) ONE step of it == one eager data-parallel step with torch.optim.Adam as
    # the reference builds it, to fp32 rounding.  Why (b) is one step: tests/test_gpu_train_path.py (two correct Adam
    # implementations drift by ~lr per element per step on this network).
    steps, lr = 3, 1e-3

    def eager_dp(opt_cls, n):
        m = fresh()
        red_ = FlatGradReducer(model=m).install()
        opt = opt_cls(m.parameters(), lr=lr)
        for _ in range(n):
            opt.zero_grad()
            loss_fn(m(*batch[:2]), batch[2])[0].backward()
            red_.finish()
            opt.step()
        red_.remove()
        torch.cuda.synchronize()
        return m

    def graphed(n):
        m = fresh()
        red_ = FlatGradReducer(model=m)
        opt = FusedAdam(m.parameters(), lr=lr)
        st = GraphedTrainStep(m, opt, loss_fn, batch[:2], batch[2], reducer=red_, warmup=3)
        for _ in range(n):
            st(batch[:2], batch[2])
        torch.cuda.synchronize()
        return m, st

    # all eager runs first, then the one captured graph (the order the bench uses: eager warm-up, capture, replays)
    say("part 2: eager FusedAdam x3")
    sd_e = eager_dp(FusedAdam, steps).state_dict()
    say("eager FusedAdam x1")
    sd_f1 = eager_dp(FusedAdam, 1).state_dict()
    say("eager torch Adam x1")
    sd_t1 = eager_dp(torch.optim.Adam, 1).state_dict()
    say("compare one step")
    for k, v in sd_t1.items():                                            # (b), via (a): graph step 1 == eager FusedAdam step 1 bitwise
        if k.endswith("num_batches_tracked"):
            continue
        err = (sd_f1[k] - v).abs()
        assert bool((err <= 2e-7 + 2e-6 * v.abs()).all()), f"rank {rank} {k}: one FusedAdam step differs from torch Adam ({float(err.max()):.3e})"
    say("graph capture + 3 replays")
    g_m, step = graphed(steps)
    say("compare graph vs eager")
    sd_g = g_m.state_dict()
    for k, v in sd_e.items():
        w = sd_g[k]
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(w) == (2 * steps if k.startswith("D.") else steps), k      # D runs twice per step
            continue
        assert torch.equal(w, v), f"rank {rank} {k}: graph path differs from eager DP with the same optimizer ({float((w - v).abs().max()):.3e})"
        if "running" not in k:                                            # replicas stay identical
            other = [torch.empty_like(w) for _ in range(world)]
            dist.all_gather(other, w.contiguous())
            assert all(torch.equal(other[0], o) for o in other), f"{k}: replicas diverged"
    dist.barrier()
    if rank == 0:
        print(f"[dp] graph path: {step.launches_per_step} libtmf launches per step, split={step.split}", flush=True)
        print("DP_NCCL_OK", flush=True)
    # a captured graph that holds NCCL kernels must be gone before the communicator is torn down (destroy_process_group()
    # waited forever with the GraphedTrainStep still alive)
    del step, g_m, sd_g
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
