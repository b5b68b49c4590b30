"""Which tensors differ between eager / graph and torch.optim.Adam / FusedAdam after one and two steps (debug aid)."""
import copy, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H
from tests.test_gpu_train_path import _model, _batches, _loss, CASES, LR
from transmf_ad_b200.optim import FusedAdam
from transmf_ad_b200.train import GraphedTrainStep

kind, kwargs, B, shape = CASES[int(sys.argv[1]) if len(sys.argv) > 1 else 0]
batches = _batches(B, shape, 3)
base = _model(kind, kwargs, seed=3)

def grads_of(m, b):
    m.zero_grad(set_to_none=True)
    _loss(m(b[0], b[1]), b[2])[0].backward()
    return {k: p.grad.clone() for k, p in m.named_parameters()}

g1 = grads_of(copy.deepcopy(base), batches[0]); g2 = grads_of(copy.deepcopy(base), batches[0])
bad = [(k, float((g1[k] - g2[k]).abs().max()), float(g1[k].abs().max())) for k in g1 if not torch.equal(g1[k], g2[k])]
print("eager run-to-run gradient differences:", bad)

def run(mode, optname, steps):
    m = copy.deepcopy(base)
    opt = FusedAdam(m.parameters(), lr=LR) if optname == "fused" else torch.optim.Adam(m.parameters(), lr=LR, capturable=(mode == "graph"))
    if mode == "graph":
        st = GraphedTrainStep(m, opt, _loss, batches[0][:2], batches[0][2], warmup=3)
        for b in batches[:steps]:
            st(b[:2], b[2])
    else:
        for b in batches[:steps]:
            opt.zero_grad(); _loss(m(b[0], b[1]), b[2])[0].backward(); opt.step()
    torch.cuda.synchronize()
    return {k: v.clone() for k, v in m.state_dict().items()}

for steps in (1, 2):
    ref = run("eager", "torch", steps)
    for mode, o in (("eager", "fused"), ("graph", "torch"), ("graph", "fused")):
        got = run(mode, o, steps)
        rows = []
        for k in ref:
            if ref[k].dtype.is_floating_point:
                d = float((ref[k] - got[k]).abs().max())
                if d > 2e-7 + 2e-6 * float(ref[k].abs().max()):
                    rows.append((k, d))
            elif int(ref[k]) != int(got[k]):
                rows.append((k, int(got[k]) - int(ref[k])))
        print(f"steps={steps} {mode}+{o}: {len(rows)} tensors differ; worst: {sorted(rows, key=lambda r: -abs(r[1]))[:8]}")
