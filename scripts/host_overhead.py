"""Host-side enqueue time of one train step (no device sync inside) vs device time; cProfile of the enqueue path."""
import cProfile, pstats, sys, time, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transmf_ad_b200.models import mymodel as M
from transmf_ad_b200.synthetic import make_labels, make_volumes, procedural_state
from transmf_ad_b200 import _lib

dev = torch.device("cuda", 0)
model = M.model_CNN_ad(128)
model.load_state_dict(procedural_state(model.state_dict(), seed=0))
model = model.to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
ce_fn = torch.nn.CrossEntropyLoss()
B = 8
label = make_labels(B).to(dev)
mri = make_volumes(B, (91, 109, 91), seed=1, labels=label.cpu()).to(dev)
pet = make_volumes(B, (91, 109, 91), seed=2, labels=label.cpu()).to(dev)
ones = torch.ones(B, dtype=torch.int64, device=dev); zeros = torch.zeros(B, dtype=torch.int64, device=dev)

def step():
    opt.zero_grad()
    o = model(mri, pet)
    loss = ce_fn(o[0], label) + (ce_fn(o[1], ones) + ce_fn(o[2], zeros)) / 2
    loss.backward()
    opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
for name in ("fwd", "bwd", "opt"):
    pass
ts = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); step(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
print("host enqueue ms / total ms per step:", [(round(a, 2), round(b, 2)) for a, b in ts])
# split
torch.cuda.synchronize()
t0 = time.perf_counter(); opt.zero_grad(); t1 = time.perf_counter(); o = model(mri, pet); t2 = time.perf_counter()
loss = ce_fn(o[0], label) + (ce_fn(o[1], ones) + ce_fn(o[2], zeros)) / 2; t3 = time.perf_counter()
loss.backward(); t4 = time.perf_counter(); opt.step(); t5 = time.perf_counter()
torch.cuda.synchronize()
print("zero_grad %.2f fwd %.2f loss %.2f bwd %.2f opt %.2f (host ms, includes any back-pressure)" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)))
n0 = _lib.launch_count(); step(); print("tmf launches per step", _lib.launch_count() - n0)
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35); print(s.getvalue()[:6000])
