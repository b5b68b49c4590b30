#!/usr/bin/env python
"""Run-to-run reproducibility stress of the inference path: N eval passes of a golden case, count passes whose logits differ
bitwise from the first.   python scripts/eval_stress.py [case] [passes]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H
from transmf_ad_b200 import evaluate as E
from transmf_ad_b200.models import mymodel as M

name = sys.argv[1] if len(sys.argv) > 1 else "model_ad_full_b8"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
gold = H.load_golden(name)
model = getattr(M, gold["kind"])(**gold["kwargs"])
model.load_state_dict(H.case_state(gold))
model = model.to("cuda").eval()
mri, pet, label = H.case_inputs(gold)
batch = {"MRI": mri, "label": label}
if gold["kind"] != "model_single":
    batch["PET"] = pet
ref = E.val_step(model, batch, "cuda")["logits"].clone()
bad, worst = 0, 0.0
junk = []
for i in range(n):
    if i % 7 == 0:                                   # churn the caching allocator: later passes see dirty memory
        junk = [torch.full((1 << 22,), float("nan"), device="cuda") for _ in range(3)]
        junk = []
    out = E.val_step(model, batch, "cuda")["logits"]
    if not torch.equal(out, ref):
        bad += 1
        worst = max(worst, float((out - ref).abs().max()))
print(f"{name}: {bad} of {n} passes differ from the first (max |diff| {worst:.3e}); env " +
      " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("TMF_")))
