#!/usr/bin/env python
"""First eval pass (BatchNorm folds computed in this pass) vs second pass (folds cached) of freshly built models, R rounds."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H
from transmf_ad_b200 import evaluate as E
from transmf_ad_b200 import _lib as L
from transmf_ad_b200.models import mymodel as M

name = sys.argv[1] if len(sys.argv) > 1 else "model_ad_full_b8"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 60
gold = H.load_golden(name)
mri, pet, label = H.case_inputs(gold)
batch = {"MRI": mri, "label": label}
if gold["kind"] != "model_single":
    batch["PET"] = pet
bad, cnt_bad, worst = 0, 0, 0.0
for r in range(R):
    model = getattr(M, gold["kind"])(**gold["kwargs"])
    model.load_state_dict(H.case_state(gold))
    for k, v in gold["buffers_after"].items():
        model.state_dict()[k].copy_(v)
    model = model.to("cuda").eval()
    if r % 3 == 0:
        junk = [torch.full((1 << 22,), float("nan"), device="cuda") for _ in range(3)]
        del junk
    n0 = L.launch_count()
    a = E.val_step(model, batch, "cuda")["logits"].clone()
    n1 = L.launch_count()
    b = E.val_step(model, batch, "cuda")["logits"]
    n2 = L.launch_count()
    if not torch.equal(a, b):
        bad += 1
        worst = max(worst, float((a - b).abs().max()))
    if not (n2 - n1) < (n1 - n0):
        cnt_bad += 1
        print("launch counts", n1 - n0, n2 - n1)
print(f"{name}: {bad} of {R} rounds with pass 1 != pass 2 (max |diff| {worst:.3e}); {cnt_bad} rounds with unexpected launch counts")
