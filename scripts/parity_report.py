"""Dump the per-tensor parity metrics of every golden case (run on the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H

names = sys.argv[1:] or ["model_single", "model_cnn", "model_cnn_ad", "model_ad_h4", "model_ad_h8", "model_transformer",
                         "model_transformer_res", "model_ad_dim64"]
for n in names:
    r = H.parity_report(n)
    print("=" * 100)
    for k, v in r.items():
        if k == "grads":
            print("  grads (worst rel_A first):")
            items = sorted(v.items(), key=lambda kv: -(kv[1].get("rel_A", 9) if not kv[1].get("conv_bias") else -1))
            for kk, e in items[:14]:
                print(f"    {kk:60s} rel_A={e.get('rel_A', -1):.3g} cos_A={e.get('cos_A', -1):.4f} cos_B={e.get('cos_B', -1):.3f} "
                      f"cosA:B={e.get('cos_A_vs_B', -1):.3f} norm={e.get('norm', -1):.3g}/{e.get('norm_A', -1):.3g}/{e.get('norm_B', -1):.3g}")
            cb = [e["absmax"] for e in v.values() if e.get("conv_bias")]
            print(f"    conv-bias grads absmax: {max(cb):.3g}; non-finite: {[kk for kk, e in v.items() if not e.get('finite', True)]}")
        elif k == "buffers":
            vals = [x for x in v.values() if isinstance(x, float)]
            print(f"  buffers: worst rel {max(vals):.3g}; nbt ok {all(x for x in v.values() if isinstance(x, bool))}")
        elif callable(v):
            print(f"  {k}(2e-2): {v(2e-2)}")
        else:
            print(f"  {k}: {v}")
