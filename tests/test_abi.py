"""The C-ABI library builds, loads without a GPU, and exports every symbol include/tmf.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tmf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tmf_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path_entry_points():
    syms = declared_symbols()
    for must in ("tmf_conv3d_fwd", "tmf_conv3d_wgrad", "tmf_conv1_fwd", "tmf_bn_act_pool_fwd", "tmf_attn_fwd",
                 "tmf_attn_bwd", "tmf_layernorm_fwd", "tmf_linear_fwd", "tmf_scale", "tmf_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    for s in declared_symbols():
        assert hasattr(built_lib, s), f"{s} declared in include/tmf.h but not exported"


def test_python_signatures_cover_the_header(built_lib):
    from transmf_ad_b200 import _lib
    bound = set(_lib.SIGNATURES) | set(_lib.PLAIN)
    assert bound == set(declared_symbols())


def test_library_reports_version_and_error_string(built_lib):
    assert built_lib.tmf_version() >= 100
    assert isinstance(built_lib.tmf_last_error(), bytes)


def test_sass_is_sm100a_only():
    """The shared object carries sm_100a code and nothing else (no multi-arch dispatch)."""
    import subprocess
    from transmf_ad_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_streamed_weight_conv_plans_for_the_bench_layers(built_lib):
    """Host-only introspection of the generic tcgen05 forward/dgrad kernel's launch plan: at the
    bench shape (B = 8, two towers) conv3.3 dgrad applies each weight box to three 128-row tiles with two issuer warps
    and one accumulator set, conv4.0 forward / dgrad tile the stacked planes of a sample; the 1x1x1 layer keeps its
    weights resident."""
    import ctypes as C
    out = (C.c_int * 8)()

    def plan(D, H, W, cin, cout, ks, B=8, ng=2):
        assert built_lib.tmf_conv3d_umma_plan_info(ng, B, D, H, W, cin, cout, ks, out) == 0
        keys = ("ok", "mt", "issuers", "acc_sets", "a_stages", "b_stages", "tiles_per_plane", "resident")
        return dict(zip(keys, list(out)))

    # "resident" is a bit field: 1 = weights resident, 2 = plane-stack tiling, 4 = kw-fused weight boxes
    p = plan(22, 27, 22, 128, 64, 3)
    assert (p["mt"], p["issuers"], p["acc_sets"], p["resident"]) == (3, 2, 1, 0) and p["b_stages"] % 2 == 0
    # conv4.0 (planes of 13 x 13 padded rows = 1.3 tiles): plane-stack tiling, 17 tiles per sample instead of 22
    p = plan(11, 13, 11, 256, 128, 3)
    assert (p["mt"], p["acc_sets"], p["resident"], p["tiles_per_plane"]) == (1, 2, 6, 17)
    p = plan(11, 13, 11, 128, 256, 3)
    assert (p["mt"], p["issuers"], p["resident"], p["tiles_per_plane"]) == (1, 1, 2, 17)
    p = plan(11, 13, 11, 256, 128, 1)
    assert p["resident"] == 1 and p["mt"] == 1
    # every plan respects the 512-column TMEM budget
    for args in ((22, 27, 22, 128, 64, 3), (11, 13, 11, 256, 128, 3), (11, 13, 11, 128, 256, 3), (45, 54, 45, 32, 64, 3),
                 (19, 23, 19, 128, 64, 3), (32, 32, 19, 128, 64, 3)):
        for B in (1, 2, 8, 32):
            p = plan(*args, B=B)
            cout = args[4]
            assert p["ok"] == 1 and p["acc_sets"] * p["issuers"] * p["mt"] * cout <= 512
    # unsupported channel counts are refused, not mis-planned
    assert built_lib.tmf_conv3d_umma_plan_info(2, 8, 22, 27, 22, 24, 64, 3, out) != 0


def test_column_conv_plans_for_the_bench_layers(built_lib, monkeypatch):
    """Host-only: the column kernel (Cin 32 / 64) takes conv2.0 .. conv3.3 forward and the matching dgrads at the bench shape;
    conv2.3 forward (32 -> 64) runs the non-stacked 64-channel variant, everything else the
    kw-stacked 32-channel slices; TMF_COL_NS=0 switches the variant off."""
    import ctypes as C
    out = (C.c_int * 6)()

    def plan(D, H, W, cin, cout, ng=2):
        rc = built_lib.tmf_conv3d_col_plan_info(ng, D, H, W, cin, cout, 3, out)
        return rc, dict(zip(("ok", "ns", "blocks", "tiles_per_plane", "slots", "slab_rows"), list(out)))

    rc, p = plan(45, 54, 45, 32, 64)                      # conv2.3 forward
    assert rc == 0 and (p["ns"], p["blocks"], p["tiles_per_plane"]) == (1, 1, 20) and p["slots"] >= 4
    rc, p = plan(45, 54, 45, 32, 32)                      # conv2.0 forward / dgrad
    assert rc == 0 and (p["ns"], p["blocks"], p["tiles_per_plane"]) == (0, 1, 20)
    rc, p = plan(45, 54, 45, 64, 32)                      # conv2.3 dgrad
    assert rc == 0 and (p["ns"], p["blocks"]) == (0, 1)
    rc, p = plan(22, 27, 22, 64, 128)                     # conv3.3 forward: four 32-channel slices per tower
    assert rc == 0 and (p["ns"], p["blocks"]) == (0, 4)
    assert plan(22, 27, 22, 128, 64)[0] != 0              # Cin = 128: the streamed-weight kernel's layer
    monkeypatch.setenv("TMF_COL_NS", "0")
    rc, p = plan(45, 54, 45, 32, 64)
    assert rc == 0 and (p["ns"], p["blocks"]) == (0, 2)
