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
