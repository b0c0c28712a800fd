"""CPU: the C-ABI library builds, loads, and exports exactly the symbols include/desco_b200.h declares (no compute)."""
import ctypes
import os
import re

from desco_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "desco_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(desco_[a-z0-9_]+)\s*\(", text)))


def test_header_and_loader_agree():
    syms = _header_symbols()
    assert len(syms) >= 20
    assert sorted(_lib.SIGNATURES) == syms


def test_library_exports_every_header_symbol():
    _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _header_symbols():
        assert hasattr(lib, name), name
    loaded = _lib.load(auto_build=False)
    assert loaded.desco_version().decode().startswith("desco_b200")
    assert loaded.desco_kernel_launches() >= 0
    for fn, args in (("desco_shmp_layer_weight_floats", ()), ("desco_gossip_weight_floats", ()),
                     ("desco_gossip_query_weight_floats", ()), ("desco_shmp_workspace_bytes", (1000, 100, 8)),
                     ("desco_gossip_workspace_bytes", (1000, 29)), ("desco_count_head_workspace_bytes", (100, 29))):
        assert getattr(loaded, fn)(*args) > 0  # pure host arithmetic: safe without a GPU


def test_sass_contains_blackwell_tensor_core_ops():
    """The built library must carry tcgen05 (UTC*MMA), TMEM loads (LDTM) and bulk async copies (UBLKCP) in SASS."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest

        pytest.skip("cuobjdump not available")
    _lib.build()
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP"):
        assert mnemonic in sass, mnemonic


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "desco_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("the fp32 oracle", ""), f
