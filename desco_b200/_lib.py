"""Loader / builder of the C-ABI shared library ``desco_b200/_C/libdesco_b200.so`` (declared in include/desco_b200.h).

There is NO CPU fallback: if the library cannot be loaded, every product entry point raises.  ``build()`` compiles all
``csrc/*.cu`` for sm_100a with nvcc (works without a GPU; the .so is git-ignored but travels to the GPU box).
"""
from __future__ import annotations

import ctypes
import glob
import os
import shutil
import subprocess
from typing import List, Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_OUT_DIR = os.path.join(_HERE, "_C")
LIB_PATH = os.path.join(_OUT_DIR, "libdesco_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wno-deprecated-declarations", "-Wno-deprecated-declarations",
]
# no --use_fast_math: the fp32 parity path keeps IEEE division / denormals; fast intrinsics are chosen per call site.

_lib: Optional[ctypes.CDLL] = None


def _sources() -> List[str]:
    return sorted(glob.glob(os.path.join(_CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _sources() + glob.glob(os.path.join(_CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE_DIR, "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into one shared library (object per source, then link)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build desco_b200 CUDA library")
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(_OUT_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(_OUT_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(
            os.path.getmtime(p) for p in [src] + glob.glob(os.path.join(_CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE_DIR, "*.h"))
        ):
            continue
        cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE_DIR, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out)
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB_PATH


_I32P = ctypes.c_void_p  # device pointers are passed as integers
_VP = ctypes.c_void_p
_I = ctypes.c_int32
_L = ctypes.c_int64
_F = ctypes.c_float

# name -> (restype, argtypes); must list every symbol of include/desco_b200.h (tests/test_cabi.py checks this)
SIGNATURES = {
    "desco_version": (ctypes.c_char_p, []),
    "desco_kernel_launches": (_L, []),
    "desco_profile_enable": (_I, [_I]),
    "desco_profile_read": (_I, [_VP, _VP]),
    "desco_partition_count": (_I, [_VP, _VP, _VP, _I, _VP, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP]),
    "desco_partition_scan_workspace_bytes": (_L, [_I]),
    "desco_partition_scan": (_I, [_VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _L, _VP]),
    "desco_partition_fill": (_I, [_VP, _VP, _VP, _I, _VP, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "desco_partition_batch_workspace_bytes": (_L, [_I]),
    "desco_partition_batch": (_I, [_VP, _VP, _VP, _I, _VP, _I, _I, _I, _I, _VP, _L, _VP, _VP, _VP, _VP, _VP, _VP, _L, _VP, _VP, _L, _VP, _VP]),
    "desco_partition_batch_async": (_I, [_VP, _VP, _VP, _I, _VP, _I, _I, _I, _I, _VP, _L, _VP, _VP, _VP, _VP, _VP, _VP, _L, _VP, _VP, _L, _VP, _VP]),
    "desco_partition_large_workspace_bytes": (_L, [_I, _I]),
    "desco_partition_large_count": (_I, [_VP, _VP, _VP, _I, _VP, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _L, _VP]),
    "desco_partition_large_fill": (_I, [_VP, _VP, _VP, _I, _VP, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _L, _VP]),
    "desco_partition_large_set_caps": (_I, [_I, _I, _I, _I, _I, _I]),
    "desco_partition_large_phase_cycles": (_I, [_VP, _I]),
    "desco_shmp_edge_types": (_I, [_VP, _VP, _I, _VP, _VP]),
    "desco_shmp_workspace_bytes": (_L, [_I, _I, _I]),
    "desco_shmp_layer_weight_floats": (_L, []),
    "desco_shmp_tc_layer_bytes": (_L, []),
    "desco_shmp_forward": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _VP, _I, _VP, _VP, _VP, _VP, _VP, _I, _I, _VP, _VP, _L, _I, _VP, _VP]),
    "desco_shmp_forward_dev": (_I, [_VP, _VP, _VP, _VP, _I, _I, _VP, _I, _VP, _I, _VP, _VP, _VP, _VP, _I, _I, _VP, _VP, _L, _I, _VP, _VP]),
    "desco_count_head_dev": (_I, [_VP, _I, _VP, _VP, _I, _VP, _I, _VP, _VP, _VP, _L, _VP]),
    "desco_shmp_mt_layer_bytes": (_L, []),
    "desco_shmp_forward_mt": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _VP, _I, _VP, _VP, _VP, _VP, _VP, _I, _I, _VP, _VP, _L, _I, _VP, _VP]),
    "desco_count_head_workspace_bytes": (_L, [_I, _I]),
    "desco_count_head": (_I, [_VP, _I, _VP, _I, _VP, _VP, _I, _VP, _VP, _VP, _L, _I, _VP, _VP]),
    "desco_gossip_tc_phase_cycles": (_I, [_VP, _I]),
    "desco_gossip_weight_floats": (_L, []),
    "desco_gossip_query_weight_floats": (_L, []),
    "desco_gossip_workspace_bytes": (_L, [_I, _I]),
    "desco_gossip_prepare_queries": (_I, [_VP, _I, _VP, _VP, _VP, _VP]),
    "desco_gossip_layer0": (_I, [_VP, _VP, _I, _I, _VP, _I, _VP, _VP, _VP]),
    "desco_gossip_layer1": (_I, [_VP, _VP, _I, _I, _VP, _I, _VP, _VP, _VP, _I, _VP, _L, _VP]),
    "desco_gossip_layer0_grouped": (_I, [_VP, _VP, _I, _I, _VP, _I, _VP, _VP, _I, _L, _VP]),
    "desco_gossip_layer1_group": (_I, [_VP, _VP, _I, _I, _VP, _I, _I, _VP, _VP, _VP, _I, _I, _VP, _L, _VP]),
    "desco_gossip_layer1_workspace_bytes": (_L, [_I, _I, _I]),
    "desco_groundtruth_count": (_I, [_VP, _VP, _VP, _I, _I, _VP, _VP, _VP, _I, _I, _VP, _VP, _VP]),
    "desco_gossip_gated_mix": (_I, [_VP, _VP, _VP, _VP, _L, _VP]),
    "desco_gossip_gate_grad": (_I, [_VP, _VP, _VP, _L, _VP, _VP]),
    "desco_gossip_gate_backward": (_I, [_VP, _I, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "desco_train_dropout": (_I, [_VP, _VP, _F, _L, _VP]),
    "desco_gossip_loss": (_I, [_VP, _I, _VP, _I, _VP, _I, _I, _VP, _VP, _I, _VP, _VP]),
    "desco_spmm_sum": (_I, [_VP, _VP, _VP, _I, _VP, _I, _I, _VP, _I, _VP]),
    "desco_gossip_gate": (_I, [_VP, _I, _I, _VP, _VP, _I, _VP, _VP, _VP, _VP]),
    "desco_shmp_fused_phase_cycles": (_I, [_VP, _I]),
    "desco_train_plan": (_I, [_VP, _I, _I, _I, _VP, _VP, _VP, _VP]),
    "desco_train_aggregate": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _VP, _I, _VP, _I, _VP, _I, _VP, _I, _VP]),
    "desco_train_dense": (_I, [_VP, _VP, _VP, _VP, _I, _I, _VP, _I, _VP, _I, _I, _I, _I, _F, _I, _VP]),
    "desco_train_wgrad": (_I, [_VP, _I, _I, _VP, _I, _I, _I, _VP, _VP, _VP, _I, _VP]),
    "desco_train_act_backward": (_I, [_VP, _I, _VP, _I, _I, _I, _I, _F, _VP]),
    "desco_train_fill_rows": (_I, [_VP, _I, _I, _I, _VP, _VP]),
    "desco_train_colsum": (_I, [_VP, _I, _I, _I, _VP, _VP]),
    "desco_train_pool": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _VP, _I, _VP, _I, _VP, _I, _VP]),
    "desco_train_head_loss": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _VP, _VP, _VP, _VP]),
    "desco_train_head_backward": (_I, [_VP, _VP, _VP, _VP, _I, _I, _VP, _VP, _VP, _VP, _VP]),
    "desco_train_adam": (_I, [_VP, _VP, _VP, _VP, _L, _F, _F, _F, _F, _F, _I, _VP]),
    "desco_tc_selftest": (_I, [_VP, _VP, _I, _I, _VP, _VP, _VP]),
    "desco_gossip_forward": (_I, [_VP, _VP, _I, _VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _L, _I, _VP]),
}


def load(auto_build: bool = True) -> ctypes.CDLL:
    """Return the loaded library; raises (never falls back) when it is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if auto_build and _stale() and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        build()
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"desco_b200: CUDA library {LIB_PATH} is missing and could not be built; there is no CPU fallback "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class DescoError(RuntimeError):
    code = 0


_ERR = {-22: "EINVAL (bad argument)", -12: "ENOMEM", -5: "ECUDA (CUDA runtime error)", -34: "ERANGE (size limit exceeded)",
        -105: "ENOBUFS (output capacity too small)"}


ENOBUFS = -105
ERANGE = -34


def check(code: int, what: str) -> None:
    if code != 0:
        err = DescoError(f"{what} failed: {code} {_ERR.get(code, '')}")
        err.code = code
        raise err
