"""ctypes binding of ``librabi_b200.so`` (C ABI declared in include/rabi_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails, an exception is raised.
``build()`` compiles the library in-tree with nvcc for sm_100a (cross-compiles without a GPU).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "librabi_b200.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
# translation units and the headers each one depends on (an object is rebuilt when one of them is newer)
_COMMON = ["rabi_internal.h", "maf_image.h", "maf_tc.h"]
# object -> (source, extra defines, headers)
SOURCES = {
    "rabi_maf": ("rabi_maf.cu", [], _COMMON + ["maf_pairs.cuh", "maf_fast.cuh", "maf_fast2.cuh", "maf_deep.cuh", "maf_warp.cuh"]),
    "maf_tc": ("maf_tc.cu", [], _COMMON),                               # host code + the DP = 4 kernels
    "maf_tc_dp8": ("maf_tc.cu", ["-DRABI_TC_SIDE_DP=8"], _COMMON),      # the same source: only the DP = 8 kernels
    "maf_tc_dp12": ("maf_tc.cu", ["-DRABI_TC_SIDE_DP=12"], _COMMON),
    "maf_tcd": ("maf_tcd.cu", [], _COMMON),
    "maf_train": ("maf_train.cu", [], _COMMON),
}

EXPORTS = [
    "rabi_build_info", "rabi_last_error", "rabi_maf_create", "rabi_maf_destroy", "rabi_maf_set_permutation",
    "rabi_maf_set_mask", "rabi_maf_set_post_permutation", "rabi_maf_finalize_structure", "rabi_maf_set_weights", "rabi_maf_ctx_project",
    "rabi_maf_ctx_project_bwd", "rabi_maf_log_prob_fwd", "rabi_maf_rsample_fwd", "rabi_rkl_fwd",
    "rabi_rkl_input_grad", "rabi_rkl_pgd_step", "rabi_rkl_pgd_run", "rabi_launch_count",
    "rabi_profile", "rabi_profile_read", "rabi_fma_peak_tflops", "rabi_sgemm", "rabi_sgemm_launch_count",
    "rabi_fkl_input_grad", "rabi_fkl_pgd_run", "rabi_pgd_update", "rabi_pgd_set_row_scales", "rabi_pgd_set_column_mask", "rabi_maf_log_prob_bwd", "rabi_maf_rsample_bwd", "rabi_fisher_trace_fwd", "rabi_fisher_trace_bwd", "rabi_train_emissions",
]


class RabiError(RuntimeError):
    pass


def _nvcc() -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return nvcc if os.path.exists(nvcc) else "nvcc"


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> _lib/obj/*.o -> _lib/librabi_b200.so (sm_100a; cross-compiles without a GPU).  Every
    translation unit is rebuilt only when it or one of its headers is newer than its object; the units compile in
    parallel.  ``RABI_NVCC_EXTRA`` adds flags (e.g. ``-DRABI_LEAN`` for kernel experiments)."""
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    extra = os.environ.get("RABI_NVCC_EXTRA", "").split()
    stamp = os.path.join(obj_dir, "flags.txt")
    flags_now = " ".join(NVCC_FLAGS + extra)
    if not os.path.exists(stamp) or open(stamp).read() != flags_now:
        force = True
    jobs, objs = [], []
    for name, (src, defines, hdrs) in SOURCES.items():
        obj = os.path.join(obj_dir, name + ".o")
        objs.append(obj)
        deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, x) for x in hdrs] + [os.path.join(INCLUDE, "rabi_b200.h")]
        if not force and os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps):
            continue
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, *defines, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        jobs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, proc in jobs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RabiError("nvcc failed:\n" + " ".join(cmd) + "\n" + out)
        if verbose:
            print(out)
    if jobs or not os.path.exists(LIB_PATH) or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs):
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH, *objs]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RabiError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    with open(stamp, "w") as f:
        f.write(flags_now)
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RabiError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(rabi_b200 has no CPU / PyTorch fallback)."
        )
    L = ctypes.CDLL(LIB_PATH)
    fp = POINTER(c_float)
    vp = c_void_p
    L.rabi_build_info.restype = c_char_p
    L.rabi_build_info.argtypes = []
    L.rabi_last_error.restype = c_char_p
    L.rabi_last_error.argtypes = [vp]
    L.rabi_launch_count.restype = c_ulonglong
    L.rabi_launch_count.argtypes = [vp]
    L.rabi_maf_create.argtypes = [POINTER(vp), c_int, c_int, c_int, c_int, c_int, POINTER(c_int), c_float, c_float]
    L.rabi_maf_destroy.argtypes = [vp]
    L.rabi_maf_set_permutation.argtypes = [vp, c_int, POINTER(c_int64)]
    L.rabi_maf_set_post_permutation.argtypes = [vp, c_int, POINTER(c_int64)]
    L.rabi_maf_set_mask.argtypes = [vp, c_int, c_int, fp, c_int, c_int]
    L.rabi_maf_finalize_structure.argtypes = [vp]
    L.rabi_maf_set_weights.argtypes = [vp, c_int, c_int, vp, vp, vp]
    L.rabi_maf_ctx_project.argtypes = [vp, vp, vp, vp, c_int, vp, vp]
    L.rabi_maf_ctx_project_bwd.argtypes = [vp, vp, vp, c_int, vp, vp]
    L.rabi_maf_log_prob_fwd.argtypes = [vp, vp, vp, c_int, c_int, vp, vp]
    L.rabi_maf_rsample_fwd.argtypes = [vp, vp, vp, c_int, c_int, vp, vp, vp]
    L.rabi_rkl_fwd.argtypes = [vp, vp, vp, vp, c_int, c_int, vp, vp]
    L.rabi_rkl_input_grad.argtypes = [vp, vp, vp, vp, vp, vp, c_int, c_int, c_float, vp, vp, vp]
    L.rabi_rkl_pgd_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, c_int, c_int, c_float, c_float, c_float, c_float,
                                    c_float, c_float, vp, vp]
    L.rabi_rkl_pgd_run.argtypes = [vp, vp, vp, vp, vp, vp, vp, c_int, c_int, c_int, c_float, c_float, c_float, c_float,
                                   c_float, c_float, vp, vp]
    L.rabi_profile.argtypes = [vp, c_int]
    L.rabi_profile_read.argtypes = [vp, POINTER(ctypes.c_double), POINTER(c_ulonglong)]
    L.rabi_fma_peak_tflops.argtypes = [c_int, c_int]
    L.rabi_fma_peak_tflops.restype = ctypes.c_double
    L.rabi_fkl_input_grad.argtypes = L.rabi_rkl_input_grad.argtypes
    L.rabi_fkl_pgd_run.argtypes = L.rabi_rkl_pgd_run.argtypes
    L.rabi_fkl_input_grad.restype = L.rabi_fkl_pgd_run.restype = c_int
    L.rabi_pgd_update.argtypes = [vp, vp, vp, vp, c_int, c_int, c_float, c_float, c_float, c_float, c_float, vp]
    L.rabi_pgd_update.restype = c_int
    L.rabi_pgd_set_row_scales.argtypes = [vp, vp, vp]
    L.rabi_pgd_set_row_scales.restype = c_int
    L.rabi_pgd_set_column_mask.argtypes = [vp, vp]
    L.rabi_pgd_set_column_mask.restype = c_int
    L.rabi_maf_log_prob_bwd.argtypes = [vp, vp, vp, vp, vp, c_int, c_int, vp, vp, vp, vp, vp, vp]
    L.rabi_fisher_trace_bwd.argtypes = [vp, vp, vp, vp, vp, c_int, c_int, vp, vp, vp, vp, vp, vp, vp]
    L.rabi_fisher_trace_bwd.restype = c_int
    L.rabi_train_emissions.argtypes = [vp, c_int, c_int, vp, vp]
    L.rabi_train_emissions.restype = c_int
    L.rabi_fisher_trace_fwd.argtypes = [vp, vp, vp, vp, vp, c_int, c_int, vp, vp, vp, vp]
    L.rabi_maf_rsample_bwd.argtypes = [vp, vp, vp, vp, vp, vp, c_int, c_int, vp, vp, vp, vp, vp, vp]
    L.rabi_sgemm.argtypes = [c_int, c_int, c_int, c_int, c_int, c_int, vp, c_int, vp, c_int, vp, c_int, c_float, vp]
    L.rabi_sgemm.restype = c_int
    L.rabi_sgemm_launch_count.argtypes = []
    L.rabi_sgemm_launch_count.restype = c_ulonglong
    for name in EXPORTS:
        fn = getattr(L, name)
        if name.startswith(("rabi_maf_", "rabi_rkl_", "rabi_profile")):
            fn.restype = c_int
    _lib = L
    return L
