"""ctypes binding of libgme_b200.so (include/gme_b200.h).

There is no CPU fallback: if the CUDA library has not been built, importing this module
raises, and every wrapper raises when a call returns an error code.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgme_b200.so")

GME_OK = 0
GME_ERR_INVALID_ARGUMENT = -1
GME_ERR_UNSUPPORTED = -2
GME_ERR_ALIGNMENT = -3
GME_ERR_CUDA = -4
GME_ERR_WORKSPACE = -5

SEARCH_EXHAUSTIVE, SEARCH_THREESTEP, SEARCH_TWODLOG, SEARCH_DIAMOND = 0, 1, 2, 3
PNORM_MAE, PNORM_MSE = 0, 1

# every symbol include/gme_b200.h declares: (name, restype, argtypes)
_p, _sz, _i, _d = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_double
SYMBOLS = {
    "gme_version": (_i, []),
    "gme_error_string": (ctypes.c_char_p, [_i]),
    "gme_last_cuda_error": (_i, []),
    "gme_bbme_motion_field": (_i, [_p, _sz, _p, _sz, _i, _i, _i, _sz, _i, _i, _i, _i, _p, _p]),
    "gme_pyr_down": (_i, [_p, _sz, _sz, _p, _sz, _sz, _i, _i, _i, _p]),
    "gme_first_parameters": (_i, [_p, _i, _i, _i, _p, _p]),
    "gme_affine_fit": (_i, [_p, _i, _i, _i, _i, _i, _d, _i, _i, _p, _p, _p, _p, _p, _p]),
    "gme_affine_field": (_i, [_p, _i, _i, _i, _p, _p]),
    "gme_compensate": (_i, [_p, _sz, _sz, _p, _i, _i, _i, _p, _sz, _sz, _p, _sz, _sz, _i, _i, _i, _p, _p]),
    "gme_compensate_diffs": (_i, [_p, _sz, _sz, _p, _i, _i, _i, _p, _sz, _sz, _p, _sz, _sz, _p, _p, _sz, _sz, _i, _i, _i, _p, _p]),
    "gme_hier_merge": (_i, [_p, _i, _i, _i, _p, _i, _i, _i, _p, _p]),
    "gme_sse": (_i, [_p, _sz, _sz, _p, _sz, _sz, _i, _i, _i, _p, _p]),
    "gme_pipeline_workspace_bytes": (_sz, [_i, _i, _i]),
    "gme_pipeline": (_i, [_p, _sz, _p, _sz, _i, _i, _i, _sz, _i, _i, _d, _p, _sz, _p, _sz, _sz, _p, _sz, _p, _p, _sz, _p]),
    "gme_pipeline_workspace_ptr": (_p, [_p, _i, _i, _i, _i]),
    "gme_stage_timing_enable": (_i, [_i]),
    "gme_stage_timing_read": (_i, [_p, _p]),
    "gme_sad_peak_probe": (_i, [_i, _i, _i, _p, _p, _p]),
    "gme_launch_count": (ctypes.c_uint64, []),
}


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first "
            "(python global-motion-estimation_b200/build.py, or __graft_entry__.build()). "
            "There is no CPU fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()
ABI_VERSION = lib.gme_version()


class GmeError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc == GME_OK:
        return
    msg = f"{what}: {lib.gme_error_string(rc).decode()} (code {rc})"
    if rc in (GME_ERR_INVALID_ARGUMENT, GME_ERR_UNSUPPORTED, GME_ERR_ALIGNMENT):
        raise ValueError(msg)
    if rc == GME_ERR_CUDA:
        msg += f", cudaError {lib.gme_last_cuda_error()}"
    raise GmeError(msg)


def launch_count() -> int:
    return int(lib.gme_launch_count())


PIPELINE_STAGES = 6
STAGE_NAMES = ("pyramids", "bbme_dense_l0", "bbme_l1", "bbme_l2", "fit", "compensate_psnr")


def stage_timing_enable(on: bool) -> None:
    check(lib.gme_stage_timing_enable(int(on)), "gme_stage_timing_enable")


def stage_timing_read():
    """-> (ms per stage summed over the recorded gme_pipeline calls, number of calls)."""
    ms = (ctypes.c_double * PIPELINE_STAGES)()
    calls = ctypes.c_int(0)
    check(lib.gme_stage_timing_read(ms, ctypes.byref(calls)), "gme_stage_timing_read")
    return list(ms), calls.value


def singular_matrix_error():
    """The exception the reference raises from np.linalg.inv (motion.py:262) on a singular normal matrix."""
    return np.linalg.LinAlgError("Singular matrix")
