"""CPU ORACLE -- test infrastructure, NOT the product.

Python face of ``oracle/gme_oracle.c``: the same function names as the reference's
``bbme.py`` / ``motion.py`` / ``utils.py`` hot path, computed on the CPU by the C
restatement (plus NumPy for the 3x3 inverse, which the reference also leaves to
``np.linalg.inv``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module.

Parity is pinned: ``tests/test_oracle_golden.py`` checks every function here against
outputs of the unmodified reference (``tests/golden/``, made by ``oracle/make_golden.py``)
and the known-answer vectors of SURVEY.md Appendix B.

Reference citations are relative to /root/reference/global_motion_estimation/.
"""
from __future__ import annotations

import cmath
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libgme_oracle.so")

BBME_BLOCK_SIZE = 16                                   # motion.py:9
MOTION_VECTOR_ERROR_THRESHOLD_PERCENTAGE = .3          # motion.py:10


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "gme_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        u8p, i32p, i16p, i64p, f64p = (ctypes.c_void_p,) * 5
        ci, cl, cd = ctypes.c_int, ctypes.c_long, ctypes.c_double
        L.oracle_motion_field.argtypes = [u8p, u8p, ci, ci, cl, ci, ci, ci, ci, i32p, i64p, ci, ci]
        L.oracle_motion_field.restype = ci
        L.oracle_pyr_down.argtypes = [u8p, ci, ci, cl, u8p, cl]
        L.oracle_pyr_down.restype = ci
        L.oracle_affine_field.argtypes = [ci, ci, f64p, i16p]
        L.oracle_affine_field.restype = ci
        L.oracle_outlier_mask.argtypes = [i32p, i16p, ci, ci, cd, u8p, i32p]
        L.oracle_outlier_mask.restype = ci
        L.oracle_normal_equations.argtypes = [i32p, u8p, ci, ci, cd, f64p, f64p, f64p]
        L.oracle_normal_equations.restype = ci
        L.oracle_compensate.argtypes = [u8p, ci, ci, cl, i32p, ci, ci, u8p, cl]
        L.oracle_compensate.restype = ci
        L.oracle_sse.argtypes = [u8p, cl, u8p, cl, ci, ci]
        L.oracle_sse.restype = ctypes.c_int64
        L.oracle_exhaustive_candidates.argtypes = [ci, ci, ci, ci]
        L.oracle_exhaustive_candidates.restype = ctypes.c_int64
        _lib = L
    return _lib


def _u8(img) -> np.ndarray:
    a = np.ascontiguousarray(img, dtype=np.uint8)
    if a.ndim != 2:
        raise ValueError("expected a 2-D grayscale frame")
    return a


# --------------------------------------------------------------------------- bbme.py
def get_motion_field(previous, current, block_size=4, search_window=2, searching_procedure=1,
                     pnorm_distance=1, threads=1, return_candidates=False):
    """bbme.get_motion_field (bbme.py:12-38) -> int32[H//bs, W//bs, 2]."""
    if not 0 <= int(searching_procedure) <= 3:
        raise IndexError("list index out of range")        # searching_procedures[idx], bbme.py:27
    if not 0 <= int(pnorm_distance) <= 1:
        raise IndexError("list index out of range")        # pnorm_distances[idx], bbme.py:60
    prev, cur = _u8(previous), _u8(current)
    H, W = prev.shape
    bs = int(block_size)
    R, C = int(H / bs), int(W / bs)
    field = np.zeros((R, C, 2), np.int32)
    L = lib()

    def run(r0, r1):
        n = ctypes.c_int64(0)
        rc = L.oracle_motion_field(prev.ctypes.data, cur.ctypes.data, H, W, W, bs, int(search_window),
                                   int(searching_procedure), int(pnorm_distance), field.ctypes.data,
                                   ctypes.byref(n), r0, r1)
        if rc != 0:
            raise ValueError(f"oracle_motion_field rc={rc}")
        return n.value

    threads = max(1, min(int(threads), R))
    if threads == 1 or R == 0:
        total = run(0, R) if R else 0
    else:
        bounds = np.linspace(0, R, threads * 4 + 1).astype(int)
        spans = [(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        with ThreadPoolExecutor(threads) as ex:           # ctypes drops the GIL
            total = sum(ex.map(lambda s: run(*s), spans))
    return (field, total) if return_candidates else field


def rescale_motion_field(motion_field, scale=2):
    """bbme.rescale_motion_field (bbme.py:537-546): nearest upsample, values *2 (hard-coded)."""
    mf = np.repeat(np.repeat(np.asarray(motion_field), scale, axis=0), scale, axis=1).astype(np.int32)
    return mf * 2


def hierarchical_wrapper(previous, current, block_size=10, search_window=4, searching_procedure=3):
    """bbme.hierarchical_wrapper (bbme.py:549-605) -> float64 field."""
    ppyr, cpyr = get_pyramids(previous, 3), get_pyramids(current, 3)
    mf = get_motion_field(ppyr[0], cpyr[0], block_size=block_size,
                          searching_procedure=searching_procedure, search_window=search_window)
    for level in range(1, 3):
        mf = rescale_motion_field(mf, 2)
        new = get_motion_field(ppyr[level], cpyr[level], block_size=block_size,
                               searching_procedure=3, search_window=search_window)
        if mf.shape != new.shape:                          # bbme.py:596-602: one row OR one column
            if mf.shape[0] != new.shape[0]:
                mf = np.vstack([mf, np.zeros((1, mf.shape[1], 2), np.int32)])
            else:
                mf = np.hstack([mf, np.zeros((mf.shape[0], 1, 2), np.int32)])
        mf = (mf + new) / 2
    return mf


# --------------------------------------------------------------------------- utils.py
def pyr_down(img) -> np.ndarray:
    """cv2.pyrDown on uint8 (utils.py:48)."""
    a = _u8(img)
    H, W = a.shape
    out = np.empty(((H + 1) // 2, (W + 1) // 2), np.uint8)
    lib().oracle_pyr_down(a.ctypes.data, H, W, W, out.ctypes.data, out.shape[1])
    return out


def get_pyramids(original_image, levels=3):
    """utils.get_pyramids (utils.py:34-51): coarsest first."""
    pyr = [original_image]
    cur = original_image
    for _ in range(1, levels):
        cur = pyr_down(cur)
        pyr.insert(0, cur)
    return pyr


def sse(a, b) -> int:
    a, b = _u8(a), _u8(b)
    return int(lib().oracle_sse(a.ctypes.data, a.shape[1], b.ctypes.data, b.shape[1], a.shape[0], a.shape[1]))


def PSNR(original, noisy):
    """utils.PSNR (utils.py:100-116): complex result, or int -1 when identical."""
    mse = sse(original, noisy) / float(np.asarray(original).size)
    if mse == 0:
        return -1
    return 20 * cmath.log10(255.0 / cmath.sqrt(mse))


# --------------------------------------------------------------------------- motion.py
def dense_motion_estimation(previous, current):
    """motion.dense_motion_estimation (motion.py:13-30): diamond, bs=2, MSE."""
    return get_motion_field(previous, current, block_size=2, searching_procedure=3)


def compute_first_parameters(dense_motion_field):
    """motion.compute_first_parameters (motion.py:176-188)."""
    a0 = np.mean(dense_motion_field[:, :, 0])
    b0 = np.mean(dense_motion_field[:, :, 1])
    return np.array([a0, 0.0, 0.0, b0, 0.0, 0.0], dtype=np.float32)


def first_parameter_estimation(previous, current):
    return compute_first_parameters(dense_motion_estimation(previous, current))


def parameter_projection(parameters):
    """motion.parameter_projection (motion.py:191-207): in place."""
    parameters[0] = parameters[0] * 2
    parameters[3] = parameters[3] * 2
    return parameters


def get_motion_field_affine(shape, parameters):
    """motion.get_motion_field_affine (motion.py:139-157) -> int16[R, C, 2]."""
    R, C = int(shape[0]), int(shape[1])
    p = np.ascontiguousarray(np.asarray(parameters), dtype=np.float64)   # float32 promotes exactly
    out = np.zeros((R, C, 2), np.int16)
    lib().oracle_affine_field(R, C, p.ctypes.data, out.ctypes.data)
    return out


def outlier_mask(gt_motion_field, old_params_motion_field, pct=MOTION_VECTOR_ERROR_THRESHOLD_PERCENTAGE):
    """motion.py:234-244 -> (bool[R, C] outlier mask, threshold)."""
    gt = np.ascontiguousarray(gt_motion_field, dtype=np.int32)
    model = np.ascontiguousarray(old_params_motion_field, dtype=np.int16)
    R, C = gt.shape[:2]
    mask = np.zeros((R, C), np.uint8)
    thr = ctypes.c_int32(0)
    lib().oracle_outlier_mask(gt.ctypes.data, model.ctypes.data, R, C, float(pct), mask.ctypes.data,
                              ctypes.byref(thr))
    return mask.astype(bool), int(thr.value)


def _solve(gt, mask, level_shape):
    """motion.py:246-286: masked normal equations accumulated in C, np.linalg.inv + matmul here."""
    gt = np.ascontiguousarray(gt, dtype=np.int32)
    R, C = gt.shape[:2]
    w = 1 / (level_shape[0] * level_shape[1])
    first, s0, s1 = np.zeros((3, 3)), np.zeros((3, 1)), np.zeros((3, 1))
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    lib().oracle_normal_equations(gt.ctypes.data, None if m is None else m.ctypes.data, R, C, w,
                                  first.ctypes.data, s0.ctypes.data, s1.ctypes.data)
    finv = np.array(np.linalg.inv(first))
    ax = np.matmul(finv, s0).reshape((3,))
    ay = np.matmul(finv, s1).reshape((3,))
    return np.concatenate([ax, ay])


def best_affine_parameters_robust(previous, current, old_parameters, procedure=3, window=2,
                                  return_intermediates=False, threads=1):
    """motion.best_affine_parameters_robust (motion.py:210-286).

    ``procedure``/``window`` default to the values the reference hard-codes (diamond, MSE)."""
    gt = get_motion_field(previous, current, block_size=BBME_BLOCK_SIZE, searching_procedure=procedure,
                          search_window=window, threads=threads)
    model = get_motion_field_affine(gt.shape, old_parameters)
    mask, thr = outlier_mask(gt, model)
    a = _solve(gt, mask, np.asarray(previous).shape)
    if return_intermediates:
        return a, dict(gt=gt, model=model, outlier=mask, threshold=thr)
    return a


def best_affine_parameters(previous, current):
    """motion.best_affine_parameters (motion.py:33-88): no outlier mask."""
    gt = get_motion_field(previous, current, block_size=BBME_BLOCK_SIZE, searching_procedure=3)
    return _solve(gt, None, np.asarray(previous).shape)


def global_motion_estimation(previous, current, procedure=3, window=2, return_intermediates=False, threads=1):
    """motion.global_motion_estimation (motion.py:109-136)."""
    ppyr, cpyr = get_pyramids(previous), get_pyramids(current)
    dense = get_motion_field(ppyr[0], cpyr[0], block_size=2, searching_procedure=3, threads=threads)
    parameters = compute_first_parameters(dense)
    inter = [dict(dense=dense, first=parameters.copy())]
    for i in range(1, len(ppyr)):
        parameters = parameter_projection(parameters)
        parameters, d = best_affine_parameters_robust(ppyr[i], cpyr[i], parameters, procedure, window,
                                                      return_intermediates=True, threads=threads)
        d["params"] = parameters.copy()
        inter.append(d)
    return (parameters, inter) if return_intermediates else parameters


def compensate_frame(frame, motion_field):
    """motion.compensate_frame (motion.py:289-321)."""
    f = _u8(frame)
    mf = np.ascontiguousarray(motion_field, dtype=np.int32)
    out = np.empty_like(f)
    lib().oracle_compensate(f.ctypes.data, f.shape[0], f.shape[1], f.shape[1], mf.ctypes.data,
                            mf.shape[0], mf.shape[1], out.ctypes.data, out.shape[1])
    return out


def motion_compensation(previous, current):
    """motion.motion_compensation (motion.py:324-341)."""
    parameters = global_motion_estimation(previous, current)
    prev = np.asarray(previous)
    shape = (prev.shape[0] // BBME_BLOCK_SIZE, prev.shape[1] // BBME_BLOCK_SIZE)
    return compensate_frame(prev, get_motion_field_affine(shape, parameters))


def exhaustive_sad_ops(H, W, bs, sw) -> int:
    """exact pixel-pair operations of one exhaustive search."""
    return int(lib().oracle_exhaustive_candidates(H, W, bs, sw)) * bs * bs
