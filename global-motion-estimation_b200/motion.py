"""Drop-in for the reference's ``motion.py``: the global-motion-estimation pipeline on the B200.

Same names, arguments, return shapes and dtypes as
/root/reference/global_motion_estimation/motion.py.  The reference hard-codes the search of the
pipeline (diamond, MSE); README:137-141 tells users to edit module constants, so the search is
exposed the same way here, defaulting to the reference's values.
"""
import numpy as np
import torch

import gme_device as _dev
import gme_native as _native
from bbme import get_motion_field
from utils import get_pyramids, timer  # noqa: F401  (re-exported like the reference)

BBME_BLOCK_SIZE = 16                                   # motion.py:9
MOTION_VECTOR_ERROR_THRESHOLD_PERCENTAGE = .3          # motion.py:10
# search used for the block_size-16 fields (motion.py:46-51, 224-229); reference: diamond (3), window ignored
BBME_SEARCHING_PROCEDURE = 3
BBME_SEARCH_WINDOW = 2


def dense_motion_estimation(previous, current):
    """motion.py:13-30 -- dense field: diamond search on 2x2 blocks."""
    return get_motion_field(previous, current, block_size=2, searching_procedure=3)


def _gt_field(previous, current):
    return get_motion_field(previous=previous, current=current, block_size=BBME_BLOCK_SIZE,
                            searching_procedure=BBME_SEARCHING_PROCEDURE, search_window=BBME_SEARCH_WINDOW)


def _fit(gt_motion_field, level_shape, old_parameters, robust):
    dev = _dev.require_cuda()
    gt = torch.from_numpy(np.ascontiguousarray(gt_motion_field, dtype=np.int32)).unsqueeze(0).to(dev)
    start = np.zeros(6) if old_parameters is None else np.asarray(old_parameters, dtype=np.float64)
    params = torch.from_numpy(np.ascontiguousarray(start)).reshape(1, 6).to(dev)
    params, status = _dev.affine_fit(gt, level_shape, params, robust=robust, project=False,
                                     pct=MOTION_VECTOR_ERROR_THRESHOLD_PERCENTAGE)
    if int(status.item()) != 0:
        raise _native.singular_matrix_error()          # np.linalg.inv in the reference (motion.py:65, 262)
    return params[0].cpu().numpy()


def best_affine_parameters(previous, current):
    """motion.py:33-88 -- least-squares affine parameters [a0,a1,a2,b0,b1,b2] without outlier rejection."""
    return _fit(_gt_field(previous, current), previous.shape, None, robust=False)


def affine_model(x, y, parameters):
    """motion.py:91-105 -- displacement of position (x, y) under the affine model (float64[2])."""
    A = np.asarray([[1, x, y, 0, 0, 0], [0, 0, 0, 1, x, y]], dtype=np.int32)
    return np.matmul(A, np.transpose(parameters))


def global_motion_estimation(previous, current):
    """motion.py:109-136 -- hierarchical robust affine GME of one frame pair -> float64[6]."""
    if BBME_BLOCK_SIZE != 16:
        # the fused pipeline is built for the reference's block size; an edited constant takes the level-by-level path
        prev_pyr, curr_pyr = get_pyramids(previous), get_pyramids(current)
        parameters = first_parameter_estimation(prev_pyr[0], curr_pyr[0])
        for i in range(1, len(prev_pyr)):
            parameters = parameter_projection(parameters)
            parameters = best_affine_parameters_robust(prev_pyr[i], curr_pyr[i], parameters)
        return parameters
    previous, current = np.asarray(previous), np.asarray(current)
    if previous.ndim == 2 and previous.shape == current.shape and previous.dtype == current.dtype == np.uint8:
        # the path results.py drives pair by pair: persistent per-geometry session (pinned staging, CUDA graph)
        return _dev.pair_session(*previous.shape).gme(previous, current, BBME_SEARCHING_PROCEDURE, BBME_SEARCH_WINDOW,
                                                      MOTION_VECTOR_ERROR_THRESHOLD_PERCENTAGE)
    params, _, _ = _dev.gme_pairs(previous[None], current[None],
                                  procedure=BBME_SEARCHING_PROCEDURE, window=BBME_SEARCH_WINDOW, want_comp=False,
                                  outlier_fraction=MOTION_VECTOR_ERROR_THRESHOLD_PERCENTAGE)
    return params[0]


def get_motion_field_affine(shape, parameters):
    """motion.py:139-157 -- int16[shape[0], shape[1], 2] model motion field."""
    dev = _dev.require_cuda()
    p = torch.from_numpy(np.ascontiguousarray(np.asarray(parameters), dtype=np.float64)).reshape(1, 6).to(dev)
    return _dev.affine_field(p, int(shape[0]), int(shape[1]))[0].cpu().numpy()


def first_parameter_estimation(previous, current):
    """motion.py:160-173."""
    return compute_first_parameters(dense_motion_estimation(previous, current))


def compute_first_parameters(dense_motion_field):
    """motion.py:176-188 -- float32[6] = [mean dx, 0, 0, mean dy, 0, 0]."""
    dev = _dev.require_cuda()
    dense = torch.from_numpy(np.ascontiguousarray(dense_motion_field, dtype=np.int32)).unsqueeze(0).to(dev)
    return _dev.first_parameters(dense)[0].cpu().numpy().astype(np.float32)


def parameter_projection(parameters):
    """motion.py:191-207 -- level l -> l+1: a0 and b0 doubled IN PLACE, same object returned."""
    parameters[0] = parameters[0] * 2
    parameters[3] = parameters[3] * 2
    return parameters


def best_affine_parameters_robust(previous, current, old_parameters):
    """motion.py:210-286 -- outlier-masked least-squares affine parameters -> float64[6]."""
    return _fit(_gt_field(previous, current), previous.shape, old_parameters, robust=True)


def compensate_frame(frame, motion_field):
    """motion.py:289-321 -- block-wise translation of ``frame`` by ``motion_field`` -> uint8[H, W]."""
    frame = np.asarray(frame)
    mf = np.asarray(motion_field)
    if not np.issubdtype(mf.dtype, np.integer):
        # a float field makes every index a float in the reference: each access raises inside its bare
        # try/except, so nothing is moved
        return np.copy(frame)
    if frame.ndim == 2 and frame.dtype == np.uint8 and mf.ndim == 3 and mf.shape[0] > 0 and mf.shape[1] > 0:
        return _dev.pair_session(*frame.shape).compensate(frame, mf[..., :2])
    dev = _dev.require_cuda()
    field = torch.from_numpy(np.ascontiguousarray(mf[..., :2], dtype=np.int32)).unsqueeze(0).to(dev)
    comp, _ = _dev.compensate(_dev.Planes.from_host(frame), field)
    return comp.to_host()[0]


def motion_compensation(previous, current):
    """motion.py:324-341 -- GME, model field at block size 16, compensated previous frame."""
    parameters = global_motion_estimation(previous, current)
    shape = (previous.shape[0] // BBME_BLOCK_SIZE, previous.shape[1] // BBME_BLOCK_SIZE)
    return compensate_frame(previous, get_motion_field_affine(shape, parameters))
