"""Drop-in for the reference's ``bbme.py``: block-based motion estimation on the B200.

Same function names, arguments, defaults, return shapes and quirks as
/root/reference/global_motion_estimation/bbme.py; the searches run in libgme_b200.so
(exhaustive: gme_bbme_exhaustive.cu, three-step / 2D-log / diamond: gme_bbme_pattern.cu).
NumPy arrays in, fresh NumPy arrays out.
"""
import argparse
import os

import cv2
import numpy as np

import gme_device as _dev
from utils import draw_motion_field, get_video_frames, get_pyramids


def _search(procedure, previous, current, mf, height, width, pnorm_distance, block_size, search_window):
    """Shared body of the four ``*_search`` functions: fills and returns ``mf`` (bbme.py:105-534)."""
    pnorm_distance = range(len(pnorm_distances))[pnorm_distance]      # list indexing: IndexError, negatives wrap
    prev = np.asarray(previous)[:height, :width]
    cur = np.asarray(current)[:height, :width]
    field = _dev.motion_field(_dev.Planes.from_host(prev), _dev.Planes.from_host(cur), block_size, search_window,
                              procedure, pnorm_distance)[0].cpu().numpy()
    rows, cols = min(field.shape[0], mf.shape[0]), min(field.shape[1], mf.shape[1])
    mf[:rows, :cols] = field[:rows, :cols]
    return mf


def get_motion_field(previous, current, block_size=4, search_window=2, searching_procedure=1, pnorm_distance=1
                     ) -> np.ndarray:
    """bbme.py:12-38 -> int32[int(H/bs), int(W/bs), 2]; [..., 0] horizontal, [..., 1] vertical."""
    height, width = previous.shape[0], previous.shape[1]
    motion_field = np.zeros((int(height / block_size), int(width / block_size), 2), dtype=np.int32)
    search = searching_procedures[searching_procedure]
    return search(previous, current, motion_field, height, width, pnorm_distance, block_size, search_window)


def compute_dfd(block_1, block_2, pnorm_index=0):
    """bbme.py:41-64 -- displaced frame difference of two equally shaped blocks as np.float32.

    Host helper kept for API completeness; the kernels evaluate the same sums as exact integers."""
    assert block_1.shape == block_2.shape
    pnorm = pnorm_distances[pnorm_index]
    return pnorm(np.array(block_1, dtype=np.float32) - np.array(block_2, dtype=np.float32))


def mae(diff_block):
    """bbme.py:67-79 -- sum of absolute values."""
    return np.sum(np.abs(diff_block))


def mse(diff_block):
    """bbme.py:82-94 -- sum of squares."""
    return np.sum(diff_block * diff_block)


def compute_current_target_block_corners(br, bl, wr, wc, bs):
    """bbme.py:97-102 -- ((top, left), (bottom, right)) of the bs x bs block whose top-left is (wr, wc)."""
    return (wr, wc), (wr + bs - 1, wc + bs - 1)


def exhaustive_search(previous, current, mf, height, width, pnorm_distance=0, block_size=4, search_window=2):
    """bbme.py:105-179."""
    return _search(0, previous, current, mf, height, width, pnorm_distance, block_size, search_window)


def threestep_search(previous, current, mf, height, width, pnorm_distance=0, block_size=4, search_window=12):
    """bbme.py:182-341."""
    return _search(1, previous, current, mf, height, width, pnorm_distance, block_size, search_window)


def twodlog_search(previous, current, mf, height, width, pnorm_function, block_size=4, search_window=12):
    """bbme.py:344-433."""
    return _search(2, previous, current, mf, height, width, pnorm_function, block_size, search_window)


def diamond_search(previous, current, mf, height, width, pnorm_distance=0, block_size=12, search_window=-1):
    """bbme.py:436-534 (search_window is ignored, as in the reference)."""
    return _search(3, previous, current, mf, height, width, pnorm_distance, block_size, 0)


def rescale_motion_field(motion_field, scale=2):
    """bbme.py:537-546 -- nearest-neighbour upsampling by ``scale``; the vectors are doubled (hard-coded 2)."""
    mf = np.asarray(motion_field)
    up = np.repeat(np.repeat(mf, scale, axis=0), scale, axis=1).astype(np.int32)
    return up * 2


def hierarchical_wrapper(previous, current, block_size=10, search_window=4, searching_procedure=3):
    """bbme.py:549-605 -- three-level BBME, each finer level averaged with the upsampled coarser one -> float64.

    When the level shapes merge the way bbme.py:596-604 expects, everything (pyramids, the three searches, the two
    merges) runs on the device with one upload and one download; odd geometries take the step-by-step host path
    below, which raises or broadcasts exactly like the reference's NumPy code."""
    prev, cur = np.asarray(previous), np.asarray(current)
    if prev.ndim == 2 and prev.shape == cur.shape and block_size > 0 and 0 <= searching_procedure <= 3:
        shapes = []
        h, w = prev.shape
        for _ in range(3):
            shapes.insert(0, (int(h / block_size), int(w / block_size)))
            h, w = (h + 1) // 2, (w + 1) // 2
        if all(r > 0 and c > 0 for r, c in shapes) and all(
                _dev.hierarchical_shapes_merge(a, b) for a, b in zip(shapes, shapes[1:])):
            out = _dev.hierarchical_field(_dev.Planes.from_host(prev), _dev.Planes.from_host(cur), block_size,
                                          search_window, searching_procedure)
            return out[0].cpu().numpy()
    previous_pyr = get_pyramids(previous, levels=3)
    current_pyr = get_pyramids(current, levels=3)
    motion_field = get_motion_field(previous_pyr[0], current_pyr[0], block_size=block_size,
                                    searching_procedure=searching_procedure, search_window=search_window)
    for level in range(1, len(previous_pyr)):
        motion_field = rescale_motion_field(motion_field, scale=2)
        finer = get_motion_field(previous_pyr[level], current_pyr[level], block_size=block_size,
                                 searching_procedure=3, search_window=search_window)
        if motion_field.shape != finer.shape:
            # integer rounding between levels: the reference pads one row OR one column (bbme.py:596-602)
            if motion_field.shape[0] != finer.shape[0]:
                motion_field = np.vstack([motion_field, np.zeros((1, motion_field.shape[1], 2), dtype=np.int32)])
            else:
                motion_field = np.hstack([motion_field, np.zeros((motion_field.shape[0], 1, 2), dtype=np.int32)])
        motion_field = (motion_field + finer) / 2
    return motion_field


pnorm_distances = [mae, mse]
searching_procedures = [exhaustive_search, threestep_search, twodlog_search, diamond_search]


def main(args):
    """bbme.py:617-647 -- frames fi-3 and fi; note that the p-norm flag is parsed but not forwarded."""
    frames = get_video_frames(args.path)
    previous, current = frames[args.fi - 3], frames[args.fi]
    motion_field = get_motion_field(previous, current, block_size=args.block_size,
                                    searching_procedure=args.searching_procedure, search_window=args.search_window)
    motion_field_hierarchical = hierarchical_wrapper(previous, current, block_size=args.block_size,
                                                     search_window=args.search_window,
                                                     searching_procedure=args.searching_procedure)
    flat = draw_motion_field(current, motion_field)
    hier = draw_motion_field(previous, motion_field_hierarchical)
    cv2.imwrite(os.path.join("resources", "images", f"{args.searching_procedure}-res.png"), flat)
    cv2.imwrite(os.path.join("resources", "images", f"{args.searching_procedure}h-res.png"), hier)


def _parser():
    parser = argparse.ArgumentParser(
        description="Computes motion field between two frames using block matching algorithms")
    parser.add_argument("-p", "--video-path", dest="path", type=str, required=True,
                        help="path of the video to analyze")
    parser.add_argument("-fi", "--frame-index", dest="fi", type=int, required=True,
                        help="index of the current frame to analyze in the video")
    parser.add_argument("-pn", "--p-norm", dest="pnorm", type=int, default=0, help="0: 1-norm (mae), 1: 2-norm (mse)")
    parser.add_argument("-bs", "--block-size", dest="block_size", type=int, default=12, help="size of the block")
    parser.add_argument("-sw", "--search-window", dest="search_window", type=int, default=8,
                        help="size of the search window")
    parser.add_argument("-sp", "--searching-procedure", dest="searching_procedure", type=int, default=1,
                        help="0: Exhaustive search,\n1: Three Step search,\n2: 2D Log search,\n3: Diamond search")
    return parser


if __name__ == "__main__":
    main(_parser().parse_args())
