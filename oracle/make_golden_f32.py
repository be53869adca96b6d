"""Golden cases for MSE with block_size > 16, where the reference's float32 cost ROUNDS.

    python oracle/make_golden_f32.py        # writes tests/golden/bbme_f32_rounding.npz (build container only)

compute_dfd sums the squared differences with np.sum over a float32 array (bbme.py:61-64,94): NumPy's pairwise
float32 reduction.  Its partial sums stay exact integers up to 2^24; beyond that -- possible only for squared
differences with block_size >= 17 -- they round to even at every node of the pairwise tree, and near-equal candidates
can compare differently from their exact integer costs.  The inputs here are high-contrast (two grey levels 255
apart, plus a little noise) so that block costs do exceed 2^24 and ties/near-ties between candidates are common.
Test infrastructure: runs the UNMODIFIED reference through oracle/ref_shim.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def inputs():
    rng = np.random.default_rng(77)
    out = []
    for (H, W) in ((70, 96), (64, 64), (97, 131)):
        a = (rng.integers(0, 2, (H + 8, W + 8)) * 255).astype(np.int32)
        dy, dx = int(rng.integers(-3, 4)), int(rng.integers(-3, 4))
        p = a[4:4 + H, 4:4 + W]
        c = a[4 - dy:4 - dy + H, 4 - dx:4 - dx + W] ^ (rng.integers(0, 40, (H, W)) == 0) * 255   # 2.5 % of pixels flipped
        p = np.clip(p + rng.integers(-1, 2, p.shape) * (p > 0), 0, 255)                         # 253..255: near-ties
        out.append((np.ascontiguousarray(p, dtype=np.uint8), np.ascontiguousarray(np.clip(c, 0, 255), dtype=np.uint8)))
    # exact integer TIES that the float32 tree breaks: anchor blocks all 0, candidates 255 with sparse 254s -- the cost of
    # a candidate depends only on how many 254s it holds, so many candidates tie exactly (first in scan order would win),
    # but sums above 2^24 round differently depending on WHERE the 254s sit in the pairwise tree
    for (H, W, density) in ((60, 84, 10), (80, 72, 4), (130, 170, 6), (100, 200, 16)):
        c = (255 - (rng.integers(0, density, (H, W)) == 0)).astype(np.uint8)
        out.append((np.zeros((H, W), np.uint8), c))
    return out


def main():
    utils, bbme, motion = ref_shim.load()
    pack, k = {}, 0
    for i, (p, c) in enumerate(inputs()):
        pack[f"p{i}"], pack[f"c{i}"] = p, c
        for bs, sw in ((20, 3), (24, 5), (32, 2), (17, 4)) + (((40, 3), (28, 6)) if i >= 5 else ()):
            if bs >= min(p.shape):
                continue
            for sp in (0, 1, 2, 3):
                mf = bbme.get_motion_field(p, c, block_size=bs, search_window=sw, searching_procedure=sp, pnorm_distance=1)
                pack[f"mf{k}"], pack[f"a{k}"] = mf, np.array([bs, sw, sp, 1, i])
                k += 1
    pack["n"] = np.array(k)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bbme_f32_rounding.npz"), **pack)
    print(k, "cases")


if __name__ == "__main__":
    main()
