"""Generates tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python oracle/make_golden.py            # rewrites tests/golden/

The reference is pure Python and cannot travel to the GPU box, so its outputs on small
seeded inputs are committed as fixtures together with this script.  Inputs are stored in
the fixtures too, so the tests never depend on regenerating them.  Test infrastructure.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "global-motion-estimation_b200"))

import gme_synth as S          # noqa: E402
import ref_shim                # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def capture_locals(func_name, names):
    """Profile hook that records locals of the reference function `func_name` when it returns."""
    store = []

    def prof(frame, event, arg):
        if event == "return" and frame.f_code.co_name == func_name:
            store.append({n: np.array(frame.f_locals[n]) for n in names if n in frame.f_locals})

    return store, prof


def gme_with_intermediates(motion, previous, current):
    """Runs motion.global_motion_estimation and records, per robust level, the BBME field, the
    model field, the outlier mask and the threshold straight from the reference's own locals."""
    store, prof = capture_locals(
        "best_affine_parameters_robust",
        ["gt_motion_field", "old_params_motion_field", "outlier", "threshold_value"])
    dense_store, dense_prof = capture_locals("compute_first_parameters", ["dense_motion_field"])

    def both(frame, event, arg):
        prof(frame, event, arg)
        dense_prof(frame, event, arg)

    sys.setprofile(both)
    try:
        params = motion.global_motion_estimation(previous, current)
    finally:
        sys.setprofile(None)
    return params, dense_store[0]["dense_motion_field"], store


def main():
    utils, bbme, motion = ref_shim.load()
    os.makedirs(GOLDEN, exist_ok=True)

    # ---- 1. SURVEY Appendix B known-answer input: all procedures x norms ------------
    prev, cur = S.pan_pair(240, 320, 5, -3)
    out = {"prev": prev, "cur": cur}
    for pn in (0, 1):
        for sp in (0, 1, 2, 3):
            out[f"mf_pn{pn}_sp{sp}"] = bbme.get_motion_field(
                prev, cur, block_size=12, search_window=12, searching_procedure=sp, pnorm_distance=pn)
    out["static_bs12"] = bbme.get_motion_field(prev, prev, 12, 12, 3, 0)
    out["static_bs16"] = bbme.get_motion_field(prev, prev, 16, 12, 3, 1)
    np.savez_compressed(os.path.join(GOLDEN, "bbme_kat.npz"), **out)
    print("bbme_kat done")

    # ---- 2. random geometries: sizes not multiples of bs, all block sizes/windows ----
    rng = np.random.default_rng(2024)
    cases = []
    geoms = [(48, 64, 2, 2), (50, 70, 4, 4), (61, 83, 8, 7), (72, 96, 10, 4), (75, 101, 12, 12),
             (96, 128, 16, 16), (64, 80, 16, 5), (40, 56, 4, 1), (90, 60, 12, 3), (33, 47, 16, 8),
             (70, 90, 6, 6), (57, 64, 3, 9)]
    for gi, (H, W, bs, sw) in enumerate(geoms):
        for kind in ("noise", "smooth"):
            if kind == "noise":
                a = rng.integers(0, 256, (H + 16, W + 16), dtype=np.uint8)
            else:
                a = S.texture(H + 16, W + 16, seed=100 + gi)
            dy, dx = int(rng.integers(-5, 6)), int(rng.integers(-5, 6))
            p = np.ascontiguousarray(a[8:8 + H, 8:8 + W])
            c = np.ascontiguousarray(a[8 - dy:8 - dy + H, 8 - dx:8 - dx + W])
            if kind == "noise":      # also some per-pixel noise so ties are rarer but present
                c = np.clip(c.astype(int) + rng.integers(-2, 3, c.shape), 0, 255).astype(np.uint8)
            for pn in (0, 1):
                for sp in (0, 1, 2, 3):
                    mf = bbme.get_motion_field(p, c, block_size=bs, search_window=sw,
                                               searching_procedure=sp, pnorm_distance=pn)
                    cases.append((p, c, bs, sw, sp, pn, mf))
    # flat / tie-heavy inputs: every candidate ties, first in scan order must win
    flat = np.full((40, 52), 77, np.uint8)
    steps = (np.arange(52)[None, :] // 8 * 16 + np.zeros((40, 1))).astype(np.uint8)
    for p, c in ((flat, flat), (steps, steps), (flat, steps)):
        for bs, sw in ((4, 3), (8, 8)):
            for pn in (0, 1):
                for sp in (0, 1, 2, 3):
                    mf = bbme.get_motion_field(p, c, block_size=bs, search_window=sw,
                                               searching_procedure=sp, pnorm_distance=pn)
                    cases.append((p, c, bs, sw, sp, pn, mf))
    pack = {"n": np.array(len(cases))}
    inputs = {}
    for k, (p, c, bs, sw, sp, pn, mf) in enumerate(cases):
        key = (p.tobytes(), c.tobytes(), p.shape)
        if key not in inputs:
            inputs[key] = len(inputs)
            pack[f"p{inputs[key]}"], pack[f"c{inputs[key]}"] = p, c
        pack[f"mf{k}"] = mf
        pack[f"a{k}"] = np.array([bs, sw, sp, pn, inputs[key]])
    np.savez_compressed(os.path.join(GOLDEN, "bbme_random.npz"), **pack)
    print("bbme_random done:", len(cases), "cases")

    # ---- 3. pyramid (cv2.pyrDown through utils.get_pyramids) -------------------------
    pack = {}
    for k, (H, W) in enumerate([(240, 320), (7, 9), (33, 47), (64, 64), (101, 75), (2, 2), (1, 5), (120, 180)]):
        img = rng.integers(0, 256, (H, W), dtype=np.uint8)
        pyr = utils.get_pyramids(img)
        pack[f"img{k}"], pack[f"l1_{k}"], pack[f"l0_{k}"] = img, pyr[1], pyr[0]
    pack["n"] = np.array(k + 1)
    np.savez_compressed(os.path.join(GOLDEN, "pyramid.npz"), **pack)
    print("pyramid done")

    # ---- 4. full GME pipeline with intermediates ---------------------------------------
    seqs = {
        "kat": S.pan_pair(240, 320, 5, -3),
        "pan": tuple(S.pan_sequence(4, 240, 352, step=(2, 1), seed=3)[[0, 3]]),
        "zoomrot": tuple(S.zoom_rotate_sequence(4, 208, 288, 0.004, 0.3, seed=4)[[0, 3]]),
        "odd": tuple(S.affine_sequence(4, 250, 330, seed=5, max_corner_px=6.0)[[0, 3]]),
    }
    pack = {"names": np.array(list(seqs))}
    for name, (p, c) in seqs.items():
        params, dense, levels = gme_with_intermediates(motion, p, c)
        shape = (p.shape[0] // motion.BBME_BLOCK_SIZE, p.shape[1] // motion.BBME_BLOCK_SIZE)
        model = motion.get_motion_field_affine(shape, params)
        comp = motion.compensate_frame(p, model)
        psnr = utils.PSNR(c, comp)
        pack[f"{name}_prev"], pack[f"{name}_cur"] = p, c
        pack[f"{name}_params"], pack[f"{name}_dense"] = params, dense
        pack[f"{name}_first"] = motion.compute_first_parameters(dense)
        for li, lv in enumerate(levels, start=1):
            pack[f"{name}_gt{li}"] = lv["gt_motion_field"]
            pack[f"{name}_model{li}"] = lv["old_params_motion_field"]
            pack[f"{name}_outlier{li}"] = lv["outlier"]
            pack[f"{name}_thr{li}"] = lv["threshold_value"]
        pack[f"{name}_modelfield"], pack[f"{name}_comp"] = model, comp
        pack[f"{name}_psnr"] = np.array([psnr.real, psnr.imag]) if isinstance(psnr, complex) else np.array([-1.0, 0.0])
        pack[f"{name}_mc"] = motion.motion_compensation(p, c)
        pack[f"{name}_nonrobust"] = motion.best_affine_parameters(p, c)
        print("gme", name, params)
    # search overrides on the bs=16 levels (configs 4/5 of BASELINE.json)
    p, c = seqs["zoomrot"]
    for sp, sw in ((0, 6), (1, 16), (2, 16)):
        undo = ref_shim.with_search_override(motion, bbme, sp, sw)
        try:
            params, dense, levels = gme_with_intermediates(motion, p, c)
        finally:
            undo()
        pack[f"ovr_sp{sp}_sw{sw}_params"] = params
        for li, lv in enumerate(levels, start=1):
            pack[f"ovr_sp{sp}_sw{sw}_gt{li}"] = lv["gt_motion_field"]
            pack[f"ovr_sp{sp}_sw{sw}_outlier{li}"] = lv["outlier"]
        print("override", sp, sw, params)
    np.savez_compressed(os.path.join(GOLDEN, "gme_pipeline.npz"), **pack)

    # ---- 5. compensation / PSNR / model field / hierarchical wrapper -------------------
    pack = {}
    frame = S.texture(100, 132, seed=9)
    other = S.texture(100, 132, seed=10)
    k = 0
    for (R, C, lo, hi) in [(6, 8, -20, 21), (12, 16, -9, 10), (3, 4, -140, 141), (100, 132, -3, 4), (7, 9, -5, 6)]:
        mf = rng.integers(lo, hi, (R, C, 2)).astype(np.int16)
        pack[f"cf_mf{k}"], pack[f"cf_out{k}"] = mf, motion.compensate_frame(frame, mf)
        k += 1
    pack["cf_n"], pack["cf_frame"], pack["cf_other"] = np.array(k), frame, other
    ps = utils.PSNR(frame, other)
    pack["psnr_fo"] = np.array([ps.real, ps.imag])
    pack["psnr_same"] = np.array(utils.PSNR(frame, frame))
    plist = [np.array([3.4, 0.017, -0.024, -1.73, -0.025, 0.0134]),
             np.array([0.5, 0.5, 0.25, -0.5, 1.5, -2.5]),            # exact .5 ties -> half-to-even
             np.array([1.25, 0, 0, -7.5, 0, 0], dtype=np.float32),
             np.array([-31.7, 0.31, 0.77, 12.2, -0.9, 0.05])]
    for k, p in enumerate(plist):
        pack[f"af_p{k}"], pack[f"af_out{k}"] = p, motion.get_motion_field_affine((9, 13), p)
    pack["af_n"] = np.array(len(plist))
    p, c = seqs["pan"]
    for k, (bs, sw, sp) in enumerate([(10, 4, 3), (12, 8, 1), (16, 4, 2), (8, 3, 0), (11, 5, 2)]):
        pack[f"hw_a{k}"] = np.array([bs, sw, sp])
        try:
            pack[f"hw_out{k}"] = bbme.hierarchical_wrapper(p, c, block_size=bs, search_window=sw,
                                                           searching_procedure=sp)
        except ValueError:      # bbme.py:596-604 pads one row OR one column; both differing raises
            pack[f"hw_out{k}"] = np.array("ValueError")
    pack["hw_n"], pack["hw_prev"], pack["hw_cur"] = np.array(k + 1), p, c
    np.savez_compressed(os.path.join(GOLDEN, "misc.npz"), **pack)
    print("misc done")
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
