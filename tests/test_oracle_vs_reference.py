"""CPU, build container only: the oracle against the LIVE reference on fresh random inputs
(skipped where /root/reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest

import gme_oracle as O
import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


def test_random_fields_all_procedures(ref):
    utils, bbme, motion = ref
    rng = np.random.default_rng(11)
    for trial in range(6):
        H, W = int(rng.integers(30, 70)), int(rng.integers(30, 80))
        bs = int(rng.choice([2, 3, 4, 5, 8, 12, 16, 18, 20]))
        sw = int(rng.integers(0, 9))
        base = rng.integers(0, 256, (H + 8, W + 8), dtype=np.uint8)
        dy, dx = rng.integers(-3, 4, 2)
        prev = np.ascontiguousarray(base[4:4 + H, 4:4 + W])
        cur = np.ascontiguousarray(base[4 - dy:4 - dy + H, 4 - dx:4 - dx + W])
        for sp in range(4):
            if sp == 3 and (H <= bs or W <= bs):
                continue
            for pn in (0, 1):
                want = bbme.get_motion_field(prev, cur, bs, sw, sp, pn)
                got = O.get_motion_field(prev, cur, bs, sw, sp, pn)
                np.testing.assert_array_equal(got, want, err_msg=f"H={H} W={W} bs={bs} sw={sw} sp={sp} pn={pn}")


def test_compensate_and_psnr(ref):
    utils, bbme, motion = ref
    rng = np.random.default_rng(12)
    frame = rng.integers(0, 256, (50, 66), dtype=np.uint8)
    other = rng.integers(0, 256, (50, 66), dtype=np.uint8)
    for R, C in ((3, 4), (5, 5), (50, 66), (7, 20)):
        mf = rng.integers(-30, 31, (R, C, 2)).astype(np.int16)
        np.testing.assert_array_equal(O.compensate_frame(frame, mf), motion.compensate_frame(frame, mf))
    assert O.PSNR(frame, other) == utils.PSNR(frame, other)
