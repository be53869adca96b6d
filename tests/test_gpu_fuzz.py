"""GPU: seeded random geometries against the oracle -- frame sizes that are not multiples of anything, windows
larger than the frame, flat and tie-heavy content, every procedure / norm / tiled block size, and whole pipelines."""
import numpy as np
import pytest
import torch

import gme_oracle as O
import gme_synth as S

pytestmark = pytest.mark.gpu

import os

PARAM_TOL = dict(atol=1e-9, rtol=1e-9)
SCALE = int(os.environ.get("GME_FUZZ_SCALE", "1"))       # GME_FUZZ_SCALE=20 for a long soak run


@pytest.fixture(scope="module")
def D():
    import gme_device
    gme_device.require_cuda()
    return gme_device


def _content(rng, H, W, kind):
    if kind == 0:                                         # smooth texture, small pan
        base = S.texture(H + 16, W + 16, seed=int(rng.integers(1 << 30)))
        dy, dx = (int(v) for v in rng.integers(-6, 7, 2))
        return (np.ascontiguousarray(base[8:8 + H, 8:8 + W]),
                np.ascontiguousarray(base[8 - dy:8 - dy + H, 8 - dx:8 - dx + W]))
    if kind == 1:                                         # white noise: many local minima
        return (rng.integers(0, 256, (H, W), dtype=np.uint8), rng.integers(0, 256, (H, W), dtype=np.uint8))
    if kind == 2:                                         # few grey levels: ties everywhere
        a = (rng.integers(0, 3, (H, W)) * 100).astype(np.uint8)
        return a, np.roll(a, (int(rng.integers(-3, 4)), int(rng.integers(-3, 4))), (0, 1))
    flat = np.full((H, W), int(rng.integers(0, 256)), np.uint8)        # constant: every cost ties
    return flat, flat.copy()


@pytest.mark.parametrize("seed", range(12 * SCALE))
def test_bbme_fuzz(D, seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(6):
        bs = int(rng.choice([2, 4, 8, 12, 16, 16, 16, 5, 20]))
        H = int(rng.integers(bs + 1, 150))
        W = int(rng.integers(bs + 1, 200))
        sw = int(rng.integers(0, 20))
        prev, cur = _content(rng, H, W, int(rng.integers(0, 4)))
        pp, cp = D.Planes.from_host(prev), D.Planes.from_host(cur)
        for sp in range(4):
            pn = int(rng.integers(0, 2))                  # (MSE beyond bs 16: float32-rounded like the reference)
            got = D.motion_field(pp, cp, bs, sw, sp, pn)[0].cpu().numpy()
            want = O.get_motion_field(prev, cur, bs, sw, sp, pn)
            np.testing.assert_array_equal(got, want, err_msg=f"H={H} W={W} bs={bs} sw={sw} sp={sp} pn={pn}")


def _near_half(p, R, C, eps):
    ii, jj = np.mgrid[0:R, 0:C]
    return any(np.abs(a - np.floor(a) - 0.5).min() < eps
               for a in (p[0] + p[1] * ii + p[2] * jj, p[3] + p[4] * ii + p[5] * jj))


def _rounding_tie(inter, final=None, final_shape=None, eps=1e-9):
    """True when some model vector a0 + a1*i + a2*j of level 1 or 2 lies within eps of a half-integer: Python's round()
    then depends on the last bits of the previous level's parameters, which differ between any two float64
    implementations of the fit (NumPy/BLAS builds included)."""
    prev = np.asarray(inter[0]["first"], dtype=np.float64)
    for level in (1, 2):
        p = prev.copy()
        p[0] *= 2
        p[3] *= 2
        R, C = inter[level]["gt"].shape[:2]
        # level 1 starts from the first estimate [2*float32(mean), 0, 0, ...]: a0 + 0*i + 0*j is EXACT in float64, so a
        # tie there is an exact .5 that round-half-even resolves identically everywhere -- a hard check, never skipped
        exact = level == 1 and not p[[1, 2, 4, 5]].any()
        if not exact and _near_half(p, R, C, eps):
            return True
        prev = np.asarray(inter[level]["params"], dtype=np.float64)
    # ... and the model field of the FINAL parameters, which steers the compensation (results.py:52-54)
    return final is not None and _near_half(np.asarray(final, dtype=np.float64), final_shape[0], final_shape[1], eps)


_TIES = {"pairs": 0, "skipped": 0}


@pytest.mark.parametrize("seed", range(6 * SCALE))
def test_pipeline_fuzz(D, seed):
    rng = np.random.default_rng(2000 + seed)
    H, W = int(rng.integers(140, 400)), int(rng.integers(140, 520))      # both levels need at least one 16 x 16 block
    n, d = 3, int(rng.integers(1, 3))
    seq = S.zoom_rotate_sequence(n + d, H, W, zoom_per_frame=float(rng.uniform(0, 0.01)),
                                 deg_per_frame=float(rng.uniform(-0.5, 0.5)), seed=int(rng.integers(1 << 30)))
    sp, sw = [(3, 2), (1, 7), (2, 9), (0, 5)][seed % 4]
    pipe = D.gme_sequence(D.Planes.from_host(seq), d, procedure=sp, window=sw)
    torch.cuda.synchronize()
    for k in range(n):
        want, inter = O.global_motion_estimation(seq[k], seq[k + d], procedure=sp, window=sw, return_intermediates=True)
        np.testing.assert_array_equal(pipe.intermediate(0)[k].cpu().numpy(), inter[0]["dense"])
        np.testing.assert_array_equal(pipe.intermediate(2)[k].cpu().numpy(), inter[2]["gt"])
        _TIES["pairs"] += 1
        if _rounding_tie(inter, want, (H // 16, W // 16)):
            _TIES["skipped"] += 1
            continue      # a model vector sits on a .5 tie: float64 round-off decides it, in the reference too
        np.testing.assert_array_equal(pipe.intermediate(4)[k].cpu().numpy().astype(bool), inter[2]["outlier"])
        np.testing.assert_allclose(pipe.params[k].cpu().numpy(), want, **PARAM_TOL)
        comp = O.compensate_frame(seq[k], O.get_motion_field_affine((H // 16, W // 16), want))
        np.testing.assert_array_equal(pipe.comp.to_host()[k], comp)
        assert int(pipe.sse[k].item()) == O.sse(seq[k + d], comp)


def test_rounding_tie_skips_stay_rare():
    """(runs after test_pipeline_fuzz) the float comparisons may be skipped only for the rare pairs whose model vector
    sits on an inexact .5 tie; if that ever becomes common the guard is hiding something."""
    assert _TIES["pairs"] > 0
    assert _TIES["skipped"] <= max(1, _TIES["pairs"] // 20), _TIES
