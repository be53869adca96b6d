"""Synthetic inputs for the GME hot path (NumPy only, deterministic on every host).

The reference's sample videos live on Google Drive and are not available offline, so
parity tests and the bench feed textured frames warped by known motion to both the
oracle and the CUDA path.  Everything here is integer / float64 NumPy arithmetic with no
library filters, so the same seed gives the same bytes in the build container and on the
GPU box.  The generator's arithmetic is not part of parity; only the bytes are.
"""
from __future__ import annotations

import numpy as np


def _box_blur_u32(a: np.ndarray, box: int) -> np.ndarray:
    """box x box mean filter (valid region) with round-half-up, via 2-D cumulative sums."""
    c = np.zeros((a.shape[0] + 1, a.shape[1] + 1), np.uint64)
    c[1:, 1:] = a.astype(np.uint64).cumsum(0).cumsum(1)
    s = c[box:, box:] - c[:-box, box:] - c[box:, :-box] + c[:-box, :-box]
    return ((s + (box * box) // 2) // (box * box)).astype(np.uint32)


def pan_pair(H, W, dx, dy, seed=0, pad=64, box=4):
    """The known-answer input of SURVEY.md Appendix B (4x4 box-filtered noise, exact crop pan)."""
    rng = np.random.default_rng(seed)
    n = rng.integers(0, 256, (H + 2 * pad + box, W + 2 * pad + box), dtype=np.uint8).astype(np.uint32)
    c = np.zeros((n.shape[0] + 1, n.shape[1] + 1), np.uint32)
    c[1:, 1:] = n.cumsum(0).cumsum(1)
    s = c[box:, box:] - c[:-box, box:] - c[box:, :-box] + c[:-box, :-box]
    canvas = ((s + box * box // 2) // (box * box)).astype(np.uint8)[:H + 2 * pad, :W + 2 * pad]
    prev = np.ascontiguousarray(canvas[pad:pad + H, pad:pad + W])
    cur = np.ascontiguousarray(canvas[pad - dy:pad - dy + H, pad - dx:pad - dx + W])
    return prev, cur


def texture(H, W, seed=0) -> np.ndarray:
    """Smooth multi-octave texture, uint8[H, W].

    Three octaves of uniform noise, each blurred by three passes of a box filter
    (a close approximation of a Gaussian of sigma ~ 1, 3, 9 px), mixed .25/.35/.40 and
    stretched to 0..255.  A smooth texture matters: pattern searches get trapped on
    white-ish noise, which is still valid for parity but a poor demonstration."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((H, W), np.float64)
    for box, weight in ((2, .25), (6, .35), (18, .40)):
        m = 3 * (box - 1)
        a = rng.integers(0, 1 << 16, (H + m, W + m), dtype=np.uint32)
        for _ in range(3):
            a = _box_blur_u32(a, box)
        a = a[:H, :W].astype(np.float64)
        a = (a - a.mean()) / (a.std() + 1e-12)
        acc += weight * a
    lo, hi = np.percentile(acc, .5), np.percentile(acc, 99.5)
    return np.clip(np.rint((acc - lo) * (255.0 / (hi - lo))), 0, 255).astype(np.uint8)


def pan_sequence(n_frames, H, W, step=(2, 1), seed=3) -> np.ndarray:
    """Exact-crop panning sequence uint8[n, H, W]: the window slides step=(cols, rows) px/frame."""
    sx, sy = step
    canvas = texture(H + abs(sy) * n_frames, W + abs(sx) * n_frames, seed)
    out = np.empty((n_frames, H, W), np.uint8)
    for k in range(n_frames):
        y0 = k * sy if sy >= 0 else (n_frames - 1 - k) * -sy
        x0 = k * sx if sx >= 0 else (n_frames - 1 - k) * -sx
        out[k] = canvas[y0:y0 + H, x0:x0 + W]
    return out


def warp_affine(src: np.ndarray, M: np.ndarray, out_shape=None) -> np.ndarray:
    """dst(y, x) = bilinear src at (M @ [x, y, 1]) with reflect-101 borders, float64 -> uint8."""
    H, W = out_shape or src.shape
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    sx = M[0, 0] * xs + M[0, 1] * ys + M[0, 2]
    sy = M[1, 0] * xs + M[1, 1] * ys + M[1, 2]
    x0, y0 = np.floor(sx), np.floor(sy)
    fx, fy = sx - x0, sy - y0

    def refl(i, n):
        i = np.abs(i.astype(np.int64))
        period = 2 * n - 2
        i = i % period
        return np.where(i >= n, period - i, i)

    x0i, x1i = refl(x0, src.shape[1]), refl(x0 + 1, src.shape[1])
    y0i, y1i = refl(y0, src.shape[0]), refl(y0 + 1, src.shape[0])
    s = src.astype(np.float64)
    top = s[y0i, x0i] * (1 - fx) + s[y0i, x1i] * fx
    bot = s[y1i, x0i] * (1 - fx) + s[y1i, x1i] * fx
    return np.clip(np.rint(top * (1 - fy) + bot * fy), 0, 255).astype(np.uint8)


def zoom_rotate_sequence(n_frames, H, W, zoom_per_frame=0.002, deg_per_frame=0.1, seed=4) -> np.ndarray:
    """Frame k = texture zoomed by (1 + zoom*k) and rotated by deg*k about the centre."""
    base = texture(H, W, seed)
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    out = np.empty((n_frames, H, W), np.uint8)
    for k in range(n_frames):
        s = 1.0 / (1.0 + zoom_per_frame * k)
        t = np.deg2rad(deg_per_frame * k)
        A = s * np.array([[np.cos(t), np.sin(t)], [-np.sin(t), np.cos(t)]])
        M = np.empty((2, 3))
        M[:, :2] = A
        M[:, 2] = np.array([cx, cy]) - A @ np.array([cx, cy])
        out[k] = warp_affine(base, M)
    return out


def affine_sequence(n_frames, H, W, seed=5, max_corner_px=12.0) -> np.ndarray:
    """General affine motion per frame (translation, scale, rotation, shear from default_rng(seed)),
    scaled so that no frame corner moves by more than ``max_corner_px`` per frame-distance 1."""
    rng = np.random.default_rng(seed)
    base = texture(H, W, seed)
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    tx, ty = rng.uniform(-1, 1, 2)
    sc, rot, sh = rng.uniform(-1, 1, 3)
    out = np.empty((n_frames, H, W), np.uint8)
    half_diag = np.hypot(cx, cy)
    unit = max_corner_px / 3.0
    for k in range(n_frames):
        a = 1.0 + (sc * unit / half_diag) * k
        t = (rot * unit / half_diag) * k
        h = (sh * unit / half_diag) * k
        A = a * np.array([[np.cos(t), np.sin(t) + h], [-np.sin(t), np.cos(t)]])
        M = np.empty((2, 3))
        M[:, :2] = A
        M[:, 2] = np.array([cx, cy]) - A @ np.array([cx, cy]) + np.array([tx, ty]) * unit * k / 3.0
        out[k] = warp_affine(base, M)
    return out
