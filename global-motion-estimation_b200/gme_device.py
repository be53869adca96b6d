"""Device-side host logic of the B200 GME path: PyTorch owns the HBM buffers and the streams,
the kernels live in libgme_b200.so (gme_native).  Everything here needs a CUDA device; there
is no CPU fallback.

Data layout in HBM
  * frames: ``Planes`` = uint8[n, H, pitch] with pitch = W rounded up to 16 bytes, so that every
    row is 16-byte aligned (128-bit loads, TMA boxes).  A sequence is ONE such buffer; the
    (previous, current) pairs at frame distance d are two views of it offset by d planes, so
    no frame is ever copied to form a pair.
  * motion fields: int32[n, R, C, 2] (channel 0 = column displacement, 1 = row displacement).
  * affine parameters: float64[n, 6] = [a0, a1, a2, b0, b1, b2] per pair.
"""
from __future__ import annotations

import cmath
import ctypes
from dataclasses import dataclass

import numpy as np
import torch

import gme_native as N

PITCH_ALIGN = 16
OUTLIER_FRACTION = .3            # motion.MOTION_VECTOR_ERROR_THRESHOLD_PERCENTAGE (motion.py:10)


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("the GME kernels need a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _round_up(v: int, a: int) -> int:
    return (v + a - 1) // a * a


@dataclass
class Planes:
    """n grayscale planes resident in HBM: ``t`` is uint8[n, H, pitch]; the first W columns are pixels."""
    t: torch.Tensor
    W: int

    @property
    def n(self) -> int:
        return self.t.shape[0]

    @property
    def H(self) -> int:
        return self.t.shape[1]

    @property
    def pitch(self) -> int:
        return self.t.stride(1)

    @property
    def stride(self) -> int:
        return self.t.stride(0)

    @property
    def ptr(self) -> int:
        return self.t.data_ptr()

    @staticmethod
    def empty(n: int, H: int, W: int, device=None) -> "Planes":
        device = device or require_cuda()
        return Planes(torch.empty((n, H, _round_up(W, PITCH_ALIGN)), dtype=torch.uint8, device=device), W)

    @staticmethod
    def from_host(frames, device=None, non_blocking: bool = False) -> "Planes":
        """frames: uint8 ndarray/tensor [H, W] or [n, H, W] (host).  One host->device copy."""
        device = device or require_cuda()
        if isinstance(frames, np.ndarray):
            a = np.ascontiguousarray(frames, dtype=np.uint8)
            src = torch.from_numpy(a)
        else:
            src = frames
        if src.dim() == 2:
            src = src.unsqueeze(0)
        if src.dim() != 3 or src.dtype != torch.uint8:
            raise ValueError("expected uint8 frames of shape [H, W] or [n, H, W]")
        n, H, W = src.shape
        p = Planes.empty(n, H, W, device)
        p.t[:, :, :W].copy_(src, non_blocking=non_blocking)
        return p

    def view(self, start: int, stop: int) -> "Planes":
        return Planes(self.t[start:stop], self.W)

    def pixels(self) -> torch.Tensor:
        return self.t[:, :, :self.W]

    def to_host(self) -> np.ndarray:
        return self.pixels().cpu().numpy()


# --------------------------------------------------------------------------- single kernels
def motion_field(prev: Planes, cur: Planes, block_size: int, search_window: int, procedure: int,
                 pnorm: int) -> torch.Tensor:
    """bbme.get_motion_field for n pairs -> int32[n, H//bs, W//bs, 2] on the device."""
    if prev.t.shape != cur.t.shape or prev.W != cur.W or prev.pitch != cur.pitch:
        raise ValueError("previous and current must have the same geometry")
    if not 0 <= int(procedure) <= 3 or not 0 <= int(pnorm) <= 1:
        raise IndexError("list index out of range")      # bbme.searching_procedures / pnorm_distances
    bs = int(block_size)
    if bs <= 0:
        raise ZeroDivisionError("division by zero") if bs == 0 else ValueError("block_size must be positive")
    R, C = int(prev.H / bs), int(prev.W / bs)
    field = torch.zeros((prev.n, R, C, 2), dtype=torch.int32, device=prev.t.device)
    if R * C:
        N.check(N.lib.gme_bbme_motion_field(prev.ptr, prev.stride, cur.ptr, cur.stride, prev.n, prev.H, prev.W,
                                            prev.pitch, bs, int(search_window), int(procedure), int(pnorm),
                                            field.data_ptr(), _stream()), "gme_bbme_motion_field")
    return field


def pyr_down(src: Planes) -> Planes:
    """One cv2.pyrDown level for n planes."""
    dst = Planes.empty(src.n, (src.H + 1) // 2, (src.W + 1) // 2, src.t.device)
    N.check(N.lib.gme_pyr_down(src.ptr, src.pitch, src.stride, dst.ptr, dst.pitch, dst.stride, src.n, src.H, src.W,
                               _stream()), "gme_pyr_down")
    return dst


def first_parameters(dense_field: torch.Tensor) -> torch.Tensor:
    n, R, C, _ = dense_field.shape
    params = torch.empty((n, 6), dtype=torch.float64, device=dense_field.device)
    N.check(N.lib.gme_first_parameters(dense_field.data_ptr(), n, R, C, params.data_ptr(), _stream()),
            "gme_first_parameters")
    return params


def affine_fit(gt_field: torch.Tensor, level_shape, params: torch.Tensor, robust: bool = True,
               project: bool = False, pct: float = .3, intermediates: bool = False):
    """In-place update of params (float64[n, 6]).  Returns (params, status[, outlier, threshold, model])."""
    n, R, C, _ = gt_field.shape
    dev = gt_field.device
    status = torch.zeros((n,), dtype=torch.int32, device=dev)
    outlier = threshold = model = None
    if intermediates:
        outlier = torch.zeros((n, R, C), dtype=torch.uint8, device=dev)
        threshold = torch.zeros((n,), dtype=torch.int32, device=dev)
        model = torch.zeros((n, R, C, 2), dtype=torch.int16, device=dev)
    N.check(N.lib.gme_affine_fit(gt_field.data_ptr(), n, R, C, int(level_shape[0]), int(level_shape[1]), float(pct),
                                 int(robust), int(project), params.data_ptr(),
                                 outlier.data_ptr() if intermediates else None,
                                 threshold.data_ptr() if intermediates else None,
                                 model.data_ptr() if intermediates else None, status.data_ptr(), _stream()),
            "gme_affine_fit")
    if intermediates:
        return params, status, outlier, threshold, model
    return params, status


def affine_field(params: torch.Tensor, R: int, C: int) -> torch.Tensor:
    params = params.contiguous()                         # (Pipeline.params is a strided view of its row buffer)
    n = params.shape[0]
    field = torch.zeros((n, R, C, 2), dtype=torch.int16, device=params.device)
    N.check(N.lib.gme_affine_field(params.data_ptr(), n, R, C, field.data_ptr(), _stream()), "gme_affine_field")
    return field


def compensate(frame: Planes, field: torch.Tensor, cur: Planes | None = None, want_diffs: bool = False):
    """motion.compensate_frame for n planes; with ``cur`` also the squared error sums (int64[n]); with
    ``want_diffs`` also the two difference images of results.py:78-83 (|cur - frame|, |cur - comp|)."""
    if field.dtype not in (torch.int16, torch.int32):
        raise TypeError("motion field must be int16 or int32")
    field = field.contiguous()
    n, R, C = field.shape[0], field.shape[1], field.shape[2]
    comp = Planes.empty(frame.n, frame.H, frame.W, frame.t.device)
    sse = torch.zeros((frame.n,), dtype=torch.int64, device=frame.t.device) if cur is not None else None
    if want_diffs:
        if cur is None:
            raise ValueError("the difference images need the current frame")
        dprev = Planes.empty(frame.n, frame.H, frame.W, frame.t.device)
        dcomp = Planes.empty(frame.n, frame.H, frame.W, frame.t.device)
        N.check(N.lib.gme_compensate_diffs(frame.ptr, frame.pitch, frame.stride, field.data_ptr(),
                                           int(field.dtype == torch.int16), R, C, cur.ptr, cur.pitch, cur.stride,
                                           comp.ptr, comp.pitch, comp.stride, dprev.ptr, dcomp.ptr, dprev.pitch,
                                           dprev.stride, frame.n, frame.H, frame.W, sse.data_ptr(), _stream()),
                "gme_compensate_diffs")
        return comp, sse, dprev, dcomp
    N.check(N.lib.gme_compensate(frame.ptr, frame.pitch, frame.stride, field.data_ptr(),
                                 int(field.dtype == torch.int16), R, C,
                                 cur.ptr if cur is not None else None, cur.pitch if cur is not None else 0,
                                 cur.stride if cur is not None else 0, comp.ptr, comp.pitch, comp.stride,
                                 frame.n, frame.H, frame.W, sse.data_ptr() if sse is not None else None, _stream()),
            "gme_compensate")
    return comp, sse


def hierarchical_shapes_merge(coarse_shape, fine_shape) -> bool:
    """True when bbme.hierarchical_wrapper can merge the two levels (bbme.py:596-604): the upsampled coarse field
    equals the fine one, or lacks exactly one row, or else exactly one column."""
    (rc, cc), (r, c) = coarse_shape, fine_shape
    return (2 * rc, 2 * cc) == (r, c) or (2 * rc + 1, 2 * cc) == (r, c) or (2 * rc, 2 * cc + 1) == (r, c)


def hierarchical_field(prev: Planes, cur: Planes, block_size: int, search_window: int, procedure: int) -> torch.Tensor:
    """bbme.hierarchical_wrapper (bbme.py:549-605) for n pairs, entirely on the device: two pyramid levels per frame,
    the coarsest field with ``procedure``, the finer two with diamond search, each merged with the upsampled coarser
    one -> float64[n, H//bs, W//bs, 2].  Raises ValueError for geometries the reference cannot merge either."""
    p1, c1 = pyr_down(prev), pyr_down(cur)
    p0, c0 = pyr_down(p1), pyr_down(c1)
    field = motion_field(p0, c0, block_size, search_window, procedure, N.PNORM_MSE)
    for lp, lc in ((p1, c1), (prev, cur)):
        fine = motion_field(lp, lc, block_size, search_window, N.SEARCH_DIAMOND, N.PNORM_MSE)
        if not hierarchical_shapes_merge(field.shape[1:3], fine.shape[1:3]):
            raise ValueError(f"operands could not be broadcast together with shapes {tuple(field.shape[1:])} {tuple(fine.shape[1:])}")
        out = torch.empty(fine.shape, dtype=torch.float64, device=fine.device)
        N.check(N.lib.gme_hier_merge(field.data_ptr(), int(field.dtype == torch.float64), field.shape[1], field.shape[2],
                                     fine.data_ptr(), fine.shape[1], fine.shape[2], fine.shape[0], out.data_ptr(),
                                     _stream()), "gme_hier_merge")
        field = out
    return field


def results_batch(frames: Planes, distance: int, procedure: int = N.SEARCH_DIAMOND, window: int = 2,
                  outlier_fraction: float = OUTLIER_FRACTION) -> dict:
    """Everything one iteration of the reference's results.py loop computes (results.py:47-59, 78-83, 109), for every
    pair (k, k + distance) of a device-resident sequence, left on the device: affine parameters, model field at block
    size 16, compensated previous frame, the two difference images and the squared-error sums behind the PSNR."""
    n = frames.n - distance
    if n <= 0:
        raise ValueError("sequence shorter than the frame distance")
    prev, cur = frames.view(0, n), frames.view(distance, distance + n)
    pipe = Pipeline(n, frames.H, frames.W, frames.t.device, want_comp=False)
    pipe.run(prev, cur, procedure, window, outlier_fraction)
    model = affine_field(pipe.params, frames.H // 16, frames.W // 16)
    comp, sse_, dprev, dcomp = compensate(prev, model, cur, want_diffs=True)
    return {"params": pipe.params, "status": pipe.status, "model_field": model, "compensated": comp, "sse": sse_,
            "diff_curr_prev": dprev, "diff_curr_comp": dcomp}


def sse(a: Planes, b: Planes) -> torch.Tensor:
    out = torch.zeros((a.n,), dtype=torch.int64, device=a.t.device)
    N.check(N.lib.gme_sse(a.ptr, a.pitch, a.stride, b.ptr, b.pitch, b.stride, a.n, a.H, a.W, out.data_ptr(),
                          _stream()), "gme_sse")
    return out


def psnr_from_sse(sse_value: int, npix: int):
    """utils.PSNR's tail (utils.py:111-116): complex result, int -1 for identical images."""
    mse = sse_value / float(npix)
    if mse == 0:
        return -1
    return 20 * cmath.log10(255.0 / cmath.sqrt(mse))


# --------------------------------------------------------------------------- whole pipeline
class Pipeline:
    """motion.global_motion_estimation + model field + compensation + PSNR for batches of n pairs.

    Owns the workspace and the output buffers; ``run`` enqueues the whole pipeline (7 kernel
    launches) on the current stream without any host synchronisation, so it can be captured in
    a CUDA graph (``capture``) and replayed.

    ``lanes`` > 1 splits the batch into that many contiguous sub-batches, each one ``gme_pipeline`` call with its own
    workspace on its own stream, forked from and joined to the current stream with events (capturable).  Pairs are
    independent, so the results are the same bits; what it buys is overlap: the stages of one call depend on each other,
    so a single call leaves the GPU nearly idle during the three small fit kernels and in the last wave of every
    kernel, and the other lane's block matching fills those holes (measured: 1080p 0.611 -> 0.584 ms per 64 pairs,
    720x480 0.549 -> 0.517 ms per 256; more than two lanes gives nothing more)."""

    def __init__(self, n: int, H: int, W: int, device=None, want_comp: bool = True, lanes: int = 1):
        self.device = device or require_cuda()
        self.n, self.H, self.W = n, H, W
        lanes = max(1, min(int(lanes), n))
        per = -(-n // lanes) if n else 0
        self._cuts = [(a, min(a + per, n)) for a in range(0, n, per)] if n else [(0, 0)]
        self._workspaces = [torch.empty((max(N.lib.gme_pipeline_workspace_bytes(b - a, H, W), 16),), dtype=torch.uint8,
                                        device=self.device) for a, b in self._cuts]
        self.workspace = self._workspaces[0]
        self._side = [torch.cuda.Stream(self.device) for _ in self._cuts[1:]]
        # one row of seven 8-byte words per pair: six float64 parameters + the uint64 squared-error sum.  The kernels
        # write both straight into it (strided outputs of gme_pipeline), so the multi-GPU gather ships `rows` as is.
        self.rows = torch.zeros((n, 7), dtype=torch.float64, device=self.device)
        self.params = self.rows[:, :6]
        self.status = torch.zeros((n,), dtype=torch.int32, device=self.device)
        self.comp = Planes.empty(n, H, W, self.device) if want_comp else None
        self.sse = self.rows.view(torch.int64)[:, 6] if want_comp else None
        self.graph = None
        self._graph_key = None

    @property
    def lanes(self) -> int:
        return len(self._cuts)

    def _launch(self, lane: int, prev: Planes, cur: Planes, procedure: int, window: int, outlier_fraction: float):
        a, b = self._cuts[lane]
        c, ws = self.comp, self._workspaces[lane]
        N.check(N.lib.gme_pipeline(prev.ptr + a * prev.stride, prev.stride, cur.ptr + a * cur.stride, cur.stride, b - a,
                                   self.H, self.W, prev.pitch, int(procedure), int(window), float(outlier_fraction),
                                   self.rows.data_ptr() + a * 56, 7,
                                   c.ptr + a * c.stride if c else None, c.pitch if c else 0, c.stride if c else 0,
                                   self.rows.data_ptr() + a * 56 + 48 if c else None, 7, self.status.data_ptr() + a * 4,
                                   ws.data_ptr(), ws.numel(), _stream()), "gme_pipeline")

    def run(self, prev: Planes, cur: Planes, procedure: int = N.SEARCH_DIAMOND, window: int = 2,
            outlier_fraction: float = OUTLIER_FRACTION):
        if prev.n != self.n or prev.H != self.H or prev.W != self.W or cur.t.shape != prev.t.shape:
            raise ValueError("pipeline geometry mismatch")
        if prev.pitch != cur.pitch:
            raise ValueError("previous and current must share one pitch")
        if self._side:
            main = torch.cuda.current_stream(self.device)
            fork = torch.cuda.Event()
            fork.record(main)
            joins = []
            for lane, side in enumerate(self._side, start=1):
                side.wait_event(fork)
                with torch.cuda.stream(side):
                    self._launch(lane, prev, cur, procedure, window, outlier_fraction)
                    done = torch.cuda.Event()
                    done.record(side)
                joins.append(done)
            self._launch(0, prev, cur, procedure, window, outlier_fraction)
            for done in joins:
                main.wait_event(done)
        else:
            self._launch(0, prev, cur, procedure, window, outlier_fraction)
        return self.params, self.sse, self.status

    def capture(self, prev: Planes, cur: Planes, procedure: int = N.SEARCH_DIAMOND, window: int = 2,
                outlier_fraction: float = OUTLIER_FRACTION):
        """Captures one run on (prev, cur) -- whose storage must stay alive and be refilled in place --
        into a CUDA graph; ``replay`` then costs one graph launch."""
        self.run(prev, cur, procedure, window, outlier_fraction)   # warm-up outside capture (module load, attributes)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        # thread_local: other threads of the process (a decode thread, NCCL's watchdog) may call the CUDA runtime
        # while this one captures
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self.run(prev, cur, procedure, window, outlier_fraction)
        self.graph = g
        self._graph_key = (prev.ptr, cur.ptr, procedure, window, outlier_fraction)
        return g

    def replay(self):
        self.graph.replay()
        return self.params, self.sse, self.status

    def intermediate(self, which: int) -> torch.Tensor:
        """Views into the workspace (tests): 0 dense field, 1/2 L1/L2 fields, 3/4 L1/L2 outlier masks, 5 model field
        (several lanes: the lanes' views concatenated, a copy)."""
        l1 = ((self.H + 1) // 2, (self.W + 1) // 2)
        l0 = ((l1[0] + 1) // 2, (l1[1] + 1) // 2)
        shapes = {0: ((l0[0] // 2, l0[1] // 2, 2), torch.int32),
                  1: ((l1[0] // 16, l1[1] // 16, 2), torch.int32),
                  2: ((self.H // 16, self.W // 16, 2), torch.int32),
                  3: ((l1[0] // 16, l1[1] // 16), torch.uint8),
                  4: ((self.H // 16, self.W // 16), torch.uint8),
                  5: ((self.H // 16, self.W // 16, 2), torch.int16)}
        tail, dtype = shapes[which]
        parts = []
        for (a, b), ws in zip(self._cuts, self._workspaces):
            shape = (b - a,) + tail
            ptr = N.lib.gme_pipeline_workspace_ptr(ws.data_ptr(), b - a, self.H, self.W, which)
            off = ptr - ws.data_ptr()
            count = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
            parts.append(ws[off:off + count].view(dtype).view(shape))
        return parts[0] if len(parts) == 1 else torch.cat(parts)

    def psnr(self):
        """Host-side tail of utils.PSNR for every pair (one device->host read of n int64)."""
        return [psnr_from_sse(int(s), self.H * self.W) for s in self.sse.cpu().tolist()]


def gme_pairs(prev, cur, procedure: int = N.SEARCH_DIAMOND, window: int = 2, want_comp: bool = True,
              outlier_fraction: float = OUTLIER_FRACTION):
    """Public batched entry point with HOST buffers: uint8 [n, H, W] previous and current frames in,
    (params float64[n, 6], psnr list, compensated uint8[n, H, W] | None) out.  Raises
    numpy.linalg.LinAlgError like the reference if any pair hits a singular normal matrix."""
    p, c = Planes.from_host(prev), Planes.from_host(cur)
    pipe = Pipeline(p.n, p.H, p.W, p.t.device, want_comp)
    pipe.run(p, c, procedure, window, outlier_fraction)
    if int(pipe.status.max().item()) != 0:
        raise N.singular_matrix_error()
    params = pipe.params.cpu().numpy()
    if not want_comp:
        return params, None, None
    return params, pipe.psnr(), pipe.comp.to_host()


def gme_sequence(frames: Planes, distance: int, procedure: int = N.SEARCH_DIAMOND, window: int = 2,
                 pipeline: Pipeline | None = None):
    """All pairs (k, k + distance) of a device-resident sequence, the loop of results.py:41-59,109.
    The two operands are views of the same buffer; nothing is copied."""
    n = frames.n - distance
    if n <= 0:
        raise ValueError("sequence shorter than the frame distance")
    pipe = pipeline or Pipeline(n, frames.H, frames.W, frames.t.device)
    pipe.run(frames.view(0, n), frames.view(distance, distance + n), procedure, window)
    return pipe


# --------------------------------------------------------------------------- the per-pair drop-in surface
class PairSession:
    """State behind the reference-facing per-pair calls (motion.global_motion_estimation, get_motion_field_affine,
    compensate_frame, utils.PSNR -- what results.py:50-59,109 drives), one per frame geometry and device.

    A call through the module surface costs host->device copies of its NumPy arguments, a handful of kernels on
    KB..MB-scale data and a device->host read of its result, so everything that is not data movement is kept out of
    it: the pipeline workspace, the device frame slots and the pinned staging buffers are allocated once; the frames
    of a call travel as ONE copy from pinned memory; results come back through pinned buffers with ONE stream
    synchronisation per call; the seven kernels of the pipeline are replayed from a CUDA graph.
    Frames are NOT cached across calls by buffer identity: a caller may legally rewrite an array in place between
    two calls, and verifying a cached copy (a host memcmp) costs as much as the pinned staging copy it would save.
    """

    def __init__(self, H: int, W: int, device=None):
        self.device = device or require_cuda()
        self.H, self.W = H, W
        self.frames = Planes.empty(3, H, W, self.device)                 # slots: previous, current, (compensated input)
        pitch = self.frames.pitch
        self.stage = torch.empty((3, H, pitch), dtype=torch.uint8, pin_memory=True)
        self.stage_np = self.stage.numpy()
        self.pipe = Pipeline(1, H, W, self.device, want_comp=False)
        self.out_host = torch.empty((8,), dtype=torch.float64, pin_memory=True)   # 6 parameters, status, spare
        self.out_dev = torch.empty((8,), dtype=torch.float64, device=self.device)
        self.sse_host = torch.empty((1,), dtype=torch.int64, pin_memory=True)
        self.sse_dev = torch.zeros((1,), dtype=torch.int64, device=self.device)
        self.comp = Planes.empty(1, H, W, self.device)
        self.comp_host = torch.empty((H, pitch), dtype=torch.uint8, pin_memory=True)
        self.graphs = {}                                                   # (procedure, window, pct) -> CUDAGraph

    def upload(self, *frames) -> None:
        """frames[k] (uint8 ndarray [H, W]) -> device slot k; one pinned staging copy each, one H2D for all.

        The staging copy goes through torch's ``copy_``: its intra-op thread pool splits a 2 MB frame over the host
        cores (measured here: 63 us against 318 us for NumPy's single-thread memcpy, which was most of a call; a
        Python-level thread pool costs more in dispatch than the copy takes)."""
        for k, f in enumerate(frames):
            if f.flags.c_contiguous and f.flags.writeable:
                self.stage[k, :, :self.W].copy_(torch.from_numpy(f))
            else:                                                          # read-only or strided caller arrays
                self.stage_np[k, :, :self.W] = f
        n = len(frames)
        self.frames.t[:n].copy_(self.stage[:n], non_blocking=True)

    def gme(self, previous, current, procedure: int, window: int, pct: float) -> np.ndarray:
        self.upload(previous, current)
        key = (int(procedure), int(window), float(pct))
        prev, cur = self.frames.view(0, 1), self.frames.view(1, 2)
        g = self.graphs.get(key)
        if g is None:
            self.pipe.run(prev, cur, *key)                                 # warm-up / validation outside capture
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self.pipe.run(prev, cur, *key)
                self.out_dev[:6] = self.pipe.params[0]
                self.out_dev[6] = self.pipe.status[0].to(torch.float64)
            self.graphs[key] = g
        g.replay()
        self.out_host.copy_(self.out_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if self.out_host[6].item() != 0:
            raise N.singular_matrix_error()
        return self.out_host[:6].numpy().copy()

    def compensate(self, frame, field: np.ndarray) -> np.ndarray:
        self.upload(frame)
        f = torch.from_numpy(np.ascontiguousarray(field, dtype=np.int32)).unsqueeze(0).to(self.device, non_blocking=True)
        n, R, C = f.shape[0], f.shape[1], f.shape[2]
        src = self.frames.view(0, 1)
        N.check(N.lib.gme_compensate(src.ptr, src.pitch, src.stride, f.data_ptr(), 0, R, C, None, 0, 0, self.comp.ptr,
                                     self.comp.pitch, self.comp.stride, 1, self.H, self.W, None, _stream()), "gme_compensate")
        self.comp_host.copy_(self.comp.t[0], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out = torch.empty((self.H, self.W), dtype=torch.uint8)             # the caller's own array (threaded copy, as above)
        out.copy_(self.comp_host[:, :self.W])
        return out.numpy()

    def sse(self, a, b) -> int:
        self.upload(a, b)
        self.sse_dev.zero_()
        x, y = self.frames.view(0, 1), self.frames.view(1, 2)
        N.check(N.lib.gme_sse(x.ptr, x.pitch, x.stride, y.ptr, y.pitch, y.stride, 1, self.H, self.W,
                              self.sse_dev.data_ptr(), _stream()), "gme_sse")
        self.sse_host.copy_(self.sse_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return int(self.sse_host.item())


_pair_sessions: dict = {}


def pair_session(H: int, W: int) -> PairSession:
    dev = require_cuda()
    key = (H, W, dev.index)
    s = _pair_sessions.get(key)
    if s is None:
        if len(_pair_sessions) >= 4:                                       # a handful of geometries at most stay resident
            _pair_sessions.pop(next(iter(_pair_sessions)))
        s = _pair_sessions[key] = PairSession(H, W, dev)
    return s


class HostSequenceRunner:
    """The loop of results.py:41-59,109 for a sequence that lives in HOST memory.

    ``run(host_frames)`` takes pinned uint8[n_frames, H, W], returns pinned float64[n_pairs, 8] rows
    (6 affine parameters, squared-error sum against the current frame, status) that are valid once the
    current stream has been synchronised.  Frames travel host->device exactly once each, on a copy stream,
    in chunk order; the pipeline of chunk j (``chunk`` pairs) starts as soon as its last frame has landed,
    so the PCIe transfer of chunk j+1 overlaps the kernels of chunk j.  Two device sequence buffers
    alternate between calls, so the upload of the next call overlaps the tail of this one."""

    def __init__(self, n_frames: int, H: int, W: int, distance: int, chunk: int = 16,
                 procedure: int = N.SEARCH_DIAMOND, window: int = 2, want_comp: bool = True, device=None):
        self.device = device or require_cuda()
        self.n_frames, self.H, self.W, self.distance = n_frames, H, W, distance
        self.n_pairs = n_frames - distance
        if self.n_pairs <= 0:
            raise ValueError("sequence shorter than the frame distance")
        self.chunk = max(1, min(chunk, self.n_pairs))
        self.procedure, self.window = procedure, window
        self.buffers = [Planes.empty(n_frames, H, W, self.device) for _ in range(2)]
        self.buffer_free = [None, None]                 # event: last kernel that read the buffer has finished
        self.turn = 0
        self.copy_stream = torch.cuda.Stream(self.device)
        self.pipe = Pipeline(self.chunk, H, W, self.device, want_comp)
        tail = self.n_pairs % self.chunk
        self.tail_pipe = Pipeline(tail, H, W, self.device, want_comp) if tail else None
        self.dev_rows = torch.zeros((self.n_pairs, 8), dtype=torch.float64, device=self.device)
        self.host_rows = torch.empty((self.n_pairs, 8), dtype=torch.float64, pin_memory=True)
        self.h2d_bytes = n_frames * H * W
        self.d2h_bytes = self.host_rows.numel() * 8

    def run(self, host_frames: torch.Tensor) -> torch.Tensor:
        if tuple(host_frames.shape) != (self.n_frames, self.H, self.W) or host_frames.dtype != torch.uint8:
            raise ValueError("expected uint8 frames of shape [n_frames, H, W]")
        compute = torch.cuda.current_stream(self.device)
        planes = self.buffers[self.turn]
        if self.buffer_free[self.turn] is not None:
            self.copy_stream.wait_event(self.buffer_free[self.turn])
        else:
            self.copy_stream.wait_stream(compute)
        uploaded = 0
        for start in range(0, self.n_pairs, self.chunk):
            stop = min(start + self.chunk, self.n_pairs)
            need = stop + self.distance                  # frames [0, need) must be resident for pairs [start, stop)
            with torch.cuda.stream(self.copy_stream):
                planes.pixels()[uploaded:need].copy_(host_frames[uploaded:need], non_blocking=True)
                landed = torch.cuda.Event()
                landed.record(self.copy_stream)
            uploaded = need
            compute.wait_event(landed)
            pipe = self.pipe if stop - start == self.chunk else self.tail_pipe
            pipe.run(planes.view(start, stop), planes.view(start + self.distance, stop + self.distance),
                     self.procedure, self.window)
            rows = self.dev_rows[start:stop]
            rows[:, :6] = pipe.params
            rows[:, 6] = pipe.sse.to(torch.float64) if pipe.sse is not None else 0
            rows[:, 7] = pipe.status.to(torch.float64)
        done = torch.cuda.Event()
        done.record(compute)
        self.buffer_free[self.turn] = done
        self.turn ^= 1
        self.host_rows.copy_(self.dev_rows, non_blocking=True)
        return self.host_rows
