#!/usr/bin/env python
"""bench.py -- GME frame-pairs/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the whole hot path (pyramids, dense + block BBME, robust affine fit, model field,
compensation, PSNR sums: gme_pipeline of include/gme_b200.h) over one batch of frame pairs
(k, k + 3) of a synthetic sequence.  `value` has the sequence resident in HBM; `e2e` is the same step
driven with HOST (pinned) frames: host->device copy of the sequence and device->host read of the
per-pair affine parameters + squared-error sums inside the timed region.  Frame pairs shard across
ranks with no data-path collective; the only exchange is the all-gather of [pairs, 7] float64 rows
(6 parameters + PSNR) at the end of each step.

`--impl reference` (and the cpu_baseline object of the default arm) times the CPU port of the
reference's pipeline (oracle/gme_oracle.c, kind "port": the reference itself is pure Python and cannot
travel to the GPU box) on the host cores, one pair per thread.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "global-motion-estimation_b200")
sys.path[:0] = [PKG]

import numpy as np  # noqa: E402

DISTANCE = 3                                   # results.py -f 3 (README default)

WORKLOADS = {
    # name: (H, W, motion, procedure, window, pairs per step per GPU, description)
    "gme_1080p": (1080, 1920, "zoomrot", 3, 2, 64,
                  "full GME pipeline, synthetic 1920x1080 zoom+rotate sequence, frame distance 3, reference-default "
                  "search (diamond, MSE, bs 2/16/16)"),
    "gme_1080p_3step": (1080, 1920, "zoomrot", 1, 16, 64,
                        "config 4: full GME pipeline 1920x1080 zoom+rotate, three-step sw=16 on the bs-16 levels"),
    "gme_1080p_2dlog": (1080, 1920, "zoomrot", 2, 16, 64,
                        "config 4: full GME pipeline 1920x1080 zoom+rotate, 2D-log sw=16 on the bs-16 levels"),
    "gme_480p": (480, 720, "pan", 3, 2, 256,
                 "config 3: full GME pipeline, synthetic 720x480 panning sequence, frame distance 3 (diamond)"),
    "gme_4k_exh32": (2160, 3840, "affine", 0, 32, 8,
                     "config 5: full GME pipeline 3840x2160 affine sequence, exhaustive sw=32 on the bs-16 levels"),
}


# ----------------------------------------------------------------------------------------------
# synthetic sequences (generated on the device with torch: float64 bilinear warp of a NumPy texture)
# ----------------------------------------------------------------------------------------------
def make_sequence(n_frames, H, W, motion, seed, device):
    import torch
    import gme_synth as S
    if motion == "pan":
        return torch.from_numpy(S.pan_sequence(n_frames, H, W, step=(2, 1), seed=seed)).to(device)
    base = torch.from_numpy(S.texture(H, W, seed)).to(device=device, dtype=torch.float64)
    ys, xs = torch.meshgrid(torch.arange(H, device=device, dtype=torch.float64),
                            torch.arange(W, device=device, dtype=torch.float64), indexing="ij")
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    rng = np.random.default_rng(seed)
    tx, ty, sc, rot, sh = rng.uniform(-1, 1, 5)
    unit = 12.0 / 3.0 / np.hypot(cx, cy)
    out = torch.empty((n_frames, H, W), dtype=torch.uint8, device=device)

    def refl(i, n):
        i = i.abs().to(torch.int64) % (2 * n - 2)
        return torch.where(i >= n, 2 * n - 2 - i, i)

    for k in range(n_frames):
        if motion == "zoomrot":                      # config 4: scale 1 + 0.002 k, angle 0.1 deg * k about the centre
            s, t, h, dx, dy = 1.0 / (1.0 + 0.002 * k), np.deg2rad(0.1 * k), 0.0, 0.0, 0.0
        else:                                        # config 5: general affine, corner displacement <= 12 px / 3 frames
            s, t, h = 1.0 + sc * unit * k, rot * unit * k, sh * unit * k
            dx, dy = tx * 4.0 * k / 3.0, ty * 4.0 * k / 3.0
        A = s * np.array([[np.cos(t), np.sin(t) + h], [-np.sin(t), np.cos(t)]])
        b = np.array([cx, cy]) - A @ np.array([cx, cy]) + np.array([dx, dy])
        sx = A[0, 0] * xs + A[0, 1] * ys + b[0]
        sy = A[1, 0] * xs + A[1, 1] * ys + b[1]
        x0, y0 = sx.floor(), sy.floor()
        fx, fy = sx - x0, sy - y0
        x0i, x1i, y0i, y1i = refl(x0, W), refl(x0 + 1, W), refl(y0, H), refl(y0 + 1, H)
        top = base[y0i, x0i] * (1 - fx) + base[y0i, x1i] * fx
        bot = base[y1i, x0i] * (1 - fx) + base[y1i, x1i] * fx
        out[k] = (top * (1 - fy) + bot * fy).round().clamp(0, 255).to(torch.uint8)
    return out


# ----------------------------------------------------------------------------------------------
# clocks (sampled during the timed regions)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _sample(self):
        try:
            self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _loop(self):
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        if self._thread is not None:
            self._sample()                      # at least one sample while the last kernels are still in flight
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# the CPU port of the reference pipeline (oracle; the only place bench.py runs anything under oracle/)
# ----------------------------------------------------------------------------------------------
def cpu_pairs_per_s(frames: np.ndarray, procedure: int, window: int, n_pairs: int, cores: int):
    """Runs n_pairs frame pairs (drawn cyclically from `frames`) through the C port of the reference's
    pipeline -- global_motion_estimation + get_motion_field_affine + compensate_frame + PSNR
    (results.py:50-59,109) -- one pair per thread on `cores` threads.  Returns (pairs/s, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from concurrent.futures import ThreadPoolExecutor
    import gme_oracle as O
    O.lib()
    H, W = frames.shape[1:]
    n_avail = frames.shape[0] - DISTANCE

    def one(k):
        prev, cur = frames[k % n_avail], frames[k % n_avail + DISTANCE]
        p = O.global_motion_estimation(prev, cur, procedure=procedure, window=window)
        comp = O.compensate_frame(prev, O.get_motion_field_affine((H // 16, W // 16), p))
        return p, O.PSNR(cur, comp)

    one(0)                                              # warm-up (library load, page-in)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(one, range(n_pairs)))
    dt = time.perf_counter() - t0
    return n_pairs / dt, dt


def dropin_pairs_per_s(frames, n_pairs: int, reps: int = 1):
    """The per-pair surface results.py drives (results.py:50-59,109): NumPy frames in, NumPy results out, four calls
    per pair through the drop-in modules -- motion.global_motion_estimation, motion.get_motion_field_affine,
    motion.compensate_frame, utils.PSNR.  Host wall clock (every call synchronises).  Returns (pairs/s, seconds)."""
    import motion
    import utils
    H, W = frames[0].shape
    shape = (int(H / motion.BBME_BLOCK_SIZE), int(W / motion.BBME_BLOCK_SIZE), 2)
    n_avail = len(frames) - DISTANCE

    def one(k):
        previous, current = frames[k % n_avail], frames[k % n_avail + DISTANCE]
        params = motion.global_motion_estimation(previous, current)
        model = motion.get_motion_field_affine(shape, parameters=params)
        compensated = motion.compensate_frame(previous, model)
        return params, utils.PSNR(current, compensated)

    one(0)
    t0 = time.perf_counter()
    for _ in range(reps):
        for k in range(n_pairs):
            one(k)
    dt = time.perf_counter() - t0
    return n_pairs * reps / dt, dt


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------
def algorithmic_bytes_per_step(H, W, pairs, frames):
    """bytes each stage must move per step (uint8 pixels, int32 fields).  The pyramids are built
    once per FRAME of the sequence (each frame serves as `previous` of one pair and `current` of another), so
    they are counted at 1.5625*H*W per frame, not 3.125*H*W per pair as the reference does it."""
    H1, W1 = (H + 1) // 2, (W + 1) // 2
    H0, W0 = (H1 + 1) // 2, (W1 + 1) // 2
    nb0, nb1, nb2 = (H0 // 2) * (W0 // 2), (H1 // 16) * (W1 // 16), (H // 16) * (W // 16)
    return {
        "pyramids": frames * (H * W + 2 * H1 * W1 + H0 * W0),       # read L2, write + read L1, write L0
        "bbme_dense_l0": pairs * (2 * H0 * W0 + 8 * nb0),
        "bbme_l1": pairs * (2 * H1 * W1 + 8 * nb1),
        "bbme_l2": pairs * (2 * H * W + 8 * nb2),
        "fit": pairs * (8 * (nb1 + nb2) + nb1 + nb2 + 48),
        "compensate_psnr": pairs * (3 * H * W + 4 * nb2),
    }


def exhaustive_ops(H, W, bs, sw):
    """pixel-pair operations of one exhaustive search = sum over blocks of nv_r * nv_c * bs^2, with
    nv(p) = min(p + sw + bs - 1, N - bs) - max(p - sw, 0) + 1 the in-frame offsets along one axis."""
    def nv(p, n):
        return min(p + sw + bs - 1, n - bs) - max(p - sw, 0) + 1
    rows = sum(nv(r, H) for r in range(0, H - bs + 1, bs))
    cols = sum(nv(c, W) for c in range(0, W - bs + 1, bs))
    return rows * cols * bs * bs


def bench_exhaustive(D, N, torch, planes, dev, sm_mhz=1965.0):
    """BBME SAD Gop/s vs the integer-pipe roofline (BASELINE.json metric, second half): the exhaustive kernel on
    config-1 and config-5 shaped inputs, against (a) the ceiling of the packed-byte pipe, 148 SMs x 64 lanes x 4 pixels
    x f_SM pixel-pairs/s (one VABSDIFF4 per lane and clock; 74.4 T/s at 1965 MHz), and (b) the rate the search's own
    instruction mix sustains on registers, measured by the probe kernel (gme_sad_peak_probe)."""
    import ctypes
    out = {}
    scratch = torch.zeros(1024, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3 / reps

    peaks = {}
    pipe_peak = 148 * 64 * 4 * sm_mhz * 1e6
    out["pipe_peak_tops"] = pipe_peak / 1e12
    out["pipe_peak_formula"] = f"148 SMs x 64 lanes x 4 pixels x {sm_mhz:.0f} MHz"
    for pn, name in ((0, "sad"), (1, "ssd")):
        npix = ctypes.c_uint64(0)
        t = timed(lambda: N.check(N.lib.gme_sad_peak_probe(pn, 148 * 16, 2048, scratch.data_ptr(), ctypes.byref(npix), stream)), 5)
        peaks[pn] = npix.value / t
        out[f"probe_{name}_tops"] = peaks[pn] / 1e12
    H, W = planes.H, planes.W
    cases = [("cfg1_mae_bs12_sw12_320x240_x256", 240, 320, 12, 12, 0, 256), ("cfg5_mse_bs16_sw32_%dx%d_x4" % (W, H), H, W, 16, 32, 1, 4)]
    for name, h, w, bs, sw, pn, n in cases:
        if h == H and w == W:
            prev, cur = planes.view(0, n), planes.view(DISTANCE, DISTANCE + n)
        else:
            crop = D.Planes.empty(n + DISTANCE, h, w, dev)
            src = planes.pixels()
            for k in range(n + DISTANCE):
                crop.pixels()[k].copy_(src[k % src.shape[0], 100:100 + h, 200:200 + w])
            prev, cur = crop.view(0, n), crop.view(DISTANCE, DISTANCE + n)
        t = timed(lambda: D.motion_field(prev, cur, bs, sw, 0, pn), 5)
        ops = exhaustive_ops(h, w, bs, sw) * n
        out[name] = {"tops": ops / t / 1e12, "ms": t * 1e3, "pixel_pair_ops": ops, "frac_of_pipe_peak": ops / t / pipe_peak,
                     "frac_of_probe_peak": ops / t / peaks[pn]}
    return out


def emit(obj, real_stdout):
    """The ONE JSON line goes to the real stdout; everything else a library prints (NCCL banners ...) was sent to stderr."""
    os.write(real_stdout, (json.dumps(obj) + "\n").encode())


def main():
    # keep stdout clean for the single JSON line: fd 1 is pointed at stderr while the bench runs
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "dropin"])
    ap.add_argument("--workload", default="gme_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="frame pairs per step per GPU (default: per workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="end", choices=["end", "step"],
                    help="N > 1: one all-gather of every step's rows at the end of the timed region, or one per step")
    ap.add_argument("--lanes", type=int, default=2, help="sub-batches of one step on forked streams (Pipeline(lanes=...))")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="upload/compute chunks per step of the host-frames arm")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    H, W, motion, procedure, window, pairs, desc = WORKLOADS[args.workload]
    pairs = args.pairs or pairs
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric, unit = "gme_frame_pairs_per_sec", "frame-pairs/s"
    config = {"workload": args.workload, "description": desc, "height": H, "width": W, "frame_distance": DISTANCE,
              "pairs_per_step_per_gpu": pairs, "search_procedure": procedure, "search_window": window,
              "l2_policy": "inputs larger than L2 (sequence + outputs per step exceed 126 MB)" if
              (pairs + DISTANCE) * H * W * 2 > 126e6 else "L2 flushed between steps (256 MB memset)"}

    if args.impl == "reference":
        if rank != 0:
            return
        cores = host_cores()
        import gme_synth as S
        nf = 8 + DISTANCE
        frames = (S.pan_sequence(nf, H, W, step=(2, 1), seed=3) if motion == "pan"
                  else S.zoom_rotate_sequence(nf, H, W, seed=4) if motion == "zoomrot"
                  else S.affine_sequence(nf, H, W, seed=5))
        rate = 0.0
        for _ in range(max(1, args.warmup)):
            rate, _ = cpu_pairs_per_s(frames, procedure, window, cores, cores)
        # a step = about one second of CPU work (a bounded sample of the workload), at least one pair per core and at
        # most 16: long enough that thread start-up and the last-thread tail do not understate the reference
        per_step = int(min(16 * cores, max(cores, round(rate / cores) * cores)))
        t = 0.0
        for _ in range(args.steps):
            _, dt = cpu_pairs_per_s(frames, procedure, window, per_step, cores)
            t += dt
        value = per_step * args.steps / t
        sample = f"{per_step} pairs per step drawn cyclically from an {nf}-frame sequence, one pair per thread"
        config["pairs_per_step_per_gpu"] = per_step        # what this arm ran per step (a bounded sample of the workload)
        config["l2_policy"] = "n/a (CPU arm)"
        emit({
            "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/f64",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}, real_stdout)
        return

    if args.impl == "dropin":
        # the reference-facing per-pair surface, the path results.py drives: one frame pair per call, host NumPy in/out
        if rank != 0:
            return
        import torch
        torch.cuda.set_device(local_rank)
        n = min(pairs, 32)
        seq = make_sequence(n + DISTANCE, H, W, motion, seed=4, device=torch.device("cuda", local_rank)).cpu().numpy()
        frames = [np.ascontiguousarray(f) for f in seq]
        for _ in range(max(1, args.warmup)):
            dropin_pairs_per_s(frames, min(n, 4))
        sampler = ClockSampler(local_rank)
        with sampler:
            v, dt = dropin_pairs_per_s(frames, n, reps=max(1, args.steps))
        config["pairs_per_step_per_gpu"] = n
        nbytes = 6 * H * W + 48 + (H // 16) * (W // 16) * 8          # frames uploaded by the four calls + parameters + field
        emit({"impl": "dropin", "metric": metric, "value": v, "unit": unit, "n_gpus": 1, "steps": max(1, args.steps),
              "warmup": max(1, args.warmup), "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 (fit: int64 sums + f64 solve)", "data": "synthetic",
              "config": config, "clocks": sampler.summary(),
              "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": nbytes * n, "d2h_bytes_per_step": (H * W + 48 + (H // 16) * (W // 16) * 4 + 8) * n,
                      "api": "motion.global_motion_estimation + get_motion_field_affine + compensate_frame + utils.PSNR per pair (results.py:50-59,109)"}},
             real_stdout)
        return

    # ------------------------------------------------------------------ the B200 arm
    import torch
    import torch.distributed as dist
    import gme_device as D
    import gme_distributed as GD
    import gme_native as N

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = "unchanged"
    try:
        # pinned host frames should live on the NUMA node this GPU hangs off: bind the process to the GPU's ideal
        # CPUs before the first pinned allocation (first touch decides the node)
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        numa = "bound to the GPU's ideal CPUs (%d cores)" % len(os.sched_getaffinity(0))
    except Exception as exc:                                                      # noqa: BLE001
        numa = f"unchanged ({type(exc).__name__})"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nf = pairs + DISTANCE
    seq = make_sequence(nf, H, W, motion, seed=4 + rank, device=dev)            # uint8[nf, H, W] on the device
    planes = D.Planes.empty(nf, H, W, dev)
    planes.pixels().copy_(seq)
    host_frames = torch.empty((nf, H, W), dtype=torch.uint8, pin_memory=True)
    host_frames.copy_(seq)
    del seq
    prev, cur = planes.view(0, pairs), planes.view(DISTANCE, nf)
    # The kernels write the [pairs, 7] rows (6 affine parameters + squared-error sum) straight into pipe.rows.  On several
    # GPUs the rows of every rank go to every rank -- the one exchange of the path (one collective at the end
    # of the job).  --gather end (default): every step's rows are copied into a [steps, pairs, 7] device buffer and ONE
    # all-gather closes the timed region.  --gather step: an all-gather per step.  Either way the copy / gather of step k
    # runs on another stream while step k + 1 computes into a second pipeline (a set is reused only after its rows have
    # been taken).
    # Each pipeline runs as args.lanes sub-batches on forked streams (the other lane's block matching fills the holes the
    # small dependent fit kernels and every kernel's last wave leave) and is captured ONCE in a CUDA graph; a step is
    # one graph launch (Pipeline.capture / replay, the public API).
    per_step_gather = world > 1 and args.gather == "step"
    pipes = [D.Pipeline(pairs, H, W, dev, lanes=args.lanes) for _ in range(2 if world > 1 else 1)]
    launches_before = N.launch_count()
    pipes[0].run(prev, cur, procedure, window)          # eager once: kernels per step, counted by the library
    launches_per_step = N.launch_count() - launches_before
    for p in pipes:
        p.capture(prev, cur, procedure, window)
    # per-stage device times (the roofline's kernel durations) come from a SEPARATE timed pass of the same steps on one
    # lane, launched eagerly, where gme_pipeline brackets every stage with CUDA events on its stream -- lanes overlap
    # stages of different sub-batches, so events inside them would not time one kernel alone
    stage_pipe = D.Pipeline(pairs, H, W, dev)
    total_pairs = pairs * world
    gathered = [torch.empty((total_pairs, 7), dtype=torch.float64, device=dev) for _ in pipes]
    pending = [None for _ in pipes]
    turn = [0]
    job_rows = job_gathered = None
    if world > 1 and not per_step_gather:
        job_rows = torch.zeros((args.steps, pairs, 7), dtype=torch.float64, device=dev)
        job_gathered = torch.empty((world * args.steps * pairs, 7), dtype=torch.float64, device=dev)
    done_steps = [0]
    side_stream = torch.cuda.Stream(dev)

    class RowsTaken:                                    # same protocol as the NCCL work handle: .wait() orders `main` after it
        def __init__(self, main):
            self.main, self.event = main, torch.cuda.Event()
            self.event.record(torch.cuda.current_stream())

        def wait(self):
            self.main.wait_event(self.event)
    flush = None if "inputs larger" in config["l2_policy"] else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    runner = D.HostSequenceRunner(nf, H, W, DISTANCE, chunk=max(1, -(-pairs // max(1, args.e2e_chunks))), procedure=procedure, window=window,
                                  device=dev)

    def step(e2e):
        if flush is not None:
            flush.zero_()
        if e2e == "stages":
            stage_pipe.run(prev, cur, procedure, window)
            return None
        if e2e:
            # public host API: pinned frames in, pinned per-pair rows out; uploads overlap the kernels chunk by chunk
            out = runner.run(host_frames)
            if world > 1:
                GD.gather_rows(runner.dev_rows[:, :7], total_pairs)
            return out
        i = turn[0]
        turn[0] = (i + 1) % len(pipes)
        if pending[i] is not None:
            pending[i].wait()                           # (stream-level) the gather that still reads this set's rows
            pending[i] = None
        pipes[i].replay()                               # results: .rows (.params / .sse) / .status / .comp on the device
        if per_step_gather:
            pending[i] = GD.gather_rows_async(pipes[i].rows, gathered[i])
        elif job_rows is not None:                      # 3.5 KB device copy on a side stream; the gather comes once, in finish()
            main = torch.cuda.current_stream()
            ready = torch.cuda.Event()
            ready.record(main)
            side_stream.wait_event(ready)
            with torch.cuda.stream(side_stream):
                job_rows[done_steps[0] % args.steps].copy_(pipes[i].rows)
                pending[i] = RowsTaken(main)
            done_steps[0] += 1

    def finish(e2e):                                    # closes a timed region
        drain()
        if job_rows is not None and e2e is False:
            dist.all_gather_into_tensor(job_gathered, job_rows.view(-1, 7))

    def drain():                                        # orders the current stream after every gather still in flight
        for i, w in enumerate(pending):
            if w is not None:
                w.wait()
                pending[i] = None

    def barrier():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e, steps: int, sampler, after_warmup=None):
        for _ in range(args.warmup):
            step(e2e)
        barrier()
        if after_warmup:
            after_warmup()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with sampler:
            start.record()
            for _ in range(steps):
                step(e2e)
            finish(e2e)                                 # the gather(s) belong to the timed region
            end.record()
            barrier()
        ms = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    # device-resident arm
    ms_total = timed(False, args.steps, sampler)
    value = total_pairs * args.steps / (ms_total * 1e-3)

    # the same steps once more for the per-stage times (see stage_pipe above)
    ms_stage_pass = timed("stages", args.steps, sampler, lambda: N.stage_timing_enable(True))
    stage_ms, calls = N.stage_timing_read()
    N.stage_timing_enable(False)
    assert calls == args.steps, (calls, args.steps)

    # end-to-end arm: host frames in, per-pair results back on the host
    ms_e2e = timed(True, args.steps, sampler)
    e2e_value = total_pairs * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # parity spot check of what was just timed (outside the timed regions; the oracle is the checker)
    parity = None
    l2_ref_ops_per_pair = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import gme_oracle as O
        fr = host_frames.numpy()
        got = runner.host_rows.numpy().copy()
        ok = True
        for k in (0, pairs - 1):
            want = O.global_motion_estimation(fr[k], fr[k + DISTANCE], procedure=procedure, window=window)
            comp = O.compensate_frame(fr[k], O.get_motion_field_affine((H // 16, W // 16), want))
            ok &= bool(np.allclose(got[k, :6], want, atol=1e-9, rtol=1e-9)) and int(got[k, 6]) == O.sse(fr[k + DISTANCE], comp)
        parity = "ok" if ok else "MISMATCH"
        # block cost evaluations the REFERENCE algorithm performs on the full-resolution level (instrumented in the
        # oracle): the algorithmic work of the dominant kernel in pixel-pair operations
        cands = [O.get_motion_field(fr[k], fr[k + DISTANCE], 16, window, procedure, 1, return_candidates=True)[1]
                 for k in (0, pairs - 1)]
        l2_ref_ops_per_pair = float(np.mean(cands)) * 256.0
    except Exception as exc:                                                      # noqa: BLE001
        parity = f"not checked: {exc}"

    # roofline of the dominant stage, measured inside the timed region
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    bytes_ps = algorithmic_bytes_per_step(H, W, pairs, nf)
    stages = {}
    for name, ms in zip(N.STAGE_NAMES, stage_ms):
        per_call = ms / args.steps
        gbs = bytes_ps[name] / (per_call * 1e-3) / 1e9 if per_call > 0 else 0.0
        stages[name] = {"ms_per_step": per_call, "share": ms / max(sum(stage_ms), 1e-12), "algorithmic_gbs": gbs,
                        "frac_of_hbm_peak": gbs / hbm_peak}
    dom = max(stages, key=lambda k: stages[k]["ms_per_step"])
    traffic, ncu_dom = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {})
        traffic = tj.get(dom)
        det = tj.get("_detail", {}).get(dom, {})
        ncu_dom = {k: det[k] for k in ("issue_active", "alu_pipe_pct", "fma_pipe_pct", "lsu_wavefronts_pct", "warp_instructions")
                   if k in det} or None
    except Exception:
        pass
    kernel_names = {"pyramids": "pyr_down_kernel", "bbme_dense_l0": "bbme_diamond2_kernel",
                    "bbme_l1": "bbme_diamond16_kernel" if procedure == 3 else "bbme_exhaustive2_kernel" if procedure == 0 else "bbme_pattern_kernel",
                    "bbme_l2": "bbme_diamond16_kernel" if procedure == 3 else "bbme_exhaustive2_kernel" if procedure == 0 else "bbme_pattern_kernel",
                    "fit": "affine_fit_kernel", "compensate_psnr": "compensate16_kernel"}
    roofline = {"bound": "hbm", "kernel": dom, "kernel_name": kernel_names[dom], "achieved": stages[dom]["algorithmic_gbs"], "peak": hbm_peak,
                "unit": "GB/s", "frac": stages[dom]["frac_of_hbm_peak"], "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_ps[dom],
                # what actually bounds the kernel (ncu of one warm step, profiles/traffic.json): share of issue slots used
                # and pipe utilisation -- the pattern searches are instruction-issue bound, not HBM bound
                "ncu": ncu_dom,
                "whole_step_algorithmic_gbs": sum(bytes_ps.values()) / (ms_total / args.steps * 1e-3) / 1e9}
    exhaustive = None
    try:
        exhaustive = bench_exhaustive(D, N, torch, planes, dev, float(sampler.max_mhz or 1965))
    except Exception as exc:                                                      # noqa: BLE001
        exhaustive = {"error": str(exc)}
    if l2_ref_ops_per_pair and isinstance(exhaustive, dict) and "probe_ssd_tops" in exhaustive:
        # the block-matching stages are bound by the integer pipes, not by HBM: the reference's pixel-pair operations
        # on the full-resolution level per second, against the measured rate of the SSD instruction mix
        t = stages["bbme_l2"]["ms_per_step"] * 1e-3
        tops = l2_ref_ops_per_pair * pairs / t / 1e12
        roofline["integer_pipe"] = {"kernel": "bbme_l2", "achieved": tops, "peak": exhaustive["pipe_peak_tops"],
                                    "unit": "T pixel-pair ops/s", "frac": tops / exhaustive["pipe_peak_tops"],
                                    "frac_of_probe": tops / exhaustive["probe_ssd_tops"],
                                    "reference_ops_per_pair": l2_ref_ops_per_pair,
                                    "note": "operations the reference algorithm performs (oracle-instrumented); the kernel "
                                            "skips candidates whose cost it already holds"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        fr = host_frames.numpy()
        n_cpu = max(2 * cores, 16)
        v, dt = cpu_pairs_per_s(fr, procedure, window, n_cpu, cores)
        while dt < 8.0 and n_cpu < 64 * cores:                                   # bounded sample of ~10-30 s
            n_cpu *= 4
            v, dt = cpu_pairs_per_s(fr, procedure, window, n_cpu, cores)
        cpu = {"value": v, "unit": unit, "cores": cores, "kind": "port",
               "sample": f"{n_cpu} pairs of the same sequence (cyclic), one pair per thread, {dt:.1f} s of wall time"}

    dropin = None
    if world == 1:
        try:
            fr = [np.ascontiguousarray(f) for f in host_frames.numpy()[:16 + DISTANCE]]
            v, dt = dropin_pairs_per_s(fr, 16, reps=2)
            dropin = {"value": v, "unit": unit, "sample": "16 pairs x 2 passes, one pair per call, NumPy in / NumPy out",
                      "api": "motion.global_motion_estimation + get_motion_field_affine + compensate_frame + utils.PSNR "
                             "(the calls of results.py:50-59,109)", "ms_per_pair": 1e3 / v}
        except Exception as exc:                                                  # noqa: BLE001
            dropin = {"error": str(exc)}

    out = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32 (fit: int64 sums + f64 solve)", "data": "synthetic", "config": config,
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": runner.h2d_bytes,
                "d2h_bytes_per_step": runner.d2h_bytes, "ms_per_step": ms_e2e / args.steps,
                "api": "gme_device.HostSequenceRunner.run(pinned uint8[frames,H,W]) -> pinned float64[pairs,8]",
                "cpu_affinity": numa,
                "h2d_gbs": runner.h2d_bytes / (ms_e2e / args.steps * 1e-3) / 1e9},
        "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
        "launch": f"one CUDA graph per step: {pipes[0].lanes} lane(s) of gme_pipeline on forked streams ({launches_per_step} kernels)",
        "exchange": (None if world == 1 else "one all_gather_into_tensor per step, overlapped with the next step" if per_step_gather
                     else f"one all_gather_into_tensor of all {args.steps} steps' [pairs, 7] rows at the end of the timed region"),
        "stage_pass": {"ms_per_step": ms_stage_pass / args.steps, "lanes": 1, "launch": "eager, stages bracketed by CUDA events",
                       "note": "separate timed pass of the same steps; source of `stages` and of the roofline's kernel durations"},
        "clocks": sampler.summary(), "roofline": roofline, "stages": stages, "bbme_exhaustive": exhaustive,
        "cpu_baseline": cpu, "parity": parity, "dropin_per_pair": dropin,
    }
    emit(out, real_stdout)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
