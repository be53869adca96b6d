"""results.py's whole job (results.py:14-112) as a streaming pipeline: decode -> pinned memory -> B200 -> PNG / JSON.

    python global-motion-estimation_b200/gme_results.py -v <video under resources/videos/> -f 3

writes exactly the tree the reference's ``results.py -v <video> -f 3`` writes -- frames/, compensated/,
curr_prev_diff/, curr_comp_diff/, model_motion_field/ and psnr_records.json, byte for byte (tests/
test_gpu_results_driver.py compares it with the hashes of the reference's own run) -- but none of the three host
stages serialises the GPU:

  * input (replaces utils.get_video_frames, utils.py:9-31): a decoder thread converts every frame
    straight into a PINNED chunk buffer (``FramePrefetcher``; cv2.cvtColor writes into the pinned slice, no per-frame
    arrays, no pageable list), chunks are handed over through a bounded queue, so decode of chunk j+1 overlaps the
    upload and the kernels of chunk j;
  * compute: one ``gme_device.results_batch`` per chunk on a device-resident sliding window (the last ``distance``
    frames of a chunk stay on the device as the first of the next): parameters, model field, compensated frame, both
    difference images and the squared errors of every pair of the chunk, all left in HBM;
  * output (replaces the cv2.imwrite calls of results.py:64-106): the images of a chunk come back in ONE
    device->host copy per kind into pinned buffers (two sets, alternating) and are PNG-encoded by a pool of host threads
    (cv2.imwrite releases the GIL) while the GPU is already on the next chunk.

Host-side orchestration only: every number comes from the CUDA kernels behind gme_device; there is no CPU fallback.
"""
from __future__ import annotations

import argparse
import json
import os
import queue
import shutil
import threading
from concurrent.futures import ThreadPoolExecutor

import cv2
import numpy as np
import torch

import gme_device as D
import gme_native as N
import utils


class FramePrefetcher:
    """Decodes a video on a background thread into pinned chunk buffers.

    Iterating yields (pinned uint8[k, H, W] tensor view, k) per chunk, in order; a chunk's storage is recycled once
    ``release`` has been called for it, so at most ``depth`` chunks are in flight.  Grayscale conversion follows
    utils.get_video_frames (utils.py:20-28): cv2.COLOR_BGR2GRAY of the decoded BGR frame."""

    def __init__(self, path: str, chunk: int = 16, depth: int = 3):
        self.capture = cv2.VideoCapture(path)
        self.chunk, self.depth = chunk, depth
        self.shape = None
        self.ready: "queue.Queue" = queue.Queue()
        self.free: "queue.Queue" = queue.Queue()
        self.buffers = []
        self.error = None
        self.thread = threading.Thread(target=self._decode, daemon=True)
        self.thread.start()

    def _decode(self):
        try:
            slot, fill, buf = None, 0, None
            bgr = None
            while self.capture.grab():
                ok, bgr = self.capture.retrieve(bgr)
                if self.shape is None:
                    self.shape = bgr.shape[:2]
                    for _ in range(self.depth):
                        self.buffers.append(torch.empty((self.chunk,) + tuple(self.shape), dtype=torch.uint8, pin_memory=True))
                        self.free.put(len(self.buffers) - 1)
                if slot is None:
                    slot, fill = self.free.get(), 0
                    buf = self.buffers[slot].numpy()
                if bgr.ndim == 3 and bgr.shape[2] == 3:
                    cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY, dst=buf[fill])
                else:
                    buf[fill] = bgr.reshape(self.shape)
                fill += 1
                if fill == self.chunk or not ok:
                    self.ready.put((slot, fill))
                    slot = None
                if not ok:
                    break
            if slot is not None and fill:
                self.ready.put((slot, fill))
        except Exception as exc:                                                 # noqa: BLE001
            self.error = exc
        finally:
            self.ready.put(None)

    def __iter__(self):
        while True:
            item = self.ready.get()
            if item is None:
                if self.error is not None:
                    raise self.error
                return
            slot, fill = item
            yield slot, self.buffers[slot][:fill], fill

    def release(self, slot: int, after: torch.cuda.Event | None = None):
        """Hands a chunk buffer back to the decoder (once ``after`` -- the upload that read it -- has completed)."""
        if after is not None:
            after.synchronize()
        self.free.put(slot)


class _HostSet:
    """Pinned landing buffers for the images of one chunk + the event that says they have arrived."""

    def __init__(self, chunk, H, W, pitch, R, C):
        def mk(*shape, dtype=torch.uint8):
            return torch.empty(shape, dtype=dtype, pin_memory=True)
        self.previous = mk(chunk, H, pitch)
        self.compensated, self.diff_prev, self.diff_comp = mk(chunk, H, pitch), mk(chunk, H, pitch), mk(chunk, H, pitch)
        self.model = mk(chunk, R, C, 2, dtype=torch.int16)
        self.sse = mk(chunk, dtype=torch.int64)
        self.status = mk(chunk, dtype=torch.int32)
        self.landed = torch.cuda.Event()
        self.pending = []                        # futures of the encoders still reading this set


def dump_results(video_path: str, save_path: str, distance: int, chunk: int = 16, workers: int | None = None,
                 procedure: int = N.SEARCH_DIAMOND, window: int = 2, outlier_fraction: float = D.OUTLIER_FRACTION,
                 progress=None) -> dict:
    """The loop of results.py:41-112 over ``video_path`` -> the same files under ``save_path`` and the PSNR dict."""
    dev = D.require_cuda()
    for sub in ("frames", "compensated", "curr_prev_diff", "model_motion_field", "curr_comp_diff"):
        os.makedirs(os.path.join(save_path, sub), exist_ok=True)
    pool = ThreadPoolExecutor(workers or min(32, (os.cpu_count() or 4)))
    copy_stream = torch.cuda.Stream(dev)
    compute = torch.cuda.current_stream(dev)
    psnr_dict, order = {}, []
    state = {"window": None, "have": 0, "sets": None, "turn": 0, "frames_seen": 0}

    def write_pair(hs: _HostSet, j: int, idx: int, W: int):
        # results.py:64-112 for one pair: file names and contents as the reference writes them
        def img(t):
            return np.ascontiguousarray(t[j].numpy()[:, :W])
        previous = img(hs.previous)
        cv2.imwrite(os.path.join(save_path, "frames", "") + str(idx - 5) + ".png", previous)
        cv2.imwrite(os.path.join(save_path, "compensated", "") + str(idx - 5) + ".png", img(hs.compensated))
        cv2.imwrite(os.path.join(save_path, "curr_prev_diff", "") + str(idx) + ".png", img(hs.diff_prev))
        cv2.imwrite(os.path.join(save_path, "curr_comp_diff", "") + str(idx) + ".png", img(hs.diff_comp))
        draw = utils.draw_motion_field(previous, hs.model[j].numpy())
        cv2.imwrite(os.path.join(save_path, "model_motion_field", "") + str(idx) + ".png", draw)
        return str(idx), str(D.psnr_from_sse(int(hs.sse[j]), previous.shape[0] * W))

    def flush(hs: _HostSet):
        for f in hs.pending:
            k, v = f.result()
            psnr_dict[k] = v
        hs.pending = []

    prefetch = FramePrefetcher(video_path, chunk=chunk)
    for slot, host_chunk, k in prefetch:
        H, W = host_chunk.shape[1:]
        if state["window"] is None:
            state["window"] = D.Planes.empty(chunk + distance, H, W, dev)
            R, C = H // 16, W // 16
            state["sets"] = [_HostSet(chunk, H, W, state["window"].pitch, R, C) for _ in range(2)]
        win = state["window"]
        have = state["have"]
        # upload the new frames behind the `have` frames kept from the previous chunk
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_stream(compute)                         # the window is free once the last kernels have read it
            win.pixels()[have:have + k].copy_(host_chunk, non_blocking=True)
            uploaded = torch.cuda.Event()
            uploaded.record(copy_stream)
        compute.wait_event(uploaded)
        pool.submit(prefetch.release, slot, uploaded)
        total = have + k
        n_pairs = total - distance
        first_idx = state["frames_seen"] - have + distance              # frame index of the first pair's current frame
        state["frames_seen"] += k
        if n_pairs > 0:
            hs = state["sets"][state["turn"]]
            state["turn"] ^= 1
            flush(hs)                                                    # its previous contents have been encoded
            res = D.results_batch(win.view(0, total), distance, procedure, window, outlier_fraction)
            hs.previous[:n_pairs].copy_(win.t[:n_pairs], non_blocking=True)
            hs.compensated[:n_pairs].copy_(res["compensated"].t, non_blocking=True)
            hs.diff_prev[:n_pairs].copy_(res["diff_curr_prev"].t, non_blocking=True)
            hs.diff_comp[:n_pairs].copy_(res["diff_curr_comp"].t, non_blocking=True)
            hs.model[:n_pairs].copy_(res["model_field"], non_blocking=True)
            hs.sse[:n_pairs].copy_(res["sse"], non_blocking=True)
            hs.status[:n_pairs].copy_(res["status"], non_blocking=True)
            hs.landed.record(compute)
            # keep the last `distance` frames as the head of the next window (device-to-device, after the kernels)
            keep = win.t[total - distance:total].clone()
            win.t[:distance].copy_(keep)
            state["have"] = distance

            def encode(hs=hs, n_pairs=n_pairs, first_idx=first_idx, W=W):
                hs.landed.synchronize()
                if int(hs.status[:n_pairs].abs().max()) != 0:
                    raise N.singular_matrix_error()                     # np.linalg.inv in the reference (motion.py:262)
                return [pool.submit(write_pair, hs, j, first_idx + j, W) for j in range(n_pairs)]

            hs.pending = encode()
            order.extend(str(first_idx + j) for j in range(n_pairs))
            if progress:
                progress(first_idx + n_pairs - 1)
        else:
            state["have"] = total
    for hs in (state["sets"] or []):
        flush(hs)
    pool.shutdown(wait=True)
    if state["window"] is None:
        raise Exception("Error reading video file: check the name of the video!")     # results.py:37-40
    ordered = {k: psnr_dict[k] for k in order}
    with open(os.path.join(save_path, "psnr_records.json"), "w") as outfile:
        json.dump(ordered, outfile)
    return ordered


def main(argv=None):
    parser = argparse.ArgumentParser(description="Launches GME and yields results (streaming B200 version of results.py)")
    parser.add_argument("-v", "--video-name", dest="path", type=str, required=True, help="name of the video to analyze")
    parser.add_argument("-f", "--frame-distance", dest="fd", type=str, required=True, help="frame displacement")
    parser.add_argument("--chunk", type=int, default=16, help="frame pairs per GPU batch")
    args = parser.parse_args(argv)
    video_path = os.path.join("resources", "videos", args.path)                  # results.py:20-34
    results_path = os.path.join("results", "")
    save_path = os.path.join(results_path, args.path.replace(".mp4", ""), "")
    if os.path.isdir(save_path):
        shutil.rmtree(save_path)
    os.makedirs(save_path)
    dump_results(video_path, save_path, int(args.fd), chunk=args.chunk)


if __name__ == "__main__":
    main()
