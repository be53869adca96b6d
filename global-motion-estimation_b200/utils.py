"""Drop-in for the reference's ``utils.py`` (same names, arguments and return types).

``get_pyramids`` and ``PSNR`` run on the B200 (gme_device / libgme_b200.so); the video,
drawing and reporting helpers are host-side OpenCV glue and are out of the hot path.
Reference line numbers: /root/reference/global_motion_estimation/utils.py.
"""
import cmath
import json
import os
import time

import cv2
import numpy as np

import gme_device as _dev


def get_video_frames(path):
    """utils.py:9-31 -- every frame of the video at ``path`` as a grayscale uint8 array."""
    capture = cv2.VideoCapture(path)
    frames = []
    while capture.grab():
        ok, frame = capture.retrieve()
        if frame.shape[2] == 3:
            frame = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
        frames.append(frame)
        if not ok:
            break
    return frames


def get_pyramids(original_image, levels=3):
    """utils.py:34-51 -- Gaussian pyramid, list ordered coarsest first, last entry the input itself."""
    pyramid = [original_image]
    if levels > 1:
        planes = _dev.Planes.from_host(np.asarray(original_image))
        for _ in range(1, levels):
            planes = _dev.pyr_down(planes)
            pyramid.insert(0, planes.to_host()[0])
    return pyramid


def draw_motion_field(frame, motion_field):
    """utils.py:54-76 -- needle diagram: one arrow per block, anchored at the block centre."""
    canvas = cv2.cvtColor(frame, cv2.COLOR_GRAY2RGB)
    rows, cols = motion_field.shape[0], motion_field.shape[1]
    block = frame.shape[0] // rows
    half = block // 2
    for by in range(rows):
        for bx in range(cols):
            vx, vy = motion_field[by][bx]
            tail = (bx * block + half, by * block + half)
            head = (int(tail[0] + vx), int(tail[1] + vy))
            cv2.arrowedLine(canvas, tail, head, (0, 0, 255), 1, line_type=cv2.LINE_AA)
    return canvas


def timer(func):
    """utils.py:79-97 -- decorator printing the wall time of ``func`` in whole seconds."""

    def wrapper(*args, **kwargs):
        started = int(time.time())
        result = func(*args, **kwargs)
        print(f"Execution of '{func.__name__}' in {int(time.time()) - started}s")
        return result

    return wrapper


def PSNR(original, noisy):
    """utils.py:100-116 -- peak signal-to-noise ratio; ``complex`` (cmath), or int -1 for identical images.

    The squared-error sum is an exact integer computed on the device."""
    original, noisy = np.asarray(original), np.asarray(noisy)
    if original.ndim == 2 and original.shape == noisy.shape and original.dtype == noisy.dtype == np.uint8:
        total = _dev.pair_session(*original.shape).sse(original, noisy)      # persistent session: pinned staging, one sync
        return _dev.psnr_from_sse(total, original.shape[0] * original.shape[1])
    a = _dev.Planes.from_host(original.astype(np.uint8, copy=False))
    b = _dev.Planes.from_host(noisy.astype(np.uint8, copy=False))
    total = int(_dev.sse(a, b).item())
    return _dev.psnr_from_sse(total, a.H * a.W)


def create_video_from_frames(frame_path, num_frames, video_name, fps=30):
    """utils.py:119-136 -- stitches '<i-3>-<i>.png' images into a video (reporting helper)."""
    images = [cv2.imread(f"{frame_path}{i - 3}-{i}.png") for i in range(3, num_frames)]
    height, width, _ = images[0].shape
    writer = cv2.VideoWriter(video_name, 0, fps, (width, height))
    for image in images:
        writer.write(image)
    cv2.destroyAllWindows()
    writer.release()


def some_data(psnr_path: str) -> None:
    """utils.py:138-164 -- average / variance / extrema of a psnr_records.json ('(<re>+<im>j)' strings)."""
    with open(psnr_path, "r") as f:
        records = json.load(f)
    values = np.zeros(shape=[len(records), 1])
    for k, text in enumerate(records.values()):
        values[k] = text[1:text.index("+")]
    avg = values.sum() / len(records)
    var = ((values - avg) ** 2).sum() / len(records)
    print("Average: {:.3f}".format(avg))
    print("Variance: {:.3f}".format(var))
    print("Standard deviation: {:.3f}".format(var ** (1 / 2)))
    print("Highest: {:.3f}".format(values.max()))
    print("Lowest: {:.3f}".format(values.min()))


if __name__ == "__main__":
    for d in os.listdir("results"):
        print(f"video {d}")
        some_data(os.path.join("results", d, "psnr_records.json"))
        print("======================")
