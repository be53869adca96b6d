"""Runs the reference's OWN drivers, unmodified, in the build container and records what they write.

    python oracle/make_results_golden.py        # rewrites tests/golden/results_*.json (about 4 minutes)

Test infrastructure (the reference is pure Python and cannot travel to the GPU box): the GPU test
tests/test_gpu_results_driver.py runs the same driver files -- results.py (results.py:14-112) and
"test scripts/motion_compensation.py" -- against the B200 drop-in modules on the same clips and compares
every file they write with the hashes recorded here: PNG bytes, decoded PNG pixels and psnr_records.json.

Clips
  * synth_pan.mp4  -- config 3 of BASELINE.json: 64-frame 720x480 exact-crop panning sequence
    (gme_synth.pan_sequence, seed 3, +2 columns / +1 row per frame), FFV1 in an .mp4 container, which
    cv2.VideoCapture decodes losslessly (checked on every use: the decoded frames are hashed).
  * pan240.mp4     -- the reference's own sample video (resources/videos/pan240.mp4, 207 frames 320x240, H.264);
    decode depends on the FFmpeg build of the image, so the fixture also holds the hash of the decoded frames and
    the test skips that case when the box decodes differently.
"""
from __future__ import annotations

import hashlib
import json
import os
import runpy
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
STAGED = os.path.join(ROOT, "baseline", "_ref", "global_motion_estimation")     # git-ignored; travels to the GPU box
REFERENCE = "/root/reference/global_motion_estimation"

SYNTH_CLIP = "synth_pan.mp4"
SYNTH_FRAMES, SYNTH_H, SYNTH_W, SYNTH_DISTANCE = 64, 480, 720, 3


def synth_frames() -> np.ndarray:
    sys.path.insert(0, os.path.join(ROOT, "global-motion-estimation_b200"))
    import gme_synth as S
    return S.pan_sequence(SYNTH_FRAMES, SYNTH_H, SYNTH_W, step=(2, 1), seed=3)


def write_lossless_clip(path: str, frames: np.ndarray) -> bool:
    """FFV1 (lossless) clip of grayscale frames; False when this OpenCV build has no such writer."""
    import cv2
    writer = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 30, (frames.shape[2], frames.shape[1]), isColor=True)
    if not writer.isOpened():
        return False
    for f in frames:
        writer.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
    writer.release()
    return True


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()[:24]


def frames_digest(frames) -> str:
    h = hashlib.sha256()
    for f in frames:
        h.update(np.ascontiguousarray(f).tobytes())
    return h.hexdigest()[:24]


def digest_tree(top: str) -> dict:
    """{relative path: [sha of the file bytes, sha of the decoded pixels + shape]} for every PNG under `top`."""
    import cv2
    out = {}
    for d, _, files in sorted(os.walk(top)):
        for name in sorted(files):
            if not name.endswith(".png"):
                continue
            p = os.path.join(d, name)
            raw = open(p, "rb").read()
            img = cv2.imread(p, cv2.IMREAD_UNCHANGED)
            out[os.path.relpath(p, top)] = [sha(raw), sha(np.ascontiguousarray(img).tobytes() + str(img.shape).encode())]
    return out


STAGED_FILES = ("results.py", os.path.join("test scripts", "motion_compensation.py"),
                os.path.join("resources", "videos", "pan240.mp4"))


def stage_reference() -> bool:
    """Copies the reference's DRIVER scripts and its sample video -- the callers of the path, not the modules the
    drop-in replaces -- to baseline/_ref/ (git-ignored, not gpurun-ignored) so the GPU box can run them unchanged
    against the drop-in modules.  Without /root/reference: True when a staged copy is already there."""
    if not os.path.isdir(REFERENCE):
        return os.path.isfile(os.path.join(STAGED, "results.py"))
    if os.path.isdir(STAGED):
        shutil.rmtree(STAGED)
    for rel in STAGED_FILES:
        os.makedirs(os.path.dirname(os.path.join(STAGED, rel)), exist_ok=True)
        shutil.copy(os.path.join(REFERENCE, rel), os.path.join(STAGED, rel))
    return True


def run_driver(script: str, argv: list, module_dir: str, cwd: str) -> None:
    """runpy of one driver file with `module_dir` first on sys.path (bbme / motion / utils resolve there)."""
    saved_path, saved_argv, saved_cwd = list(sys.path), list(sys.argv), os.getcwd()
    saved_mods = {k: sys.modules.pop(k, None) for k in ("utils", "bbme", "motion")}
    sys.path.insert(0, module_dir)
    sys.argv = [script] + argv
    os.chdir(cwd)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)
            runpy.run_path(script, run_name="__main__")
    finally:
        os.chdir(saved_cwd)
        sys.path[:] = saved_path
        sys.argv[:] = saved_argv
        for k in ("utils", "bbme", "motion"):
            sys.modules.pop(k, None)
            if saved_mods[k] is not None:
                sys.modules[k] = saved_mods[k]


def main():
    import types
    t = types.ModuleType("tkinter")
    t.image_names = lambda *a, **k: ()
    sys.modules.setdefault("tkinter", t)                 # utils.py:2
    if not hasattr(np, "infty"):
        np.infty = np.inf                                # bbme.py:143 ...
    import cv2
    os.makedirs(GOLDEN, exist_ok=True)
    work = tempfile.mkdtemp(prefix="gme_results_")
    os.makedirs(os.path.join(work, "resources", "videos"))
    sys.path.insert(0, REFERENCE)
    import utils as ref_utils                            # the reference's decoder, for the frame digests
    sys.path.remove(REFERENCE)
    for k in ("utils",):
        sys.modules.pop(k, None)

    # ---- config 3: results.py -v synth_pan.mp4 -f 3 ---------------------------------------------------------
    seq = synth_frames()
    clip = os.path.join(work, "resources", "videos", SYNTH_CLIP)
    assert write_lossless_clip(clip, seq), "no FFV1 writer"
    decoded = ref_utils.get_video_frames(clip)
    assert len(decoded) == SYNTH_FRAMES and all(np.array_equal(a, b) for a, b in zip(decoded, seq)), "clip is not lossless"
    run_driver(os.path.join(REFERENCE, "results.py"), ["-v", SYNTH_CLIP, "-f", str(SYNTH_DISTANCE)], REFERENCE, work)
    top = os.path.join(work, "results", SYNTH_CLIP.replace(".mp4", ""))
    fixture = {"clip": SYNTH_CLIP, "frames": SYNTH_FRAMES, "height": SYNTH_H, "width": SYNTH_W, "distance": SYNTH_DISTANCE,
               "frames_digest": frames_digest(seq), "cv2": cv2.__version__, "numpy": np.__version__,
               "psnr_records": json.load(open(os.path.join(top, "psnr_records.json"))), "files": digest_tree(top)}
    json.dump(fixture, open(os.path.join(GOLDEN, "results_synth_pan.json"), "w"), indent=0, sort_keys=True)
    print("results.py synth_pan:", len(fixture["files"]), "PNGs,", len(fixture["psnr_records"]), "PSNR records")

    # ---- the reference's sample video through results.py -f 3 and "test scripts/motion_compensation.py" ----------
    shutil.copy(os.path.join(REFERENCE, "resources", "videos", "pan240.mp4"), os.path.join(work, "resources", "videos", "pan240.mp4"))
    decoded = ref_utils.get_video_frames(os.path.join(work, "resources", "videos", "pan240.mp4"))
    run_driver(os.path.join(REFERENCE, "results.py"), ["-v", "pan240.mp4", "-f", "3"], REFERENCE, work)
    top = os.path.join(work, "results", "pan240")
    fixture = {"clip": "pan240.mp4", "frames": len(decoded), "height": int(decoded[0].shape[0]), "width": int(decoded[0].shape[1]),
               "distance": 3, "frames_digest": frames_digest(decoded), "cv2": cv2.__version__, "numpy": np.__version__,
               "psnr_records": json.load(open(os.path.join(top, "psnr_records.json"))), "files": digest_tree(top)}
    json.dump(fixture, open(os.path.join(GOLDEN, "results_pan240.json"), "w"), indent=0, sort_keys=True)
    print("results.py pan240:", len(fixture["files"]), "PNGs,", len(fixture["psnr_records"]), "PSNR records")

    os.makedirs(os.path.join(work, "results", "pan240_mse"), exist_ok=True)      # the script only creates the leaf
    run_driver(os.path.join(REFERENCE, "test scripts", "motion_compensation.py"), [], REFERENCE, work)
    top = os.path.join(work, "results", "pan240_mse")
    fixture = {"clip": "pan240.mp4", "frames_digest": frames_digest(decoded),
               "psnr_records": json.load(open(os.path.join(top, "psnr_records.json"))), "files": digest_tree(top)}
    json.dump(fixture, open(os.path.join(GOLDEN, "results_motion_compensation_script.json"), "w"), indent=0, sort_keys=True)
    print("motion_compensation.py pan240:", len(fixture["files"]), "PNGs,", len(fixture["psnr_records"]), "PSNR records")
    shutil.rmtree(work)
    stage_reference()


if __name__ == "__main__":
    main()
