"""GPU: the reference's OWN driver files, unmodified, run against the B200 drop-in modules.

BASELINE.json north_star: "results.py and the bbme.py CLI run unchanged as a drop-in".  The driver files are the
reference's (staged under the git-ignored baseline/_ref/ by __graft_entry__.build(); /root/reference does not exist
on the GPU box); bbme / motion / utils resolve to global-motion-estimation_b200/ through dropin_run.py.  Every PNG
the driver writes must equal, byte for byte, what the same driver wrote on top of the real reference modules in the
build container (hashes in tests/golden/results_*.json, made by oracle/make_results_golden.py), and
psnr_records.json must hold the same strings.
"""
import json
import os
import shutil
import subprocess
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref", "global_motion_estimation")
LAUNCHER = os.path.join(ROOT, "global-motion-estimation_b200", "dropin_run.py")


def _fixture(name):
    with open(os.path.join(ROOT, "tests", "golden", name + ".json")) as f:
        return json.load(f)


def _staged(*parts):
    import make_results_golden as G
    G.stage_reference()                                  # copies from /root/reference when that exists (build container)
    p = os.path.join(STAGED, *parts)
    if not os.path.isfile(p):
        pytest.skip(f"reference driver not staged at {p} (run __graft_entry__.build() where /root/reference exists)")
    return p


def _run_driver(script, argv, cwd):
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, LAUNCHER, script] + argv, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-4000:]
    return time.perf_counter() - t0


def _compare(top, fixture):
    import make_results_golden as G
    got = G.digest_tree(top)
    assert sorted(got) == sorted(fixture["files"]), "the driver wrote a different set of files"
    pixel_diff = [k for k in got if got[k][1] != fixture["files"][k][1]]
    assert not pixel_diff, f"{len(pixel_diff)} PNGs hold different pixels, e.g. {pixel_diff[:5]}"
    byte_diff = [k for k in got if got[k][0] != fixture["files"][k][0]]
    assert not byte_diff, f"{len(byte_diff)} PNGs have the same pixels but different bytes, e.g. {byte_diff[:5]}"
    with open(os.path.join(top, "psnr_records.json")) as f:
        psnr = json.load(f)
    assert sorted(psnr) == sorted(fixture["psnr_records"])
    for k, want in fixture["psnr_records"].items():
        assert psnr[k] == want, (k, psnr[k], want)       # str(complex): exact integer SSE / N, then the same cmath tail


def test_results_py_unchanged_on_the_config3_clip(tmp_path):
    """Config 3: results.py -v synth_pan.mp4 -f 3 (results.py:41-112), 61 pairs of a 64-frame 720x480 pan."""
    import make_results_golden as G
    fx = _fixture("results_synth_pan")
    script = _staged("results.py")
    seq = G.synth_frames()
    assert G.frames_digest(seq) == fx["frames_digest"], "the synthetic sequence differs from the build container's"
    os.makedirs(tmp_path / "resources" / "videos")
    clip = str(tmp_path / "resources" / "videos" / fx["clip"])
    if not G.write_lossless_clip(clip, seq):
        pytest.skip("no lossless (FFV1) video writer in this OpenCV build")
    import utils
    decoded = utils.get_video_frames(clip)
    if len(decoded) != len(seq) or G.frames_digest(decoded) != fx["frames_digest"]:
        pytest.skip("the clip did not decode losslessly on this box")
    dt = _run_driver(script, ["-v", fx["clip"], "-f", str(fx["distance"])], str(tmp_path))
    _compare(str(tmp_path / "results" / fx["clip"].replace(".mp4", "")), fx)
    print(f"results.py on the drop-in: {len(fx['psnr_records'])} pairs in {dt:.1f} s (incl. import, decode, PNG writes)")


def _stage_sample_video(tmp_path, fx):
    import make_results_golden as G
    video = _staged("resources", "videos", "pan240.mp4")
    os.makedirs(tmp_path / "resources" / "videos", exist_ok=True)
    shutil.copy(video, tmp_path / "resources" / "videos" / "pan240.mp4")
    import utils
    decoded = utils.get_video_frames(str(tmp_path / "resources" / "videos" / "pan240.mp4"))
    if G.frames_digest(decoded) != fx["frames_digest"]:
        pytest.skip("this box decodes pan240.mp4 (H.264) differently from the build container")


def test_results_py_unchanged_on_the_reference_sample_video(tmp_path):
    """results.py -v pan240.mp4 -f 3 on the reference's own sample video: 204 pairs, 1020 PNGs."""
    fx = _fixture("results_pan240")
    script = _staged("results.py")
    _stage_sample_video(tmp_path, fx)
    _run_driver(script, ["-v", "pan240.mp4", "-f", "3"], str(tmp_path))
    _compare(str(tmp_path / "results" / "pan240"), fx)


def test_motion_compensation_script_unchanged(tmp_path):
    """"test scripts/motion_compensation.py" -- the only other reference script on the live API:
    motion.motion_compensation + utils.PSNR over pan240.mp4 at distance 3."""
    fx = _fixture("results_motion_compensation_script")
    script = _staged("test scripts", "motion_compensation.py")
    _stage_sample_video(tmp_path, fx)
    os.makedirs(tmp_path / "results" / "pan240_mse")
    _run_driver(script, [], str(tmp_path))
    _compare(str(tmp_path / "results" / "pan240_mse"), fx)


def test_batched_api_equals_the_per_pair_driver_on_config3():
    """Config 3 through the batched surface: all 61 pairs of the same clip in one gme_pipeline call against the oracle
    (fields, masks, parameters, compensated frames, squared errors) -- and the PSNR strings of results.py's JSON."""
    import torch
    import gme_device as D
    import gme_oracle as O
    import make_results_golden as G
    fx = _fixture("results_synth_pan")
    seq = G.synth_frames()
    d = fx["distance"]
    n = len(seq) - d
    planes = D.Planes.from_host(seq)
    pipe = D.Pipeline(n, seq.shape[1], seq.shape[2])
    pipe.run(planes.view(0, n), planes.view(d, d + n))
    torch.cuda.synchronize()
    assert int(pipe.status.abs().max().item()) == 0
    params = pipe.params.cpu().numpy()
    f1, f2 = pipe.intermediate(1).cpu().numpy(), pipe.intermediate(2).cpu().numpy()
    m1, m2 = pipe.intermediate(3).cpu().numpy(), pipe.intermediate(4).cpu().numpy()
    comp = pipe.comp.to_host()
    psnr = pipe.psnr()
    for k in range(n):
        want, inter = O.global_motion_estimation(seq[k], seq[k + d], return_intermediates=True)
        np.testing.assert_array_equal(f1[k], inter[1]["gt"])
        np.testing.assert_array_equal(f2[k], inter[2]["gt"])
        np.testing.assert_array_equal(m1[k].astype(bool), inter[1]["outlier"])
        np.testing.assert_array_equal(m2[k].astype(bool), inter[2]["outlier"])
        np.testing.assert_allclose(params[k], want, atol=1e-9, rtol=1e-9)
        model = O.get_motion_field_affine((seq.shape[1] // 16, seq.shape[2] // 16), want)
        np.testing.assert_array_equal(comp[k], O.compensate_frame(seq[k], model))
        assert str(psnr[k]) == fx["psnr_records"][str(k + d)]


def test_streaming_results_dump_writes_the_same_tree(tmp_path):
    """gme_results.dump_results -- decode thread -> pinned chunks -> batched kernels -> PNG encoder pool
    -- writes, byte for byte, the tree the reference's results.py wrote for the config-3 clip."""
    import make_results_golden as G
    import gme_results
    fx = _fixture("results_synth_pan")
    seq = G.synth_frames()
    clip = str(tmp_path / fx["clip"])
    if not G.write_lossless_clip(clip, seq):
        pytest.skip("no lossless (FFV1) video writer in this OpenCV build")
    import utils
    if G.frames_digest(utils.get_video_frames(clip)) != fx["frames_digest"]:
        pytest.skip("the clip did not decode losslessly on this box")
    out = str(tmp_path / "out")
    os.makedirs(out)
    t0 = time.perf_counter()
    psnr = gme_results.dump_results(clip, os.path.join(out, ""), fx["distance"], chunk=16)
    dt = time.perf_counter() - t0
    assert list(psnr) == [str(k) for k in range(fx["distance"], fx["frames"])]      # the order results.py inserts them
    _compare(out, fx)
    print(f"streaming dump: {len(psnr)} pairs, {5 * len(psnr)} PNGs in {dt:.2f} s")


def test_frame_prefetcher_equals_get_video_frames(tmp_path):
    """The pinned-memory decoder yields exactly the frames utils.get_video_frames returns (utils.py:9-31), in order,
    for chunk sizes that do and do not divide the frame count."""
    import make_results_golden as G
    import gme_results
    import gme_synth as S
    import utils
    seq = S.pan_sequence(11, 96, 160, step=(2, 1), seed=9)
    clip = str(tmp_path / "c.mp4")
    if not G.write_lossless_clip(clip, seq):
        pytest.skip("no lossless (FFV1) video writer in this OpenCV build")
    want = utils.get_video_frames(clip)
    for chunk in (4, 11, 16):
        pf = gme_results.FramePrefetcher(clip, chunk=chunk)
        got = []
        for slot, frames, k in pf:
            assert frames.is_pinned() and frames.shape[0] == k
            got.extend(f.numpy().copy() for f in frames)
            pf.release(slot)
        assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))


def test_streaming_dump_equals_the_unchanged_driver_on_an_odd_geometry(tmp_path):
    """A 250 x 330 clip (neither dimension a multiple of 16: padded pitch, partial last tiles, chunks that do not divide
    the pair count): the streaming dump and the reference's unchanged results.py on top of the drop-in modules must
    write the same bytes."""
    import make_results_golden as G
    import gme_results
    import gme_synth as S
    script = _staged("results.py")
    seq = S.zoom_rotate_sequence(13, 250, 330, zoom_per_frame=0.004, deg_per_frame=0.3, seed=31)
    os.makedirs(tmp_path / "resources" / "videos")
    clip = str(tmp_path / "resources" / "videos" / "odd.mp4")
    if not G.write_lossless_clip(clip, seq):
        pytest.skip("no lossless (FFV1) video writer in this OpenCV build")
    _run_driver(script, ["-v", "odd.mp4", "-f", "2"], str(tmp_path))
    want = G.digest_tree(str(tmp_path / "results" / "odd"))
    out = str(tmp_path / "stream")
    os.makedirs(out)
    psnr = gme_results.dump_results(clip, os.path.join(out, ""), 2, chunk=4)
    assert G.digest_tree(out) == want and len(want) == 5 * 11
    with open(tmp_path / "results" / "odd" / "psnr_records.json") as f:
        assert json.load(f) == psnr
