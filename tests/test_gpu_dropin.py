"""GPU: the drop-in modules (bbme / motion / utils with the reference's names and signatures) against
the reference's golden outputs -- these read like calls into the reference itself."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PARAM_TOL = dict(atol=1e-9, rtol=1e-9)


@pytest.fixture(scope="module")
def mods():
    import gme_device
    gme_device.require_cuda()
    import bbme
    import motion
    import utils
    return utils, bbme, motion


def test_bbme_api(mods, golden):
    utils, bbme, motion = mods
    g = golden("bbme_kat")
    prev, cur = g["prev"], g["cur"]
    for pn in (0, 1):
        for sp in (0, 1, 2, 3):
            mf = bbme.get_motion_field(prev, cur, block_size=12, search_window=12, searching_procedure=sp,
                                       pnorm_distance=pn)
            assert mf.dtype == np.int32 and mf.shape == (20, 26, 2)
            np.testing.assert_array_equal(mf, g[f"mf_pn{pn}_sp{sp}"])
    # the four search functions fill and return the array they are given
    mf = np.zeros((20, 26, 2), np.int32)
    out = bbme.diamond_search(prev, cur, mf, 240, 320, 1, 12)
    assert out is mf
    np.testing.assert_array_equal(mf, g["mf_pn1_sp3"])
    with pytest.raises(IndexError):
        bbme.get_motion_field(prev, cur, searching_procedure=7)
    assert bbme.compute_dfd(prev[:4, :4], cur[:4, :4], 1).dtype == np.float32
    assert bbme.searching_procedures[3] is bbme.diamond_search and bbme.pnorm_distances[0] is bbme.mae


def test_hierarchical_wrapper(mods, golden):
    utils, bbme, motion = mods
    g = golden("misc")
    for k in range(int(g["hw_n"])):
        bs, sw, sp = (int(v) for v in g[f"hw_a{k}"])
        want = g[f"hw_out{k}"]
        if want.dtype.kind == "U":
            with pytest.raises(ValueError):
                bbme.hierarchical_wrapper(g["hw_prev"], g["hw_cur"], bs, sw, sp)
        else:
            got = bbme.hierarchical_wrapper(g["hw_prev"], g["hw_cur"], block_size=bs, search_window=sw,
                                            searching_procedure=sp)
            assert got.dtype == np.float64
            np.testing.assert_array_equal(got, want)


def test_utils_api(mods, golden):
    utils, bbme, motion = mods
    g = golden("pyramid")
    for k in range(int(g["n"])):
        pyr = utils.get_pyramids(g[f"img{k}"])
        assert len(pyr) == 3 and pyr[2] is g[f"img{k}"] or np.array_equal(pyr[2], g[f"img{k}"])
        np.testing.assert_array_equal(pyr[1], g[f"l1_{k}"])
        np.testing.assert_array_equal(pyr[0], g[f"l0_{k}"])
    m = golden("misc")
    ps = utils.PSNR(m["cf_frame"], m["cf_other"])
    assert isinstance(ps, complex) and (ps.real, ps.imag) == tuple(m["psnr_fo"])
    assert utils.PSNR(m["cf_frame"], m["cf_frame"]) == -1
    draw = utils.draw_motion_field(m["cf_frame"], np.zeros((6, 8, 2), np.int32))
    assert draw.shape == (100, 132, 3)


@pytest.mark.parametrize("name", ["kat", "pan", "zoomrot", "odd"])
def test_motion_api(mods, golden, name):
    utils, bbme, motion = mods
    g = golden("gme_pipeline")
    prev, cur = g[f"{name}_prev"], g[f"{name}_cur"]
    assert motion.BBME_BLOCK_SIZE == 16 and motion.MOTION_VECTOR_ERROR_THRESHOLD_PERCENTAGE == .3
    # the dense first estimate runs on the coarsest pyramid level (motion.py:123-128,171)
    np.testing.assert_array_equal(motion.dense_motion_estimation(utils.get_pyramids(prev)[0], utils.get_pyramids(cur)[0]),
                                  g[f"{name}_dense"])
    first = motion.compute_first_parameters(g[f"{name}_dense"])
    assert first.dtype == np.float32
    np.testing.assert_array_equal(first, g[f"{name}_first"])
    params = motion.global_motion_estimation(prev, cur)
    assert params.dtype == np.float64 and params.shape == (6,)
    np.testing.assert_allclose(params, g[f"{name}_params"], **PARAM_TOL)
    shape = (prev.shape[0] // 16, prev.shape[1] // 16, 2)
    model = motion.get_motion_field_affine(shape, g[f"{name}_params"])
    assert model.dtype == np.int16
    np.testing.assert_array_equal(model, g[f"{name}_modelfield"])
    comp = motion.compensate_frame(prev, model)
    np.testing.assert_array_equal(comp, g[f"{name}_comp"])
    np.testing.assert_array_equal(motion.motion_compensation(prev, cur), g[f"{name}_mc"])
    np.testing.assert_allclose(motion.best_affine_parameters(prev, cur), g[f"{name}_nonrobust"], **PARAM_TOL)
    p = np.array([1.5, 0, 0, -2.25, 0, 0], dtype=np.float32)
    assert motion.parameter_projection(p) is p and p[0] == 3.0 and p[3] == -4.5
    # host NumPy matmul exactly as the reference (motion.py:102-104); BLAS may order the sum differently (1 ulp)
    np.testing.assert_allclose(motion.affine_model(2, 3, params),
                               [params[0] + 2 * params[1] + 3 * params[2], params[3] + 2 * params[4] + 3 * params[5]],
                               rtol=1e-14, atol=0)


def test_bbme_cli_on_a_lossless_clip(mods, tmp_path, monkeypatch):
    """Config 2 through the reference's CLI surface (bbme.py:617-712): frames fi-3 and fi of a lossless clip, diamond,
    bs 16 -- the two PNGs must be the needle diagrams of the oracle's fields.  Also pins the CLI quirk that -pn is
    parsed but never forwarded (always MSE)."""
    import cv2
    import gme_oracle as O
    import gme_synth as S
    utils, bbme, motion = mods
    seq = S.pan_sequence(6, 240, 320, step=(2, 1), seed=6)        # a geometry the reference own wrapper accepts
    clip = str(tmp_path / "clip.avi")
    writer = cv2.VideoWriter(clip, cv2.VideoWriter_fourcc(*"FFV1"), 30, (320, 240), isColor=True)
    if not writer.isOpened():
        pytest.skip("no lossless (FFV1) video writer in this OpenCV build")
    for f in seq:
        writer.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
    writer.release()
    frames = utils.get_video_frames(clip)
    if len(frames) != 6 or not all(np.array_equal(a, b) for a, b in zip(frames, seq)):
        pytest.skip("the clip did not decode losslessly on this box")
    monkeypatch.chdir(tmp_path)
    os.makedirs("resources/images")
    args = bbme._parser().parse_args(["-p", clip, "-fi", "4", "-pn", "0", "-bs", "16", "-sw", "16", "-sp", "3"])
    bbme.main(args)
    prev, cur = seq[1], seq[4]
    flat = utils.draw_motion_field(cur, O.get_motion_field(prev, cur, 16, 16, 3, 1))          # MSE despite -pn 0
    hier = utils.draw_motion_field(prev, O.hierarchical_wrapper(prev, cur, 16, 16, 3))
    np.testing.assert_array_equal(cv2.imread("resources/images/3-res.png"), flat)
    np.testing.assert_array_equal(cv2.imread("resources/images/3h-res.png"), hier)


_RANK_WORKER = r'''
import os, sys
root, port, rank, world = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
sys.path[:0] = [os.path.join(root, "global-motion-estimation_b200")]
import numpy as np, torch, torch.distributed as dist
import gme_device as D, gme_synth as S
from gme_distributed import run_sharded
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
seq = S.pan_sequence(10, 96, 160, step=(2, 1), seed=4)
d, n_pairs = 3, 7
planes = D.Planes.from_host(seq)                     # every rank holds the frames of its own pair range only in a real run

def compute(start, stop):                            # this rank's contiguous pair range on the GPU
    if stop == start:
        return torch.zeros((0, 7), dtype=torch.float64)
    pipe = D.Pipeline(stop - start, 96, 160)
    pipe.run(planes.view(start, stop), planes.view(start + d, stop + d))
    torch.cuda.synchronize()
    rows = torch.zeros((stop - start, 7), dtype=torch.float64)
    rows[:, :6] = pipe.params.cpu()
    rows[:, 6] = torch.tensor([p.real if p != -1 else -1.0 for p in pipe.psnr()], dtype=torch.float64)
    return rows

rows = run_sharded(n_pairs, compute)
if rank == 0:
    np.save(sys.argv[5], rows.numpy())
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [1, 2, 3])
def test_pair_sharding_gives_identical_rows(world, tmp_path):
    """SURVEY 8(e) determinism check: the gathered [pairs, 7] rows (6 parameters + PSNR) do not depend on how many
    ranks the pairs are sharded over.  The ranks share cuda:0 here (no rank waits on another inside a kernel); the
    collective runs over gloo, the same gather code the NCCL path uses."""
    import socket
    import subprocess
    import sys as _sys
    script = tmp_path / "rank.py"
    script.write_text(_RANK_WORKER)
    outs = {}
    for w in sorted({1, world}):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        out = tmp_path / f"rows_{w}.npy"
        procs = [subprocess.Popen([_sys.executable, str(script), ROOT, str(port), str(r), str(w), str(out)],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(w)]
        logs = [p.communicate(timeout=300)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), logs
        outs[w] = np.load(out)
    np.testing.assert_array_equal(outs[world], outs[1])
    assert outs[1].shape == (7, 7) and np.isfinite(outs[1]).all()


def test_edited_module_constants_are_honoured(mods, monkeypatch):
    """The reference is configured by editing module constants (README:137-141).  An edited outlier percentage reaches
    the fused pipeline; an edited block size takes the level-by-level path.  Both against the oracle."""
    import gme_oracle as O
    import gme_synth as S
    utils, bbme, motion = mods
    seq = S.zoom_rotate_sequence(2, 256, 384, zoom_per_frame=0.01, deg_per_frame=0.5, seed=17)
    prev, cur = seq[0], seq[1]

    def oracle_gme(pct):
        ppyr, cpyr = O.get_pyramids(prev), O.get_pyramids(cur)
        p = O.first_parameter_estimation(ppyr[0], cpyr[0])
        for i in (1, 2):
            p = O.parameter_projection(p)
            gt = O.get_motion_field(ppyr[i], cpyr[i], block_size=O.BBME_BLOCK_SIZE, searching_procedure=3)
            mask, _ = O.outlier_mask(gt, O.get_motion_field_affine(gt.shape, p), pct=pct)
            p = O._solve(gt, mask, ppyr[i].shape)
        return p

    monkeypatch.setattr(motion, "MOTION_VECTOR_ERROR_THRESHOLD_PERCENTAGE", .2)
    np.testing.assert_allclose(motion.global_motion_estimation(prev, cur), oracle_gme(.2), **PARAM_TOL)
    monkeypatch.setattr(motion, "MOTION_VECTOR_ERROR_THRESHOLD_PERCENTAGE", .3)
    np.testing.assert_allclose(motion.global_motion_estimation(prev, cur), oracle_gme(.3), **PARAM_TOL)
    assert not np.allclose(oracle_gme(.2), oracle_gme(.3), atol=1e-12)           # the constant does matter here

    monkeypatch.setattr(motion, "BBME_BLOCK_SIZE", 8)
    monkeypatch.setattr(O, "BBME_BLOCK_SIZE", 8)
    np.testing.assert_allclose(motion.global_motion_estimation(prev, cur), oracle_gme(.3), **PARAM_TOL)
