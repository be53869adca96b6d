"""GPU: the CUDA path (through the C-ABI library, via gme_device's ctypes wrappers) against the
committed golden outputs of the reference and against the CPU oracle on seeded inputs.

Bars: motion fields, outlier masks, thresholds, model fields, compensated frames and squared-error
sums are BIT-EXACT; affine parameters within atol = rtol = 1e-9 (PARAM_TOL; the reference accumulates
the normal equations in float64 term by term, the kernel sums exact integers and rounds once); PSNR
within 1e-12 relative."""
import os

import numpy as np
import pytest
import torch

import gme_oracle as O
import gme_synth as S

pytestmark = pytest.mark.gpu

PARAM_TOL = dict(atol=1e-9, rtol=1e-9)


@pytest.fixture(scope="module")
def D():
    import gme_device
    gme_device.require_cuda()
    return gme_device


def field_of(D, prev, cur, bs, sw, sp, pn):
    f = D.motion_field(D.Planes.from_host(prev), D.Planes.from_host(cur), bs, sw, sp, pn)
    torch.cuda.synchronize()
    return f.cpu().numpy()


# ------------------------------------------------------------------ BBME vs golden (reference outputs)
@pytest.mark.parametrize("pn", [0, 1])
@pytest.mark.parametrize("sp", [0, 1, 2, 3])
def test_bbme_kat(D, golden, pn, sp):
    g = golden("bbme_kat")
    got = field_of(D, g["prev"], g["cur"], 12, 12, sp, pn)[0]
    np.testing.assert_array_equal(got, g[f"mf_pn{pn}_sp{sp}"])


def test_bbme_static_quirk(D, golden):
    g = golden("bbme_kat")
    np.testing.assert_array_equal(field_of(D, g["prev"], g["prev"], 12, 12, 3, 0)[0], g["static_bs12"])
    np.testing.assert_array_equal(field_of(D, g["prev"], g["prev"], 16, 12, 3, 1)[0], g["static_bs16"])


def test_bbme_random_geometries(D, golden):
    """240 reference cases: all procedures x norms, block sizes 2..16 incl. the generic path (3, 6, 10),
    frames that are not multiples of bs, windows 1..16, flat / tie-heavy inputs."""
    g = golden("bbme_random")
    bad = []
    for k in range(int(g["n"])):
        bs, sw, sp, pn, inp = (int(v) for v in g[f"a{k}"])
        got = field_of(D, g[f"p{inp}"], g[f"c{inp}"], bs, sw, sp, pn)[0]
        if not np.array_equal(got, g[f"mf{k}"]):
            bad.append((k, bs, sw, sp, pn, int((got != g[f"mf{k}"]).sum())))
    assert not bad, bad


def test_bbme_mse_block_sizes_beyond_16_round_like_float32(D, golden):
    """128 reference cases with MSE and block sizes 17..40, where the reference's float32 cost exceeds 2^24 and rounds
    at the nodes of NumPy's pairwise sum: high-contrast inputs, some built so that candidates tie exactly
    as integers and the rounding decides the winner (an exact-integer argmin gets those blocks wrong)."""
    g = golden("bbme_f32_rounding")
    bad = []
    for k in range(int(g["n"])):
        bs, sw, sp, pn, inp = (int(v) for v in g[f"a{k}"])
        got = field_of(D, g[f"p{inp}"], g[f"c{inp}"], bs, sw, sp, pn)[0]
        if not np.array_equal(got, g[f"mf{k}"]):
            bad.append((k, bs, sw, sp, int((got != g[f"mf{k}"]).sum())))
    assert not bad, bad


# ------------------------------------------------------------------ BBME vs oracle, larger seeded inputs
@pytest.mark.parametrize("H,W,bs,sw,sp,pn", [
    (240, 320, 12, 12, 0, 0),          # config 1
    (480, 720, 16, 16, 3, 1),          # config 2
    (480, 720, 16, 16, 0, 1),
    (270, 480, 2, 2, 3, 1),            # dense L0 of 1080p
    (540, 960, 16, 16, 1, 1),
    (540, 960, 16, 16, 2, 1),
    (1080, 1920, 16, 2, 3, 1),
    (1080, 1920, 16, 16, 1, 0),
    (1080, 1920, 16, 16, 2, 1),
    (250, 333, 8, 5, 0, 0), (250, 333, 8, 5, 3, 0), (250, 333, 4, 3, 1, 1), (250, 333, 4, 6, 2, 0),
    (120, 180, 16, 32, 0, 1),          # window larger than the frame in places
    (100, 100, 20, 4, 0, 0), (100, 100, 20, 4, 3, 0), (100, 100, 7, 3, 1, 0), (64, 96, 5, 9, 2, 0),
    (200, 260, 16, 60, 0, 1),          # window copies exceed the shared-memory budget: first-generation tiled kernel
    (64, 300, 8, 130, 0, 0),           # more than 256 row offsets: key packing of the second kernel does not apply
    (48, 64, 4, 6, 0, 1), (40, 52, 2, 5, 0, 0),      # block sizes 4 and 2: first-generation tiled kernel
])
def test_bbme_vs_oracle(D, H, W, bs, sw, sp, pn):
    seq = S.zoom_rotate_sequence(2, H, W, zoom_per_frame=0.01, deg_per_frame=0.6, seed=H + bs + sp)
    prev, cur = seq[0], np.roll(seq[1], (3, -5), (0, 1))
    want = O.get_motion_field(prev, cur, bs, sw, sp, pn, threads=8)
    got = field_of(D, prev, cur, bs, sw, sp, pn)[0]
    np.testing.assert_array_equal(got, want)


def test_bbme_large_motion_leaves_staged_window(D):
    """Vectors far beyond the staged margin: the global-memory path must give the same integers."""
    base = S.texture(300, 700, seed=21)
    prev, cur = np.ascontiguousarray(base[20:276, 100:612]), np.ascontiguousarray(base[20:276, 30:542])  # 70 px pan
    for sp, sw in ((3, 0), (2, 64), (1, 60)):
        for bs in (16, 8, 2):
            want = O.get_motion_field(prev, cur, bs, sw, sp, 1, threads=8)
            got = field_of(D, prev, cur, bs, sw, sp, 1)[0]
            np.testing.assert_array_equal(got, want, err_msg=f"sp={sp} bs={bs}")
    # the walks really left the staged window (margin 32 px around the tile)
    assert np.abs(field_of(D, prev, cur, 16, 0, 3, 1)[0][..., 0]).max() > 32
    assert np.abs(field_of(D, prev, cur, 16, 64, 2, 1)[0][..., 0]).max() > 32


def test_bbme_batch_equals_single_and_unaligned_pitch(D):
    """A batch gives what each pair gives alone; W % 16 != 0 (cooperative loader instead of TMA) agrees too."""
    seq = S.pan_sequence(6, 96, 150, step=(3, -1), seed=8)       # W = 150: pitch padded to 160
    planes = D.Planes.from_host(seq)
    prev, cur = planes.view(0, 4), planes.view(2, 6)             # two views of one buffer, distance 2
    for sp, bs, sw in ((0, 8, 6), (3, 16, 0), (1, 12, 8), (2, 4, 7)):
        batch = D.motion_field(prev, cur, bs, sw, sp, 0).cpu().numpy()
        for k in range(4):
            np.testing.assert_array_equal(batch[k], O.get_motion_field(seq[k], seq[k + 2], bs, sw, sp, 0))
    # a 4-byte aligned but not 16-byte aligned pitch takes the non-TMA path
    t = torch.zeros((2, 96, 156), dtype=torch.uint8, device="cuda")
    t[:, :, :150] = torch.from_numpy(seq[:2]).cuda()
    odd = D.Planes(t, 150)
    a = D.motion_field(odd.view(0, 1), odd.view(1, 2), 16, 8, 0, 1).cpu().numpy()[0]
    np.testing.assert_array_equal(a, O.get_motion_field(seq[0], seq[1], 16, 8, 0, 1))
    b = D.motion_field(odd.view(0, 1), odd.view(1, 2), 16, 0, 3, 1).cpu().numpy()[0]
    np.testing.assert_array_equal(b, O.get_motion_field(seq[0], seq[1], 16, 0, 3, 1))


def test_bbme_error_conventions(D):
    p = D.Planes.from_host(np.zeros((16, 40), np.uint8))
    with pytest.raises(IndexError):
        D.motion_field(p, p, 4, 2, 4, 0)
    with pytest.raises(IndexError):
        D.motion_field(p, p, 4, 2, 0, 2)
    with pytest.raises(ValueError):
        D.motion_field(p, p, 16, 2, 3, 0)        # diamond with H <= bs: undefined in the reference
    assert D.motion_field(p, p, 64, 2, 0, 0).shape == (1, 0, 0, 2)


# ------------------------------------------------------------------ pyramid
def test_pyramid_golden_and_cv2(D, golden):
    import cv2
    g = golden("pyramid")
    for k in range(int(g["n"])):
        img = g[f"img{k}"]
        l1 = D.pyr_down(D.Planes.from_host(img))
        l0 = D.pyr_down(l1)
        np.testing.assert_array_equal(l1.to_host()[0], g[f"l1_{k}"])
        np.testing.assert_array_equal(l0.to_host()[0], g[f"l0_{k}"])
    rng = np.random.default_rng(3)
    for H, W in ((1080, 1920), (480, 720), (2160, 3840), (1079, 1917), (33, 16), (5, 40)):
        batch = rng.integers(0, 256, (3, H, W), dtype=np.uint8)
        got = D.pyr_down(D.Planes.from_host(batch)).to_host()
        for k in range(3):
            np.testing.assert_array_equal(got[k], cv2.pyrDown(batch[k]))
            if H * W < 10 ** 6:
                np.testing.assert_array_equal(got[k], O.pyr_down(batch[k]))


# ------------------------------------------------------------------ fit pieces
def test_affine_field_golden(D, golden):
    g = golden("misc")
    for k in range(int(g["af_n"])):
        p = torch.from_numpy(g[f"af_p{k}"].astype(np.float64)).reshape(1, 6).cuda()
        got = D.affine_field(p, 9, 13)[0].cpu().numpy()
        assert got.dtype == np.int16
        np.testing.assert_array_equal(got, g[f"af_out{k}"])


def test_first_parameters_and_fit_golden(D, golden):
    g = golden("gme_pipeline")
    for name in ("kat", "pan", "zoomrot", "odd"):
        dense = torch.from_numpy(g[f"{name}_dense"]).unsqueeze(0).cuda()
        first = D.first_parameters(dense)[0].cpu().numpy()
        np.testing.assert_array_equal(first.astype(np.float32), g[f"{name}_first"])
        params = torch.from_numpy(first.copy()).reshape(1, 6).cuda()
        prev = g[f"{name}_prev"]
        l1 = ((prev.shape[0] + 1) // 2, (prev.shape[1] + 1) // 2)
        for li, shape in ((1, l1), (2, prev.shape)):
            gt = torch.from_numpy(g[f"{name}_gt{li}"]).unsqueeze(0).cuda()
            params, status, outlier, thr, model = D.affine_fit(gt, shape, params, robust=True, project=True,
                                                               intermediates=True)
            assert int(status.item()) == 0
            np.testing.assert_array_equal(model[0].cpu().numpy(), g[f"{name}_model{li}"])
            np.testing.assert_array_equal(outlier[0].cpu().numpy().astype(bool), g[f"{name}_outlier{li}"])
            assert int(thr.item()) == int(g[f"{name}_thr{li}"])
        np.testing.assert_allclose(params[0].cpu().numpy(), g[f"{name}_params"], **PARAM_TOL)
        # non-robust variant (motion.best_affine_parameters)
        gt2 = torch.from_numpy(g[f"{name}_gt2"]).unsqueeze(0).cuda()
        p0 = torch.zeros((1, 6), dtype=torch.float64, device="cuda")
        p0, status = D.affine_fit(gt2, prev.shape, p0, robust=False)
        np.testing.assert_allclose(p0[0].cpu().numpy(), g[f"{name}_nonrobust"], **PARAM_TOL)


def test_fit_order_statistic_edge_cases(D):
    """Thresholds on tiny fields (int(.3*N) == 0 -> Python's [-0]), heavy ties and large differences."""
    rng = np.random.default_rng(5)
    for R, C, lo, hi in ((1, 3, -3, 4), (2, 2, -1, 2), (3, 5, 0, 1), (9, 13, -5000, 5000), (40, 70, -2, 3)):
        gt = rng.integers(lo, hi, (R, C, 2)).astype(np.int32)
        old = np.array([0.3, 0.01, -0.02, -0.4, 0.0, 0.015])
        model = O.get_motion_field_affine((R, C), old)
        want_mask, want_thr = O.outlier_mask(gt, model)
        p = torch.from_numpy(old.copy()).reshape(1, 6).cuda()
        _, status, outlier, thr, m = D.affine_fit(torch.from_numpy(gt).unsqueeze(0).cuda(), (R * 16, C * 16), p,
                                                  robust=True, project=False, intermediates=True)
        np.testing.assert_array_equal(m[0].cpu().numpy(), model)
        assert int(thr.item()) == want_thr
        np.testing.assert_array_equal(outlier[0].cpu().numpy().astype(bool), want_mask)


def test_fit_singular_matrix_flag(D):
    """All kept blocks collinear (one block row): np.linalg.inv raises LinAlgError in the reference."""
    gt = torch.zeros((1, 1, 6, 2), dtype=torch.int32, device="cuda")
    p = torch.zeros((1, 6), dtype=torch.float64, device="cuda")
    _, status = D.affine_fit(gt, (16, 96), p, robust=True)
    assert int(status.item()) == 1
    with pytest.raises(np.linalg.LinAlgError):
        O._solve(gt[0].cpu().numpy(), None, (16, 96))


# ------------------------------------------------------------------ compensation + PSNR
def test_compensate_golden_and_oracle(D, golden):
    g = golden("misc")
    frame, other = g["cf_frame"], g["cf_other"]
    fp, op = D.Planes.from_host(frame), D.Planes.from_host(other)
    for k in range(int(g["cf_n"])):
        mf = g[f"cf_mf{k}"]
        for dtype in (torch.int16, torch.int32):
            comp, sse = D.compensate(fp, torch.from_numpy(mf.astype(np.int32)).to(dtype).unsqueeze(0).cuda(), op)
            np.testing.assert_array_equal(comp.to_host()[0], g[f"cf_out{k}"])
            assert int(sse.item()) == O.sse(other, g[f"cf_out{k}"])
    assert int(D.sse(fp, op).item()) == O.sse(frame, other)
    ps = D.psnr_from_sse(int(D.sse(fp, op).item()), frame.size)
    assert (ps.real, ps.imag) == tuple(g["psnr_fo"])
    assert D.psnr_from_sse(int(D.sse(fp, fp).item()), frame.size) == -1
    rng = np.random.default_rng(9)
    for H, W, R, C in ((1080, 1920, 67, 120), (480, 720, 30, 45), (250, 333, 15, 20), (96, 100, 8, 8), (64, 64, 70, 3)):
        f = rng.integers(0, 256, (2, H, W), dtype=np.uint8)
        c = rng.integers(0, 256, (2, H, W), dtype=np.uint8)
        mf = rng.integers(-40, 41, (2, R, C, 2)).astype(np.int16)
        comp, sse = D.compensate(D.Planes.from_host(f), torch.from_numpy(mf).cuda(), D.Planes.from_host(c))
        got = comp.to_host()
        for k in range(2):
            want = O.compensate_frame(f[k], mf[k])
            np.testing.assert_array_equal(got[k], want)
            assert int(sse[k].item()) == O.sse(c[k], want)


# ------------------------------------------------------------------ whole pipeline
@pytest.mark.parametrize("name", ["kat", "pan", "zoomrot", "odd"])
def test_pipeline_golden(D, golden, name):
    g = golden("gme_pipeline")
    prev, cur = g[f"{name}_prev"], g[f"{name}_cur"]
    pipe = D.Pipeline(1, prev.shape[0], prev.shape[1])
    pipe.run(D.Planes.from_host(prev), D.Planes.from_host(cur))
    torch.cuda.synchronize()
    assert int(pipe.status.item()) == 0
    np.testing.assert_array_equal(pipe.intermediate(0)[0].cpu().numpy(), g[f"{name}_dense"])
    np.testing.assert_array_equal(pipe.intermediate(1)[0].cpu().numpy(), g[f"{name}_gt1"])
    np.testing.assert_array_equal(pipe.intermediate(2)[0].cpu().numpy(), g[f"{name}_gt2"])
    np.testing.assert_array_equal(pipe.intermediate(3)[0].cpu().numpy().astype(bool), g[f"{name}_outlier1"])
    np.testing.assert_array_equal(pipe.intermediate(4)[0].cpu().numpy().astype(bool), g[f"{name}_outlier2"])
    np.testing.assert_allclose(pipe.params[0].cpu().numpy(), g[f"{name}_params"], **PARAM_TOL)
    np.testing.assert_array_equal(pipe.intermediate(5)[0].cpu().numpy(), g[f"{name}_modelfield"])
    np.testing.assert_array_equal(pipe.comp.to_host()[0], g[f"{name}_comp"])
    ps = pipe.psnr()[0]
    assert isinstance(ps, complex) and abs(ps.real - g[f"{name}_psnr"][0]) <= 1e-12 * abs(ps.real)


@pytest.mark.parametrize("sp,sw", [(0, 6), (1, 16), (2, 16)])
def test_pipeline_search_override_golden(D, golden, sp, sw):
    g = golden("gme_pipeline")
    prev, cur = g["zoomrot_prev"], g["zoomrot_cur"]
    pipe = D.Pipeline(1, prev.shape[0], prev.shape[1])
    pipe.run(D.Planes.from_host(prev), D.Planes.from_host(cur), procedure=sp, window=sw)
    np.testing.assert_array_equal(pipe.intermediate(1)[0].cpu().numpy(), g[f"ovr_sp{sp}_sw{sw}_gt1"])
    np.testing.assert_array_equal(pipe.intermediate(2)[0].cpu().numpy(), g[f"ovr_sp{sp}_sw{sw}_gt2"])
    np.testing.assert_array_equal(pipe.intermediate(4)[0].cpu().numpy().astype(bool), g[f"ovr_sp{sp}_sw{sw}_outlier2"])
    np.testing.assert_allclose(pipe.params[0].cpu().numpy(), g[f"ovr_sp{sp}_sw{sw}_params"], **PARAM_TOL)


@pytest.mark.parametrize("H,W,kind,sp,sw", [
    (480, 720, "pan", 3, 2),               # config 3
    (1080, 1920, "zoomrot", 1, 16),        # config 4, three-step
    (1080, 1920, "zoomrot", 2, 16),        # config 4, 2D-log
    (1080, 1920, "zoomrot", 3, 2),
    (540, 960, "affine", 0, 8),            # exhaustive on the bs=16 levels (config 5 shape, reduced size/window)
])
def test_pipeline_sequence_vs_oracle(D, H, W, kind, sp, sw):
    """A device-resident sequence at frame distance 3 (pairs are views of one buffer) against the oracle."""
    n_frames, d = 6, 3
    seq = {"pan": lambda: S.pan_sequence(n_frames, H, W, (2, 1), seed=3),
           "zoomrot": lambda: S.zoom_rotate_sequence(n_frames, H, W, seed=4),
           "affine": lambda: S.affine_sequence(n_frames, H, W, seed=5)}[kind]()
    pipe = D.gme_sequence(D.Planes.from_host(seq), d, procedure=sp, window=sw)
    torch.cuda.synchronize()
    params, comp, psnr = pipe.params.cpu().numpy(), pipe.comp.to_host(), pipe.psnr()
    dense, f1, f2 = (pipe.intermediate(i).cpu().numpy() for i in (0, 1, 2))
    o1, o2 = (pipe.intermediate(i).cpu().numpy().astype(bool) for i in (3, 4))
    for k in range(n_frames - d):
        want, inter = O.global_motion_estimation(seq[k], seq[k + d], procedure=sp, window=sw,
                                                 return_intermediates=True, threads=8)
        np.testing.assert_array_equal(dense[k], inter[0]["dense"])
        np.testing.assert_array_equal(f1[k], inter[1]["gt"])
        np.testing.assert_array_equal(f2[k], inter[2]["gt"])
        np.testing.assert_array_equal(o1[k], inter[1]["outlier"])
        np.testing.assert_array_equal(o2[k], inter[2]["outlier"])
        np.testing.assert_allclose(params[k], want, **PARAM_TOL)
        model = O.get_motion_field_affine((H // 16, W // 16), want)
        want_comp = O.compensate_frame(seq[k], model)
        np.testing.assert_array_equal(comp[k], want_comp)
        wp = O.PSNR(seq[k + d], want_comp)
        assert abs(psnr[k].real - wp.real) <= 1e-12 * abs(wp.real)


def test_pipeline_properties_4k(D):
    """Full-size (2160x3840) size-independent properties: an exact 16k-pixel pan is recovered exactly by the
    exhaustive search away from the borders, identical frames give PSNR -1 ... and graph replay == eager."""
    H, W = 2160, 3840
    base = S.texture(H + 32, W + 64, seed=12)
    prev = np.ascontiguousarray(base[16:16 + H, 32:32 + W])
    cur = np.ascontiguousarray(base[16 - 7:16 - 7 + H, 32 + 9:32 + 9 + W])        # content moves (-9, +7)
    pp, cp = D.Planes.from_host(prev), D.Planes.from_host(cur)
    f = D.motion_field(pp, cp, 16, 32, 0, 0)[0].cpu().numpy()                    # config-5 search: exhaustive, sw=32
    inner = f[3:-3, 3:-3]
    assert (inner[..., 0] == -9).all() and (inner[..., 1] == 7).all()
    fm = D.motion_field(pp, cp, 16, 32, 0, 1)[0].cpu().numpy()
    np.testing.assert_array_equal(fm[3:-3, 3:-3], inner)
    # exhaustive field on a stripe of the frame equals the oracle there (full 4K on one CPU thread takes minutes)
    sl = slice(0, 16 * 6)
    want = O.get_motion_field(prev[sl, :640], cur[sl, :640], 16, 32, 0, 0, threads=8)
    got = D.motion_field(D.Planes.from_host(prev[sl, :640].copy()), D.Planes.from_host(cur[sl, :640].copy()),
                         16, 32, 0, 0)[0].cpu().numpy()
    np.testing.assert_array_equal(got, want)
    pipe = D.Pipeline(1, H, W)
    pipe.run(pp, cp)
    eager = pipe.params.clone()
    assert int(pipe.status.item()) == 0
    pv = pipe.params[0].cpu().numpy()
    # diamond search gets trapped on this texture (a0 = -6.08, not -9): the bar is the oracle, not the true motion
    np.testing.assert_allclose(pv, O.global_motion_estimation(prev, cur, threads=8), **PARAM_TOL)
    assert np.abs(pv[[1, 2, 4, 5]]).max() < 1e-2
    pipe.capture(pp, cp)
    pipe.params.zero_()
    pipe.replay()
    torch.cuda.synchronize()
    assert torch.equal(pipe.params, eager)
    same = D.Pipeline(1, H, W)
    same.run(pp, pp)
    assert same.psnr()[0] == -1 or same.psnr()[0].real > 40      # static content: the -1 edge vectors cost little


def test_pipeline_config5_full_size(D):
    """Config 5 AT SIZE: one 2160x3840 general-affine pair (frame distance 3) through gme_pipeline with exhaustive
    search, sw = 32, on the bs-16 levels, against the threaded oracle: dense field, both exhaustive fields
    (64.7 G pixel-pair operations), both outlier masks, parameters, compensated frame and squared error."""
    H, W, d = 2160, 3840, 3
    seq = S.affine_sequence(d + 1, H, W, seed=5)
    prev, cur = seq[0], seq[d]
    pipe = D.Pipeline(1, H, W)
    pipe.run(D.Planes.from_host(prev), D.Planes.from_host(cur), procedure=0, window=32)
    torch.cuda.synchronize()
    assert int(pipe.status.item()) == 0
    want, inter = O.global_motion_estimation(prev, cur, procedure=0, window=32, return_intermediates=True,
                                             threads=len(os.sched_getaffinity(0)))
    np.testing.assert_array_equal(pipe.intermediate(0)[0].cpu().numpy(), inter[0]["dense"])
    np.testing.assert_array_equal(pipe.intermediate(1)[0].cpu().numpy(), inter[1]["gt"])
    np.testing.assert_array_equal(pipe.intermediate(2)[0].cpu().numpy(), inter[2]["gt"])
    np.testing.assert_array_equal(pipe.intermediate(3)[0].cpu().numpy().astype(bool), inter[1]["outlier"])
    np.testing.assert_array_equal(pipe.intermediate(4)[0].cpu().numpy().astype(bool), inter[2]["outlier"])
    np.testing.assert_allclose(pipe.params[0].cpu().numpy(), want, **PARAM_TOL)
    comp = O.compensate_frame(prev, O.get_motion_field_affine((H // 16, W // 16), want))
    np.testing.assert_array_equal(pipe.comp.to_host()[0], comp)
    assert int(pipe.sse.item()) == O.sse(cur, comp)


# ------------------------------------------------------------------ sequence mode, host runner, stage timing
def test_sequence_aliasing_equals_separate_buffers(D):
    """gme_pipeline builds each frame's pyramid once when prev/cur are two views of one sequence buffer; the
    results must equal the run on two unrelated buffers (pyramids built per pair, as the reference does)."""
    seq = S.zoom_rotate_sequence(7, 272, 400, zoom_per_frame=0.004, deg_per_frame=0.3, seed=31)
    d, n = 2, 5
    planes = D.Planes.from_host(seq)
    a = D.Pipeline(n, 272, 400)
    a.run(planes.view(0, n), planes.view(d, d + n))
    prev, cur = D.Planes.from_host(seq[:n].copy()), D.Planes.from_host(seq[d:d + n].copy())
    b = D.Pipeline(n, 272, 400)
    b.run(prev, cur)
    torch.cuda.synchronize()
    assert torch.equal(a.params, b.params) and torch.equal(a.sse, b.sse)
    assert torch.equal(a.comp.pixels(), b.comp.pixels())
    for which in range(6):
        assert torch.equal(a.intermediate(which), b.intermediate(which)), which
    want = O.global_motion_estimation(seq[1], seq[1 + d])
    np.testing.assert_allclose(a.params[1].cpu().numpy(), want, **PARAM_TOL)


@pytest.mark.parametrize("lanes", [2, 3, 16])
def test_pipeline_lanes_give_the_same_bits(D, lanes):
    """Pipeline(lanes=k) runs the batch as k sub-batches on forked streams (own workspace each): every output and every
    intermediate equals the single-call run, eagerly and as a captured CUDA graph replayed on refilled inputs; a
    ragged split (7 pairs over 3 lanes) and more lanes than pairs are covered."""
    seq = S.zoom_rotate_sequence(9, 272, 400, zoom_per_frame=0.004, deg_per_frame=0.3, seed=33)
    d, n = 2, 7
    planes = D.Planes.from_host(seq)
    one = D.Pipeline(n, 272, 400)
    one.run(planes.view(0, n), planes.view(d, d + n))
    many = D.Pipeline(n, 272, 400, lanes=lanes)
    assert many.lanes == min(lanes, n)
    many.run(planes.view(0, n), planes.view(d, d + n))
    torch.cuda.synchronize()
    assert torch.equal(one.rows, many.rows) and torch.equal(one.status, many.status)
    assert torch.equal(one.comp.pixels(), many.comp.pixels())
    for which in range(6):
        assert torch.equal(one.intermediate(which), many.intermediate(which)), which
    # captured: refill the same storage with other frames, replay, compare with a fresh eager run
    many.capture(planes.view(0, n), planes.view(d, d + n))
    other = S.pan_sequence(9, 272, 400, step=(1, -2), seed=5)
    planes.pixels().copy_(torch.from_numpy(other).to(planes.t.device))
    many.replay()
    one.run(planes.view(0, n), planes.view(d, d + n))
    torch.cuda.synchronize()
    assert torch.equal(one.rows, many.rows) and torch.equal(one.comp.pixels(), many.comp.pixels())
    want = O.global_motion_estimation(other[n - 1], other[n - 1 + d])
    np.testing.assert_allclose(many.params[n - 1].cpu().numpy(), want, **PARAM_TOL)


def test_host_sequence_runner_matches_pipeline(D):
    """The overlapped host path (chunked uploads on a copy stream, two alternating device buffers) returns the
    rows the resident pipeline computes, call after call, including a ragged last chunk."""
    seq = S.pan_sequence(14, 160, 208, step=(2, 1), seed=9)
    d = 3
    n = seq.shape[0] - d
    planes = D.Planes.from_host(seq)
    pipe = D.Pipeline(n, 160, 208)
    pipe.run(planes.view(0, n), planes.view(d, d + n))
    runner = D.HostSequenceRunner(seq.shape[0], 160, 208, d, chunk=4)
    host = torch.from_numpy(seq).pin_memory()
    for _ in range(3):                                   # alternates the two device buffers
        rows = runner.run(host)
        torch.cuda.synchronize()
        got = rows.numpy()
        np.testing.assert_array_equal(got[:, :6], pipe.params.cpu().numpy())
        np.testing.assert_array_equal(got[:, 6].astype(np.int64), pipe.sse.cpu().numpy())
        assert (got[:, 7] == 0).all()
    assert runner.h2d_bytes == seq.size and runner.d2h_bytes == n * 8 * 8


def test_stage_timing_api(D):
    import gme_native as N
    seq = S.pan_sequence(6, 96, 160, seed=2)
    planes = D.Planes.from_host(seq)
    pipe = D.Pipeline(4, 96, 160)
    N.stage_timing_enable(True)
    for _ in range(3):
        pipe.run(planes.view(0, 4), planes.view(2, 6))
    ms, calls = N.stage_timing_read()
    N.stage_timing_enable(False)
    assert calls == 3 and len(ms) == N.PIPELINE_STAGES and all(m > 0 for m in ms)
    pipe.run(planes.view(0, 4), planes.view(2, 6))
    assert N.stage_timing_read()[1] == 0                 # disabled: nothing recorded


@pytest.mark.parametrize("H,W", [(17, 17), (18, 40), (33, 33), (48, 400), (130, 150), (64, 1000)])
def test_diamond16_small_and_border_frames(D, H, W):
    """The register-neighbourhood diamond kernel: frames barely larger than a block (every centre is clamped),
    frames narrower than a tile, and the H % 16 == 0 last-row quirk -- all against the oracle, both norms."""
    rng = np.random.default_rng(H * 1000 + W)
    base = S.texture(H + 12, W + 12, seed=H + W)
    prev = np.ascontiguousarray(base[6:6 + H, 6:6 + W])
    dy, dx = (int(v) for v in rng.integers(-5, 6, 2))
    cur = np.ascontiguousarray(base[6 - dy:6 - dy + H, 6 - dx:6 - dx + W])
    for pn in (0, 1):
        want = O.get_motion_field(prev, cur, 16, 0, 3, pn)
        np.testing.assert_array_equal(field_of(D, prev, cur, 16, 0, 3, pn)[0], want)
    np.testing.assert_array_equal(field_of(D, prev, prev, 16, 0, 3, 1)[0], O.get_motion_field(prev, prev, 16, 0, 3, 1))


def test_sad_probe_counts(D):
    import ctypes
    import gme_native as N
    scratch = torch.zeros(1024, dtype=torch.int32, device="cuda")
    n = ctypes.c_uint64(0)
    N.check(N.lib.gme_sad_peak_probe(1, 148, 16, scratch.data_ptr(), ctypes.byref(n), None))
    torch.cuda.synchronize()
    assert n.value == 148 * 256 * 16 * 4 * 8 * 4


def test_unaligned_pitch_takes_the_cooperative_paths(D):
    """Planes whose pitch is a multiple of 4 but not of 16 cannot travel by TMA / 128-bit accesses: every kernel
    then takes its cooperative-load / per-pixel path, and the results must not change."""
    H, W, n, d = 100, 150, 3, 2
    seq = S.pan_sequence(n + d, H, W, step=(3, -1), seed=14)
    odd = torch.zeros((n + d, H, 152), dtype=torch.uint8, device="cuda")           # pitch 152: 8 (mod 16)
    odd[:, :, :W].copy_(torch.from_numpy(seq))
    planes = D.Planes(odd, W)
    assert planes.pitch == 152
    for bs, sw, sp, pn in ((16, 0, 3, 1), (2, 0, 3, 1), (16, 6, 0, 0), (12, 4, 0, 1), (8, 5, 1, 1), (16, 8, 2, 0)):
        got = D.motion_field(planes.view(0, n), planes.view(d, d + n), bs, sw, sp, pn).cpu().numpy()
        for k in range(n):
            np.testing.assert_array_equal(got[k], O.get_motion_field(seq[k], seq[k + d], bs, sw, sp, pn),
                                          err_msg=f"bs={bs} sw={sw} sp={sp} pn={pn} pair={k}")
    np.testing.assert_array_equal(D.pyr_down(planes).to_host()[1], O.pyr_down(seq[1]))
    pipe = D.Pipeline(n, H, W)
    comp_odd = D.Planes(torch.zeros((n, H, 152), dtype=torch.uint8, device="cuda"), W)
    pipe.comp = comp_odd                                                            # unaligned output planes too
    pipe.run(planes.view(0, n), planes.view(d, d + n))
    torch.cuda.synchronize()
    for k in range(n):
        want = O.global_motion_estimation(seq[k], seq[k + d])
        np.testing.assert_allclose(pipe.params[k].cpu().numpy(), want, **PARAM_TOL)
        comp = O.compensate_frame(seq[k], O.get_motion_field_affine((H // 16, W // 16), want))
        np.testing.assert_array_equal(pipe.comp.to_host()[k], comp)
        assert int(pipe.sse[k].item()) == O.sse(seq[k + d], comp)


# ------------------------------------------------------------------ the rows next to the hot path
def test_results_batch_matches_the_results_loop(D):
    """gme_device.results_batch = one iteration of results.py's loop per pair (results.py:47-59,78-83,109), with the
    difference images fused into the compensation kernel: everything against the oracle / the NumPy formulas."""
    H, W, d = 272, 400, 3                       # 272 = 17 * 16: the last block row exists; W % 16 == 0: fused kernel
    seq = S.zoom_rotate_sequence(6, H, W, zoom_per_frame=0.004, deg_per_frame=0.3, seed=21)
    out = D.results_batch(D.Planes.from_host(seq), d)
    torch.cuda.synchronize()
    comp, dp, dc = out["compensated"].to_host(), out["diff_curr_prev"].to_host(), out["diff_curr_comp"].to_host()
    for k in range(seq.shape[0] - d):
        prev, cur = seq[k], seq[k + d]
        want = O.global_motion_estimation(prev, cur)
        np.testing.assert_allclose(out["params"][k].cpu().numpy(), want, **PARAM_TOL)
        model = O.get_motion_field_affine((H // 16, W // 16), want)
        np.testing.assert_array_equal(out["model_field"][k].cpu().numpy(), model)
        wc = O.compensate_frame(prev, model)
        np.testing.assert_array_equal(comp[k], wc)
        np.testing.assert_array_equal(dp[k], np.absolute(cur.astype("int") - prev.astype("int")).astype("uint8"))
        np.testing.assert_array_equal(dc[k], np.absolute(cur.astype("int") - wc.astype("int")).astype("uint8"))
        assert int(out["sse"][k].item()) == O.sse(cur, wc)
    # a geometry the fused kernel does not take (W % 16 != 0): separate difference kernels, same images
    seq2 = S.pan_sequence(5, 100, 150, step=(2, 1), seed=5)
    out2 = D.results_batch(D.Planes.from_host(seq2), 2)
    wc = O.compensate_frame(seq2[0], out2["model_field"][0].cpu().numpy())
    np.testing.assert_array_equal(out2["compensated"].to_host()[0], wc)
    np.testing.assert_array_equal(out2["diff_curr_comp"].to_host()[0],
                                  np.absolute(seq2[2].astype("int") - wc.astype("int")).astype("uint8"))
    np.testing.assert_array_equal(out2["diff_curr_prev"].to_host()[0],
                                  np.absolute(seq2[2].astype("int") - seq2[0].astype("int")).astype("uint8"))


@pytest.mark.parametrize("H,W,bs,sw,sp", [(240, 320, 12, 8, 1), (480, 720, 16, 4, 3), (250, 322, 10, 6, 0), (96, 128, 8, 5, 2)])
def test_hierarchical_field_on_device(D, H, W, bs, sw, sp):
    """bbme.hierarchical_wrapper with pyramids, searches and merges all on the device, batched, against the oracle."""
    seq = S.zoom_rotate_sequence(4, H, W, zoom_per_frame=0.006, deg_per_frame=0.4, seed=H + bs)
    planes = D.Planes.from_host(seq)
    try:
        want = [O.hierarchical_wrapper(seq[k], seq[k + 1], bs, sw, sp) for k in range(3)]
    except ValueError:
        with pytest.raises(ValueError):
            D.hierarchical_field(planes.view(0, 3), planes.view(1, 4), bs, sw, sp)
        return
    got = D.hierarchical_field(planes.view(0, 3), planes.view(1, 4), bs, sw, sp).cpu().numpy()
    assert got.dtype == np.float64
    for k in range(3):
        np.testing.assert_array_equal(got[k], want[k])
