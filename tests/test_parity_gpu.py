"""Parity of the CUDA path (libsurfb200.so, called through the C-ABI) against
  (1) the CPU oracle (oracle/liboracle.so),
  (2) the unmodified reference built for sm_100a (oracle/_ref/libsurfref.so) when it is present,
  (3) the committed golden vectors of the reference (tests/golden/*.npz).
Bars (BASELINE.json north_star): integral bit-exact; Hessian maps within 1e-5 relative (bit-exact in
practice); >= 99 % of keypoints within 0.1 px and 0.05 in scale, disagreements listed; descriptors
within 1e-3 L2; match indices equal.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
import ref_lib
from helpers import describe_misses, keypoint_parity, load_golden, load_pair, parity_entry

pytestmark = pytest.mark.gpu

# Descriptor bar of BASELINE.json: 1e-3 L2, asserted on the MAXIMUM over all compared rows.
TOL = 1e-3
# Rotated path: the orientation feeds sin/cos of every sample position, so a 1-ulp difference in `ori` moves samples
# across cell / window boundaries. The reference's own orientation is not reproducible run to run (shared-memory float
# atomics, surfd.cu:1795-1805): the comparisons with the reference measure its run-to-run spread on the
# same frame in the test itself (rotated_bar) and use max(1e-3, that spread); against the deterministic oracle the bar is 1e-3.
TOL_ROT = 1e-3

REF = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref/libsurfref.so not built")


def _torch():
    import torch
    return torch


def _sb():
    import cuda_surf_b200 as sb
    return sb


def make_det(w, h, noctaves=4, thresh=4.0, upright=True, extend=False, max_pts=32768, batch=1, doubled=False):
    sb = _sb()
    det = sb.Surfor()
    det.init(noctaves, thresh, doubled, 9, 2, upright, extend, 4, w, h, max_pts=max_pts, batch=batch)
    return det


def upload(img, pitch=None):
    torch = _torch()
    sb = _sb()
    h, w = img.shape
    pitch = pitch or sb.iAlignUp(w, 128)
    buf = np.zeros((h, pitch), np.uint8)
    buf[:, :w] = img
    return torch.from_numpy(buf).cuda(), (w, h, pitch)


def run_detect(det, img, max_pts=32768, desc=True, pitch=None):
    sb = _sb()
    d_img, whp = upload(img, pitch)
    data = sb.initSurfData(max_pts, True, True)
    dd = det.detectAndCompute(d_img, data, whp, desc=desc)
    pts = data.host_points()
    de = dd[: data.num_pts].cpu().numpy() if desc else None
    return data, pts, de



def rotated_bar(ref, img, rpts, rdesc, pts, desc, max_pts, nruns=2):
    """The rotated path of the reference is not reproducible run to run (order-dependent shared-memory float atomics in
    its orientation histogram and descriptor sums, surfd.cu:1795-1805, 1222-1266): an orientation that lands one ulp
    elsewhere moves samples across cell boundaries. Runs the reference `nruns` more times on the same frame and returns
    (spread, l2_best): the largest descriptor distance between two of ITS OWN runs, and per keypoint of the first run the
    distance from this library's descriptor to the nearest of the reference's runs. The bar of the rotated comparisons
    is max(1e-3, spread), as recorded in the parity report."""
    _, _, ok0, idx0, _, _ = keypoint_parity(rpts, pts)
    best = np.full(len(rpts), np.inf)
    best[ok0] = np.linalg.norm(desc[idx0[ok0]] - rdesc[ok0], axis=1)
    spread = 0.0
    for _ in range(nruns):
        p2, d2 = ref.detect(img, max_pts=max_pts)
        _, _, okr, idxr, _, _ = keypoint_parity(rpts, p2)
        if okr.any():
            spread = max(spread, float(np.linalg.norm(d2[idxr[okr]] - rdesc[okr], axis=1).max()))
        _, _, ok2, idx2, _, _ = keypoint_parity(p2, pts)  # this library against that run, mapped back to the first run's rows
        l2 = np.full(len(p2), np.inf)
        l2[ok2] = np.linalg.norm(desc[idx2[ok2]] - d2[ok2], axis=1)
        both = okr.copy()
        both[okr] = ok2[idxr[okr]]
        best[both] = np.minimum(best[both], l2[idxr[both]])
    return spread, best[ok0]



def assert_ambiguity(got, want, what):
    """ambiguity = second / best, exact except where a group's second and third candidate are closer than the ranking
    resolution of the tensor-core pass (3e-5: split-bf16 product, packed keys): the exact re-score then sees the third and
    the ratio moves by < 1e-4. At most two such rows per match are tolerated (offline estimate: one row in ~10 matches of
    5 k x 5 k descriptors); everything else must agree to 1e-6."""
    d = np.abs(got["ambiguity"] - want["ambiguity"])
    near = np.nonzero(d > 1e-6)[0]
    assert len(near) <= 2 and (len(d) == 0 or d.max() <= 1e-4), f"{what}: ambiguity differs in rows {near[:10]} by up to {d.max():.3e}"
    return near, d


def assert_resp_close(got, want, what):
    assert got.shape == want.shape, what
    if np.array_equal(got, want):
        return
    denom = np.maximum(np.abs(want), 1e-3)
    rel = np.abs(got - want) / denom
    assert rel.max() <= 1e-5, f"{what}: max rel err {rel.max():.3e} at {np.unravel_index(rel.argmax(), rel.shape)}"


CASES = [  # (w, h, seed, noctaves)  -- ragged sizes on purpose
    (320, 240, 3, 3),
    (333, 251, 4, 3),
    (640, 480, 5000, 4),
    (1281, 723, 11, 4),
    (257, 131, 12, 2),
]


@pytest.mark.parametrize("w,h,seed,noct", CASES)
def test_integral_and_hessian_vs_oracle(w, h, seed, noct):
    sb = _sb()
    img = sb.synth_frame(w, h, seed)
    det = make_det(w, h, noct)
    run_detect(det, img, desc=False)
    orc = ol.Oracle(noct, 4.0, False, 9, 2, True, False, 4)
    I = orc.integral(img)
    got_I = det.get_integral()
    assert np.array_equal(got_I, I), f"integral differs at {np.argwhere(got_I != I)[:5]}"
    want = orc.hessian(I)
    got = det.get_response()
    for o, (g, wnt) in enumerate(zip(det.split_response(got), orc.split_resp(want, w, h))):
        assert_resp_close(g, wnt, f"octave {o}")


@pytest.mark.parametrize("pitch_pad", [0, 3])
def test_integral_unaligned_pitch(pitch_pad):
    """pitch not a multiple of 8 takes the byte-load path of the integral kernels"""
    sb = _sb()
    w, h = 300, 200
    img = sb.synth_frame(w, h, 21)
    det = make_det(w, h, 2)
    run_detect(det, img, desc=False, pitch=w + pitch_pad)
    assert np.array_equal(det.get_integral(), ol.Oracle(2).integral(img))


@pytest.mark.parametrize("w,h,seed,noct", CASES[:4])
@pytest.mark.parametrize("upright", [True, False])
def test_keypoints_and_descriptors_vs_oracle(w, h, seed, noct, upright, report):
    sb = _sb()
    img = sb.synth_frame(w, h, seed)
    det = make_det(w, h, noct, upright=upright)
    data, pts, desc = run_detect(det, img)
    orc = ol.Oracle(noct, 4.0, False, 9, 2, upright, False, 4)
    opts, odesc = orc.detect_and_compute(img)
    fr, fg, ok, idx, miss_r, miss_g = keypoint_parity(opts, pts)
    assert abs(len(pts) - len(opts)) <= max(2, len(opts) // 100), (len(pts), len(opts))
    assert fr >= 0.99 and fg >= 0.99, f"matched {fr:.4f}/{fg:.4f}\nmissing:\n{describe_misses(miss_r, 4.0)}"
    assert np.array_equal(pts["laplace"][idx[ok]], opts["laplace"][ok])
    # descriptors of matched keypoints (positions agree to float round-off, so descriptors do too)
    exact = ok & (np.abs(pts["x"][idx] - opts["x"]) < 1e-4) & (np.abs(pts["y"][idx] - opts["y"]) < 1e-4) & \
        (np.abs(pts["scale"][idx] - opts["scale"]) < 1e-5)
    dori = None
    if not upright:
        dori = np.abs(np.angle(np.exp(1j * (pts["ori"][idx] - opts["ori"]))))
        assert np.median(dori[ok]) < 1e-4
    # every matched keypoint is compared; the ones whose position differs by more than round-off are LISTED
    l2_all = np.linalg.norm(desc[idx[ok]] - odesc[ok], axis=1)
    report(**parity_entry(f"{w}x{h} seed {seed} upright={upright} vs oracle", opts, pts, ok, idx, miss_r, miss_g, l2_all,
                          extra={"position_differs_over_1e-4": int((ok & ~exact).sum()),
                                 "ori_diff_max": None if dori is None else float(dori[ok].max())}))
    assert (ok & ~exact).sum() <= max(1, ok.sum() // 100), f"{(ok & ~exact).sum()} matched keypoints differ in position by > 1e-4"
    l2 = np.linalg.norm(desc[idx[exact]] - odesc[exact], axis=1)
    assert l2.max() <= (TOL if upright else TOL_ROT), f"descriptor L2 max {l2.max():.3e}"


def test_describe_given_points_matches_oracle_exact_inputs():
    """same keypoints in, descriptors out: isolates the descriptor kernel from detection"""
    torch = _torch()
    sb = _sb()
    w, h = 640, 480
    img = sb.synth_frame(w, h, 5000)
    for upright, extend in [(True, False), (False, False), (True, True)]:
        det = make_det(w, h, 4, upright=upright, extend=extend)
        run_detect(det, img, desc=False)
        orc = ol.Oracle(4, 4.0, False, 9, 2, upright, extend, 4)
        I = orc.integral(img)
        opts = orc.keypoints(I, orc.hessian(I))
        if not upright:
            opts = orc.orientation(I, opts)
        want = orc.describe(I, opts)
        d_pts = torch.from_numpy(opts.view(np.uint8).copy()).cuda()
        d_desc = torch.zeros((len(opts), det.nfeatures), dtype=torch.float32, device="cuda")
        det.describe(d_pts, len(opts), d_desc)
        got = d_desc.cpu().numpy()
        if not upright:
            gori = d_pts.cpu().numpy().view(sb.POINT_DTYPE)["ori"]
            dori = np.abs(np.angle(np.exp(1j * (gori - opts["ori"]))))
            assert np.nanmax(dori) < 1e-3, f"orientation differs by {np.nanmax(dori)}"
        l2 = np.linalg.norm(got - want, axis=1)
        bad = np.isnan(l2)
        assert not bad.any()
        # rotated descriptors inherit the orientation's round-off through sin/cos
        assert l2.max() <= (TOL if upright else TOL_ROT), f"upright={upright} extend={extend}: L2 max {l2.max():.3e}"


@REF
@pytest.mark.parametrize("which", ["left", "right"])
@pytest.mark.parametrize("upright", [True, False])
def test_bundled_pair_vs_reference(which, upright, report):
    """config 1 of BASELINE.md: main.cpp defaults on the reference's own stereo pair (and once with upright=false)"""
    left, right = load_pair()
    img = left if which == "left" else right
    h, w = img.shape
    ref = ref_lib.Reference(w, h, 4, 4.0, False, 9, 2, upright, False, 4)
    ref.detect(img, max_pts=32768)  # the reference's first call reads uninitialised scratch
    rpts, rdesc = ref.detect(img, max_pts=32768)
    rI, rlayers, _ = ref.stages(img)
    det = make_det(w, h, 4, upright=upright)
    data, pts, desc = run_detect(det, img)
    assert np.array_equal(det.get_integral(), rI)
    for o, (g, wnt) in enumerate(zip(det.split_response(det.get_response()), rlayers)):
        assert_resp_close(g, wnt, f"octave {o}")
    fr, fg, ok, idx, miss_r, miss_g = keypoint_parity(rpts, pts)
    assert fr >= 0.99 and fg >= 0.99, f"{fr:.4f}/{fg:.4f}\nreference-only:\n{describe_misses(miss_r, 4.0)}\nours-only:\n{describe_misses(miss_g, 4.0)}"
    assert np.array_equal(pts["laplace"][idx[ok]], rpts["laplace"][ok])
    l2 = np.linalg.norm(desc[idx[ok]] - rdesc[ok], axis=1)
    spread, bar = 0.0, TOL
    if not upright:
        spread, l2 = rotated_bar(ref, img, rpts, rdesc, pts, desc, 32768)
        bar = max(TOL_ROT, spread)
    ref.close()
    report(**parity_entry(f"bundled {which} upright={upright} vs reference", rpts, pts, ok, idx, miss_r, miss_g, l2,
                          extra={"reference_run_to_run_desc_l2_max": spread, "bar": bar}))
    assert len(rpts) == (2739 if which == "left" else 3443)
    assert l2.max() <= bar, f"descriptor L2: max {l2.max():.3e}, rows over: {(l2 > TOL).sum()}, reference's own spread {spread:.3e}"


@REF
@pytest.mark.parametrize("upright,extend", [(True, False), (False, False), (True, True)])
def test_describe_vs_reference_same_points(upright, extend):
    torch = _torch()
    sb = _sb()
    w, h = 640, 480
    img = sb.synth_frame(w, h, 5000)
    ref = ref_lib.Reference(w, h, 4, 4.0, False, 9, 2, upright, extend, 4)
    rpts, rdesc = ref.detect(img)
    ref.close()
    det = make_det(w, h, 4, upright=upright, extend=extend)
    run_detect(det, img, desc=False)
    d_pts = torch.from_numpy(rpts.view(np.uint8).copy()).cuda()
    d_desc = torch.zeros((len(rpts), det.nfeatures), dtype=torch.float32, device="cuda")
    det.describe(d_pts, len(rpts), d_desc)
    got = d_desc.cpu().numpy()
    if not upright:
        gori = d_pts.cpu().numpy().view(sb.POINT_DTYPE)["ori"]
        dori = np.abs(np.angle(np.exp(1j * (gori - rpts["ori"]))))
        assert np.nanmax(dori) < 1e-3, f"orientation differs by {np.nanmax(dori)}"
    l2 = np.linalg.norm(got - rdesc, axis=1)
    assert np.isfinite(l2).all()
    assert l2.max() <= (TOL if upright else TOL_ROT), f"L2 max {l2.max():.3e}"


@REF
def test_match_vs_reference():
    sb = _sb()
    torch = _torch()
    w, h = 640, 480
    sl = sb.synth_frame(w, h, 5000)
    sr = sb.synth_frame(w, h, 5000, 12, 2, 5000 ^ 0xA5A5)
    det = make_det(w, h, 4)
    d1, p1, f1 = run_detect(det, sl)
    d1_desc = torch.from_numpy(f1).cuda()
    d2, p2, f2 = run_detect(det, sr)
    d2_desc = torch.from_numpy(f2).cuda()
    det.match(d1, d2, d1_desc, d2_desc)
    got = d1.host_points()
    ref = ref_lib.Reference(w, h, 4)
    want = ref.match(p1, f1, p2, f2)
    ref.close()
    assert np.array_equal(got["match"], want["match"])
    assert np.array_equal(got["score"], want["score"])
    assert_ambiguity(got, want, "bundled pair")
    assert np.array_equal(got["match_x"], want["match_x"]) and np.array_equal(got["match_y"], want["match_y"])
    # host copy of the five match fields (surf.cpp:421-425)
    assert np.array_equal(d1.h_data["match"][: d1.num_pts], got["match"])


def test_match_vs_oracle_and_tail_rule():
    sb = _sb()
    torch = _torch()
    rng = np.random.default_rng(5)
    for n1, n2 in [(100, 95), (33, 64), (257, 1000), (5, 31)]:
        f1 = rng.standard_normal((n1, 64)).astype(np.float32)
        f2 = rng.standard_normal((n2, 64)).astype(np.float32)
        f1 /= np.linalg.norm(f1, axis=1, keepdims=True)
        f2 /= np.linalg.norm(f2, axis=1, keepdims=True)
        p1 = np.zeros(n1, ol.POINT_DTYPE)
        p2 = np.zeros(n2, ol.POINT_DTYPE)
        p2["x"] = rng.random(n2).astype(np.float32)
        p2["y"] = rng.random(n2).astype(np.float32)
        want = ol.match(p1, f1, p2, f2)
        det = make_det(64, 64, 1, max_pts=2048)
        a = sb.initSurfData(2048)
        b = sb.initSurfData(2048)
        a.num_pts, b.num_pts = n1, n2
        a.d_data[: n1 * 48] = torch.from_numpy(p1.view(np.uint8)).cuda()
        b.d_data[: n2 * 48] = torch.from_numpy(p2.view(np.uint8)).cuda()
        det.match(a, b, torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda())
        got = a.host_points()
        assert np.array_equal(got["match"], want["match"]), (n1, n2)
        assert (got["match"] < n2 - n2 % 32).all()  # the last n2 % 32 descriptors are never candidates
        assert np.array_equal(got["score"], want["score"])
        assert_ambiguity(got, want, f"{n1} x {n2}")


@pytest.mark.parametrize("nf,n1,n2", [(64, 2739, 3443), (128, 700, 1500), (64, 128, 256), (128, 1, 40)])
def test_match_large_and_surf128_vs_oracle(nf, n1, n2):
    """several row blocks, column tiles and splits of the tensor-core matcher; 128-d uses the 128-column tile"""
    sb = _sb()
    torch = _torch()
    rng = np.random.default_rng(nf + n1)
    # correlated descriptors (a few shared directions + noise) so that best / second are close, as on real data
    basis = rng.standard_normal((24, nf)).astype(np.float32)
    f1 = (rng.random((n1, 24)).astype(np.float32) @ basis + 0.3 * rng.standard_normal((n1, nf)).astype(np.float32))
    f2 = (rng.random((n2, 24)).astype(np.float32) @ basis + 0.3 * rng.standard_normal((n2, nf)).astype(np.float32))
    f1 = np.abs(f1) / np.linalg.norm(f1, axis=1, keepdims=True)
    f2 = np.abs(f2) / np.linalg.norm(f2, axis=1, keepdims=True)
    p1 = np.zeros(n1, ol.POINT_DTYPE)
    p2 = np.zeros(n2, ol.POINT_DTYPE)
    p2["x"] = rng.random(n2).astype(np.float32) * 100
    p2["y"] = rng.random(n2).astype(np.float32) * 100
    want = ol.match(p1, f1, p2, f2)
    det = make_det(64, 64, 1, extend=(nf == 128), max_pts=4096)
    a = sb.initSurfData(4096)
    b = sb.initSurfData(4096)
    a.num_pts, b.num_pts = n1, n2
    a.d_data[: n1 * 48] = torch.from_numpy(p1.view(np.uint8)).cuda()
    b.d_data[: n2 * 48] = torch.from_numpy(p2.view(np.uint8)).cuda()
    det.match(a, b, torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda())
    got = a.host_points()
    same = got["match"] == want["match"]
    assert same.mean() >= 0.999, f"match index differs for {(~same).sum()} of {n1} rows"
    assert np.array_equal(got["score"][same], want["score"][same])
    assert np.array_equal(got["match_x"][same], want["match_x"][same])
    close = np.abs(got["ambiguity"] - want["ambiguity"]) <= 1e-4
    assert close.mean() >= 0.999, f"ambiguity differs for {(~close).sum()} rows, max {np.abs(got['ambiguity'] - want['ambiguity']).max()}"


def test_golden_pair_counts():
    """committed golden vectors of the reference (skips until they exist)"""
    g = load_golden("pair_left_upright")
    if g is None:
        pytest.skip("tests/golden/pair_left_upright.npz not generated yet")
    left, _ = load_pair()
    det = make_det(1280, 960, 4)
    data, pts, desc = run_detect(det, left)
    fr, fg, ok, idx, miss_r, miss_g = keypoint_parity(g["pts"], pts)
    assert fr >= 0.99 and fg >= 0.99


# ------------------------------------------------------------------ properties at full size

@pytest.mark.parametrize("w,h,seed", [(1920, 1080, 1), (3840, 2160, 2)])
def test_full_size_properties(w, h, seed):
    """BASELINE configs 2 and 3: checks that do not need the (slow) oracle beyond the integral"""
    sb = _sb()
    img = sb.synth_frame(w, h, seed)
    det = make_det(w, h, 5, max_pts=65536)
    data, pts, desc = run_detect(det, img, max_pts=65536)
    I = det.get_integral()
    assert I[-1, -1] == int(img.astype(np.int64).sum())
    assert np.array_equal(I, ol.Oracle(5).integral(img))
    assert not (I[0, :].any() or I[:, 0].any())
    # idempotence: the same frame again gives the same set (order may differ)
    data2, pts2, desc2 = run_detect(det, img, max_pts=65536)
    assert len(pts) == len(pts2)
    k1 = np.sort(pts, order=["x", "y", "scale"])
    k2 = np.sort(pts2, order=["x", "y", "scale"])
    assert np.array_equal(k1["x"], k2["x"]) and np.array_equal(k1["strength"], k2["strength"])
    # descriptors are unit vectors, keypoints inside the frame, strengths above threshold
    nrm = np.linalg.norm(desc, axis=1)
    assert np.allclose(nrm, 1.0, atol=1e-5)
    assert (pts["strength"] >= 4.0).all() and (pts["x"] > 0).all() and (pts["x"] < w).all() and (pts["y"] < h).all()
    assert 1.5 < len(pts) / (w * h / 1000.0) < 3.5, len(pts)  # synth_v1 density
    if w == 1920:
        opts, _ = ol.Oracle(5, 4.0, False, 9, 2, True, False, 4).detect_and_compute(img, desc=False)
        fr, fg, *_ = keypoint_parity(opts, pts)
        assert fr >= 0.99 and fg >= 0.99


def test_batch_equals_single_and_host_path():
    sb = _sb()
    torch = _torch()
    w, h, nb = 640, 480, 9  # 9 frames: the host path pipelines them as chunks of 4 + 4 + 1
    frames = np.stack([sb.synth_frame(w, h, 100 + i) for i in range(nb)])
    det = make_det(w, h, 4, max_pts=4096, batch=nb)
    pitch = sb.iAlignUp(w, 128)
    buf = np.zeros((nb, h, pitch), np.uint8)
    buf[:, :, :w] = frames
    d_imgs = torch.from_numpy(buf).cuda()
    pts = torch.zeros((nb, 4096 * 48), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(nb, dtype=torch.int32, device="cuda")
    desc = torch.zeros((nb, 4096, 64), dtype=torch.float32, device="cuda")
    det.detect_batch(d_imgs, pitch, pts, cnt, desc)
    torch.cuda.synchronize()
    counts = cnt.cpu().numpy()
    # host-buffer end-to-end entry point
    hp = np.zeros((nb, 4096), sb.POINT_DTYPE)
    hc = np.zeros(nb, np.int32)
    hd = np.zeros((nb, 4096, 64), np.float32)
    det.detect_batch_host(frames, hp, hc, hd)
    assert np.array_equal(hc, counts)
    # a second call reuses the streams and events; pinned buffers and a partial batch
    hp2 = torch.zeros((nb, 4096 * 48), dtype=torch.uint8).pin_memory()
    hc2 = torch.zeros(nb, dtype=torch.int32).pin_memory()
    det.detect_batch_host(torch.from_numpy(frames).pin_memory()[:3], hp2, hc2, None)
    assert np.array_equal(hc2.numpy()[:3], counts[:3])
    assert np.array_equal(np.sort(hp2[1].numpy().view(sb.POINT_DTYPE)[: counts[1]]["x"]), np.sort(hp[1, : counts[1]]["x"]))
    # several batches in flight (submit / wait): different sizes, results equal to the synchronous call
    fa, fb = np.ascontiguousarray(frames[:9]), np.ascontiguousarray(frames[2:7])
    ta = det.submit_batch_host(fa)
    tb = det.submit_batch_host(fb)
    td = det.submit_batch_host(fb)
    with pytest.raises(sb.SurfError):
        det.submit_batch_host(fa)  # a fourth outstanding batch is refused
    pa, ca, da = np.zeros((9, 4096), sb.POINT_DTYPE), np.zeros(9, np.int32), np.zeros((9, 4096, 64), np.float32)
    pb, cb, db = np.zeros((5, 4096), sb.POINT_DTYPE), np.zeros(5, np.int32), np.zeros((5, 4096, 64), np.float32)
    det.wait_batch_host(ta, pa, ca, da)
    tc = det.submit_batch_host(fa)  # the first set is free again
    det.wait_batch_host(tb, pb, cb, db)
    pc, cc, dc = np.zeros((9, 4096), sb.POINT_DTYPE), np.zeros(9, np.int32), np.zeros((9, 4096, 64), np.float32)
    det.wait_batch_host(td, pb.copy(), cb.copy(), None)
    det.wait_batch_host(tc, pc, cc, dc)
    assert np.array_equal(ca, counts) and np.array_equal(cb, counts[2:7]) and np.array_equal(cc, counts)
    for f in range(5):
        n = counts[2 + f]
        o1 = np.argsort(pb[f, :n], order=["x", "y", "scale"]); o2 = np.argsort(hp[2 + f, :n], order=["x", "y", "scale"])
        assert np.array_equal(pb[f, :n]["x"][o1], hp[2 + f, :n]["x"][o2]) and np.array_equal(db[f, :n][o1], hd[2 + f, :n][o2])
        o3 = np.argsort(pc[2 + f, :n], order=["x", "y", "scale"])
        assert np.array_equal(dc[2 + f, :n][o3], hd[2 + f, :n][o2])
    single = make_det(w, h, 4, max_pts=4096)
    for f in range(nb):
        data, spts, sdesc = run_detect(single, frames[f], max_pts=4096)
        assert data.num_pts == counts[f]
        bp = pts[f].cpu().numpy().view(sb.POINT_DTYPE)[: counts[f]]
        bd = desc[f, : counts[f]].cpu().numpy()
        for got_p, got_d in ((bp, bd), (hp[f, : hc[f]], hd[f, : hc[f]])):
            o1 = np.argsort(got_p, order=["x", "y", "scale"])
            o2 = np.argsort(spts, order=["x", "y", "scale"])
            assert np.array_equal(got_p["x"][o1], spts["x"][o2])
            assert np.array_equal(got_d[o1], sdesc[o2])


def test_edge_cases():
    sb = _sb()
    # blank frame: no keypoints, nothing crashes
    det = make_det(128, 96, 2, max_pts=64)
    data, pts, desc = run_detect(det, np.full((96, 128), 77, np.uint8), max_pts=64)
    assert data.num_pts == 0
    # keypoint cap: count clamps to max_pts, nothing is written past the buffer
    img = sb.synth_frame(640, 480, 9)
    det = make_det(640, 480, 4, max_pts=100)
    torch = _torch()
    d_img, whp = upload(img)
    data = sb.initSurfData(100)
    guard = torch.full((4800 + 4800,), 0xAB, dtype=torch.uint8, device="cuda")
    data.d_data = guard[:4800]
    det.detectAndCompute(d_img, data, whp)
    assert data.num_pts == 100
    assert (guard[4800:] == 0xAB).all()
    # smallest frame the library accepts; 8 octaves on a small frame is refused, not crashed
    det = make_det(32, 32, 1, max_pts=16)
    run_detect(det, sb.synth_frame(32, 32, 1), max_pts=16)
    with pytest.raises(sb.SurfError):
        make_det(64, 64, 8)
    # wrong frame size for the context is an error (the reference would read uninitialised scratch)
    det = make_det(128, 96, 2, max_pts=64)
    d_img, whp = upload(sb.synth_frame(100, 96, 1))
    with pytest.raises(sb.SurfError):
        det.detectAndCompute(d_img, sb.initSurfData(64), whp)


# ---------------------------------------------------------------------------------------- doubled=true

@pytest.mark.parametrize("w,h,upright", [(640, 480, True), (333, 251, False)])
def test_doubled_vs_oracle(w, h, upright):
    """Surfor::init(..., doubled=true): integral of the 2x frame and Hessian maps bit-exact, keypoints / descriptors
    as for the plain path."""
    sb = _sb()
    img = sb.synth_frame(w, h, 21)
    det = make_det(w, h, 3, upright=upright, doubled=True)
    orc = ol.Oracle(3, 4.0, True, 9, 2, upright, False, 4)
    data, pts, desc = run_detect(det, img)
    I = det.get_integral()
    Iw = orc.integral(img)
    assert I.shape == (2 * h - 1, 2 * w - 1) and np.array_equal(I, Iw)
    assert np.array_equal(det.get_response().view(np.uint32), orc.hessian(Iw).view(np.uint32))
    opts, odesc = orc.detect_and_compute(img)
    fr, fg, ok, idx, miss_r, miss_g = keypoint_parity(opts, pts)
    assert len(opts) > 100 and fr >= 0.99 and fg >= 0.99, f"{fr:.4f}/{fg:.4f}\n{describe_misses(miss_r, 4.0)}"
    l2 = np.linalg.norm(desc[idx[ok]] - odesc[ok], axis=1)
    assert l2.max() <= (TOL if upright else TOL_ROT), f"descriptor L2 max {l2.max():.3e}"


@REF
def test_doubled_vs_reference():
    """The reference's own doubled path (cuIntegralDoubleU4 + the same pipeline), stage dump and public API."""
    sb = _sb()
    w, h = 640, 480
    img = sb.synth_frame(w, h, 22)
    ref = ref_lib.Reference(w, h, 3, 4.0, True, 9, 2, True, False, 4)
    rI, rresp, rflat = ref.stages(img)
    rpts, rdesc = ref.detect(img)
    ref.close()
    det = make_det(w, h, 3, doubled=True)
    data, pts, desc = run_detect(det, img)
    assert np.array_equal(det.get_integral(), rI), "integral of the 2x frame differs from cuIntegralDoubleU4"
    assert np.array_equal(det.get_response().view(np.uint32), rflat.view(np.uint32))
    fr, fg, ok, idx, miss_r, miss_g = keypoint_parity(rpts, pts)
    assert len(rpts) > 100 and fr >= 0.99 and fg >= 0.99, f"{fr:.4f}/{fg:.4f}\n{describe_misses(miss_r, 4.0)}"
    l2 = np.linalg.norm(desc[idx[ok]] - rdesc[ok], axis=1)
    assert l2.max() <= TOL, f"descriptor L2 max {l2.max():.3e}"


def test_match_filter_ratio_laplace_cross():
    """Consumer-side acceptance: ratio test on `ambiguity`, Laplacian-sign check, symmetric cross-check --
    against the same filters in numpy on the matcher's own output."""
    sb = _sb()
    w, h = 640, 480
    left = sb.synth_frame(w, h, 5000)
    right = sb.synth_frame(w, h, 5000, 12, 2, 5000 ^ 0xA5A5)
    det = make_det(w, h, 4)
    d1, p1, f1 = run_detect(det, left)
    d_img, whp = upload(right)
    d2 = sb.initSurfData(32768, True, True)
    f2 = det.detectAndCompute(d_img, d2, whp)
    # run_detect's descriptor tensor belongs to d1's call; recompute on device for the match
    d_img1, _ = upload(left)
    d1 = sb.initSurfData(32768, True, True)
    f1 = det.detectAndCompute(d_img1, d1, whp)
    det.match(d1, d2, f1, f2)
    det.match(d2, d1, f2, f1)
    a, b = d1.host_points(), d2.host_points()
    thr = float(np.quantile(a["ambiguity"][a["match"] >= 0], 0.3))  # keeps ~30 % of the rows
    for laplace, cross in [(False, False), (True, False), (False, True), (True, True)]:
        got = det.match_filter(d1, d2, thr, laplace, cross)
        keep = (a["match"] >= 0) & (a["ambiguity"] < thr)
        j = np.clip(a["match"], 0, len(b) - 1)
        if laplace:
            keep &= b["laplace"][j] == a["laplace"]
        if cross:
            keep &= b["match"][j] == np.arange(len(a))
        want = np.nonzero(keep)[0]
        assert len(want) > 10
        assert np.array_equal(got["idx1"], want) and np.array_equal(got["idx2"], a["match"][want])
        assert np.array_equal(got["ambiguity"], a["ambiguity"][want]) and np.array_equal(got["score"], a["score"][want])


# ---------------------------------------------------------------------------------------- non-default Surfor::init arguments

@pytest.mark.parametrize("kw", [
    dict(init_mask_size=6),                  # lobe 2 -> 4 layers per octave, even lobes
    dict(init_mask_size=12),                 # lobe 4 -> 6 layers per octave
    dict(sampling_step=1, noctaves=3),       # octave 0 sampled at every pixel
    dict(sampling_step=3, noctaves=3),
    dict(desc_wsz=2),                        # 2x2 cells, 16-d descriptors, 6x wider cell spacing
    dict(desc_wsz=3, upright=False),         # 36-d, rotated
    dict(noctaves=1),
    dict(thresh=0.5, noctaves=2),            # many weak keypoints
    dict(extend=True, desc_wsz=2),           # 32-d SURF-128 layout on a 2x2 grid
])
def test_init_argument_variants_vs_oracle(kw):
    """Every Surfor::init argument (surf.h:27-29) away from the main.cpp defaults: maps bit-exact, keypoints and
    descriptors as for the default configuration. These take the generic kernels (no octave-0 fast path, row tables
    wider than 48 entries, other descriptor sizes)."""
    sb = _sb()
    w, h = 400, 300
    a = dict(noctaves=4, thresh=4.0, init_mask_size=9, sampling_step=2, upright=True, extend=False, desc_wsz=4)
    a.update(kw)
    img = sb.synth_frame(w, h, 77)
    det = sb.Surfor()
    det.init(a["noctaves"], a["thresh"], False, a["init_mask_size"], a["sampling_step"], a["upright"], a["extend"], a["desc_wsz"],
             w, h, max_pts=32768)
    orc = ol.Oracle(a["noctaves"], a["thresh"], False, a["init_mask_size"], a["sampling_step"], a["upright"], a["extend"], a["desc_wsz"])
    data, pts, desc = run_detect(det, img)
    I = orc.integral(img)
    assert np.array_equal(det.get_integral(), I)
    assert np.array_equal(det.get_response().view(np.uint32), orc.hessian(I).view(np.uint32)), "Hessian maps differ"
    opts, odesc = orc.detect_and_compute(img)
    assert len(opts) > 10, len(opts)
    fr, fg, ok, idx, miss_r, miss_g = keypoint_parity(opts, pts)
    assert fr >= 0.99 and fg >= 0.99, f"{fr:.4f}/{fg:.4f}\n{describe_misses(miss_r, a['thresh'])}"
    assert desc.shape[1] == odesc.shape[1] == a["desc_wsz"] ** 2 * (8 if a["extend"] else 4)
    l2 = np.linalg.norm(desc[idx[ok]] - odesc[ok], axis=1)
    assert l2.max() <= (TOL if a["upright"] else TOL_ROT), f"descriptor L2 max {l2.max():.3e}"
    # matching for this descriptor size (16/32/36-d take the generic kernel, 64/128-d the tensor-core one)
    torch = _torch()
    d_img, whp = upload(img)
    d1 = sb.initSurfData(32768, True, True)
    f1 = det.detectAndCompute(d_img, d1, whp)
    d_img2, _ = upload(sb.synth_frame(w, h, 77, 6, 2, 99))
    d2 = sb.initSurfData(32768, True, True)
    f2 = det.detectAndCompute(d_img2, d2, whp)
    if d2.num_pts >= 32:
        det.match(d1, d2, f1, f2)
        got = d1.host_points()
        want = ol.match(got, f1[: d1.num_pts].cpu().numpy(), d2.host_points(), f2[: d2.num_pts].cpu().numpy())
        assert np.array_equal(got["match"], want["match"])
        assert np.allclose(got["score"], want["score"], rtol=0, atol=1e-6)
        assert_ambiguity(got, want, "init variants")


# ---------------------------------------------------------------------------------------- full-size parity vs the reference

@REF
@pytest.mark.parametrize("w,h,seed", [(1920, 1080, 1), (3840, 2160, 2)])
@pytest.mark.parametrize("upright", [True, False])
def test_full_size_vs_reference(w, h, seed, upright, report):
    """BASELINE configs[1] (1920x1080, seed 1) and configs[2] (3840x2160, seed 2): keypoints and descriptors against the
    unmodified reference on the same frame, upright and rotated, every disagreement listed in the parity report."""
    sb = _sb()
    img = sb.synth_frame(w, h, seed)
    ref = ref_lib.Reference(w, h, 5, 4.0, False, 9, 2, upright, False, 4)
    ref.detect(img, max_pts=65536)  # discard the first call
    rpts, rdesc = ref.detect(img, max_pts=65536)
    rI, rlayers, _ = ref.stages(img)
    det = make_det(w, h, 5, upright=upright, max_pts=65536)
    data, pts, desc = run_detect(det, img, max_pts=65536)
    assert np.array_equal(det.get_integral(), rI)
    for o, (g, wnt) in enumerate(zip(det.split_response(det.get_response()), rlayers)):
        assert_resp_close(g, wnt, f"octave {o}")
    fr, fg, ok, idx, miss_r, miss_g = keypoint_parity(rpts, pts)
    l2 = np.linalg.norm(desc[idx[ok]] - rdesc[ok], axis=1)
    dori = np.abs(np.angle(np.exp(1j * (pts["ori"][idx[ok]] - rpts["ori"][ok])))) if not upright else np.zeros(1)
    spread, bar = 0.0, TOL
    if not upright:
        spread, l2 = rotated_bar(ref, img, rpts, rdesc, pts, desc, 65536)
        bar = max(TOL_ROT, spread)
    ref.close()
    report(**parity_entry(f"{w}x{h} seed {seed} upright={upright} vs reference", rpts, pts, ok, idx, miss_r, miss_g, l2,
                          extra={"ori_diff_max": float(dori.max()), "reference_run_to_run_desc_l2_max": spread, "bar": bar}))
    assert len(rpts) > 4000 and fr >= 0.99 and fg >= 0.99, f"{fr:.4f}/{fg:.4f}\nreference-only:\n{describe_misses(miss_r, 4.0)}\nours-only:\n{describe_misses(miss_g, 4.0)}"
    assert np.array_equal(pts["laplace"][idx[ok]], rpts["laplace"][ok])
    assert l2.max() <= bar, f"descriptor L2 max {l2.max():.3e}, rows over: {(l2 > TOL).sum()} of {len(l2)}, reference's own spread {spread:.3e}"


@REF
def test_stereo_1080p_match_vs_reference(report):
    """BASELINE configs[4] at full size: two 1080p stereo pairs (left seed 5000+p, right = the same texture 12 px to the
    side + noise), detect + describe here, matched here and by the reference's Surfor::match on the SAME descriptors:
    indices, scores and ambiguities must be equal row by row."""
    sb = _sb()
    torch = _torch()
    w, h = 1920, 1080
    det = make_det(w, h, 5, max_pts=32768)
    ref = ref_lib.Reference(w, h, 5)
    for p in range(2):
        sl = sb.synth_frame(w, h, 5000 + p)
        sr = sb.synth_frame(w, h, 5000 + p, 12, 2, (5000 + p) ^ 0xA5A5)
        d1, p1, f1 = run_detect(det, sl)
        d2, p2, f2 = run_detect(det, sr)
        det.match(d1, d2, torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda())
        got = d1.host_points()
        want = ref.match(p1, f1, p2, f2)
        diff = np.nonzero(got["match"] != want["match"])[0]
        report(case=f"stereo pair {p} 1080p", n1=int(len(p1)), n2=int(len(p2)), match_rows_differing=[int(i) for i in diff[:40]],
               match_rows_differing_count=int(len(diff)), accepted_ambiguity_lt_0p8=int((got["ambiguity"] < 0.8).sum()),
               score_max_abs_diff=float(np.abs(got["score"] - want["score"]).max()),
               ambiguity_max_abs_diff=float(np.abs(got["ambiguity"] - want["ambiguity"]).max()))
        assert len(p1) > 4000 and len(p2) > 4000
        assert len(diff) == 0, f"pair {p}: match index differs in rows {diff[:10]}"
        assert np.array_equal(got["score"], want["score"])
        # ambiguity = second / best. `second` is the reference's group rule applied to the TWO candidates per group that
        # the tensor-core pass ranks highest (2e-5 resolution: split-bf16 product); when a group's second and third
        # candidate are closer than that, the exact re-score may see the third: |delta ambiguity| <= 1e-4 then. Such rows
        # are rare (offline estimate: one row in ~10 pairs), listed in the parity report, and bounded here.
        near, damb = assert_ambiguity(got, want, f"stereo pair {p}")
        report(**{"case": f"stereo pair {p} 1080p ambiguity", "rows_over_1e-6": [int(i) for i in near[:20]], "max_abs_diff": float(damb.max())})
        assert np.array_equal(got["match_x"], want["match_x"]) and np.array_equal(got["match_y"], want["match_y"])
    ref.close()


def test_low_threshold_candidate_queue_cannot_overflow(report):
    """thresh 0.05 at 1080p: an order of magnitude more candidates than at the default threshold. The candidate queue has
    one slot per NMS cell (sb_info.cand_capacity), so nothing is dropped: the keypoint set equals the oracle's."""
    sb = _sb()
    w, h = 1920, 1080
    img = sb.synth_frame(w, h, 1)
    cap = 262144
    det = sb.Surfor()
    det.init(5, 0.05, False, 9, 2, True, False, 4, w, h, max_pts=cap)
    assert det.info.cand_capacity >= 300000  # 0.32 M cells at 1080p
    data, pts, _ = run_detect(det, img, max_pts=cap, desc=False)
    opts, _ = ol.Oracle(5, 0.05, False, 9, 2, True, False, 4).detect_and_compute(img, max_pts=cap, desc=False)
    fr, fg, ok, idx, miss_r, miss_g = keypoint_parity(opts, pts)
    report(**parity_entry("1080p seed 1 thresh 0.05 vs oracle", opts, pts, ok, idx, miss_r, miss_g, thresh=0.05))
    assert len(opts) > 10000 and len(pts) < cap, (len(opts), len(pts))
    assert abs(len(pts) - len(opts)) <= len(opts) // 200, (len(pts), len(opts))
    assert fr >= 0.99 and fg >= 0.99, f"{fr:.4f}/{fg:.4f}\n{describe_misses(miss_r, 0.05)}"


def test_tma_descriptor_path_equals_gather_path(monkeypatch):
    """SURFB200_DESCRIBE_TMA=1 (opt-in, read by sb_create): the keypoints with sampling step 2 and R <= 15 are described by
    describe_upright_tma_kernel from TMA-staged patches of the integral image, the rest by the gather kernel. Same
    keypoints (detection is untouched), descriptors equal to the default path's to float round-off (the two kernels sum
    the samples in different orders), incl. a batch with keypoints at the frame borders."""
    sb = _sb()
    torch = _torch()
    for (w, h, seed, noct) in [(640, 480, 5000, 4), (1920, 1080, 1, 5)]:
        img = sb.synth_frame(w, h, seed)
        det0 = make_det(w, h, noct)
        _, p0, f0 = run_detect(det0, img)
        nk0 = det0.info.kernels_per_frame
        det0.close()
        monkeypatch.setenv("SURFB200_DESCRIBE_TMA", "1")
        det1 = make_det(w, h, noct)
        monkeypatch.delenv("SURFB200_DESCRIBE_TMA")
        assert det1.info.kernels_per_frame == nk0 + 2
        for rep in range(2):  # twice: the class counters must re-arm
            _, p1, f1 = run_detect(det1, img)
            # (the append order of keypoints differs from run to run: compare in sorted order)
            k0, k1 = np.lexsort((p0["scale"], p0["y"], p0["x"])), np.lexsort((p1["scale"], p1["y"], p1["x"]))
            assert len(p1) == len(p0) and np.array_equal(p1["x"][k1], p0["x"][k0]) and np.array_equal(p1["scale"][k1], p0["scale"][k0])
            l2 = np.linalg.norm(f1[k1] - f0[k0], axis=1)
            assert l2.max() <= 1e-5, f"{w}x{h} rep {rep}: L2 max {l2.max():.3e} at row {l2.argmax()}"
        det1.close()


def test_tma_descriptor_path_batch_and_doubled(monkeypatch):
    """The opt-in TMA descriptor path with several frame slots (the tensor maps index the slot as their third coordinate,
    the class lists are per slot) and with doubled=true (the lattice lives on the 2x frame): descriptors equal to the
    default path's, frame by frame."""
    sb = _sb()
    torch = _torch()
    w, h, nb = 400, 300, 3
    frames = np.stack([sb.synth_frame(w, h, 40 + i) for i in range(nb)])
    pitch = sb.iAlignUp(w, 128)
    buf = np.zeros((nb, h, pitch), np.uint8)
    buf[:, :, :w] = frames
    d_imgs = torch.from_numpy(buf).cuda()
    for doubled in (False, True):
        out = []
        for tma in (0, 1):
            if tma:
                monkeypatch.setenv("SURFB200_DESCRIBE_TMA", "1")
            det = make_det(w, h, 3, max_pts=8192, batch=nb, doubled=doubled)
            if tma:
                monkeypatch.delenv("SURFB200_DESCRIBE_TMA")
            pts = torch.zeros((nb, 8192 * 48), dtype=torch.uint8, device="cuda")
            cnt = torch.zeros(nb, dtype=torch.int32, device="cuda")
            desc = torch.zeros((nb, 8192, 64), dtype=torch.float32, device="cuda")
            for _ in range(2):
                det.detect_batch(d_imgs, pitch, pts, cnt, desc)
            torch.cuda.synchronize()
            c = cnt.cpu().numpy()
            res = []
            for f in range(nb):
                p = pts[f].cpu().numpy().view(sb.POINT_DTYPE)[: c[f]]
                d = desc[f, : c[f]].cpu().numpy()
                o = np.lexsort((p["scale"], p["y"], p["x"]))
                res.append((p[o], d[o]))
            out.append(res)
            det.close()
        for f in range(nb):
            (p0, d0), (p1, d1) = out[0][f], out[1][f]
            assert len(p0) == len(p1) > 50 and np.array_equal(p0["x"], p1["x"])
            l2 = np.linalg.norm(d0 - d1, axis=1)
            assert l2.max() <= 1e-5, f"doubled={doubled} frame {f}: L2 max {l2.max():.3e}"


def test_match_pairs_equals_match_per_pair():
    """sb_match_pairs_async (all stereo pairs of a detect batch in one launch sequence, counts read on the device) gives,
    per pair, exactly the five match fields of sb_match on the same keypoints and descriptors; a custom pair list, a
    bound below a frame's count (the first `bound` points take part) and an empty frame are covered."""
    sb = _sb()
    torch = _torch()
    w, h, NF, MAXP = 640, 480, 4, 4096
    det = make_det(w, h, 4, max_pts=MAXP, batch=2 * NF)
    pitch = sb.iAlignUp(w, 128)
    buf = np.zeros((2 * NF, h, pitch), np.uint8)
    for p in range(NF):
        buf[2 * p, :, :w] = sb.synth_frame(w, h, 5000 + p)
        buf[2 * p + 1, :, :w] = sb.synth_frame(w, h, 5000 + p, 12, 2, (5000 + p) ^ 0xA5A5)
    buf[7] = 0  # an empty right frame: pair 3 has no candidates
    d = torch.from_numpy(buf).cuda()
    pts = torch.zeros((2 * NF, MAXP * 48), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(2 * NF, dtype=torch.int32, device="cuda")
    desc = torch.zeros((2 * NF, MAXP, 64), dtype=torch.float32, device="cuda")
    det.detect_batch(d, pitch, pts, cnt, desc)
    torch.cuda.synchronize()
    counts = cnt.cpu().numpy()
    assert counts[0] > 300 and counts[7] == 0

    class View:
        def __init__(self, f, n): self.d_data = pts[f]; self.num_pts = int(n); self.h_data = None

    def fields(f, n):
        a = pts[f].cpu().numpy().view(sb.POINT_DTYPE)[:n]
        return {k: a[k].copy() for k in ("score", "match", "match_x", "match_y", "ambiguity")}

    def clear():
        v = pts.view(2 * NF, MAXP, 48)
        v[:, :, 28:48] = 0  # score, match, match_x, match_y, ambiguity
        torch.cuda.synchronize()

    for bound, pairs in [(MAXP, None), (256, None), (MAXP, [(0, 1), (2, 0), (1, 3), (5, 4)])]:
        plist = pairs or [(2 * z, 2 * z + 1) for z in range(NF)]
        want = []
        clear()
        for a, b in plist:
            det.match_async(View(a, min(counts[a], bound)), View(b, min(counts[b], bound)), desc[a], desc[b])
            torch.cuda.synchronize()
            want.append(fields(a, min(counts[a], bound)))
            clear()
        dp = torch.tensor(pairs, dtype=torch.int32, device="cuda") if pairs else None
        det.match_pairs_async(pts, cnt, desc, len(plist), bound, dp)
        torch.cuda.synchronize()
        got = [fields(a, min(counts[a], bound)) for a, _ in plist]
        for z, (g, wnt) in enumerate(zip(got, want)):
            for k in g:
                assert np.array_equal(g[k], wnt[k]), f"bound {bound} pairs {pairs} pair {z} field {k}"
    det.close()
