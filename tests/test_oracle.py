"""CPU oracle (oracle/liboracle.so): internal consistency, the host schedule against the values derived
by hand from the reference, and -- when tests/golden/*.npz exist -- parity with the
outputs of the unmodified reference run on a B200 (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import cuda_surf_b200 as sb
import oracle_lib as ol
from helpers import keypoint_parity, load_golden, load_pair

import zlib


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def test_schedule_matches_reference_constants():
    # o0 lobes [3,5,7,9,11], computed borders [6,6,6,7,9], borders[] [6,6,6,6,7], mborders [7,8];
    # o>=1: three new lobes, computed borders [8,8,9], borders[] all 8, mborders [9,9]
    o = ol.Oracle(5, 4.0, False, 9, 2, True, False, 4)
    s = o.schedule(1920, 1080)
    assert [s[0].l[i] for i in range(5)] == [3, 5, 7, 9, 11]
    assert [s[0].b1[i] for i in range(5)] == [6, 6, 6, 7, 9]
    assert [s[0].borders[i] for i in range(5)] == [6, 6, 6, 6, 7]
    assert [s[0].mb[i] for i in range(2)] == [7, 8]
    lobes = {1: [15, 19, 23], 2: [31, 39, 47], 3: [63, 79, 95], 4: [127, 159, 191]}
    for k in range(1, 5):
        assert [s[k].l[i] for i in range(3)] == lobes[k]
        assert [s[k].b1[i] for i in range(3)] == [8, 8, 9]
        assert [s[k].borders[i] for i in range(5)] == [8] * 5
        assert [s[k].mb[i] for i in range(2)] == [9, 9]
    assert [(s[k].sw, s[k].sh) for k in range(5)] == [(960, 540), (480, 270), (240, 135), (120, 67), (60, 33)]
    assert abs(s[0].norm[0] - 1.0) < 1e-7  # (9/3^2)^2


def test_integral_against_numpy():
    rng = np.random.default_rng(0)
    for (h, w) in [(1, 1), (7, 13), (96, 128), (251, 333)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        I = ol.Oracle().integral(img)
        want = np.zeros((h + 1, w + 1), np.int64)
        want[1:, 1:] = img.astype(np.int64).cumsum(0).cumsum(1)
        assert np.array_equal(I, want)


def test_hessian_of_a_blob_and_linearity_of_box_sums():
    # a bright Gaussian blob gives one strong maximum with laplace sign -1 near its centre
    y, x = np.mgrid[0:200, 0:200]
    img = (30 + 200 * np.exp(-((x - 100.3) ** 2 + (y - 90.7) ** 2) / (2 * 6.0 ** 2))).astype(np.uint8)
    o = ol.Oracle(3, 4.0, False, 9, 2, True, False, 4)
    pts, desc = o.detect_and_compute(img)
    assert len(pts) >= 1
    p = pts[np.argmax(pts["strength"])]
    assert abs(p["x"] - 100.3) < 1.0 and abs(p["y"] - 90.7) < 1.0 and p["laplace"] == -1
    assert np.allclose(np.linalg.norm(desc, axis=1), 1.0, atol=1e-5)
    # a constant image has an exactly zero response everywhere
    I = o.integral(np.full((120, 160), 91, np.uint8))
    assert not o.hessian(I).any()


def test_bundled_pair_counts():
    """an independent NumPy emulation of the reference predicts 2739 / 3443 keypoints"""
    left, right = load_pair()
    o = ol.Oracle(4, 4.0, False, 9, 2, True, False, 4)
    pl, dl = o.detect_and_compute(left)
    pr, dr = o.detect_and_compute(right)
    assert (len(pl), len(pr)) == (2739, 3443)
    assert list(np.bincount(pl["o"])) == [1928, 598, 174, 39]
    m = ol.match(pl, dl, pr, dr)
    assert (m["match"] >= 0).all() and (m["match"] < 3424).all()  # 3443 rounded down to a multiple of 32


def test_match_semantics_small():
    # group rule: second best inside the winner's group (!= group 0) is invisible to the merge
    f2 = np.zeros((32, 64), np.float32)
    f1 = np.zeros((1, 64), np.float32)
    f1[0, 0] = 1
    f2[9, 0] = 0.9   # group 2 (p2 % 32 // 4)
    f2[10, 0] = 0.8  # same group: true second best
    f2[20, 0] = 0.5  # group 5: what the reference reports as second
    p1 = np.zeros(1, ol.POINT_DTYPE)
    p2 = np.zeros(32, ol.POINT_DTYPE)
    m = ol.match(p1, f1, p2, f2)
    assert m["match"][0] == 9 and abs(m["score"][0] - 0.9) < 1e-7
    assert abs(m["ambiguity"][0] - 0.5 / (0.9 + 1e-6)) < 1e-6


GOLD = ["pair_left_upright", "pair_right_upright", "pair_left_rotated", "small_upright", "small_rotated",
        "small_extend", "ragged_upright", "stereo_left", "stereo_right", "synth1080_upright",
        "small_doubled", "small_doubled_rotated"]


def _image_for(name, g):
    if name.startswith("pair_left"):
        return load_pair()[0]
    if name.startswith("pair_right"):
        return load_pair()[1]
    if "img" in g:
        return g["img"]
    if name == "stereo_left":
        return sb.synth_frame(640, 480, 5000)
    if name == "stereo_right":
        return sb.synth_frame(640, 480, 5000, 12, 2, 5000 ^ 0xA5A5)
    if name == "synth1080_upright":
        return sb.synth_frame(1920, 1080, 1)
    raise KeyError(name)


@pytest.mark.parametrize("name", GOLD)
def test_oracle_against_reference_golden(name):
    g = load_golden(name)
    if g is None:
        pytest.skip(f"tests/golden/{name}.npz not generated yet (needs a GPU run of make_golden.py)")
    img = _image_for(name, g)
    h, w = img.shape
    assert (w, h) == (int(g["w"]), int(g["h"]))
    upright, extend, noct = bool(g["upright"]), bool(g["extend"]), int(g["noctaves"])
    doubled = bool(g["doubled"]) if "doubled" in g else False
    o = ol.Oracle(noct, float(g["thresh"]), doubled, 9, 2, upright, extend, 4)
    I = o.integral(img)
    resp = o.hessian(I)
    layers = o.split_resp(resp, I.shape[1] - 1, I.shape[0] - 1)  # the 2x frame when doubled
    if "integral" in g:
        assert np.array_equal(I, g["integral"])
        for k, L in enumerate(layers):
            assert np.array_equal(L, g[f"resp{k}"]), f"octave {k}: max abs diff {np.abs(L - g[f'resp{k}']).max()}"
    else:
        assert crc(I) == g["integral_crc"]
        for k, L in enumerate(layers):
            got = np.array([crc(L[s]) for s in range(L.shape[0])], np.uint32)
            assert np.array_equal(got, g[f"resp{k}_crc"]), f"octave {k}: Hessian maps are not bit-identical"
    rpts = g["pts"]
    pts = o.keypoints(I, resp)
    assert abs(len(pts) - len(rpts)) <= max(1, len(rpts) // 200), (len(pts), len(rpts))
    fr, fg, ok, idx, miss_r, miss_g = keypoint_parity(rpts, pts)
    assert fr >= 0.99 and fg >= 0.99, (fr, fg)
    assert np.array_equal(pts["laplace"][idx[ok]], rpts["laplace"][ok])
    # descriptors (and orientation) on the reference's own keypoints
    rdesc = g["desc"]
    sub = rpts[: len(rdesc)].copy()
    if not upright:
        mine = o.orientation(I, sub)
        dori = np.abs(np.angle(np.exp(1j * (mine["ori"] - sub["ori"]))))
        assert np.nanmax(dori) < 1e-3
    d = o.describe(I, sub)
    l2 = np.linalg.norm(d - rdesc, axis=1)
    assert (l2 <= 1e-3).mean() >= 0.99 and np.nanmax(l2) < 2e-2, (np.nanmax(l2), (l2 <= 1e-3).mean())


def test_oracle_match_against_reference_golden():
    g = load_golden("stereo_match")
    if g is None:
        pytest.skip("tests/golden/stereo_match.npz not generated yet")
    m = ol.match(g["pts1"], g["desc1"], g["pts2"], g["desc2"])
    want = g["matched"]
    assert np.array_equal(m["match"], want["match"])
    assert np.array_equal(m["score"], want["score"])
    assert np.allclose(m["ambiguity"], want["ambiguity"], atol=1e-6)


def test_doubled_upsample_and_integral():
    """doubled=true (surf.cpp:234-235): the integral is that of the (2w-2) x (2h-2) bilinear 2x frame defined by
    integralDoubleRow0U2 (surfd.cu:166-207)."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    h, w = img.shape
    orc = ol.Oracle(3, 4.0, True, 9, 2, True, False, 4)
    I = orc.integral(img)
    assert I.shape == (2 * h - 1, 2 * w - 1)
    a = img.astype(np.float32)
    up = np.zeros((2 * h - 2, 2 * w - 2), np.float32)
    up[0::2, 0::2] = a[: h - 1, : w - 1]
    up[0::2, 1::2] = np.rint((a[: h - 1, : w - 1] + a[: h - 1, 1:]) * np.float32(0.5))       # rint: ties to even, as cvt.rni
    up[1::2, 0::2] = np.rint((a[: h - 1, : w - 1] + a[1:, : w - 1]) * np.float32(0.5))
    up[1::2, 1::2] = np.rint((a[: h - 1, : w - 1] + a[: h - 1, 1:] + a[1:, : w - 1] + a[1:, 1:]) * np.float32(0.25))
    want = np.zeros_like(I)
    want[1:, 1:] = up.astype(np.int64).cumsum(0).cumsum(1)
    assert np.array_equal(I, want)
    # the whole path runs on the 2x frame: sampling doubles, positions and scales are halved back (surf.cpp:69-72)
    assert orc.p.sampling == 4 and orc.p.divisor == 0.5


def test_doubled_detect_is_consistent_with_plain_detect_on_the_2x_frame():
    """Keypoints of doubled=true on a frame == keypoints of doubled=false with sampling 4 on the 2x frame, with
    x, y, scale halved (makePoint, surfd.cu:1003-1006)."""
    import cuda_surf_b200 as sb
    img = sb.synth_frame(320, 240, 7)
    h, w = img.shape
    od = ol.Oracle(3, 4.0, True, 9, 2, True, False, 4)
    pd, dd = od.detect_and_compute(img)
    assert len(pd) > 50
    I2 = od.integral(img)
    op = ol.Oracle(3, 4.0, False, 9, 4, True, False, 4)
    pp = op.keypoints(I2, op.hessian(I2))
    assert len(pp) == len(pd)
    assert np.allclose(np.sort(pp["x"]) * 0.5, np.sort(pd["x"]), atol=1e-5)
    assert np.allclose(np.sort(pp["scale"]) * 0.5, np.sort(pd["scale"]), atol=1e-5)
    assert np.allclose(np.linalg.norm(dd, axis=1), 1.0, atol=1e-4)


def test_synth_numpy_port_equals_library_generator():
    """tests/synth_np.py (used by the reference arm of bench.py, which must not load the product library) is
    bit-identical to sb_synth_frame (host code of the library; no GPU needed)."""
    import cuda_surf_b200 as sb
    import synth_np
    for w, h, seed, extra in [(320, 240, 3, ()), (333, 251, 4, (12, 2, 99)), (640, 480, 5000, (12, 2, 5000 ^ 0xA5A5))]:
        assert np.array_equal(synth_np.synth_frame(w, h, seed, *extra), sb.synth_frame(w, h, seed, *extra))
