"""The C++ drop-in on hardware: build/surf_demo is the reference's main.cpp flow (cudaSurfDemo2, main.cpp:163-283)
compiled against include/compat/surf.h -- the reference's own surf::Surfor interface -- and linked with libsurfb200.so.
It is executed here on the bundled stereo pair and on a synthetic pair; keypoint counts, the match results and the
descriptor-buffer ownership contract (surfd.cu:3262-3266) are checked against the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from helpers import load_pair

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "build", "surf_demo")


def write_pgm(path, img):
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


def run_demo(args, env_extra, dump):
    env = dict(os.environ, SURF_DEMO_REPEATS="3", SURF_DEMO_DUMP=dump, **env_extra)
    out = subprocess.run([DEMO, "0"] + args, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    kv = dict(line.split("=", 1) for line in out.stdout.splitlines() if "=" in line and " " not in line.split("=", 1)[0])
    pts = np.fromfile(dump, ol.POINT_DTYPE)
    return kv, pts, out.stdout


def oracle_pair(left, right):
    orc = ol.Oracle(4, 4.0, False, 9, 2, True, False, 4)
    p1, f1 = orc.detect_and_compute(left)
    p2, f2 = orc.detect_and_compute(right)
    return p1, f1, p2, f2


@pytest.mark.skipif(not os.path.exists(DEMO), reason="build/surf_demo not built (make demo)")
@pytest.mark.parametrize("fresh", [0, 1])
def test_surf_demo_bundled_pair(tmp_path, fresh, report):
    left, right = load_pair()
    lp, rp = str(tmp_path / "left.pgm"), str(tmp_path / "right.pgm")
    write_pgm(lp, left)
    write_pgm(rp, right)
    kv, pts, text = run_demo([lp, rp], {"SURFB200_FRESH_DESC": str(fresh)}, str(tmp_path / "pts.bin"))
    assert int(kv["features1"]) == 2739 and int(kv["features2"]) == 3443, text
    # the *desc_addr contract: fresh_desc=1 is the reference's (new buffer per call, the first one untouched),
    # the default reuses the caller's buffer (and therefore overwrites it with the second frame's descriptors)
    assert int(kv["desc_fresh"]) == fresh and int(kv["first_desc_intact"]) == fresh, text
    # match results of the demo's final state against the oracle on ITS OWN keypoints / descriptors is not possible from
    # outside the process; instead the oracle runs the whole flow and the per-row results are compared by keypoint
    p1, f1, p2, f2 = oracle_pair(left, right)
    want = ol.match(p1, f1, p2, f2)
    assert len(pts) == len(want) == 2739
    ko = np.lexsort((want["scale"], want["y"], want["x"]))
    kg = np.lexsort((pts["scale"], pts["y"], pts["x"]))
    assert np.allclose(pts["x"][kg], want["x"][ko], atol=1e-3) and np.allclose(pts["y"][kg], want["y"][ko], atol=1e-3)
    # the matched partner is compared by position (keypoint order differs between the two implementations)
    same = (np.abs(pts["match_x"][kg] - want["match_x"][ko]) < 1e-3) & (np.abs(pts["match_y"][kg] - want["match_y"][ko]) < 1e-3)
    good_g, good_o = int((pts["ambiguity"] < 0.8).sum()), int((want["ambiguity"] < 0.8).sum())
    report(case=f"surf_demo bundled pair fresh={fresh}", features1=2739, features2=3443, match_partner_differs=int((~same).sum()),
           good_demo=good_g, good_oracle=good_o, score_max_abs_diff=float(np.abs(pts["score"][kg] - want["score"][ko]).max()))
    assert int(kv["good"]) == good_g
    # the oracle's candidate set is the first n2 - n2 % 32 points of ITS order, the demo's of the GPU's append order: the
    # 19 tail points differ, so a few rows may legitimately pick another partner
    assert (~same).sum() <= 0.02 * len(pts), (~same).sum()
    assert abs(good_g - good_o) <= max(3, good_o // 20), (good_g, good_o)


@pytest.mark.skipif(not os.path.exists(DEMO), reason="build/surf_demo not built (make demo)")
def test_surf_demo_synth_pair(tmp_path, report):
    import cuda_surf_b200 as sb
    left = sb.synth_frame(1280, 960, 5000)
    right = sb.synth_frame(1280, 960, 5000, 12, 2, 5000 ^ 0xA5A5)
    kv, pts, text = run_demo([], {}, str(tmp_path / "pts.bin"))
    p1, f1, p2, f2 = oracle_pair(left, right)
    report(case="surf_demo synth pair", features1=int(kv["features1"]), features2=int(kv["features2"]),
           oracle_features1=int(len(p1)), oracle_features2=int(len(p2)))
    assert abs(int(kv["features1"]) - len(p1)) <= max(2, len(p1) // 100) and abs(int(kv["features2"]) - len(p2)) <= max(2, len(p2) // 100), text
    assert len(pts) == int(kv["features1"])
