"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libsurfref.so, CUDA)
on a B200:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'
    cp gpurun_out/golden/*.npz tests/golden/

Inputs: the reference's bundled stereo pair (stored losslessly as tests/golden/{left,right}_1280x960.png,
made from /root/reference/data/*.pgm) and deterministic `synth_v1` frames (cuda-surf_b200/csrc/synth.cpp).
The reference ships no golden vectors of its own, so these files are what pins
both the CPU oracle (tests/test_oracle_golden.py, no GPU) and the CUDA product (tests -m gpu).
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

import cuda_surf_b200 as sb  # noqa: E402
import ref_lib  # noqa: E402

NDESC = 512  # descriptors kept per image for the large cases (keeps the fixtures small)


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def stage_sums(integral, layers):
    """bit-exact fingerprints of the big intermediates + a coarse subsample for debugging"""
    out = {"integral_crc": crc(integral), "integral_sub": integral[::16, ::16].copy()}
    for o, L in enumerate(layers):
        out[f"resp{o}_crc"] = np.array([crc(L[s]) for s in range(L.shape[0])], np.uint32)
        out[f"resp{o}_sub"] = L[:, ::8, ::8].copy()
    return out


def run_case(name, img, outdir, noctaves, upright, extend=False, full=False, thresh=4.0, max_pts=65536, doubled=False):
    h, w = img.shape
    ref = ref_lib.Reference(w, h, noctaves, thresh, doubled, 9, 2, upright, extend, 4)
    pts, desc = ref.detect(img, max_pts=max_pts, desc=True)
    d = {"w": w, "h": h, "noctaves": noctaves, "upright": int(upright), "extend": int(extend), "thresh": thresh,
         "doubled": int(doubled), "pts": pts, "npts": len(pts)}
    if full:
        d["desc"] = desc
        integral, layers, _ = ref.stages(img)
        d["integral"] = integral
        for o, L in enumerate(layers):
            d[f"resp{o}"] = L
        d["img"] = img
    else:
        d["desc"] = desc[:NDESC]
        integral, layers, _ = ref.stages(img)
        d.update(stage_sums(integral, layers))
    ref.close()
    np.savez_compressed(os.path.join(outdir, name + ".npz"), **d)
    print(name, "pts", len(pts), "desc nan", int(np.isnan(desc).sum()))
    return pts, desc


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[2] == "doubled":
        # doubled=true (surf.cpp:234-235): the reference's cuIntegralDoubleU4 path; added after the first set, so that
        # the committed files of the other cases (nondeterministic keypoint order in the reference) stay as they are
        run_case("small_doubled", sb.synth_frame(320, 240, 3), outdir, 3, True, full=True, doubled=True)
        run_case("small_doubled_rotated", sb.synth_frame(200, 152, 5), outdir, 2, False, full=True, doubled=True)
        return
    left = cv2.imread(os.path.join(HERE, "left_1280x960.png"), cv2.IMREAD_GRAYSCALE)
    right = cv2.imread(os.path.join(HERE, "right_1280x960.png"), cv2.IMREAD_GRAYSCALE)
    # config 1: main.cpp defaults (4 octaves, thresh 4, upright, 64-d), plus one rotated run
    run_case("pair_left_upright", left, outdir, 4, True)
    run_case("pair_right_upright", right, outdir, 4, True)
    run_case("pair_left_rotated", left, outdir, 4, False)
    # small synthetic frames, everything kept
    small = sb.synth_frame(320, 240, 3)
    run_case("small_upright", small, outdir, 3, True, full=True)
    run_case("small_rotated", small, outdir, 3, False, full=True)
    run_case("small_extend", small, outdir, 3, True, extend=True, full=True)
    ragged = sb.synth_frame(333, 251, 4)
    run_case("ragged_upright", ragged, outdir, 3, True, full=True)
    # stereo pair for matching: left, and the same texture 12 px to the side with +-2 grey levels of noise
    sl = sb.synth_frame(640, 480, 5000)
    sr = sb.synth_frame(640, 480, 5000, 12, 2, 5000 ^ 0xA5A5)
    pl, dl = run_case("stereo_left", sl, outdir, 4, True, full=False)
    pr, dr = run_case("stereo_right", sr, outdir, 4, True, full=False)
    ref = ref_lib.Reference(640, 480, 4, 4.0, False, 9, 2, True, False, 4)
    m = ref.match(pl, dl, pr, dr)
    ref.close()
    np.savez_compressed(os.path.join(outdir, "stereo_match.npz"), pts1=pl, desc1=dl, pts2=pr, desc2=dr, matched=m)
    print("stereo match: n1", len(pl), "n2", len(pr), "ambiguity<0.8:", int((m["ambiguity"] < 0.8).sum()))
    # the 1080p bench frame: counts only (sanity for the synthetic density)
    big = sb.synth_frame(1920, 1080, 1)
    ref = ref_lib.Reference(1920, 1080, 5, 4.0, False, 9, 2, True, False, 4)
    pts, desc = ref.detect(big, max_pts=32768, desc=True)
    integral, layers, _ = ref.stages(big)
    ref.close()
    d = {"w": 1920, "h": 1080, "noctaves": 5, "upright": 1, "extend": 0, "thresh": 4.0, "pts": pts, "npts": len(pts),
         "desc": desc[:NDESC]}
    d.update(stage_sums(integral, layers))
    np.savez_compressed(os.path.join(outdir, "synth1080_upright.npz"), **d)
    print("synth1080", len(pts))


if __name__ == "__main__":
    main()
