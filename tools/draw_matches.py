"""The visual outputs of the reference demo (main.cpp:21-71 drawKeypoints / drawMatches, 262-268) on this library:
detect + describe a pair, match left -> right on the device, write the keypoint and match pictures.

usage: python tools/draw_matches.py [left right] [--out DIR] [--rotated] [--max-ambiguity A] [--laplace] [--cross]
       (no images: the reference's bundled 1280x960 pair from tests/golden)

Consumer-side tool: it only calls the library's public interface (Surfor.init / detectAndCompute /
match / match_filter) and OpenCV for image I/O and drawing."""
import argparse
import os
import sys

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_surf_b200 as sb  # noqa: E402


def draw_keypoints(gray, pts):
    """main.cpp:21-43: a circle of the keypoint's scale (and its orientation, when there is one) per keypoint."""
    img = cv2.cvtColor(gray, cv2.COLOR_GRAY2BGR)
    for p in pts:
        c = (int(round(float(p["x"]))), int(round(float(p["y"]))))
        r = max(1, int(round(2.0 * float(p["scale"]))))
        col = (0, 255, 0) if p["laplace"] > 0 else (0, 128, 255)
        cv2.circle(img, c, r, col, 1, cv2.LINE_AA)
        if p["ori"] != 0.0:
            e = (int(round(c[0] + r * np.cos(p["ori"]))), int(round(c[1] + r * np.sin(p["ori"]))))
            cv2.line(img, c, e, col, 1, cv2.LINE_AA)
    return img


def draw_matches(left, right, pl, pr, pairs):
    """main.cpp:45-71: the two frames side by side, a line per accepted match."""
    h, w = left.shape
    img = np.zeros((max(h, right.shape[0]), w + right.shape[1], 3), np.uint8)
    img[:h, :w] = cv2.cvtColor(left, cv2.COLOR_GRAY2BGR)
    img[: right.shape[0], w:] = cv2.cvtColor(right, cv2.COLOR_GRAY2BGR)
    rng = np.random.default_rng(0)
    for m in pairs:
        a, b = pl[int(m["idx1"])], pr[int(m["idx2"])]
        col = tuple(int(v) for v in rng.integers(64, 256, 3))
        p0 = (int(round(float(a["x"]))), int(round(float(a["y"]))))
        p1 = (w + int(round(float(b["x"]))), int(round(float(b["y"]))))
        cv2.line(img, p0, p1, col, 1, cv2.LINE_AA)
        cv2.circle(img, p0, 2, col, -1)
        cv2.circle(img, p1, 2, col, -1)
    return img


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("images", nargs="*")
    ap.add_argument("--out", default="gpurun_out")
    ap.add_argument("--rotated", action="store_true")
    ap.add_argument("--max-ambiguity", type=float, default=0.95)
    ap.add_argument("--laplace", action="store_true")
    ap.add_argument("--cross", action="store_true")
    ap.add_argument("--max-lines", type=int, default=400)
    a = ap.parse_args()
    if len(a.images) == 2:
        left, right = (cv2.imread(p, cv2.IMREAD_GRAYSCALE) for p in a.images)
    else:
        g = os.path.join(ROOT, "tests", "golden")
        left = cv2.imread(os.path.join(g, "left_1280x960.png"), cv2.IMREAD_GRAYSCALE)
        right = cv2.imread(os.path.join(g, "right_1280x960.png"), cv2.IMREAD_GRAYSCALE)
    if left is None or right is None or left.shape != right.shape:
        sys.exit("need two grey images of equal size")
    h, w = left.shape
    # main.cpp:187-204 defaults
    det = sb.Surfor()
    det.init(4, 4.0, False, 9, 2, not a.rotated, False, 4, w, h, max_pts=10000)
    pitch = sb.iAlignUp(w, 128)
    data, desc, pts = [], [], []
    for img in (left, right):
        buf = np.zeros((h, pitch), np.uint8)
        buf[:, :w] = img
        d = sb.initSurfData(10000)
        desc.append(det.detectAndCompute(torch.from_numpy(buf).cuda(), d, (w, h, pitch)))
        data.append(d)
    det.match(data[0], data[1], desc[0], desc[1])
    if a.cross:
        det.match(data[1], data[0], desc[1], desc[0])
    pairs = det.match_filter(data[0], data[1], a.max_ambiguity, a.laplace, a.cross)
    pts = [d.host_points() for d in data]
    print(f"keypoints {data[0].num_pts} / {data[1].num_pts}, accepted matches {len(pairs)} "
          f"(ambiguity < {a.max_ambiguity}{', equal laplace' if a.laplace else ''}{', cross-checked' if a.cross else ''})")
    os.makedirs(a.out, exist_ok=True)
    cv2.imwrite(os.path.join(a.out, "keypoints_left.png"), draw_keypoints(left, pts[0]))
    cv2.imwrite(os.path.join(a.out, "keypoints_right.png"), draw_keypoints(right, pts[1]))
    order = np.argsort(pairs["ambiguity"])[: a.max_lines]
    cv2.imwrite(os.path.join(a.out, "matches.png"), draw_matches(left, right, pts[0], pts[1], pairs[order]))
    print("wrote", ", ".join(os.path.join(a.out, n) for n in ("keypoints_left.png", "keypoints_right.png", "matches.png")))


if __name__ == "__main__":
    main()
