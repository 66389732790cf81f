"""Shared helpers of the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    p = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(p):
        return None
    return dict(np.load(p))


def load_pair():
    import cv2
    left = cv2.imread(os.path.join(GOLDEN, "left_1280x960.png"), cv2.IMREAD_GRAYSCALE)
    right = cv2.imread(os.path.join(GOLDEN, "right_1280x960.png"), cv2.IMREAD_GRAYSCALE)
    return left, right


def match_keypoints(a, b, tol_xy=0.1, tol_scale=0.05):
    """For every point of `a` find the nearest point of `b` in (x, y). Returns (matched mask over a,
    index into b, listing of the unmatched points of a)."""
    from scipy.spatial import cKDTree
    if len(a) == 0 or len(b) == 0:
        return np.zeros(len(a), bool), np.zeros(len(a), int), a
    tree = cKDTree(np.stack([b["x"], b["y"]], 1).astype(np.float64))
    # several candidates: two keypoints can share (x, y) at different scales
    k = min(4, len(b))
    dist, idx = tree.query(np.stack([a["x"], a["y"]], 1).astype(np.float64), k=k)
    if k == 1:
        dist, idx = dist[:, None], idx[:, None]
    ok = np.zeros(len(a), bool)
    best = idx[:, 0].copy()
    for j in range(k):
        cand = (dist[:, j] <= tol_xy) & (np.abs(a["scale"] - b["scale"][idx[:, j]]) <= tol_scale)
        new = cand & ~ok
        best[new] = idx[new, j]
        ok |= cand
    return ok, best, a[~ok]


def keypoint_parity(ref, got, frac=0.99):
    """Set-based parity demanded by BASELINE.json: >= 99 % of keypoints matched within 0.1 px and
    0.05 in scale, both ways; disagreements are returned for listing."""
    ok_r, idx_r, miss_r = match_keypoints(ref, got)
    ok_g, idx_g, miss_g = match_keypoints(got, ref)
    fr = ok_r.mean() if len(ref) else 1.0
    fg = ok_g.mean() if len(got) else 1.0
    return fr, fg, ok_r, idx_r, miss_r, miss_g


def describe_misses(miss, thresh):
    lines = []
    for p in miss[:20]:
        lines.append(f"  x={p['x']:.3f} y={p['y']:.3f} scale={p['scale']:.3f} strength={p['strength']:.4f}"
                     f"{'  (threshold boundary)' if abs(p['strength'] - thresh) < 0.02 * thresh else ''}")
    return "\n".join(lines)


def point_rows(pts, limit=40):
    """keypoints as plain dicts for the parity report"""
    return [{"x": float(p["x"]), "y": float(p["y"]), "scale": float(p["scale"]), "strength": float(p["strength"])}
            for p in pts[:limit]]


def parity_entry(what, ref_pts, got_pts, ok, idx, miss_r, miss_g, l2=None, thresh=4.0, extra=None):
    """One record of the parity report: counts, unmatched keypoints both ways (flagged when their strength is within
    2 % of the threshold), descriptor rows whose L2 distance exceeds 1e-4 (listed) and the maximum."""
    e = {"case": what, "n_ref": int(len(ref_pts)), "n_ours": int(len(got_pts)), "matched": int(ok.sum()),
         "ref_only": [dict(r, threshold_boundary=bool(abs(r["strength"] - thresh) < 0.02 * thresh)) for r in point_rows(miss_r)],
         "ours_only": [dict(r, threshold_boundary=bool(abs(r["strength"] - thresh) < 0.02 * thresh)) for r in point_rows(miss_g)]}
    if l2 is not None and len(l2):
        over = np.nonzero(l2 > 1e-4)[0]
        e["desc_l2_max"] = float(l2.max())
        e["desc_rows_over_1e-4"] = [{"ref_row": int(np.nonzero(ok)[0][i]), "l2": float(l2[i])} for i in over[:40]]
        e["desc_rows_over_1e-4_count"] = int(len(over))
    if extra:
        e.update(extra)
    return e
