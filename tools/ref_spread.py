#!/usr/bin/env python
"""Run-to-run spread of the REFERENCE's own rotated path (its orientation histogram and descriptor sums use
order-dependent float atomics, surfd.cu:1795-1805, 1222-1266) next to the distance between this library and the
reference on the same frames. GPU only; writes gpurun_out/ref_spread.json (copied to profiles/ by hand).

    python tools/ref_spread.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import cuda_surf_b200 as sb  # noqa: E402
import ref_lib  # noqa: E402
from helpers import load_pair  # noqa: E402


def key(p):
    return np.stack([p["x"], p["y"], p["scale"]], 1).astype(np.float64)


def align(a, b):
    """indices into b of the points of a with identical (x, y, scale); -1 when absent"""
    kb = {tuple(v): i for i, v in enumerate(key(b))}
    return np.array([kb.get(tuple(v), -1) for v in key(a)])


def ang(a, b):
    return np.abs(np.angle(np.exp(1j * (a.astype(np.float64) - b.astype(np.float64)))))


def ours(img, noct, upright, max_pts=65536):
    h, w = img.shape
    det = sb.Surfor()
    det.init(noct, 4.0, False, 9, 2, upright, False, 4, w, h, max_pts=max_pts)
    pitch = sb.iAlignUp(w, 128)
    buf = np.zeros((h, pitch), np.uint8)
    buf[:, :w] = img
    data = sb.initSurfData(max_pts, True, True)
    dd = det.detectAndCompute(torch.from_numpy(buf).cuda(), data, (w, h, pitch))
    pts = data.host_points()
    desc = dd[: data.num_pts].cpu().numpy()
    det.close()
    return pts, desc


def main():
    left, _ = load_pair()
    cases = [("synth640", sb.synth_frame(640, 480, 5000), 4), ("pair_left", left, 4), ("synth1080", sb.synth_frame(1920, 1080, 1), 5)]
    out = {}
    for name, img, noct in cases:
        h, w = img.shape
        for upright in (True, False):
            ref = ref_lib.Reference(w, h, noct, 4.0, False, 9, 2, upright, False, 4)
            ref.detect(img)  # discard the first call
            runs = [ref.detect(img) for _ in range(4)]
            ref.close()
            p0, d0 = runs[0]
            spread_ori, spread_l2 = 0.0, 0.0
            for p, d in runs[1:]:
                ix = align(p0, p)
                ok = ix >= 0
                spread_ori = max(spread_ori, float(ang(p0["ori"][ok], p["ori"][ix[ok]]).max()))
                spread_l2 = max(spread_l2, float(np.linalg.norm(d0[ok] - d[ix[ok]], axis=1).max()))
            po, do = ours(img, noct, upright)
            ix = align(p0, po)
            ok = ix >= 0
            dori = ang(p0["ori"][ok], po["ori"][ix[ok]])
            l2 = np.linalg.norm(d0[ok] - do[ix[ok]], axis=1)
            out[f"{name}_{'upright' if upright else 'rotated'}"] = {
                "n_ref": int(len(p0)), "n_ours": int(len(po)), "identical_xy_scale": int(ok.sum()),
                "ref_vs_ref_ori_max": spread_ori, "ref_vs_ref_desc_l2_max": spread_l2,
                "ours_vs_ref_ori_max": float(dori.max()), "ours_vs_ref_ori_p999": float(np.quantile(dori, 0.999)),
                "ours_vs_ref_desc_l2_max": float(l2.max()), "ours_vs_ref_desc_l2_p999": float(np.quantile(l2, 0.999)),
                "ours_vs_ref_desc_rows_over_1e-3": int((l2 > 1e-3).sum()), "ours_vs_ref_desc_rows_over_1e-4": int((l2 > 1e-4).sum()),
            }
            print(name, upright, out[f"{name}_{'upright' if upright else 'rotated'}"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ref_spread.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
