"""Measurements for the BASELINE.json configs that are not the bench.py line (configs[0], [2], [4]) and for the
rotated / doubled variants of configs[1]. Writes one JSON object per line to stdout.
    python tools/measure_configs.py > gpurun_out/configs.jsonl"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_surf_b200 as sb
import ref_lib
from helpers import load_pair

HAVE_REF = ref_lib.available() and ref_lib.lib().ref_device_count() > 0


def upload(img):
    h, w = img.shape
    pitch = sb.iAlignUp(w, 128)
    buf = np.zeros((h, pitch), np.uint8); buf[:, :w] = img
    return torch.from_numpy(buf).cuda(), (w, h, pitch)


def latency(det, d_img, whp, max_pts, nf, iters=200, warm=20):
    data = sb.initSurfData(max_pts, True, True)
    dd = torch.zeros((max_pts, nf), dtype=torch.float32, device="cuda")
    ts = []
    for i in range(warm + iters):
        torch.cuda.synchronize(); a = time.perf_counter()
        det.detectAndCompute(d_img, data, whp, desc_out=dd)
        b = time.perf_counter()
        if i >= warm: ts.append((b - a) * 1e3)
    return float(np.percentile(ts, 50)), float(np.percentile(ts, 90)), data.num_pts


def out(**kw):
    print(json.dumps(kw), flush=True)


# ---- configs[0]: bundled pair, main.cpp flow (4 octaves, thresh 4, upright 64-d): detect both, match
left, right = load_pair()
h, w = left.shape
for upright in (True, False):
    det = sb.Surfor(); det.init(4, 4.0, False, 9, 2, upright, False, 4, w, h, max_pts=10000)
    dl, whp = upload(left); dr, _ = upload(right)
    p50l, p90l, nl = latency(det, dl, whp, 10000, 64)
    p50r, p90r, nr = latency(det, dr, whp, 10000, 64)
    rec = {"config": "configs[0] bundled 1280x960 pair", "upright": upright, "ours_ms_per_pair_p50": p50l + p50r,
           "keypoints": [nl, nr]}
    d1 = sb.initSurfData(10000); f1 = det.detectAndCompute(dl, d1, whp)
    d2 = sb.initSurfData(10000); f2 = det.detectAndCompute(dr, d2, whp)
    ts = []
    for i in range(60):
        a = time.perf_counter(); det.match(d1, d2, f1, f2); ts.append((time.perf_counter() - a) * 1e3)
    rec["ours_match_ms_p50"] = float(np.percentile(ts[10:], 50))
    rec["matches_ambiguity_lt_0.8"] = int(len(det.match_filter(d1, d2, 0.8)))
    if HAVE_REF:
        ref = ref_lib.Reference(w, h, 4, 4.0, False, 9, 2, upright, False, 4)
        ml, _ = ref.time_detect(left, 10000, 10, 100); mr, _ = ref.time_detect(right, 10000, 10, 100)
        rec["ref_ms_per_pair_p50"] = float(np.percentile(ml, 50) + np.percentile(mr, 50))
        mm = ref.time_match(d1.host_points(), f1[: d1.num_pts].cpu().numpy(), d2.host_points(), f2[: d2.num_pts].cpu().numpy(), 5, 50)
        rec["ref_match_ms_p50"] = float(np.percentile(mm, 50))
        ref.close()
    out(**rec); det.close()

# ---- configs[1] variants and configs[2]: single-frame latency
for name, (W, H), seed, kw in [("configs[1] 1080p upright", (1920, 1080), 1, dict(upright=True)),
                               ("configs[1] 1080p rotated", (1920, 1080), 1, dict(upright=False)),
                               ("configs[1] 1080p SURF-128", (1920, 1080), 1, dict(upright=True, extend=True)),
                               ("1080p doubled=true (4 octaves)", (1920, 1080), 1, dict(upright=True, doubled=True, noct=4)),
                               ("configs[2] 4K upright", (3840, 2160), 2, dict(upright=True))]:
    img = sb.synth_frame(W, H, seed)
    noct = kw.get("noct", 5)
    det = sb.Surfor()
    det.init(noct, 4.0, kw.get("doubled", False), 9, 2, kw.get("upright", True), kw.get("extend", False), 4, W, H, max_pts=65536)
    d_img, whp = upload(img)
    p50, p90, n = latency(det, d_img, whp, 65536, det.nfeatures, iters=100, warm=10)
    rec = {"config": name, "ours_ms_p50": p50, "ours_ms_p90": p90, "keypoints": n}
    # stage times of a batch of 8 distinct frames (CUDA events)
    B = 8 if W < 3000 else 4
    detb = sb.Surfor()
    detb.init(noct, 4.0, kw.get("doubled", False), 9, 2, kw.get("upright", True), kw.get("extend", False), 4, W, H, max_pts=65536, batch=B)
    buf = np.zeros((B, H, whp[2]), np.uint8)
    for f in range(B): buf[f, :, :W] = sb.synth_frame(W, H, seed + f)
    d = torch.from_numpy(buf).cuda()
    pts = torch.zeros((B, 65536 * 48), dtype=torch.uint8, device="cuda"); cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
    desc = torch.zeros((B, 65536, detb.nfeatures), dtype=torch.float32, device="cuda")
    for _ in range(3): ms = detb.detect_batch_profile(d, whp[2], pts, cnt, desc)
    rec["stage_us_per_frame_batch"] = {k: 1e3 * v / B for k, v in zip(("integral", "hessian", "nms", "describe"), ms)}
    rec["batch"] = B
    if HAVE_REF and not kw.get("doubled", False):
        ref = ref_lib.Reference(W, H, noct, 4.0, False, 9, 2, kw.get("upright", True), kw.get("extend", False), 4)
        m, nr = ref.time_detect(img, 65536, 5, 40)
        rec["ref_ms_p50"] = float(np.percentile(m, 50)); rec["ref_keypoints"] = int(nr)
        ref.close()
    out(**rec); det.close(); detb.close()

# ---- configs[4]: stereo pairs, detect + describe both, match L->R, ratio test, all on the device
W, H, NP = 1920, 1080, 32
det = sb.Surfor(); det.init(5, 4.0, False, 9, 2, True, False, 4, W, H, max_pts=16384, batch=2 * NP)
pitch = sb.iAlignUp(W, 128)
buf = np.zeros((2 * NP, H, pitch), np.uint8)
for p in range(NP):
    buf[2 * p, :, :W] = sb.synth_frame(W, H, 5000 + p)
    buf[2 * p + 1, :, :W] = sb.synth_frame(W, H, 5000 + p, 12, 2, (5000 + p) ^ 0xA5A5)
d = torch.from_numpy(buf).cuda()
pts = torch.zeros((2 * NP, 16384 * 48), dtype=torch.uint8, device="cuda"); cnt = torch.zeros(2 * NP, dtype=torch.int32, device="cuda")
desc = torch.zeros((2 * NP, 16384, 64), dtype=torch.float32, device="cuda")


class View:  # SurfData-like view of one frame of the batch
    def __init__(self, f, n): self.d_data = pts[f]; self.num_pts = n; self.h_data = None


def step():
    det.detect_batch(d, pitch, pts, cnt, desc)
    counts = cnt.cpu().numpy()  # host gather of the counts (the only host round trip)
    for p in range(NP):
        det.match_async(View(2 * p, int(counts[2 * p])), View(2 * p + 1, int(counts[2 * p + 1])), desc[2 * p], desc[2 * p + 1])
    return counts


for _ in range(3): counts = step()
torch.cuda.synchronize(); t0 = time.perf_counter(); N = 10
for _ in range(N): step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / N
hp = pts[0].cpu().numpy().view(sb.POINT_DTYPE)[: counts[0]]
out(config="configs[4] 1080p stereo pairs: detect+describe both, tcgen05 match L->R on the device", pairs_per_step=NP,
    ms_per_step=dt * 1e3, pairs_per_s=NP / dt, keypoints_per_frame=float(counts.mean()),
    pair0_rows_with_ambiguity_lt_0_8=int(((hp["ambiguity"] < 0.8) & (hp["match"] >= 0)).sum()),
    pair0_rows_matched=int((hp["match"] >= 0).sum()))
