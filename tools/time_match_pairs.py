"""Batched matcher (sb_match_pairs_async) on 32 synthetic 1080p stereo pairs: python tools/time_match_pairs.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_surf_b200 as sb
W, H, NP, MAXP, BOUND = 1920, 1080, 32, 16384, 8192
det = sb.Surfor(); det.init(5, 4.0, False, 9, 2, True, False, 4, W, H, max_pts=MAXP, batch=2 * NP)
pitch = sb.iAlignUp(W, 128)
buf = np.zeros((2 * NP, H, pitch), np.uint8)
for p in range(NP):
    buf[2 * p, :, :W] = sb.synth_frame(W, H, 5000 + p)
    buf[2 * p + 1, :, :W] = sb.synth_frame(W, H, 5000 + p, 12, 2, (5000 + p) ^ 0xA5A5)
d = torch.from_numpy(buf).cuda()
pts = torch.zeros((2 * NP, MAXP * 48), dtype=torch.uint8, device="cuda"); cnt = torch.zeros(2 * NP, dtype=torch.int32, device="cuda")
desc = torch.zeros((2 * NP, MAXP, 64), dtype=torch.float32, device="cuda")
det.detect_batch(d, pitch, pts, cnt, desc)
for _ in range(3): det.match_pairs_async(pts, cnt, desc, NP, BOUND)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5): det.match_pairs_async(pts, cnt, desc, NP, BOUND)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 5 / NP
c = cnt.cpu().numpy()
flop = sum(2.0 * int(c[2 * p]) * (int(c[2 * p + 1]) & ~31) * 64 for p in range(NP)) / NP
print(f"match_pairs {NP} pairs, {c.mean():.0f} keypoints per frame: {us:.2f} us per pair -> {flop / us / 1e6:.1f} TFLOP/s algorithmic")
