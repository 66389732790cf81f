"""Small fixed workload for ncu captures: one batch of 1080p synth frames through the hot path.
usage: python tools/prof_kernels.py [batch] [upright 0/1]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_surf_b200 as sb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
upright = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
W, H = 1920, 1080
det = sb.Surfor()
det.init(5, 4.0, False, 9, 2, upright, False, 4, W, H, max_pts=16384, batch=B)
pitch = sb.iAlignUp(W, 128)
buf = np.zeros((B, H, pitch), np.uint8)
for f in range(B):
    buf[f, :, :W] = sb.synth_frame(W, H, 1 + f)
d = torch.from_numpy(buf).cuda()
pts = torch.zeros((B, 16384 * 48), dtype=torch.uint8, device="cuda")
cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
desc = torch.zeros((B, 16384, det.nfeatures), dtype=torch.float32, device="cuda")
for it in range(3):
    ms = det.detect_batch_profile(d, pitch, pts, cnt, desc)
torch.cuda.synchronize()
print("stage ms", ms, "per frame us", [1e3 * m / B for m in ms], "kp", cnt.cpu().numpy().mean())
