"""Resident-batch throughput of sb_detect_batch_async alone: python tools/time_resident.py [batch]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_surf_b200 as sb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W, H, MAX = 1920, 1080, 16384
det = sb.Surfor(); det.init(5, 4.0, False, 9, 2, True, False, 4, W, H, max_pts=MAX, batch=B)
pitch = sb.iAlignUp(W, 128)
pool = [sb.synth_frame(W, H, 1 + i) for i in range(16)]
buf = np.zeros((B, H, pitch), np.uint8)
for f in range(B): buf[f, :, :W] = np.roll(pool[f % 16], 37 * (f // 16), axis=1)
d = torch.from_numpy(buf).cuda()
pts = torch.zeros((B, MAX * 48), dtype=torch.uint8, device="cuda"); cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
desc = torch.zeros((B, MAX, 64), dtype=torch.float32, device="cuda")
for _ in range(5): det.detect_batch(d, pitch, pts, cnt, desc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 30; e0.record()
for _ in range(N): det.detect_batch(d, pitch, pts, cnt, desc)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / N
print(f"split={os.environ.get('SB_BATCH_SPLIT','1')} batch={B}: {ms:.3f} ms/step -> {B/ms*1e3:.0f} frames/s, kp {cnt.cpu().numpy().mean():.0f}")
