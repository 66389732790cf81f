"""Time the host-buffer path (sb_detect_batch_host) alone: python tools/time_e2e.py [batch]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_surf_b200 as sb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W, H, MAX = 1920, 1080, 16384
det = sb.Surfor(); det.init(5, 4.0, False, 9, 2, True, False, 4, W, H, max_pts=MAX, batch=B)
pool = [sb.synth_frame(W, H, 1 + i) for i in range(16)]
frames = np.stack([np.roll(pool[f % 16], 37 * (f // 16), axis=1) for f in range(B)])
hf = torch.from_numpy(frames).pin_memory()
hp = torch.zeros((B, MAX * 48), dtype=torch.uint8).pin_memory()
hc = torch.zeros(B, dtype=torch.int32).pin_memory()
hd = torch.zeros((B, MAX, 64), dtype=torch.float32).pin_memory()
for _ in range(3): det.detect_batch_host(hf, hp, hc, hd)
torch.cuda.synchronize(); t0 = time.perf_counter(); N = 20
for _ in range(N): det.detect_batch_host(hf, hp, hc, hd)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / N
print(f"chunk={os.environ.get('SB_HOST_CHUNK_removed','default')} batch={B}: {dt*1e3:.3f} ms/step -> {B/dt:.0f} frames/s; kp/frame {hc.numpy().mean():.0f}")
