"""p50 of the synchronous single-frame call (BASELINE configs[1] / [2]): python tools/time_latency.py [w h seed]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_surf_b200 as sb
cases = [(1920, 1080, 1), (3840, 2160, 2)] if len(sys.argv) < 4 else [tuple(int(a) for a in sys.argv[1:4])]
for w, h, seed in cases:
    for upright in (True, False):
        det = sb.Surfor(); det.init(5, 4.0, False, 9, 2, upright, False, 4, w, h, max_pts=32768)
        pitch = sb.iAlignUp(w, 128)
        buf = np.zeros((h, pitch), np.uint8); buf[:, :w] = sb.synth_frame(w, h, seed)
        d = torch.from_numpy(buf).cuda(); data = sb.initSurfData(32768, True, True)
        dd = torch.zeros((32768, 64), dtype=torch.float32, device="cuda")
        ts = []
        for i in range(220):
            torch.cuda.synchronize(); a = time.perf_counter()
            det.detectAndCompute(d, data, (w, h, pitch), desc_out=dd)
            ts.append((time.perf_counter() - a) * 1e3)
        ts = ts[20:]
        print(f"{w}x{h} upright={upright} PDL={os.environ.get('SURFB200_PDL','0')}: p50 {np.percentile(ts,50):.4f} ms p90 {np.percentile(ts,90):.4f} ms, {data.num_pts} keypoints")
        det.close()
