"""Time the matcher (ours vs the reference) on real descriptors of the bundled pair."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_surf_b200 as sb
import ref_lib
from helpers import load_pair
left, right = load_pair()
h, w = left.shape
det = sb.Surfor(); det.init(4, 4.0, False, 9, 2, True, False, 4, w, h, max_pts=10000)
pitch = sb.iAlignUp(w, 128)
def run(img):
    buf = np.zeros((h, pitch), np.uint8); buf[:, :w] = img
    d = torch.from_numpy(buf).cuda(); data = sb.initSurfData(10000); desc = det.detectAndCompute(d, data, (w, h, pitch)); return data, desc
d1, f1 = run(left); d2, f2 = run(right)
for _ in range(5): det.match(d1, d2, f1, f2)
torch.cuda.synchronize(); ts = []
for _ in range(50):
    a = time.perf_counter(); det.match(d1, d2, f1, f2); ts.append((time.perf_counter() - a) * 1e3)
print(f"ours  match {d1.num_pts}x{d2.num_pts}: p50 {np.percentile(ts,50):.3f} ms (sync call incl. D2H of 5 fields)")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): det.match_async(d1, d2, f1, f2)
torch.cuda.synchronize(); e0.record()
for _ in range(50): det.match_async(d1, d2, f1, f2)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 50 * 1e3
flop = 2.0 * d1.num_pts * (d2.num_pts - d2.num_pts % 32) * 64
print(f"ours  match kernels only (CUDA events, 4 launches): {us:.1f} us -> {flop / us / 1e6:.2f} TFLOP/s algorithmic (fp32-equivalent 2*N1*N2*64)")
# the same three launches as ONE CUDA graph replayed back to back: device time without the host's launch rate
g = torch.cuda.CUDAGraph()
cs = torch.cuda.Stream()
with torch.cuda.stream(cs):
    det.match_async(d1, d2, f1, f2)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=cs):
        det.match_async(d1, d2, f1, f2)
    for _ in range(3): g.replay()
    torch.cuda.synchronize(); e0.record(cs)
    for _ in range(50): g.replay()
    e1.record(cs); torch.cuda.synchronize()
usg = e0.elapsed_time(e1) / 50 * 1e3
print(f"ours  match as a CUDA graph (prep + mma + final, device time): {usg:.1f} us -> {flop / usg / 1e6:.2f} TFLOP/s algorithmic = "
      f"{100 * flop / usg / 1e6 / 1373.9:.2f} % of the sustained bf16 peak")
if ref_lib.available():
    ref = ref_lib.Reference(w, h, 4)
    ms = ref.time_match(d1.host_points(), f1[:d1.num_pts].cpu().numpy(), d2.host_points(), f2[:d2.num_pts].cpu().numpy(), 5, 50)
    print(f"ref   match: p50 {np.percentile(ms,50):.3f} ms")
    want = ref.match(d1.host_points(), f1[:d1.num_pts].cpu().numpy(), d2.host_points(), f2[:d2.num_pts].cpu().numpy())
    det.match(d1, d2, f1, f2); got = d1.host_points()
    print("index equal:", np.array_equal(got["match"], want["match"]), "score equal:", np.array_equal(got["score"], want["score"]),
          "ambiguity max diff:", float(np.abs(got["ambiguity"] - want["ambiguity"]).max()))
