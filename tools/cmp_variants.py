import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import cuda_surf_b200 as sb, oracle_lib as ol, ref_lib
from helpers import keypoint_parity
w, h = 400, 300
img = sb.synth_frame(w, h, 77)
for wsz, upright in [(3, False), (3, True), (2, False), (4, False)]:
    ref = ref_lib.Reference(w, h, 3, 4.0, False, 9, 2, upright, False, wsz)
    rpts, rdesc = ref.detect(img); ref.close()
    orc = ol.Oracle(3, 4.0, False, 9, 2, upright, False, wsz)
    opts, odesc = orc.detect_and_compute(img)
    det = sb.Surfor(); det.init(3, 4.0, False, 9, 2, upright, False, wsz, w, h, max_pts=32768)
    pitch = sb.iAlignUp(w, 128); buf = np.zeros((h, pitch), np.uint8); buf[:, :w] = img
    d = torch.from_numpy(buf).cuda(); data = sb.initSurfData(32768, True, True)
    dd = det.detectAndCompute(d, data, (w, h, pitch)); gpts = data.host_points(); gdesc = dd[:data.num_pts].cpu().numpy()
    def cmp(a, ad, b, bd, name):
        fr, fg, ok, idx, *_ = keypoint_parity(a, b)
        l2 = np.linalg.norm(bd[idx[ok]] - ad[ok], axis=1)
        dori = np.abs(np.angle(np.exp(1j * (b["ori"][idx[ok]] - a["ori"][ok]))))
        print(f"wsz={wsz} upright={upright} {name}: n={len(a)}/{len(b)} match {fr:.3f}/{fg:.3f} desc L2 max {l2.max():.2e} median {np.median(l2):.2e} ori max {dori.max():.2e}")
    cmp(rpts, rdesc, opts, odesc, "ref vs oracle")
    cmp(rpts, rdesc, gpts, gdesc, "ref vs ours  ")
    det.close()
