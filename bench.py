#!/usr/bin/env python
"""bench.py -- frames/s of SURF detect+describe on synthetic 1080p frames (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (integral -> Hessian -> NMS+refine -> descriptors) over one batch
of B distinct synthetic frames per GPU. Workload = BASELINE.json configs[1] (1920x1080 `synth_v1`
frames, 5 octaves x 5 layers -- the reference's max_scale for lobe 3, SURVEY.md 2.4-1 -- thresh 4,
upright 64-d descriptors), batched as configs[3] shards it: frames are independent, ranks own
contiguous shards, no data-path collective (scaling "weak": B frames per GPU).

  value  whole-job frames/s with the batch resident in HBM, CUDA events on the launching stream,
         barrier + synchronize on both sides, max over ranks.
  e2e    the same through the host-buffer C-ABI call (sb_detect_batch_host): H2D of the frames and
         D2H of counts + keypoints + descriptors inside the timed region.
  roofline / cpu_baseline: see DESIGN.md "Measurement".
--impl reference times the reference's own implementation (oracle/_ref, the unmodified CUDA sources
built for sm_100a, through Surfor::detectAndCompute as main.cpp:239-245 calls it); if that library is
absent, the CPU port under oracle/ on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
NOCT, THRESH = 5, 4.0
MAX_PTS = 32768  # BASELINE configs[1]/[2]: keypoint buffer of 32768 per frame
POOL = 16  # distinct generated frames; the batch is filled with horizontally rolled copies


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def make_frames(n, synth):
    """n frames of the workload: POOL distinct synth_v1 frames (seeds 1..POOL), then horizontally rolled copies. `synth` is
    the generator: sb.synth_frame in our arm, the bit-identical numpy port tests/synth_np.py in the reference arm (which
    must not load the product library)."""
    pool = [synth(W, H, 1 + i) for i in range(min(POOL, n))]
    out = np.empty((n, H, W), np.uint8)
    for f in range(n):
        out[f] = np.roll(pool[f % len(pool)], 37 * (f // len(pool)), axis=1)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:  # first sample before the timed region starts
                time.sleep(0.01)
            self.rows.clear()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_config(B, world):
    """`config` of the JSON line -- identical in both arms (the reference arm processes the same B frames per step, one
    synchronous Surfor::detectAndCompute at a time: it has no batched entry point)."""
    return {"workload": "BASELINE configs[1]: synth_v1 1920x1080 (seeds 1..16, rolled copies), 5 octaves x 5 layers, thresh 4, "
                        f"upright 64-d, max_pts {MAX_PTS}; {B} distinct frames per GPU per step (configs[3] sharding)",
            "frames_per_step_per_gpu": B, "max_pts": MAX_PTS,
            "l2_policy": f"inputs+intermediates per step {B * (W * H + 23.6e6) / 1e6:.0f} MB > 126 MB L2",
            "parallelism": f"frames sharded over {world} GPU(s), no collective"}


def algorithmic_bytes(info, noct, n_kp, nfeat, w=None, h=None):
    """compulsory traffic of a staged pipeline, per frame (unpadded sizes)."""
    w, h = w or W, h or H
    b_img = w * h
    b_int = 4 * (w + 1) * (h + 1)
    b_resp = 4 * info.max_scale * sum(info.sw[o] * info.sh[o] for o in range(noct))
    b_kp, b_desc = 48 * n_kp, 4 * nfeat * n_kp
    return {"integral": b_img + b_int, "hessian": b_int + b_resp, "nms": b_resp + b_kp,
            "describe": b_int + b_kp + b_desc}


def bind_to_gpu_numa_node(torch, local_rank):
    """Run this rank (and first-touch its pinned buffers) on the NUMA node its GPU hangs off: with 8 ranks the host
    legs of `e2e` otherwise cross the socket interconnect. Best effort; returns the node or None."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def run_ours(args, rank, world, local_rank, dist):
    import torch
    import cuda_surf_b200 as sb
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    dev = torch.device("cuda", local_rank)
    B = args.batch
    det = sb.Surfor()
    det.init(NOCT, THRESH, False, 9, 2, True, False, 4, W, H, max_pts=MAX_PTS, batch=B, device=local_rank)
    nf = det.nfeatures
    # this rank's shard of the global batch (contiguous; frames independent)
    lo, hi = sb.shard_range(B * world, world, rank)
    frames = make_frames(B, sb.synth_frame) if world == 1 else np.roll(make_frames(B, sb.synth_frame), 11 * rank, axis=2)
    pitch = sb.iAlignUp(W, 128)
    h_pad = np.zeros((B, H, pitch), np.uint8)
    h_pad[:, :, :W] = frames
    d_imgs = torch.from_numpy(h_pad).to(dev)
    pts = torch.zeros((B, MAX_PTS * 48), dtype=torch.uint8, device=dev)
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    desc = torch.zeros((B, MAX_PTS, nf), dtype=torch.float32, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    for _ in range(args.warmup):
        det.detect_batch(d_imgs, pitch, pts, cnt, desc)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        det.detect_batch(d_imgs, pitch, pts, cnt, desc)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    counts = cnt.cpu().numpy()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = B * world * args.steps / (ms_max / 1e3)

    # ---- per-stage times (CUDA events at the stage boundaries, same stream) -> roofline of the top kernel
    stage = np.zeros(4)
    for _ in range(args.steps):
        stage += np.array(det.detect_batch_profile(d_imgs, pitch, pts, cnt, desc))
    stage /= args.steps
    names = ["integral", "hessian", "nms", "describe"]
    kp_mean = float(counts.mean())
    ab = algorithmic_bytes(det.info, NOCT, kp_mean, nf)
    peak, peak_kind = peaks()
    top = int(np.argmax(stage))
    ach = ab[names[top]] * B / (stage[top] / 1e3) / 1e9
    per_stage = {n: {"ms": float(stage[i]), "gbs": ab[n] * B / (stage[i] / 1e3) / 1e9,
                     "frac": ab[n] * B / (stage[i] / 1e3) / 1e9 / peak} for i, n in enumerate(names)}
    # DRAM bytes of the dominant kernel per launch: NOT measured in this run (ncu cannot run inside a timed bench); read from
    # the committed `ncu --set full` capture of the same kernel at the same batch size and labelled as such
    traffic, traffic_src = None, None
    for tf in ("r2_traffic.json", "r1_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", tf)))
            kmap = tj.get("stage_kernel", {"describe": "describe_upright_kernel<4", "nms": "nms_scan_tile_kernel", "hessian": "hessian_o0_kernel"})
            hit = [k for k in tj["kernels"] if names[top] in kmap and k.startswith(kmap[names[top]])]
            if tj.get("batch") == B and hit:
                traffic = tj["kernels"][hit[0]]["traffic_bytes"]
                traffic_src = f"committed capture profiles/{tf} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"
                break
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": names[top], "achieved": ach, "peak": peak, "peak_kind": peak_kind,
                "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": ab[names[top]] * B, "stages": per_stage,
                "pipeline_frac": sum(ab.values()) * B / (float(stage.sum()) / 1e3) / 1e9 / peak}

    # ---- end to end through the host-buffer C-ABI call (pinned host memory both ways)
    h_frames = torch.from_numpy(frames).pin_memory()
    h_pts = torch.zeros((B, MAX_PTS * 48), dtype=torch.uint8).pin_memory()
    h_cnt = torch.zeros(B, dtype=torch.int32).pin_memory()
    h_desc = torch.zeros((B, MAX_PTS, nf), dtype=torch.float32).pin_memory()
    # steady state of a streaming caller: batch k+1 is submitted (uploads + kernels enqueued) before batch k's
    # results are waited for, so only the copies themselves -- all inside the timed region -- bound the rate
    out = [(h_pts, h_cnt, h_desc),
           (torch.zeros_like(h_pts).pin_memory(), torch.zeros_like(h_cnt).pin_memory(), torch.zeros_like(h_desc).pin_memory())]

    def e2e_steps(k, with_desc=True):
        tickets = []
        for i in range(k):
            tickets.append(det.submit_batch_host(h_frames))
            if i >= 2:  # two batches submitted ahead: k downloads, k+1 computes, k+2 uploads
                det.wait_batch_host(tickets[i - 2], *(out[i & 1] if with_desc else out[i & 1][:2]))
        for i in range(max(0, k - 2), k):
            det.wait_batch_host(tickets[i], *(out[i & 1] if with_desc else out[i & 1][:2]))

    e2e_steps(max(3, args.warmup // 2))
    barrier()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    te = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = B * world * args.steps / float(te.item())
    nk = int(h_cnt.sum().item())
    # the same batches through the synchronous call (one batch in flight: its first upload and last download are exposed)
    ns = max(3, args.steps // 4)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    for _ in range(ns):
        det.detect_batch_host(h_frames, h_pts, h_cnt, h_desc)
    t3 = time.perf_counter()
    ts = torch.tensor([(t3 - t2) / ns], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    # The reference's own contract (surf.cpp:335-342): keypoints come back to the host, descriptors are computed and STAY on
    # the device (its *desc_addr is a device pointer that Surfor::match consumes there). Same streaming call, the wait
    # without a descriptor destination: 0.24 MB instead of 1.5 MB per frame travel device -> host.
    e2e_steps(3, with_desc=False)
    barrier()
    t4 = time.perf_counter()
    e2e_steps(args.steps, with_desc=False)
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    tk = torch.tensor([t5 - t4], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tk, op=dist.ReduceOp.MAX)
    e2e_kp = {"value": B * world * args.steps / float(tk.item()), "unit": "frames/s", "h2d_bytes_per_step": int(B * W * H),
              "d2h_bytes_per_step": int(4 * B + B * int(h_cnt.max().item()) * 48),
              "what": "descriptors computed and left on the device, as Surfor::detectAndCompute returns them (surfd.cu:3262-3266); "
                      "counts + keypoints to pinned host memory"}
    e2e = {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": int(B * W * H),
           # counts + one strided copy per array, as wide as the batch's largest count (what sb_wait_batch_host moves)
           "d2h_bytes_per_step": int(4 * B + B * int(h_cnt.max().item()) * (48 + nf * 4)),
           "d2h_result_bytes_per_step": int(4 * B + nk * 48 + nk * nf * 4),
           "api": "sb_submit_batch_host / sb_wait_batch_host, three batches in flight, pinned host buffers",
           "synchronous_call_value": B * world / float(ts.item())}

    # ---- single-frame latency through Surfor::detectAndCompute (BASELINE: p50 ms/frame)
    lat = None
    if rank == 0:
        one = sb.Surfor()
        one.init(NOCT, THRESH, False, 9, 2, True, False, 4, W, H, max_pts=MAX_PTS, batch=1, device=local_rank)
        data = sb.initSurfData(MAX_PTS, True, True, device=local_rank)
        dd = torch.zeros((MAX_PTS, nf), dtype=torch.float32, device=dev)
        ts = []
        for i in range(20 + 100):
            torch.cuda.synchronize()
            a = time.perf_counter()
            one.detectAndCompute(d_imgs[0], data, (W, H, pitch), desc_out=dd)
            b = time.perf_counter()
            if i >= 20:
                ts.append((b - a) * 1e3)
        lat = {"p50_ms": float(np.percentile(ts, 50)), "p90_ms": float(np.percentile(ts, 90)), "keypoints": data.num_pts}
        one.close()

    # ---- configs[3] as written: 1 024 frames in total, STRONG-scaled over the ranks (contiguous shards, batches of B)
    TOTAL = 1024
    lo3, hi3 = sb.shard_range(TOTAL, world, rank)
    n_local = hi3 - lo3

    def strong_pass():
        done = 0
        while done < n_local:
            nb = min(B, n_local - done)
            det.detect_batch(d_imgs[:nb], pitch, pts[:nb], cnt[:nb], desc[:nb])
            done += nb

    strong_pass()
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    strong_pass()
    s1.record()
    barrier()
    tst = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tst, op=dist.ReduceOp.MAX)
    strong = {"workload": f"BASELINE configs[3]: {TOTAL} 1080p frames in total, contiguous shards over {world} GPU(s), batches of {B}",
              "scaling": "strong", "value": TOTAL / (float(tst.item()) / 1e3), "unit": "frames/s", "ms_total": float(tst.item()),
              "frames_per_gpu": n_local}

    # ---- configs[2] (4K) and configs[4] (stereo pairs + matching): rank 0 at N=1, extra keys of the same line
    cfg2 = cfg4 = None
    if rank == 0 and world == 1 and not args.no_extra:
        cfg2 = measure_4k(sb, torch, dev, local_rank)
        cfg4 = measure_stereo(sb, torch, det, dev, pitch, pts, cnt, desc, B)

    # ---- CPU baseline (oracle port) on rank 0 at N=1 only: checker code, timed beside, never shipped
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol
        cores = os.cpu_count() or 1
        n_s = int(min(max(2 * cores, 8), B))
        orc = ol.Oracle(NOCT, THRESH, False, 9, 2, True, False, 4)
        sample = np.ascontiguousarray(frames[np.arange(n_s) % B])
        orc.time_frames(sample[: min(cores, n_s)], MAX_PTS, cores)  # warm the pages and the OpenMP team
        secs, done = 0.0, 0
        while secs < 12.0 and done < 200 * n_s:  # bounded sample: ~12 s of wall time on all host cores
            s_, _ = orc.time_frames(sample, MAX_PTS, cores)
            secs += s_
            done += n_s
        cpu = {"value": done / secs, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{done} frames ({done // n_s} passes over {n_s} of the same 1080p frames), one frame per OpenMP "
                         f"worker on {cores} threads, {secs:.1f} s"}

    if rank == 0:
        kpf = det.info.kernels_per_frame
        line = {"metric": "frames/s SURF detect+describe @1080p", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f32",
                "data": "synthetic", "config": workload_config(B, world),
                "keypoints_per_frame": kp_mean, "host_numa_binding": numa,
                "e2e": e2e, "e2e_descriptors_stay_on_device": e2e_kp, "host_copy_ceiling": host_ceiling(world),
                "gpu_launches": kpf * args.steps, "clocks": clocks, "roofline": roofline,
                "cpu_baseline": cpu, "latency": lat, "strong_scaling_configs3": strong, "configs2_4k": cfg2,
                "configs4_stereo": cfg4, "impl": "ours"}
        emit(json.dumps(line))
    det.close()


def host_ceiling(world):
    """Raw pinned-copy ceiling of the 8-GPU box at this rank count, from the committed measurement (not taken in this run):
    what `e2e` can reach at most when every rank moves one step's bytes both ways."""
    try:
        hj = json.load(open(os.path.join(ROOT, "profiles", "r2_host_copy_ceiling.json")))["runs"][str(world)]
        return {"frames_per_s_both_directions": hj["both_x1"]["frames_per_s_ceiling"], "h2d_gbs_total": hj["h2d_only_x1"]["gbs_total"],
                "d2h_gbs_total": hj["d2h_only_x1"]["gbs_total"], "both_gbs_total": hj["both_x1"]["gbs_total"],
                "source": "committed capture profiles/r2_host_copy_ceiling.json (tools/host_copy_ceiling.py on the 8 x B200 box)"}
    except Exception:
        return None


def measure_4k(sb, torch, dev, local_rank):
    """BASELINE configs[2]: one synthetic 3840x2160 frame (seed 2), 5 octaves, keypoint buffer capped at MAX_PTS: p50
    latency of the synchronous call, and per-stage achieved GB/s (algorithmic bytes of SURVEY.md 8d) on a 4-frame batch."""
    w4, h4, B4 = 3840, 2160, 4
    pitch = sb.iAlignUp(w4, 128)
    buf = np.zeros((B4, h4, pitch), np.uint8)
    for f in range(B4):
        buf[f, :, :w4] = sb.synth_frame(w4, h4, 2 + f)
    d = torch.from_numpy(buf).to(dev)
    one = sb.Surfor()
    one.init(NOCT, THRESH, False, 9, 2, True, False, 4, w4, h4, max_pts=MAX_PTS, batch=B4, device=local_rank)
    data = sb.initSurfData(MAX_PTS, True, True, device=local_rank)
    dd = torch.zeros((MAX_PTS, one.nfeatures), dtype=torch.float32, device=dev)
    ts = []
    for i in range(10 + 60):
        torch.cuda.synchronize()
        a = time.perf_counter()
        one.detectAndCompute(d[0], data, (w4, h4, pitch), desc_out=dd)
        if i >= 10:
            ts.append((time.perf_counter() - a) * 1e3)
    pts = torch.zeros((B4, MAX_PTS * 48), dtype=torch.uint8, device=dev)
    cnt = torch.zeros(B4, dtype=torch.int32, device=dev)
    desc = torch.zeros((B4, MAX_PTS, one.nfeatures), dtype=torch.float32, device=dev)
    stage = np.zeros(4)
    reps = 5
    for i in range(2 + reps):
        ms = np.array(one.detect_batch_profile(d, pitch, pts, cnt, desc))
        if i >= 2:
            stage += ms
    stage /= reps
    kp = float(cnt.cpu().numpy().mean())
    ab = algorithmic_bytes(one.info, NOCT, kp, one.nfeatures, w4, h4)
    peak, _ = peaks()
    names = ["integral", "hessian", "nms", "describe"]
    out = {"workload": f"BASELINE configs[2]: synth_v1 3840x2160 seed 2, 5 octaves, max_pts {MAX_PTS}",
           "p50_ms": float(np.percentile(ts, 50)), "p90_ms": float(np.percentile(ts, 90)), "keypoints": int(data.num_pts),
           "batch_for_stage_times": B4, "frames_per_s_batch": B4 / (float(stage.sum()) / 1e3),
           "stages": {n: {"us_per_frame": 1e3 * float(stage[i]) / B4, "gbs": ab[n] * B4 / (float(stage[i]) / 1e3) / 1e9,
                          "frac_of_hbm_peak": ab[n] * B4 / (float(stage[i]) / 1e3) / 1e9 / peak} for i, n in enumerate(names)}}
    one.close()
    return out


def measure_stereo(sb, torch, det, dev, pitch, pts, cnt, desc, B):
    """BASELINE configs[4]: 1080p stereo pairs (left seed 5000+p; right = the same texture 12 px to the side + noise):
    detect + describe both frames, match L->R with the tcgen05 matcher on the device, accept `ambiguity < 0.8`."""
    NP = B // 2
    buf = np.zeros((2 * NP, H, pitch), np.uint8)
    for p in range(NP):
        buf[2 * p, :, :W] = sb.synth_frame(W, H, 5000 + p)
        buf[2 * p + 1, :, :W] = sb.synth_frame(W, H, 5000 + p, 12, 2, (5000 + p) ^ 0xA5A5)
    d = torch.from_numpy(buf).to(dev)

    class View:  # SurfData-like view of one frame of the batch
        def __init__(self, f, n):
            self.d_data, self.num_pts, self.h_data = pts[f], n, None

    BOUND = 8192  # keypoints per frame that take part in the batched matching (4.9 k per frame here); sizes its scratch

    def step_per_pair():  # round 1: counts to the host, then three launches per pair
        det.detect_batch(d, pitch, pts[: 2 * NP], cnt[: 2 * NP], desc[: 2 * NP])
        counts = cnt[: 2 * NP].cpu().numpy()
        for p in range(NP):
            det.match_async(View(2 * p, int(counts[2 * p])), View(2 * p + 1, int(counts[2 * p + 1])), desc[2 * p], desc[2 * p + 1])
        return counts

    def step():  # all pairs in one launch sequence, counts read on the device: no host round trip in the step
        det.detect_batch(d, pitch, pts[: 2 * NP], cnt[: 2 * NP], desc[: 2 * NP])
        det.match_pairs_async(pts[: 2 * NP], cnt[: 2 * NP], desc[: 2 * NP], NP, BOUND)

    N = 10
    for _ in range(3):
        counts = step_per_pair()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N):
        step_per_pair()
    torch.cuda.synchronize()
    dt_pp = (time.perf_counter() - t0) / N
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / N
    assert int(counts.max()) <= BOUND
    # the batched matcher alone (CUDA events)
    eb0, eb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eb0.record()
    for _ in range(5):
        det.match_pairs_async(pts[: 2 * NP], cnt[: 2 * NP], desc[: 2 * NP], NP, BOUND)
    eb1.record()
    torch.cuda.synchronize()
    match_pairs_us = eb0.elapsed_time(eb1) * 1e3 / 5 / NP
    # the matcher alone on pair 0 (CUDA events on the launching stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    v0, v1 = View(0, int(counts[0])), View(1, int(counts[1]))
    for _ in range(3):
        det.match_async(v0, v1, desc[0], desc[1])
    e0.record()
    for _ in range(20):
        det.match_async(v0, v1, desc[0], desc[1])
    e1.record()
    torch.cuda.synchronize()
    match_us = e0.elapsed_time(e1) * 1e3 / 20
    hp = pts[0].cpu().numpy().view(sb.POINT_DTYPE)[: counts[0]]
    tensor = None
    try:
        tensor = json.load(open(os.path.join(ROOT, "profiles", "r2_match.json")))
    except Exception:
        pass
    n1, n2 = int(counts[0]), int(counts[1])
    peak_tf = 1373.9
    try:
        peak_tf = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    except Exception:
        pass
    return {"workload": f"BASELINE configs[4]: {NP} 1080p stereo pairs per step (of 512), detect+describe both, tcgen05 match L->R, ratio 0.8",
            "pairs_per_step": NP, "ms_per_step": dt * 1e3, "pairs_per_s": NP / dt, "keypoints_per_frame": float(counts.mean()),
            "api": f"sb_detect_batch_async + sb_match_pairs_async (bound {BOUND}), no host round trip inside the step",
            "pairs_per_s_match_per_pair": NP / dt_pp, "match_us_per_pair_batched": match_pairs_us,
            "match_tflops_algorithmic_batched": 2.0 * int(counts[0]) * (int(counts[1]) - int(counts[1]) % 32) * 64 / (match_pairs_us * 1e-6) / 1e12,
            "match_us_pair0": match_us, "match_shape": [n1, n2 - n2 % 32, 64],
            "match_tflops_algorithmic": 2.0 * n1 * (n2 - n2 % 32) * 64 / (match_us * 1e-6) / 1e12,
            "match_frac_of_bf16_sustained_peak": 2.0 * n1 * (n2 - n2 % 32) * 64 / (match_us * 1e-6) / 1e12 / peak_tf,
            "match_mma_tensor_pipe_pct": None if tensor is None else tensor.get("match_mma_tensor_pipe_pct"),
            "match_mma_tensor_pipe_source": None if tensor is None else "committed capture profiles/r2_match.json (ncu sm__pipe_tensor_cycles_active)",
            "pair0_rows_matched": int((hp["match"] >= 0).sum()),
            "pair0_rows_with_ambiguity_lt_0_8": int(((hp["ambiguity"] < 0.8) & (hp["match"] >= 0)).sum())}


def run_reference(args, rank, world):
    """The reference's own implementation on the same workload, rank 0 only. This arm never imports the product package:
    the frames come from the numpy port of the generator (tests/synth_np.py, bit-identical), the timed path is the
    unmodified surfd.cu + surf.cpp of oracle/_ref driven through Surfor::detectAndCompute."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth_np
    B = args.batch
    frames = make_frames(B, synth_np.synth_frame)
    cfg = workload_config(B, world)
    import ref_lib
    have_gpu = False
    if ref_lib.available():
        try:
            have_gpu = ref_lib.lib().ref_device_count() > 0
        except OSError:
            have_gpu = False
    if have_gpu:
        ref = ref_lib.Reference(W, H, NOCT, THRESH, False, 9, 2, True, False, 4)
        per = []   # ms per frame, device-resident (as main.cpp:239-245)
        per_e = []  # with H2D of the frame and D2H of points + descriptors
        for step in range(args.warmup + args.steps):
            ms_d = sum(float(ref.time_detect(f, MAX_PTS, 0, 1)[0][0]) for f in frames)
            ms_e, ref_kp = 0.0, 0
            for f in frames:
                ms1, n1 = ref.time_detect_e2e(f, MAX_PTS, 0, 1)
                ms_e += float(ms1[0]); ref_kp += int(n1)
            if step >= args.warmup:
                per.append(ms_d); per_e.append(ms_e)
        ref.close()
        ms_step = float(np.mean(per))
        value = len(frames) / (ms_step / 1e3)
        e2e_v = len(frames) / (float(np.mean(per_e)) / 1e3)
        line = {"metric": "frames/s SURF detect+describe @1080p", "value": value, "unit": "frames/s", "n_gpus": 1,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32+f32", "data": "synthetic", "config": cfg,
                "impl": "reference",
                "cpu_baseline": {"value": value, "unit": "frames/s", "cores": 1, "kind": "reference",
                                 "sample": f"{len(frames)} frames per step, one synchronous call each, through Surfor::detectAndCompute of the unmodified "
                                           "reference built for sm_100a (oracle/_ref); it is a CUDA program, so it runs on "
                                           "GPU 0 driven by one host thread"},
                "e2e": {"value": e2e_v, "unit": "frames/s", "h2d_bytes_per_step": int(len(frames) * W * H),
                        "d2h_bytes_per_step": int(ref_kp * (48 + 4 * 64))}}
        emit(json.dumps(line))
        return
    # no reference library / no GPU: the CPU port of oracle/ on all host cores
    import oracle_lib as ol
    cores = os.cpu_count() or 1
    orc = ol.Oracle(NOCT, THRESH, False, 9, 2, True, False, 4)
    n_s = int(min(max(cores, 4), 64))
    sample = np.ascontiguousarray(frames[np.arange(n_s) % len(frames)])
    secs = []
    for step in range(min(args.warmup, 1) + args.steps):
        s, _ = orc.time_frames(sample, MAX_PTS, cores)
        if step >= min(args.warmup, 1):
            secs.append(s)
    v = n_s / float(np.mean(secs))
    line = {"metric": "frames/s SURF detect+describe @1080p", "value": v, "unit": "frames/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f32", "data": "synthetic",
            "config": cfg, "impl": "reference",
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{n_s} frames per step, one per OpenMP worker"},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(json.dumps(line))


_REAL_STDOUT = None


def emit(text):
    """The JSON line goes to the process's real stdout; everything else written to fd 1 meanwhile (e.g. the
    'NCCL version' banner of the first communicator) has been sent to stderr."""
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(text + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[2] (4K) and configs[4] (stereo) legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    try:
        run_ours(args, rank, world, local_rank, dist)
    finally:
        if dist is not None:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
