"""Raw pinned host<->device copy ceiling of the box at N ranks, with the byte counts of one bench.py step per GPU
(H2D 132.7 MB of frames, D2H 98.1 MB of keypoints + descriptors): what `e2e` can reach at most when every rank drives
its own GPU. One process per GPU as in bench.py:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/host_copy_ceiling.py
Legs: H2D only, D2H only, both at once (two streams; PCIe is full duplex), each as ONE cudaMemcpyAsync per step and as the
64 per-frame copies of a streaming caller. Time = max over ranks (CUDA events), rank 0 prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

H2D = 64 * 1080 * 1920
D2H = 98116864
STEPS = 12


def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h_in = torch.empty(H2D, dtype=torch.uint8).pin_memory(); h_in.fill_(7)
    h_out = torch.empty(D2H, dtype=torch.uint8).pin_memory(); h_out.fill_(1)
    d_in = torch.empty(H2D, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(D2H, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def leg(up, down, pieces):
        def once():
            if up:
                with torch.cuda.stream(s1):
                    for a, b in zip(d_in.chunk(pieces), h_in.chunk(pieces)):
                        a.copy_(b, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    for a, b in zip(h_out.chunk(pieces), d_out.chunk(pieces)):
                        a.copy_(b, non_blocking=True)
        for _ in range(3):
            once()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0); s2.wait_event(e0)
        for _ in range(STEPS):
            once()
        cur = torch.cuda.current_stream()
        cur.wait_stream(s1); cur.wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / STEPS], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        nbytes = (H2D if up else 0) + (D2H if down else 0)
        return {"ms_per_step": round(ms, 3), "gbs_per_gpu": round(nbytes / ms / 1e6, 2), "gbs_total": round(world * nbytes / ms / 1e6, 2),
                "frames_per_s_ceiling": round(world * 64 / ms * 1e3)}

    out = {"what": "pinned host<->device copy ceiling, bytes of one 64-frame bench step per GPU", "n_gpus": world, "steps": STEPS,
           "h2d_bytes": H2D, "d2h_bytes": D2H, "cpus": os.cpu_count()}
    for pieces in (1, 64):
        out[f"h2d_only_x{pieces}"] = leg(True, False, pieces)
        out[f"d2h_only_x{pieces}"] = leg(False, True, pieces)
        out[f"both_x{pieces}"] = leg(True, True, pieces)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
