"""Turn the outputs of tools/gpu_round.sh <tag> (+ tools/measure_configs.py) under gpurun_out/ into the committed
per-round artefacts under profiles/ (files are prefixed with the tag's round: r2x -> r2_).
usage: python tools/refresh_profiles.py <tag> [configs.jsonl]"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def short(name):
    return name.split("(")[0].replace("void ", "")


def read_ncu_csv(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    return rows[hi], rows[hi + 1:]


def main():
    tag = sys.argv[1]
    R = tag[:2]  # 'r1', 'r2', ...
    cfg = sys.argv[2] if len(sys.argv) > 2 else None
    shutil.copy(os.path.join(G, f"{tag}_bench_ours.json"), os.path.join(P, f"{R}_bench_n1.json"))
    shutil.copy(os.path.join(G, f"{tag}_bench_ref.json"), os.path.join(P, f"{R}_bench_ref.json"))
    shutil.copy(os.path.join(G, f"{tag}_launches.csv"), os.path.join(P, f"{R}_launches_final.csv"))
    if cfg:
        shutil.copy(cfg, os.path.join(P, f"{R}_configs.jsonl"))
    bench = json.load(open(os.path.join(P, f"{R}_bench_n1.json")))

    # ---- launch list: the 64-frame launches of the timed steps
    hdr, data = read_ncu_csv(os.path.join(G, f"{tag}_launches.csv"))
    kn, gs, mv, mu = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] == "ns" else v * 1e3 if r[mu] == "ms" else v
        agg.setdefault((short(r[kn]), r[gs]), []).append(v)
    big = [(k, g, v) for (k, g), v in agg.items() if k.startswith("sb::") and g.strip("()").replace(" ", "").split(",")[1:] and
           "64" in g.strip("()").replace(" ", "").split(",")[1:] and "clamp" not in k]
    tot = sum(sum(v) / len(v) for _, _, v in big)
    st = bench["roofline"]["stages"]
    ms = {k: v["ms"] * 1e3 for k, v in st.items()}
    ssum = sum(ms.values())
    fam = {"integral": 0.0, "hessian": 0.0, "nms": 0.0, "describe": 0.0}
    out = ["ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 3 --no-cpu (final kernels of the round)",
           "per-launch times are cold-cache and serialised; 64-frame batch launches and their share of one step:", ""]
    for k, g, v in big:
        m = sum(v) / len(v)
        out.append(f"{k:<45s} grid {g:<18s} n={len(v):3d} mean {m:9.2f} us  share {100 * m / tot:5.1f}%")
        for f_ in fam:
            if f_ in k:
                fam[f_] += m
    out += ["", f"sum of one step's kernels: {tot:.1f} us = {tot / 64:.2f} us per frame (bench.py, warm and back to back: {bench['ms_per_step'] * 1e3:.0f} us per step)",
            f"bench.py stage times (CUDA events, the plain run of the same command): integral {ms['integral']:.0f} us, Hessian {ms['hessian']:.0f} us, NMS {ms['nms']:.0f} us, describe {ms['describe']:.0f} us per 64 frames:",
            "stage shares, bench.py / launch list: " + ", ".join(f"{n} {100 * ms[n] / ssum:.1f} % / {100 * fam[n] / tot:.1f} %" for n in ("describe", "hessian", "nms", "integral")),
            "", "(the first 200 launches of the run: the timed 64-frame batches above, then the ramped 2/4/8/16-frame chunks of the host-buffer legs and clamp_counts -- see the round's launches_final.csv)"]
    open(os.path.join(P, f"{R}_launches_final.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))

    # ---- dram traffic per 64-frame launch
    hdr, data = read_ncu_csv(os.path.join(G, f"{tag}_traffic.csv"))
    kn, mn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    t = collections.OrderedDict()
    for r in data:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9}.get(r[mu], 1)
        t.setdefault(short(r[kn]), {})[r[mn]] = v
    tr = {"batch": 64, "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, tools/prof_kernels.py 64 1 "
          "(one launch per kernel, 64 x 1080p frames)", "kernels": {}}
    for k, m in t.items():
        rd, wr = m.get("dram__bytes_read.sum", 0.0), m.get("dram__bytes_write.sum", 0.0)
        tr["kernels"][k] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr}
        print(k, round(rd / 1e6, 1), round(wr / 1e6, 1))
    json.dump(tr, open(os.path.join(P, f"{R}_traffic.json"), "w"), indent=1)

    # ---- ncu --set full summary
    title = "ncu --set full --clock-control none, tools/prof_kernels.py 8 1 (final kernels of the round, 8 x 1080p frames per launch; dram bytes at the bench batch of 64 are in the round's traffic.json)"
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, f"{tag}_full.ncu-rep"), title],
                         capture_output=True, text=True).stdout
    open(os.path.join(P, f"{R}_ncu_full_final_summary.txt"), "w").write(txt)
    print(len(txt.splitlines()), "summary lines")


if __name__ == "__main__":
    main()
