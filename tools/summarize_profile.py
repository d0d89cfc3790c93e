#!/usr/bin/env python
"""Turn one tools/gpu_profile.sh pass (gpurun_out/<tag>_*) into the tracked summaries under profiles/:
  <out>_launches.txt      per-kernel share of one bench step (ncu launch list; cold-cache, serialised: compare SHARES)
  <out>_ncu_<k>_metrics.txt  selected `ncu --set full` metrics of one launch of each shading kernel
  <out>_ncu_<k>_by_line.txt  instructions / stall samples per source line of the tile function
  <out>_traffic.json      DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), read by bench.py
usage: summarize_profile.py <tag> <out-prefix>"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, outp = sys.argv[1], sys.argv[2]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def launches():
    rows = [r for r in csv.reader(open(os.path.join(G, f"{tag}_launches.csv"))) if len(r) > 5]
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        u = d["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        k = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "")
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f"{outp}_launches.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none over `bench.py --steps 1 --warmup 3` (every launch of the\n"
                "# process: warm-up, timed and end-to-end frames).  Cold-cache, serialised times: the SHARES are the evidence.\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:60s} n={v[0]:4d} total={v[1] / 1e3:9.2f} ms share={v[1] / tot * 100:5.1f}% avg={v[1] / v[0]:9.1f} us\n")


KEEP = ["gpu__time_duration.sum", "launch__", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_",
        "sm__pipe_tensor_cycles_active.avg", "smsp__warps_active.avg", "smsp__warps_eligible.avg", "smsp__average_warps_issue_stalled",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate", "l1tex__t_sector_hit_rate", "sm__cycles_elapsed.avg",
        "sm__throughput.avg", "sm__warps_active.avg"]


def ncu_kernel(name, kern_sub):
    rep = os.path.join(G, f"{tag}_{name}.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, val = rows[0], rows[1], rows[2]
    traffic = {}
    with open(os.path.join(P, f"{outp}_ncu_{name}_metrics.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, one launch of {kern_sub[:-5]}<18> inside bench.py (one chunk = the whole 512x512x128 frame)\n")
        for h, u, v in zip(hdr, units, val):
            if any(k in h for k in KEEP) and ".max" not in h and ".min" not in h and ".sum.pct" not in h:
                f.write(f"{h:95s} {v:>20s} {u}\n")
            if h in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
                x = float(v.replace(",", ""))
                x *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u, 1)
                traffic[h] = x
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    src_csv = f"/tmp/_src_{name}.csv"
    open(src_csv, "w").write(src)
    cub = "/tmp/_cub"
    os.makedirs(cub, exist_ok=True)
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "keypointnerf_b200/lib/libkpnerf_b200.so")], cwd=cub, capture_output=True)
    sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(cub, "kpn_shade_tc.sm_100a.cubin")], capture_output=True, text=True).stdout
    open("/tmp/_tci.sass", "w").write(sass)
    lines = open(os.path.join(ROOT, "keypointnerf_b200/csrc/kpn_shade_tc.cu")).read().splitlines()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools/ncu_by_line.py"), src_csv, "/tmp/_tci.sass", kern_sub, "kpn_shade_tc.cu",
                          "1", str(len(lines))], capture_output=True, text=True).stdout
    rows = []
    head = []
    for l in out.splitlines():
        mm = re.match(r"\s*(-?\d+)\s+([\d.]+)\s+([\d.]+)\s+(\d+)\s*(.*)", l)
        if mm:
            rows.append((int(mm[1]), float(mm[2]), float(mm[3]), int(mm[4]), mm[5]))
        else:
            head.append(l)
    rows.sort(key=lambda r: -r[1])
    with open(os.path.join(P, f"{outp}_ncu_{name}_by_line.txt"), "w") as f:
        f.write(f"# {kern_sub}: warp instructions / stall samples per source line of kpn_shade_tc.cu (innermost inlined frame), lines with\n"
                "# >= 0.25 % of the instructions, heaviest first (line -1: no line info)\n" + "\n".join(head[:2]) + "\n")
        for r in rows:
            if r[1] < 0.25:
                break
            text = lines[r[0] - 1].strip()[:90] if r[0] > 0 else ""
            f.write(f"{r[0]:6d} {r[1]:6.2f} {r[2]:6.2f} {r[3]:6d}  {r[4][:60]:60s} | {text}\n")
    return traffic


def frame_dram():
    """DRAM bytes + duration of every kernel of ONE frame (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum)."""
    path = os.path.join(G, f"{tag}_frame_dram.csv")
    if not os.path.exists(path):
        return {}
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, out = None, collections.OrderedDict()
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        k = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").split("::")[-1] + "#" + d["ID"]
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u, 1)
        out.setdefault(k, {})[d["Metric Name"]] = v
    return out


launches()
tr = {"geo": ncu_kernel("geo", "shade_geo_vseq_kernelILi18"), "color": ncu_kernel("color", "shade_color_kernelILi18")}
fr = frame_dram()
tot = sum(v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0) for v in fr.values())
json.dump({"source": f"ncu captures gpurun_out/{tag}_*: `kernels` = one --set full launch of each shading kernel (a launch covers one chunk = the whole "
                     "512x512x128 frame at the default chunk size); `frame` = every kernel of one frame (scene re-layout, front, geometry, colour list, "
                     "colour, compositing) with --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum",
           "kernels": tr, "frame": fr, "frame_dram_bytes": tot},
          open(os.path.join(P, f"{outp}_traffic.json"), "w"), indent=1)
print(json.dumps(tr, indent=1), "frame DRAM bytes", tot)
