#!/usr/bin/env python
"""Benchmark of the ray-march hot path (BASELINE.json metric: rays/sec, 512x512, 128 samples/ray,
3 source views, 18 keypoints).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the hot path over one synthetic frame: re-layout of the (new) source feature
maps + ray generation + sampling + shading + compositing for every pixel of a 512x512 novel view.
With N > 1 (torchrun) each rank renders its own novel view of the same scene (BASELINE config 5,
weak scaling) and the frames are exchanged with ONE NCCL all-gather.

`value`   : inputs resident in HBM, CUDA-event time on the launch stream, max over ranks.
`e2e`     : the same metric through the reference-facing API (`KeypointNeRF.render_pifu_nerf`) with
            pinned HOST tensors in and host tensors out (H2D + D2H inside the timed region).
`roofline`: dominant (shading) kernel, algorithmic FLOPs of the samples it shades / its device time
            measured live with CUDA events inside the timed region, against the measured dense
            bf16 tensor peak in MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the CPU oracle port (oracle/, torch-CPU, all host threads) on a
            bounded sample of the same workload (one 64x64 strided pass = 4096 rays x 128 samples).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SIZE = 512
S_C = 128
N_KPT = 18
N_VIEWS = 3
WORKLOAD = "512x512 frame, 128 samples/ray (fine off), 3 source views 512^2, 18 keypoints, random maps+weights"


def flop_per_sample(n_kpt: int, n_views: int) -> int:
    """Dense-layer FLOPs (2*MAC) per evaluated sample, SURVEY.md section 8a."""
    enc = 7 * n_kpt
    geo = (enc + 64) * 128 + 128 * 128 + 136 * 120 + 120 * 64
    ibr = 4 * 16 + 16 * 35 + 105 * 64 + 64 * 32 + 32 * 32 + 32 * 33 + 32 * 32 + 32 + 37 * 16 + 16 * 8 + 8
    pooled = 128 * 64 + 64 * 64 + 64 * 2 + 128 * 24
    return 2 * (n_views * (geo + ibr) + pooled)


def ncu_traffic():
    """DRAM bytes per shading launch pair (geometry + colour kernel) from the latest committed ncu capture (profiles/*_traffic.json,
    written by tools/summarize_profile.py); None when no capture is committed."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    tot = sum(k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0) for k in d["kernels"].values())
    return tot, os.path.basename(files[-1])


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops"))), "hbm_gbs": float(d["hbm_gbs"]),
                "source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"}
    return {"tflops": 1400.0, "hbm_gbs": 6650.0, "source": "B200_PROFILING.md fallback, sustained (of fallback)"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_rays_per_s(steps: int, warmup: int, threads: int | None = None):
    """Oracle port on the host cores: one 64x64 strided pass (4096 rays x 128 samples) per step.  The torch-CPU path is bound by
    materialised intermediates and does not scale to every core of a large host, so (unless `threads` is given) the first two
    untimed passes try all cores and 16 threads and the timed passes use the faster setting; `cores` reports the threads used."""
    import torch
    from keypointnerf_b200 import synthetic as syn
    from oracle import kpnerf_oracle as O
    scene = syn.make_scene(SIZE, N_VIEWS, N_KPT)
    fw = O.fold_weights(syn.make_weights(N_KPT))
    target = syn.make_target(SIZE, azimuth=1.0)

    def one_pass(i):
        t0 = time.perf_counter()
        O.render_tile(scene, fw, target, 4, i % 8, (i // 8) % 8, S_C)
        return time.perf_counter() - t0

    with torch.no_grad():
        if threads:
            torch.set_num_threads(threads)
        else:
            all_cores = torch.get_num_threads()
            cands = sorted({all_cores, min(16, all_cores)}, reverse=True)
            if len(cands) > 1:
                one_pass(0)                      # first call of the process: allocator / thread-pool warm-up, not a measurement
                probe = {}
                for n in cands:
                    torch.set_num_threads(n)
                    probe[n] = one_pass(1)
                torch.set_num_threads(min(probe, key=probe.get))
        cores = torch.get_num_threads()
        times = []
        for i in range(warmup + steps):
            dt = one_pass(i)
            if i >= warmup:
                times.append(dt)
    rays = 64 * 64
    return rays * len(times) / sum(times), cores, sum(times) / len(times) * 1e3


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    val, cores, ms = cpu_port_rays_per_s(args.steps, args.warmup)
    sample = "one 64x64 strided pass of the 512x512 frame per step (4096 rays x 128 samples)"
    line = {"impl": "reference", "metric": "rays/sec", "value": val, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU oracle port of the reference PyTorch path (the reference is "
                       "Python and cannot travel to the GPU box); bounded sample per step"},
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--engine", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as G
    from keypointnerf_b200 import distributed as D
    from keypointnerf_b200 import synthetic as syn
    from keypointnerf_b200.testing import build_model, scene_tensors

    G.build()
    rank, world, local = D.init_from_env("nccl")
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"

    scene = syn.make_scene(SIZE, N_VIEWS, N_KPT)
    weights = syn.make_weights(N_KPT)
    target = syn.make_target(SIZE, azimuth=1.0 + rank * np.pi / 4.0)   # one novel view per rank
    net = build_model(weights, N_KPT, dev)
    net.engine = args.engine
    a = scene_tensors(scene, target, dev)
    h = scene_tensors(scene, target, "cpu", pin=True)
    m = net.marcher()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def bind(t):
        m.set_scene(KRT=t["cam"]["KRT"], extrin=t["sp_data"]["extrin"], kpt3d=t["sp_data"]["kpt3d"].reshape(-1, 3),
                    bounds=t["bounds"], feat64=t["feat_geo"][0], feat8=t["feat_geo"][1], feat_tex=t["feat_tex"],
                    img=t["img"], fg=t["fg"], width=scene["width"], height=scene["height"], znear=scene["znear"],
                    zfar=scene["zfar"], nml_scale=scene["nml_scale"])

    def step_device():
        bind(a)   # every step is a new frame: the atlases are re-packed
        res = m.render(K=a["cam_tar"]["K"], RT=a["cam_tar"]["RT"], znear=target["znear"], zfar=target["zfar"],
                       x0=0, y0=0, step=1, nx=SIZE, ny=SIZE, S_c=S_C, fine=False, out_device="cuda", engine=args.engine)
        frames = D.gather_views(res["tex_fg"], world)
        return frames

    cfgk = dict(sample_per_ray_c=S_C, sample_per_ray_f=0, fine=False, uniform=True)

    def step_e2e():
        net._scene_key = None   # new frame: host feature maps are uploaded again
        return net.render_pifu_nerf(net, h["img"], h["cam"], h["cam_tar"], level=4, sp_data=h["sp_data"],
                                    feat_geo=h["feat_geo"], feat_tex=h["feat_tex"], src_foreground_mask=h["fg"],
                                    bounds=h["bounds"], mask_at_box=None, **cfgk)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    m.stats()           # drain event pool / counters
    launches0 = m.stats()["kernel_launches"]
    m.set_profiling(True)
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (not timed)
        evs[k][0].record()
        step_device()
        evs[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    st = m.stats()
    m.set_profiling(False)
    ms_total = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    tms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_total = float(tms.item())
    rays_per_step = SIZE * SIZE * world
    value = rays_per_step * args.steps / (ms_total * 1e-3)
    launches = st["kernel_launches"] - launches0

    # end to end through the reference-facing API, host buffers in / host buffers out
    out = None
    for _ in range(3):
        out = step_e2e()      # keep the previous result alive like the timed loop does (two sets of pinned output buffers)
    barrier()
    e2e_each = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        t1 = time.perf_counter()
        out = step_e2e()      # returns host tensors: the call itself waits for the device-to-host copies
        e2e_each.append((time.perf_counter() - t1) * 1e3)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = rays_per_step * args.steps / float(te.item())
    h2d = sum(int(t.numel() * t.element_size()) for t in
              [h["feat_geo"][0], h["feat_geo"][1], h["feat_tex"], h["img"], h["fg"], h["cam"]["KRT"], h["sp_data"]["extrin"],
               h["sp_data"]["kpt3d"], h["bounds"], h["cam_tar"]["K"], h["cam_tar"]["RT"]])
    d2h = sum(int(v.numel() * v.element_size()) for v in out.values())

    fps = flop_per_sample(N_KPT, N_VIEWS)
    pk = peaks()
    valid_per_step = st["samples_valid"]
    shade_ms = st["shade_ms"]
    ach = (fps * valid_per_step * args.steps) / (shade_ms * 1e-3) / 1e12 if shade_ms > 0 else None
    traffic, traffic_src = ncu_traffic()
    roofline = {"bound": "tensor", "kernel": "shade_geo_kernel + shade_color_kernel (per-sample gather+encode+MLPs; geometry is ~85% of it)",
                "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": (ach / pk["tflops"]) if ach else None,
                "traffic": traffic, "traffic_unit": "DRAM bytes per launch pair (one chunk = one 512x512x128 frame), ncu --set full",
                "traffic_source": traffic_src,
                "peak_source": pk["source"], "flop_per_sample": fps, "valid_samples_per_step": valid_per_step,
                "valid_frac": valid_per_step / float(SIZE * SIZE * S_C), "shade_ms_per_step": shade_ms / args.steps,
                "shade_launches_per_step": st["shade_launches"] / args.steps,
                "shade_share_of_step": (shade_ms / ms_total) if ms_total else None,
                "nominal_tflops_all_samples": fps * SIZE * SIZE * S_C * args.steps / (ms_total * 1e-3) / 1e12}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, cores, ms = cpu_port_rays_per_s(args.cpu_steps, 1)
        cpu = {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_steps} x one 64x64 strided pass of the same frame (4096 rays x 128 samples), "
                         f"{ms:.0f} ms each, torch-CPU oracle port"}

    if rank == 0:
        line = {"metric": "rays/sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "fp32" if args.engine == 1 else net.marcher_dtype(), "data": "synthetic",
                "config": {"workload": WORKLOAD, "parallelism": f"one novel view per GPU x{world}, one all-gather of frames",
                           "l2": "flushed between timed iterations (256 MiB write)", "engine": args.engine,
                           "wall_ms_per_step_incl_flush": t_wall / args.steps * 1e3},
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_each": [round(x, 2) for x in e2e_each]},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
