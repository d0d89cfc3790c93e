#!/usr/bin/env python
"""Where the geometry kernel spends its cycles: renders the bench frame with the instrumented library (-DKPN_STAGE_TIMING:
cycle stamps of a few warps of block 0) and prints medians per stage.

    python tools/stage_times.py            # view-sequential kernel (engine 0 at 18 keypoints): row warp 0, issuer, producer warp 0
    python tools/stage_times.py --engine4  # row-per-view kernel: one issuer warp and its column-half partner
    python tools/stage_times.py --build    # only (re)build keypointnerf_b200/lib/libkpnerf_b200_timing.so (needs nvcc)
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402

lib_path = os.path.join(G.LIBDIR, "libkpnerf_b200_timing.so")
if "--build" in sys.argv or not os.path.exists(lib_path):
    G.build_variant("timing", ["KPN_STAGE_TIMING"])
if "--build" in sys.argv:
    raise SystemExit(0)
os.environ["KPN_LIB"] = lib_path

import numpy as np  # noqa: E402
import torch  # noqa: E402

from keypointnerf_b200 import synthetic as syn  # noqa: E402
from keypointnerf_b200.testing import build_model, scene_tensors  # noqa: E402

ROWS, TILES, WORDS, NBLK = 3, 48, 32, 256
engine = 4 if "--engine4" in sys.argv else 0
n_kpt = 18
scene, weights, target = syn.make_scene(512, 3, n_kpt), syn.make_weights(n_kpt), syn.make_target(512)
net = build_model(weights, n_kpt, "cuda:0")
a = scene_tensors(scene, target, "cuda:0")
m = net._bind_scene(a["cam"], a["feat_geo"], a["feat_tex"], a["sp_data"], a["img"], a["fg"], a["bounds"])
kw = dict(K=a["cam_tar"]["K"], RT=a["cam_tar"]["RT"], znear=2.0, zfar=5.0, x0=0, y0=0, step=1, nx=512, ny=512, S_c=128, engine=engine)
NW = ROWS * TILES * WORDS + NBLK
buf = (C.c_ulonglong * NW)()
nt = C.c_int(0)
m.lib.kpn_debug_stage_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
for it in range(3):
    m.render(**kw)
    torch.cuda.synchronize()
    rc = m.lib.kpn_debug_stage_times(m.ctx, buf, NW, C.byref(nt))
    assert rc == 0, rc
allw = np.frombuffer(buf, dtype=np.uint64).astype(np.int64)
rec = allw[:ROWS * TILES * WORDS].reshape(ROWS, TILES, WORDS)
blk = allw[ROWS * TILES * WORDS:][:148]
med = lambda x: float(np.median(x))


def vseq_report():
    """Row 0 = row warp 0, row 1 = the issuer warp, row 2 = producer warp 0 of block 0; one record per iteration."""
    n = nt.value
    r, i, p = rec[0][8:n - 1], rec[1][8:n - 1], rec[2][8:n - 1]
    rn, pn, rp = rec[0][9:n], rec[2][9:n], rec[0][7:n - 2]      # the next / previous iteration's records
    # the 14 stages of an iteration in issue order: name, row-warp stamp where its CUDA-core work starts, stamp of its signal,
    # stamp where its accumulator wait returns, stamp just before that wait is entered (>= 32: the next iteration's record)
    T = [("A4 P0|C", 1, 2, 5, 4), ("B1 v0", 3, 4, 7, 6), ("A5 P1", 5, 6, 9, 8), ("B2 v0", 7, 8, 11, 10), ("A0 v1", 9, 10, 13, 12),
         ("B3 v0", 11, 12, 15, 14), ("A1 v1", 13, 14, 17, 16), ("B0 v2", 15, 16, 19, 18), ("A2 v1", 17, 18, 21, 20),
         ("B1 v2", 19, 20, 23, 22), ("A3 v1", 21, 22, 25, 24), ("B2 v2", 23, 24, 27, 26), ("A3 v2", 27, 28, 32 + 1, 32 + 0),
         ("B0 v0'", 28, 29, 32 + 3, 32 + 2)]
    col = lambda j: r[:, j] if j < 32 else rn[:, j - 32]
    print(f"iterations recorded: {n}; medians in SM cycles; iteration period {med(rn[:, 0] - r[:, 0]):.0f}")
    print(f"{'stage':8s} {'cuda work':>10s} {'sig->issuer':>12s} {'issue':>7s} {'commit->wake':>13s} {'round trip':>11s} {'waited':>7s}")
    tot = np.zeros(6)
    for g, (nm, start, sig, wake, before) in enumerate(T):
        vals = [med(col(sig) - col(start)), med(i[:, 2 * g] - col(sig)), med(i[:, 2 * g + 1] - i[:, 2 * g]), med(col(wake) - i[:, 2 * g + 1]),
                med(col(wake) - col(sig)), med(col(wake) - col(before))]
        tot += vals
        print(f"{nm:8s} " + " ".join(f"{v:{w}.0f}" for v, w in zip(vals, (10, 12, 7, 13, 11, 7))))
    print(f"{'sum':8s} " + " ".join(f"{v:{w}.0f}" for v, w in zip(tot, (10, 12, 7, 13, 11, 7))))
    print(f"view-1 accumulate (no signal): {med(r[:, 26] - r[:, 25]):.0f}")
    # producer warp 0: [0] iteration start, [1] samples staged, per view v: [2+4v] buffers free, [3+4v] its units staged, [5+4v] arrived
    print(f"producer warp 0: iteration period {med(pn[:, 0] - p[:, 0]):.0f}; sample buffer wait + staging {med(p[:, 1] - p[:, 0]):.0f}")
    for v in range(3):
        prev_end = p[:, 1] if v == 0 else p[:, 5 + 4 * (v - 1)]
        # view 0 of iteration it is consumed by the build at the END of iteration it - 1 (stamp 29 of the previous record)
        sig = r[:, 10] if v == 1 else r[:, 16] if v == 2 else rp[:, 29]
        print(f"  view {v}: waited for buffers {med(p[:, 2 + 4 * v] - prev_end):6.0f}  units staged {med(p[:, 3 + 4 * v] - p[:, 2 + 4 * v]):6.0f}  "
              f"arrive {med(p[:, 5 + 4 * v] - p[:, 3 + 4 * v]):5.0f}   arrive -> row warp 0 has built and signalled: {med(sig - p[:, 5 + 4 * v]):7.0f}")
    print("M cycles per block (main loop), every 8th of the sorted 148: " + " ".join(f"{x / 1e6:.2f}" for x in np.sort(blk)[::8])
          + f"; max {blk.max() / 1e6:.2f}")
    if "--blocks" in sys.argv:
        print("per block: " + " ".join(f"{x / 1e6:.1f}" for x in blk))


def row_per_view_report():
    t = rec[0][4:nt.value]   # the issuer warp; skip the first tiles (cold)
    t1 = rec[1][4:nt.value]  # its column-half-1 partner (same rows, other warp)
    print(f"tiles recorded: {nt.value}; medians in SM cycles over {len(t)} tiles of one issuer warp (slot 0 of cluster 0)")
    print(f"tile period (start -> next start of the same slot): {med(t[1:, 0] - t[:-1, 0]):.0f}")
    print(f"stage-0 input build: {med(t[:, 1] - t[:, 0]):.0f}")
    names = ["L0", "L1", "L2", "L3", "P0|C", "P1"]
    tot = dict(cuda=med(t[:, 1] - t[:, 0]), arrive=0.0, gather=0.0, issue=0.0, mma=0.0)
    print(f"{'stage':6s} {'signal':>8s} {'all-arrived':>12s} {'issue':>8s} {'mma+wake':>10s} {'epilogue':>9s}")
    for s, n in enumerate(names):
        sig = med(t[:, 2 + 5 * s] - t[:, (1 if s == 0 else 6 + 5 * (s - 1))])
        arr = med(t[:, 3 + 5 * s] - t[:, 2 + 5 * s])
        iss = med(t[:, 4 + 5 * s] - t[:, 3 + 5 * s])
        mma = med(t[:, 5 + 5 * s] - t[:, 4 + 5 * s])
        epi = med(t[:, 6 + 5 * s] - t[:, 5 + 5 * s])
        print(f"{n:6s} {sig:8.0f} {arr:12.0f} {iss:8.0f} {mma:10.0f} {epi:9.0f}")
        tot["arrive"] += sig; tot["gather"] += arr; tot["issue"] += iss; tot["mma"] += mma; tot["cuda"] += epi
    print("per tile: " + ", ".join(f"{k} {v:.0f}" for k, v in tot.items()) + f"; sum {sum(tot.values()):.0f}")
    # the h = 1 partner warp: its own CUDA-core time per stage (accumulator visible -> next input signalled) next to the issuer's
    print("CUDA-core time per stage, h=0 (issuer) | h=1:  build " + f"{med(t[:, 1] - t[:, 0]):.0f} | {med(t1[:, 1] - t1[:, 0]):.0f}")
    for s, n in enumerate(names):
        print(f"  {n:6s} {med(t[:, 6 + 5 * s] - t[:, 5 + 5 * s]):6.0f} | {med(t1[:, 6 + 5 * s] - t1[:, 5 + 5 * s]):6.0f}"
              f"     signal->acc visible: {med(t[:, 5 + 5 * s] - t[:, 2 + 5 * s]):6.0f} | {med(t1[:, 5 + 5 * s] - t1[:, 2 + 5 * s]):6.0f}")
    print(f"  tile start offset h=1 vs h=0 (median): {med(t1[:, 0] - t[:, 0]):.0f}")


if engine == 0:
    vseq_report()
else:
    row_per_view_report()
