#!/usr/bin/env python
"""Where one tile of the geometry kernel spends its cycles: renders the bench frame with the instrumented library
(-DKPN_STAGE_TIMING: cycle stamps of one issuer warp, block 0 / slot 0) and prints, per stage, the medians of
  build/epilogue (CUDA cores) | own arrive -> all 16 row warps arrived | MMA issue | commit -> accumulator visible.

    python tools/stage_times.py            # builds keypointnerf_b200/lib/libkpnerf_b200_timing.so if missing (needs nvcc)
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
    sys.exit(0)
os.environ["KPN_LIB"] = lib_path

import numpy as np  # noqa: E402
import torch  # noqa: E402

from keypointnerf_b200 import synthetic as syn  # noqa: E402
from keypointnerf_b200.testing import build_model, scene_tensors  # noqa: E402

n_kpt = 18
scene, weights, target = syn.make_scene(512, 3, n_kpt), syn.make_weights(n_kpt), syn.make_target(512)
net = build_model(weights, n_kpt, "cuda:0")
a = scene_tensors(scene, target, "cuda:0")
m = net._bind_scene(a["cam"], a["feat_geo"], a["feat_tex"], a["sp_data"], a["img"], a["fg"], a["bounds"])
kw = dict(K=a["cam_tar"]["K"], RT=a["cam_tar"]["RT"], znear=2.0, zfar=5.0, x0=0, y0=0, step=1, nx=512, ny=512, S_c=128)
buf = (C.c_ulonglong * (2 * 48 * 32))()
nt = C.c_int(0)
m.lib.kpn_debug_stage_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
for it in range(3):
    m.render(**kw)
    torch.cuda.synchronize()
    rc = m.lib.kpn_debug_stage_times(m.ctx, buf, 2 * 48 * 32, C.byref(nt))
    assert rc == 0, rc
both = np.frombuffer(buf, dtype=np.uint64).reshape(2, 48, 32).astype(np.int64)
t = both[0][4:nt.value]   # the issuer warp; skip the first tiles (cold)
t1 = both[1][4:nt.value]  # its column-half-1 partner (same rows, other warp)
print(f"tiles recorded: {nt.value}; medians in SM cycles over {len(t)} tiles of one issuer warp (slot 0 of cluster 0)")
med = lambda x: float(np.median(x))
tile = med(t[1:, 0] - t[:-1, 0])
print(f"tile period (start -> next start of the same slot): {tile:.0f}")
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
