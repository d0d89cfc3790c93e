"""Diagnostic (not a pytest): per-stage wait vs compute cycles of the tensor-core row warps."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypointnerf_b200 import synthetic as syn
from keypointnerf_b200.testing import build_model, scene_tensors
scene = syn.make_scene(src_size=512, n_kpt=18); net = build_model(syn.make_weights(18), 18, "cuda:0")
target = syn.make_target(size=512); a = scene_tensors(scene, target, "cuda:0")
m = net._bind_scene(a["cam"], a["feat_geo"], a["feat_tex"], a["sp_data"], a["img"], a["fg"], a["bounds"])
names = ["L0", "L1", "L2", "L3", "P0|CMP", "P1", "BASE0", "BASE1", "VIS1A", "VIS1B", "VIS2A", "OUT0"]
for eng in (2, 0):
    kw = dict(K=a["cam_tar"]["K"], RT=a["cam_tar"]["RT"], znear=2.0, zfar=5.0, x0=0, y0=0, step=2, nx=256, ny=256, S_c=128, engine=eng)
    m.render(**kw); torch.cuda.synchronize()
    m.lib.kpn_debug_timing(m.ctx, 1, None)
    m.render(**kw); torch.cuda.synchronize()
    out = (C.c_ulonglong * 16)()
    m.lib.kpn_debug_timing(m.ctx, 0, out)
    o = np.array(list(out), dtype=np.float64)
    tiles = max(o[13], 1)
    print(f"engine {eng}: tiles recorded {int(o[13])}, cycles/tile {o[12]/tiles:.0f}, total wait/tile {o[:12].sum()/tiles:.0f} ({100*o[:12].sum()/o[12]:.1f}%)")
    print(f"   colour kernel: tiles {int(o[15])}, cycles/tile {o[14]/max(o[15],1):.0f}")
    print("   wait cycles per stage:", ", ".join(f"{n}={o[i]/tiles:.0f}" for i, n in enumerate(names)))
