"""Profiling driver (not a pytest): a few 64x64x128 strided passes through the default engine, for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypointnerf_b200 import synthetic as syn  # noqa: E402
from keypointnerf_b200.testing import build_model, scene_tensors  # noqa: E402

engine = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scene = syn.make_scene(src_size=512, n_kpt=18)
net = build_model(syn.make_weights(18), 18, "cuda:0")
target = syn.make_target(size=512)
a = scene_tensors(scene, target, "cuda:0")
m = net._bind_scene(a["cam"], a["feat_geo"], a["feat_tex"], a["sp_data"], a["img"], a["fg"], a["bounds"])
for i in range(n):
    r = m.render(K=a["cam_tar"]["K"], RT=a["cam_tar"]["RT"], znear=2.0, zfar=5.0, x0=i, y0=0, step=8, nx=64, ny=64, S_c=128,
                 engine=engine)
torch.cuda.synchronize()
print("done", m.stats())
