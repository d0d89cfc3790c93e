#!/usr/bin/env python
"""torchrun worker: N-GPU renders (BASELINE config 4 lattice-sharded frame, config 5 one view per rank) == 1-GPU renders, bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_check.py [size]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from keypointnerf_b200 import distributed as D  # noqa: E402
from keypointnerf_b200 import synthetic as syn  # noqa: E402
from keypointnerf_b200.testing import build_model, scene_tensors  # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    rank, world, local = D.init_from_env("nccl")
    dev = f"cuda:{local}"
    scene = syn.make_scene(512, 3, 18)
    weights = syn.make_weights(18)
    net = build_model(weights, 18, dev)
    tgt = syn.make_target(size, azimuth=1.0)
    a = scene_tensors(scene, tgt, dev)
    m = net._bind_scene(a["cam"], a["feat_geo"], a["feat_tex"], a["sp_data"], a["img"], a["fg"], a["bounds"])
    kw = dict(K=a["cam_tar"]["K"], RT=a["cam_tar"]["RT"], znear=2.0, zfar=5.0, S_c=64, S_f=32, fine=True)
    ok = True
    # config 4: one frame in lattice phases; every rank also renders the whole frame on its own GPU as the single-GPU result
    sharded = D.render_frame_lattice_sharded(m, width=size, height=size, rank=rank, world=world, **kw)
    whole = m.render(x0=0, y0=0, step=1, nx=size, ny=size, out_device="cuda", **kw)
    for k in ("tex_fg", "alpha", "depth", "tex_fg_fine", "alpha_fine", "sdf"):
        ok &= bool(torch.equal(sharded[k], whole[k]))
    # config 5: one view per rank, gathered; every rank re-renders all the views itself
    mine = m.render(K=torch.from_numpy(syn.make_target(size, 1.0 + rank * np.pi / 4)["K"]).to(dev),
                    RT=torch.from_numpy(syn.make_target(size, 1.0 + rank * np.pi / 4)["RT"]).to(dev),
                    znear=2.0, zfar=5.0, S_c=64, x0=0, y0=0, step=1, nx=size, ny=size, out_device="cuda")["tex_fg"]
    views = D.gather_views(mine, world)
    for r in range(world):
        t = syn.make_target(size, 1.0 + r * np.pi / 4)
        solo = m.render(K=torch.from_numpy(t["K"]).to(dev), RT=torch.from_numpy(t["RT"]).to(dev), znear=2.0, zfar=5.0, S_c=64,
                        x0=0, y0=0, step=1, nx=size, ny=size, out_device="cuda")["tex_fg"]
        ok &= bool(torch.equal(views[r], solo))
    torch.cuda.synchronize()
    m.check_health()
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    if world > 1:
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_OK" if float(flag) == 1.0 else "MULTI_GPU_MISMATCH", f"world={world} size={size}")
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0 if float(flag) == 1.0 else 1


if __name__ == "__main__":
    sys.exit(main())
