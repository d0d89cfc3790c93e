"""Where does the end-to-end (host tensors in / host tensors out) frame time go?  Times the pieces of bench.py's step_e2e."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as G
from keypointnerf_b200 import synthetic as syn
from keypointnerf_b200.testing import build_model, scene_tensors
G.build()
dev = "cuda:0"
scene = syn.make_scene(512, 3, 18); weights = syn.make_weights(18); target = syn.make_target(512, azimuth=1.0)
net = build_model(weights, 18, dev)
h = scene_tensors(scene, target, "cpu", pin=True)
cfgk = dict(sample_per_ray_c=128, sample_per_ray_f=0, fine=False, uniform=True)
def sync(): torch.cuda.synchronize()
def T(f, n=5):
    sync(); t0 = time.perf_counter()
    for _ in range(n): f()
    sync(); return (time.perf_counter() - t0) / n * 1e3
def full():
    net._scene_key = None
    return net.render_pifu_nerf(net, h["img"], h["cam"], h["cam_tar"], level=4, sp_data=h["sp_data"], feat_geo=h["feat_geo"],
                                feat_tex=h["feat_tex"], src_foreground_mask=h["fg"], bounds=h["bounds"], mask_at_box=None, **cfgk)
for _ in range(3): full()
print("full e2e            %.2f ms" % T(full))
def bind():
    net._scene_key = None
    net._bind_scene(h["cam"], h["feat_geo"], h["feat_tex"], h["sp_data"], h["img"], h["fg"], h["bounds"])
print("bind scene (host)   %.2f ms" % T(bind))
m = net.marcher()
def rend(dev_out):
    return m.render(K=h["cam_tar"]["K"], RT=h["cam_tar"]["RT"], znear=target["znear"], zfar=target["zfar"], x0=0, y0=0, step=1, nx=512, ny=512,
                    S_c=128, fine=False, out_device=dev_out, engine=0)
print("render -> cpu out   %.2f ms" % T(lambda: rend("cpu")))
print("render -> cuda out  %.2f ms" % T(lambda: rend("cuda")))
print("marcher() key check %.3f ms" % T(lambda: net.marcher(), 20))
fg = h["fg"]
print("fg -> uint8 (cpu)   %.3f ms" % T(lambda: fg.detach().reshape(3, 1, 512, 512).to(dtype=torch.uint8).contiguous(), 20))
print("pinned alloc 5MB    %.3f ms" % T(lambda: torch.empty(3, 512, 512, dtype=torch.float32, pin_memory=True), 20))
