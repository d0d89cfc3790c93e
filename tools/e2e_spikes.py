"""Find the slow iterations of the end-to-end path (host tensors in / host tensors out) and which phase they sit in."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from keypointnerf_b200 import synthetic as syn
from keypointnerf_b200 import renderer as R
from keypointnerf_b200.testing import build_model, scene_tensors
dev = "cuda:0"
scene = syn.make_scene(512, 3, 18); weights = syn.make_weights(18); target = syn.make_target(512, azimuth=1.0)
net = build_model(weights, 18, dev)
h = scene_tensors(scene, target, "cpu", pin=True)
cfgk = dict(sample_per_ray_c=128, sample_per_ray_f=0, fine=False, uniform=True)
m = net.marcher()
ph = {}
def wrap(obj, name, key):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); ph[key] = ph.get(key, 0.0) + (time.perf_counter() - t0) * 1e3; return r
    setattr(obj, name, g)
wrap(m, "set_scene", "set_scene")
wrap(m.lib, "kpn_render", "kpn_render_call")
wrap(m.lib, "kpn_set_scene", "kpn_set_scene_call")
wrap(m, "check_health", "health")
_empty = torch.empty
def empty(*a, **k):
    t0 = time.perf_counter(); r = _empty(*a, **k)
    if k.get("pin_memory"): ph["pinned_alloc"] = ph.get("pinned_alloc", 0.0) + (time.perf_counter() - t0) * 1e3
    return r
torch.empty = empty
def full():
    net._scene_key = None
    return net.render_pifu_nerf(net, h["img"], h["cam"], h["cam_tar"], level=4, sp_data=h["sp_data"], feat_geo=h["feat_geo"],
                                feat_tex=h["feat_tex"], src_foreground_mask=h["fg"], bounds=h["bounds"], mask_at_box=None, **cfgk)
out = None
for _ in range(3): out = full()
rows = []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 80):
    ph.clear()
    t0 = time.perf_counter(); out = full(); dt = (time.perf_counter() - t0) * 1e3
    rows.append((dt, dict(ph)))
ts = sorted(r[0] for r in rows)
print("e2e ms: median %.2f p90 %.2f max %.2f mean %.2f" % (ts[len(ts) // 2], ts[int(len(ts) * 0.9)], ts[-1], sum(ts) / len(ts)))
med = {k: sorted(r[1].get(k, 0.0) for r in rows)[len(rows) // 2] for k in rows[0][1]}
print("median phases:", {k: round(v, 3) for k, v in med.items()})
for i, (dt, p) in enumerate(rows):
    if dt > ts[len(ts) // 2] * 1.15:
        print("slow iteration", i, "%.2f ms" % dt, {k: round(v, 2) for k, v in p.items()})
