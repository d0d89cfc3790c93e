"""CPU emulation of the tensor-core engine's rounding (oracle with fp16-rounded GEMM inputs per layer group): where does the RGB
error budget of the 1e-3 gate go?  Run: python tools/err_budget.py (about two minutes on 8 cores)."""
import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from keypointnerf_b200 import synthetic as syn
from oracle import kpnerf_oracle as O
torch.set_num_threads(8)
scene = syn.make_scene(512, 3, 18); w = syn.make_weights(18); fw = O.fold_weights(w); target = syn.make_target(512, azimuth=1.0)
h16 = lambda x: x.half().float()
orig_lin = O.lin
GEO1 = ["mlp_geo.layers1.layers.%d.linear" % i for i in range(4)]
COLW = ["mlp_tex.ray_encoder.2","mlp_tex.base_layer.0","mlp_tex.base_layer.2","mlp_tex.vis_layer1.0","mlp_tex.vis_layer1.2","mlp_tex.vis_layer2.0","mlp_tex.out_layer.0"]
def make_lin(geo_act=False, col_act=False, col_w=False, geo_w=False, lat16=False, den_act=False, geo_layers=None):
    def lin(fw_, name, x):
        W, b = fw_[name]
        if name in GEO1:
            if geo_act or (geo_layers is not None and GEO1.index(name) in geo_layers): x = h16(x)
            if geo_w: W = h16(W); b = h16(b)
        if name in ("mlp_geo.layers2.layers.0.linear","mlp_geo.layers2.layers.1.linear") and den_act: x = h16(x)
        if name in COLW:
            if col_act: x = h16(x)
            if col_w: W = h16(W)
        y = x @ W.t() + b
        if name == "ibr_compress_gfeat" and lat16: y = h16(y)
        return y
    return lin
with torch.no_grad():
    ref = O.render_tile(scene, fw, target, 4, 0, 0, 32)
    cases = {"geo act fp16": dict(geo_act=True), "colour act fp16": dict(col_act=True), "colour W fp16": dict(col_w=True),
             "latent fp16": dict(lat16=True), "all (engine 0 model)": dict(geo_act=True, col_act=True, col_w=True, lat16=True),
             "all + geo W single (engine 2)": dict(geo_act=True, col_act=True, col_w=True, lat16=True, geo_w=True),
             "geo W single only": dict(geo_w=True), "density-tail act single": dict(den_act=True)}
    for i in range(4):
        cases[f"geo act fp16 only layer {i} input"] = dict(geo_layers=[i])
    for k, kw in cases.items():
        O.lin = make_lin(**kw)
        out = O.render_tile(scene, fw, target, 4, 0, 0, 32)
        a=(out['alpha']-ref['alpha']).abs(); ok = a < 0.05
        e = (out['tex_fg']-ref['tex_fg']).abs().amax(0)[ok]; a = a[ok]
        print(f"{k:34s} rgb max {float(e.max()):.2e} rms {float((e**2).mean().sqrt()):.2e} | alpha max {float(a.max()):.2e} | flipped {int((~ok).sum())}")
    O.lin = orig_lin
