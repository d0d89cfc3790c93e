"""Diagnostic (not a pytest): run the tiny golden query + one small render on the tensor-core engine and print the device
watchdog words (kpn_debug_timing): [flag, block, thread, tag, parity].  A non-zero flag means a barrier wait gave up."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypointnerf_b200.testing import build_model, scene_tensors
from tests.util import load_golden, scene_from_meta


def watchdog(m, tag):
    out = (C.c_ulonglong * 16)()
    m.lib.kpn_debug_timing(m.ctx, 1, out)
    w = list(out)[:5]
    print(f"{tag}: watchdog flag={w[0]} block={w[1]} thread={w[2]} (warp {w[2] // 32}) tag=0x{w[3]:x} parity={w[4]}")


for name in sys.argv[1:] or ["tiny"]:
    g, meta, sha = load_golden(name)
    scene, weights, target = scene_from_meta(meta)
    net = build_model(weights, meta["n_kpt"], "cuda:0")
    a = scene_tensors(scene, target, "cuda:0")
    net.engine = 0
    pts = torch.from_numpy(g["query_pts"]).cuda()[None]
    view = torch.from_numpy(g["query_view"]).cuda()[None]
    with torch.no_grad():
        out, valid = net.query(pts, a["cam"], a["feat_geo"], a["feat_tex"], n_views=3, sp_data=a["sp_data"], tx_data={"img": a["img"]},
                               view=view, src_foreground_mask=a["fg"], bounds=a["bounds"])
    torch.cuda.synchronize()
    v = g["query_valid"]
    o = out[0].cpu().numpy()
    print(name, "n", len(v), "valid", int(v.sum()), "rgb err", float(np.abs(o[v][:, 2:] - g["query_out"][v][:, 2:]).max()),
          "rad err", float(np.abs(o[v][:, 1] - g["query_out"][v][:, 1]).max()))
    watchdog(net.marcher(), name)
