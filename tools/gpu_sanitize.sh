#!/bin/bash
# compute-sanitizer memcheck over the smoke render (tensor-core engine, coarse + fine) and a small fp32-engine / query / ERT run.
out=gpurun_out; mkdir -p $out
timeout 280 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 3 python -c "
import __graft_entry__ as g
g.smoke()
import torch
from keypointnerf_b200 import synthetic as syn
from keypointnerf_b200.testing import build_model, scene_tensors
scene = syn.make_scene(src_size=64, n_kpt=24, fg_mode='hull'); w = syn.make_weights(24); t = syn.make_target(size=48, zoom=2.0)
net = build_model(w, 24, 'cuda:0'); a = scene_tensors(scene, t, 'cuda:0')
m = net._bind_scene(a['cam'], a['feat_geo'], a['feat_tex'], a['sp_data'], a['img'], a['fg'], a['bounds'])
kw = dict(K=a['cam_tar']['K'], RT=a['cam_tar']['RT'], znear=2.0, zfar=5.0, x0=0, y0=0, step=1, nx=48, ny=48, S_c=19, S_f=21, fine=True)
for eng in (0, 1):
    r = m.render(engine=eng, ert_eps=1e-3, debug=True, **kw)
    pts = torch.rand(3000, 3, device='cuda') - 0.5; view = torch.nn.functional.normalize(torch.randn(3000, 3, device='cuda'), dim=-1)
    o, v = m.query(pts, view, engine=eng)
torch.cuda.synchronize(); m.check_health(); print('SANITIZE_DONE', float(r['alpha_fine'].max()))
" > $out/sanitize.log 2>&1
echo "exit $?" >> $out/sanitize.log
tail -15 $out/sanitize.log
