"""Drop-in check against the reference's own caller (needs /root/reference: build container only).

The reference's ``KeypointNeRFLightningModule`` (``src/model.py:28-46``) is instantiated twice from the same config: once with the
reference's ``KeypointNeRF`` and once with ``keypointnerf_b200.model.KeypointNeRF`` patched in (the one-line change of
INTEGRATION.md).  The reference's checkpoint is loaded into the second one with the reference's own STRICT ``load_ckpt``
(``src/model.py:113-117``), then both walk ``render_novel_views`` -> ``attach_im_feat`` -> ``render_full_nerf_image`` ->
``render_pifu_nerf`` (``src/model.py:453-507``) on the same batch.  There is no GPU here, so the CUDA marcher is replaced by a
stand-in that answers ``set_scene`` / ``render`` with the CPU oracle: what is under test is everything ABOVE the C ABI -- state_dict
keys, encoders, feature caching, argument marshalling, output dict -- and the images must agree with the reference's.
"""
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (build container only)")


def _import_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

        @property
        def device(self):
            return torch.device("cpu")

    k = stub("kornia")
    k.utils = stub("kornia.utils", tensor_to_image=lambda x: x.permute(0, 2, 3, 1).cpu().numpy())
    k.metrics = stub("kornia.metrics")
    kg = stub("kornia.geometry")
    kg.conversions = stub("kornia.geometry.conversions", convert_points_to_homogeneous=lambda x: x)
    k.geometry = kg
    stub("pytorch_lightning", LightningModule=LightningModule)
    stub("pytorch_lightning.utilities")
    stub("pytorch_lightning.utilities.apply_func", move_data_to_device=lambda b, d: b)
    stub("skimage")
    stub("skimage.metrics", structural_similarity=None)
    stub("imageio")
    torch.Tensor.cuda = lambda self, *a, **kw: self
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import src.model as M
    M.VGGLoss = lambda: None
    return M


class OracleMarcher:
    """Stand-in for ``RayMarcher`` (same ``set_scene`` / ``render`` keywords) that shades with the CPU oracle."""

    def __init__(self, weights):
        from oracle import kpnerf_oracle as O
        self.O, self.fw = O, O.fold_weights(weights)
        self.scene_binds = 0

    def set_scene(self, *, KRT, extrin, kpt3d, bounds, feat64, feat8, feat_tex, img, fg, width, height, znear, zfar, nml_scale):
        n = lambda t: t.detach().float().contiguous().numpy()
        V = img.shape[0]
        self.scene = {"n_views": V, "KRT": n(KRT).reshape(V, 4, 4), "extrin": n(extrin).reshape(V, 4, 4), "kpt3d": n(kpt3d)[None],
                      "bounds": n(bounds).reshape(1, 2, 3), "feat64": n(feat64), "feat8": n(feat8), "feat_tex": n(feat_tex),
                      "img": n(img), "fg": fg.reshape(V, 1, *fg.shape[-2:]).numpy().astype(bool), "width": float(width),
                      "height": float(height), "znear": float(znear), "zfar": float(zfar), "nml_scale": float(nml_scale)}
        self.scene_binds += 1

    def render(self, *, K, RT, znear, zfar, x0, y0, step, nx, ny, S_c, S_f=0, fine=False, out_device=None, engine=0, ert_eps=0.0,
               step_y=0):
        O = self.O
        target = {"K": K.reshape(-1, 4, 4).numpy(), "RT": RT.reshape(-1, 4, 4).numpy(), "znear": znear, "zfar": zfar}
        ys = torch.arange(ny) * (step_y or step) + y0
        xs = torch.arange(nx) * step + x0
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        pix = torch.stack([xx, yy], -1).reshape(-1, 2).float()
        flat = O.render_pixels(self.scene, self.fw, target, pix, S_c, S_f, fine)
        img = lambda v: v.reshape(ny, nx, 3).permute(2, 0, 1).contiguous()
        res = {"tex_fg": img(flat["tex_fg"]), "depth": flat["depth"].reshape(ny, nx), "alpha": flat["alpha"].reshape(ny, nx)}
        if fine:
            res.update({"tex_fg_fine": img(flat["tex_fg_fine"]), "depth_fine": flat["depth_fine"].reshape(ny, nx),
                        "alpha_fine": flat["alpha_fine"].reshape(ny, nx), "sdf": flat["sdf"].reshape(ny, nx)})
        return res


def _batch(scene, tgt):
    """What ZJUDataset yields for one frame (``src/zju_dataset.py:217-343``): index 0 = target view, 1.. = source views."""
    V, s = scene["n_views"], scene["src_size"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    Rt = np.concatenate([tgt["RT"][:, :3, :4], scene["extrin"][:, :3, :4]], 0)[None]
    Ks = np.concatenate([tgt["K"][:, :3, :3], scene["K"][:, :3, :3]], 0)[None]
    imgs = np.concatenate([np.zeros((1, 3, s, s), np.float32), scene["img"]], 0)[None]
    msk = np.concatenate([np.ones((1, 1, s, s), np.float32), scene["fg"].astype(np.float32)], 0)[None]
    return {"images": t(imgs), "images_masks": t(msk), "Rt": t(Rt), "K": t(Ks), "kpt3d": t(scene["kpt3d"]),
            "bounds": t(scene["bounds"]), "mask_at_box": torch.ones(1, s * s)}


@pytest.fixture
def keep_tensor_cuda():
    orig = torch.Tensor.cuda          # the CPU harness makes Tensor.cuda the identity: undo it afterwards
    yield
    torch.Tensor.cuda = orig


def test_reference_lightning_module_runs_on_this_class(tmp_path, keep_tensor_cuda):
    from keypointnerf_b200 import synthetic as syn
    from keypointnerf_b200.config import default_cfg
    from keypointnerf_b200.model import KeypointNeRF as Mine
    M = _import_reference()
    RefNet = M.KeypointNeRF
    torch.manual_seed(0)
    n_kpt = 18
    cfg = default_cfg(n_kpt)
    cfg["models"]["KeypointNeRF"]["dr_kwargs"].update(sample_per_ray_c=12, sample_per_ray_f=8)
    cfg.update(expname="t", out_dir=str(tmp_path), training={})
    weights = syn.make_weights(n_kpt)
    scene = syn.make_scene(src_size=128, n_kpt=n_kpt, fg_mode="hull")
    tgt = syn.make_target(size=64, azimuth=1.0, zoom=2.0)
    batch = _batch(scene, tgt)
    cams = [{"intrinsics": torch.from_numpy(syn.make_target(64, az, zoom=2.0)["K"]),
             "w2cs": torch.from_numpy(syn.make_target(64, az, zoom=2.0)["RT"][0]), "im_h": 64, "im_w": 64, "znear": 2.0, "zfar": 5.0}
            for az in (1.0, 2.5)]
    try:
        ref = M.KeypointNeRFLightningModule(cfg, None).eval()
        res = ref.model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}, strict=False)
        assert not res.unexpected_keys
        sd = ref.state_dict()
        sd["model.vgg_loss.vgg.slice1.0.weight"] = torch.zeros(4)   # the real model owns a frozen VGG19 (src/model.py:608)
        path = os.path.join(tmp_path, "latest.ckpt")
        torch.save({"state_dict": sd, "epoch": 3, "global_step": 7}, path)
        with torch.no_grad():
            want = ref.render_novel_views(cams, batch, only_renderings=True)[0]

        M.KeypointNeRF = Mine                                        # the one-line change of INTEGRATION.md
        mine = M.KeypointNeRFLightningModule(cfg, None).eval()
        assert isinstance(mine.model, Mine)
        assert set(k for k in mine.state_dict()) == set(k for k in ref.state_dict()), "state_dict keys differ from the reference's"
        assert mine.load_ckpt(path) == (3, 7)                        # strict load_state_dict inside
        fake = OracleMarcher({k: v.numpy() for k, v in mine.model.state_dict().items() if not k.startswith(("geo_", "tex_", "sp_"))})
        mine.model.marcher = lambda: fake
        with torch.no_grad():
            got = mine.render_novel_views(cams, batch, only_renderings=True)[0]
    finally:
        M.KeypointNeRF = RefNet
    assert got.shape == want.shape == (2, 64, 64, 3) and got.dtype == np.uint8
    assert fake.scene_binds == 1, "the source-image set must be bound (and encoded) once per sweep, not once per camera"
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert want.max() > 30, "the test scene renders nothing"
    assert d.max() <= 1, f"8-bit renderings differ by up to {d.max()}"
