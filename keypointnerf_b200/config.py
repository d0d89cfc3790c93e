"""Model configuration in the reference's dict schema (``cfg['models']['KeypointNeRF']``,
reference ``configs/zju.json:29-120``, read at ``src/model.py:562-601``).  Only the keys the
ray-march path consumes are material here; the rest are carried so that a reference config
file can be passed unchanged."""
from __future__ import annotations

import copy
import json


def default_cfg(n_kpt: int = 24) -> dict:
    model = {
        "ds_geo": 1, "ds_tex": 1, "v_level": 3, "xy_level": -1, "z_level": 4,
        "train_out_h": 64, "train_out_w": 64,
        "sp_args": {"sp_level": 3, "sp_type": "rel_z_decay", "scale": 1.0, "sigma": 0.1, "n_kpt": n_kpt},
        "geo_args": {"n_stack": 1, "n_downsample": 4, "out_ch": 64, "hd": False},
        "mlp_geo_args": {
            "n_dims1": [9, 128, 128, 120, 64], "n_dims2": [128, 64, 64, 2],
            "skip_dims": [64, 8], "skip_layers": [0, 2],
            "nl_layer": "softplus", "norm": "weight", "pool_types": ["mean", "var"], "dualheads": False,
        },
        "tex_args": {"ngf": 64, "n_downsample": 3, "n_blocks": 4, "n_upsample": 2, "out_ch": 8, "norm": "instance"},
        "mlp_tex_args": {"args": {"in_feat_ch": 32, "n_samples": 64}, "gcompress": {"in_ch": 128, "out_ch": 24}},
        "dr_level": 5,
        "dr_kwargs": {"fine": True, "uniform": False, "blur": 3, "rand_noise_std": 0.01,
                      "sample_per_ray_c": 64, "sample_per_ray_f": 64},
    }
    return {"models": {"KeypointNeRF": model}}


def load_cfg(path: str) -> dict:
    """Read a reference-style JSON config (``src/config.py:56-70`` accepts JSON or YAML; JSON here)."""
    with open(path) as f:
        return json.load(f)


def with_n_kpt(cfg: dict, n_kpt: int) -> dict:
    cfg = copy.deepcopy(cfg)
    cfg["models"]["KeypointNeRF"]["sp_args"]["n_kpt"] = n_kpt
    return cfg
