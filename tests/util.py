"""Shared helpers for the tests: golden loading and scene reconstruction from fixture metadata."""
import hashlib
import json
import os

import numpy as np

from keypointnerf_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    meta = json.loads(bytes(g.pop("meta")).decode())
    sha = bytes(g.pop("input_sha256")).decode()
    return g, meta, sha


def scene_from_meta(meta):
    scene = syn.make_scene(src_size=meta["src_size"], n_views=3, n_kpt=meta["n_kpt"], seed=meta["scene_seed"],
                           fg_hole=meta["fg_hole"], fg_mode=meta.get("fg_mode", "ones"))
    weights = syn.make_weights(meta["n_kpt"], seed=meta["w_seed"])
    target = syn.make_target(size=meta["tgt_size"], azimuth=meta["azimuth"], zoom=meta.get("zoom", 1.0))
    return scene, weights, target


def checksum(scene, weights):
    h = hashlib.sha256()
    for k in ("feat64", "feat8", "feat_tex", "img", "fg", "kpt3d", "KRT", "extrin"):
        h.update(np.ascontiguousarray(scene[k]).tobytes())
    for k in sorted(weights):
        h.update(np.ascontiguousarray(weights[k]).tobytes())
    return h.hexdigest()


def psnr(a, b):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)
