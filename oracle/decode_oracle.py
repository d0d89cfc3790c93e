"""CPU oracle of the source-view decode.  TEST INFRASTRUCTURE ONLY.

The reference decodes every view with OpenCV on CPU workers (``/root/reference/src/zju_dataset.py:266-287``); this file restates
exactly those calls (cv2 is the third-party dependency the reference itself uses; it is present in this image and on the GPU box),
so the CUDA kernel (``keypointnerf_b200/csrc/kpn_decode.cu``) is compared with the reference's own arithmetic, bit for bit.
"""
import cv2
import numpy as np


def decode_view(img_u8: np.ndarray, msk_u8: np.ndarray, K: np.ndarray, D: np.ndarray, ratio: float):
    """One view: (H0,W0,3) uint8 RGB + (H0,W0) uint8 mask -> ((3,H,W) float32, (1,H,W) bool, K scaled)."""
    in_K, in_D = np.array(K).astype(np.float32), np.array(D).astype(np.float32)
    input_img = img_u8.astype(np.float32) / 255.                                                   # :269
    input_msk = (msk_u8 != 0).astype(np.uint8)                                                      # :200-211
    input_img, input_msk = cv2.undistort(input_img, in_K, in_D), cv2.undistort(input_msk, in_K, in_D)   # :270
    H, W = int(input_img.shape[0] * ratio), int(input_img.shape[1] * ratio)                        # :273
    input_img = cv2.resize(input_img, (W, H), interpolation=cv2.INTER_AREA)                          # :274
    input_msk = cv2.resize(input_msk, (W, H), interpolation=cv2.INTER_NEAREST)
    input_img[input_msk == 0] = 0                                                                  # :277
    input_msk = (input_msk != 0)
    in_K = in_K.copy()
    in_K[:2] = in_K[:2] * ratio                                                                     # :296
    return np.ascontiguousarray(input_img.transpose(2, 0, 1)), input_msk[None], in_K             # image2tensor: HWC -> CHW


def synthetic_views(n_views: int = 3, h: int = 256, w: int = 320, seed: int = 0):
    """Random images, blob masks (one touching the border) and mildly distorting cameras."""
    rng = np.random.default_rng(seed)
    imgs = rng.integers(0, 256, (n_views, h, w, 3), dtype=np.uint8)
    msks = np.zeros((n_views, h, w), np.uint8)
    K = np.zeros((n_views, 3, 3), np.float32)
    D = np.zeros((n_views, 5), np.float32)
    for v in range(n_views):
        cv2.ellipse(msks[v], (int(w * 0.45) + 7 * v, int(h * 0.5) - 5 * v), (int(w * 0.25), int(h * 0.4)), 20 * v, 0, 360, 255, -1)
        if v == 0:
            msks[v, : h // 6, : w // 5] = 1
        K[v] = [[w * 0.85 + 3 * v, 0, w * 0.5 + 1.3 * v], [0, w * 0.84 - 2 * v, h * 0.5 - 0.9 * v], [0, 0, 1]]
        D[v] = [-0.22 + 0.05 * v, 0.19 - 0.03 * v, 0.001 * (v + 1), -0.0007 * (v + 1), 0.03 - 0.01 * v]
    return imgs, msks, K, D
