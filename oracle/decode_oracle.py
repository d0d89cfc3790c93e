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


def decode_view_restated(img_u8: np.ndarray, msk_u8: np.ndarray, K: np.ndarray, D: np.ndarray, factor: int):
    """The arithmetic of ``kpn_decode.cu`` written out in numpy (same operation order, no OpenCV): pinned to ``decode_view`` by
    ``tests/test_decode_cpu.py``, so the kernel's specification is checked against cv2 without a GPU."""
    H0, W0 = msk_u8.shape
    Kd = np.array(K).astype(np.float32).astype(np.float64)
    fx, fy, cx, cy = Kd[0, 0], Kd[1, 1], Kd[0, 2], Kd[1, 2]
    k1, k2, p1, p2, k3 = [float(v) for v in np.array(D).astype(np.float32).astype(np.float64)[:5]]
    ir = np.linalg.inv(Kd)
    u, v = np.meshgrid(np.arange(W0, dtype=np.float64), np.arange(H0, dtype=np.float64))
    X = u * ir[0, 0] + v * ir[0, 1] + ir[0, 2]
    Y = u * ir[1, 0] + v * ir[1, 1] + ir[1, 2]
    Wc = u * ir[2, 0] + v * ir[2, 1] + ir[2, 2]
    x, y = X / Wc, Y / Wc
    x2, y2 = x * x, y * y
    r2, xy2 = x2 + y2, 2 * x * y
    kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    xd = x * kr + p1 * xy2 + p2 * (r2 + 2 * x2)
    yd = y * kr + p1 * (r2 + 2 * y2) + p2 * xy2
    iu = np.rint((fx * xd + cx) * 32).astype(np.int64)      # cvRound: half to even
    iv = np.rint((fy * yd + cy) * 32).astype(np.int64)
    sx, sy, a, b = iu >> 5, iv >> 5, iu & 31, iv & 31
    fxq, fyq = a.astype(np.float32) / 32, b.astype(np.float32) / 32

    def px(im, yy, xx):
        ok = (yy >= 0) & (yy < H0) & (xx >= 0) & (xx < W0)
        out = np.zeros(yy.shape + im.shape[2:], im.dtype)
        out[ok] = im[yy[ok], xx[ok]]
        return out

    w = [(1 - fxq) * (1 - fyq), fxq * (1 - fyq), (1 - fxq) * fyq, fxq * fyq]
    imgf = img_u8.astype(np.float32) / np.float32(255.)
    taps = [px(imgf, sy, sx), px(imgf, sy, sx + 1), px(imgf, sy + 1, sx), px(imgf, sy + 1, sx + 1)]
    und = taps[0] * w[0][..., None] + taps[1] * w[1][..., None] + taps[2] * w[2][..., None] + taps[3] * w[3][..., None]
    wi = [(32 - a) * (32 - b) * 32, a * (32 - b) * 32, (32 - a) * b * 32, a * b * 32]          # 15-bit fixed point, exact
    m = (msk_u8 != 0).astype(np.uint8)
    mt = [px(m, sy, sx), px(m, sy, sx + 1), px(m, sy + 1, sx), px(m, sy + 1, sx + 1)]
    um = (sum(t.astype(np.int64) * q for t, q in zip(mt, wi)) + (1 << 14)) >> 15
    H, W = H0 // factor, W0 // factor
    box = [und[dy:H * factor:factor, dx:W * factor:factor] for dy in range(factor) for dx in range(factor)]
    acc, k = np.zeros((H, W, 3), np.float32), 0
    while k <= len(box) - 4:                                 # OpenCV's resizeAreaFast sums four taps at a time
        acc = acc + (((box[k] + box[k + 1]) + box[k + 2]) + box[k + 3])
        k += 4
    while k < len(box):
        acc = acc + box[k]
        k += 1
    out = acc * np.float32(1.0 / (factor * factor)) if factor > 1 else acc
    fg = um[0:H * factor:factor, 0:W * factor:factor] != 0
    out = out.copy()
    out[~fg] = 0
    return np.ascontiguousarray(out.transpose(2, 0, 1)), fg[None]
