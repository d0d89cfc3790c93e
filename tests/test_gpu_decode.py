"""Source-view decode: the CUDA kernel behind ``kpn_decode_views`` against the reference's own cv2 calls
(``oracle/decode_oracle.py`` restates ``src/zju_dataset.py:266-287``): bit-identical images and masks, device and host paths."""
import numpy as np
import pytest
import torch

from keypointnerf_b200.renderer import RayMarcher
from oracle import decode_oracle as DO  # checker only

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ratio,shape", [(0.5, (256, 320)), (1.0, (128, 160)), (0.25, (256, 256)), (1.0 / 3.0, (192, 240))])
def test_decode_matches_cv2_bit_for_bit(ratio, shape):
    imgs, msks, K, D = DO.synthetic_views(3, shape[0], shape[1], seed=int(ratio * 8))
    want = [DO.decode_view(imgs[v], msks[v], K[v], D[v], ratio) for v in range(3)]
    m = RayMarcher(0)
    for dev in ("cuda:0", "cpu"):
        ti = torch.from_numpy(imgs).to(dev)
        tm = torch.from_numpy(msks).to(dev)
        img, msk, Ks = m.decode_views(ti, tm, K, D, ratio)
        torch.cuda.synchronize()
        assert img.shape == (3, 3, want[0][0].shape[1], want[0][0].shape[2]) and msk.dtype == torch.bool
        for v in range(3):
            assert np.array_equal(msk[v].cpu().numpy(), want[v][1]), (dev, v, "mask")
            d = np.abs(img[v].cpu().numpy() - want[v][0])
            assert d.max() == 0.0, (dev, v, float(d.max()), float((d > 0).mean()))
            assert np.array_equal(Ks[v].numpy(), want[v][2])
        assert 0.05 < float(msk.float().mean()) < 0.9
    # no mask: every pixel counts as foreground, the image is only undistorted and resized; where an all-ones mask survives the
    # undistortion (inside the source image) the result equals the masked decode
    img, msk, _ = m.decode_views(torch.from_numpy(imgs).cuda(), None, K, D, ratio)
    ref_img, ref_msk, _ = DO.decode_view(imgs[1], np.full_like(msks[0], 255), K[1], D[1], ratio)
    keep = np.broadcast_to(ref_msk, ref_img.shape)
    assert bool(msk.all()) and np.abs(img[1].cpu().numpy() - ref_img)[keep].max() == 0.0
