"""The arithmetic the decode kernel implements (``oracle/decode_oracle.py::decode_view_restated``, a numpy transcript of
``kpn_decode.cu``) is bit-identical to the reference's cv2 calls (``decode_view``, ``src/zju_dataset.py:266-287``)."""
import numpy as np
import pytest

from oracle import decode_oracle as DO


@pytest.mark.parametrize("factor,shape", [(2, (256, 320)), (1, (96, 128)), (4, (256, 256)), (3, (192, 240))])
def test_restated_decode_equals_cv2(factor, shape):
    imgs, msks, K, D = DO.synthetic_views(3, shape[0], shape[1], seed=factor)
    for v in range(3):
        want_img, want_msk, _ = DO.decode_view(imgs[v], msks[v], K[v], D[v], 1.0 / factor)
        got_img, got_msk = DO.decode_view_restated(imgs[v], msks[v], K[v], D[v], factor)
        assert got_img.shape == want_img.shape
        assert np.array_equal(got_msk, want_msk)
        assert np.abs(got_img - want_img).max() == 0.0
        assert 0.05 < want_msk.mean() < 0.9
