"""Host-side checks of two contracts of the view-sequential geometry kernel (keypointnerf_b200/csrc/kpn_shade_tc.cu) that a GPU run
would only show as a slowdown or as a subtle numerical drift:

  * the shared-memory staging layout `vs_f64_word` (producers write eight float4 groups of a row from eight adjacent lanes, the row
    warps read one word of 32 consecutive rows): a bijection per word, and free of bank conflicts on both sides;
  * the thread-local view pooling: mean = S1, var = S2 - S1^2 (2 - sum pw) with S1 = sum pw x, S2 = sum pw x^2 is the reference's
    weighted mean / variance (reference src/utils.py:722-748, restated in oracle/kpnerf_oracle.py) for weights that do NOT sum to 1
    exactly (they are normalised with a 1e-6 guard, reference src/model.py:757-759).
"""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "keypointnerf_b200", "csrc", "kpn_shade_tc.cu")).read()


def vs_f64_word(w, row):
    return w * 128 + (row ^ (4 * ((w >> 1) & 7)))


def test_python_restatement_matches_the_source():
    m = re.search(r"int vs_f64_word\(int w, int row\) \{ return (.*?); \}", SRC)
    assert m, "vs_f64_word not found"
    assert m.group(1).replace(" ", "") == "w*128+(row^(4*((w>>1)&7)))"
    # the producers' store address is the same function written out for words 2g, 2g+1, 2g+16, 2g+17 of group pair (g, 8+g)
    assert "f64b + (2 * l8) * 128 + (row ^ (4 * l8))" in SRC
    assert "d[16 * 128] = w0; d[17 * 128] = w1;" in SRC


def test_staging_words_are_a_bijection_per_word():
    for w in range(32):
        idx = {vs_f64_word(w, r) for r in range(128)}
        assert idx == set(range(w * 128, (w + 1) * 128))


def test_no_bank_conflicts_on_either_side():
    # consumer: lanes = 32 consecutive rows of one lane quarter, one word
    for w in range(32):
        for q4 in range(4):
            banks = {vs_f64_word(w, 32 * q4 + lane) % 32 for lane in range(32)}
            assert len(banks) == 32
    # producer: one store instruction of a warp = lanes (row 4u + l/8, group g = l%8), word 2g + i for fixed i in {0, 1, 16, 17}
    for u in range(32):
        for i in (0, 1, 16, 17):
            banks = {vs_f64_word(2 * (lane % 8) + i, 4 * u + lane // 8) % 32 for lane in range(32)}
            assert len(banks) == 32


def test_running_sum_pooling_equals_weighted_mean_and_variance():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1000, 3, 64))
    pw = rng.random((1000, 3, 1))
    pw = pw / (pw.sum(1, keepdims=True) + 1e-6)          # the reference's normalisation: sums to slightly less than 1
    mean = (pw * x).sum(1)
    var = (pw * (x - mean[:, None]) ** 2).sum(1)
    s1 = (pw * x).sum(1)
    s2 = (pw * x * x).sum(1)
    var2 = s2 - s1 * s1 * (2.0 - pw.sum(1))
    np.testing.assert_allclose(s1, mean, rtol=0, atol=0)
    np.testing.assert_allclose(var2, var, rtol=1e-9, atol=1e-12)
