"""Second, independent pin for block unpack: python `gguf.quants.dequantize` (ggml layouts)
must agree BIT-EXACTLY with the oracle's dequantize on random valid blocks (SURVEY §8c,
Appendix D-5), and vec_dot must agree with the dot of the dequantized operands."""
import numpy as np
import pytest

from oracle import oracle as oc
from tests.blockgen import random_blocks

gguf = pytest.importorskip("gguf")
from gguf import GGMLQuantizationType as GQ, quants  # noqa: E402

GG = {oc.Q4_0: GQ.Q4_0, oc.Q4_1: GQ.Q4_1, oc.Q5_0: GQ.Q5_0, oc.Q5_1: GQ.Q5_1, oc.Q8_0: GQ.Q8_0,
      oc.Q2_K: GQ.Q2_K, oc.Q3_K: GQ.Q3_K, oc.Q4_K: GQ.Q4_K, oc.Q5_K: GQ.Q5_K, oc.Q6_K: GQ.Q6_K}


@pytest.mark.parametrize("t", sorted(GG))
def test_dequantize_bit_exact_vs_ggufpy(t):
    rng = np.random.default_rng(1000 + t)
    nb = 64
    raw = random_blocks(t, nb, rng)
    n = nb * oc.block_elems(t)
    ours = oc.dequantize(t, raw, n)
    theirs = quants.dequantize(raw.reshape(nb, -1), GG[t]).reshape(-1).astype(np.float32)
    assert np.isfinite(ours).all()
    if t in (oc.Q2_K, oc.Q3_K, oc.Q4_K, oc.Q5_K, oc.Q6_K, oc.Q4_1, oc.Q5_1):
        # gguf-py associates (d*sc)*q - (dmin*m) exactly like the reference; equality is bitwise
        # except where numpy evaluates d*q+m with an fma-free but differently ordered expression
        np.testing.assert_array_equal(ours.view(np.uint32), theirs.view(np.uint32))
    else:
        np.testing.assert_array_equal(ours.view(np.uint32), theirs.view(np.uint32))


@pytest.mark.parametrize("t", sorted(GG) + [oc.Q8_K])
def test_vec_dot_matches_dequantized_dot(t):
    rng = np.random.default_rng(2000 + t)
    k = 2048
    w = random_blocks(t, k // oc.block_elems(t), rng, scale=0.02)
    x = rng.standard_normal(k).astype(np.float32)
    act = oc.quantize(oc.rhs_type(t), x)
    got = float(oc.vec_dot(t, w, act, k))
    wd = oc.dequantize(t, w, k).astype(np.float64)
    ad = oc.dequantize(oc.rhs_type(t), act, k).astype(np.float64)
    want = float((wd * ad).sum())
    scale = float(np.abs(wd * ad).sum())
    tol = 2e-3 if t in (oc.Q4_1, oc.Q5_1) else 1e-5      # Q4_1/Q5_1 round d*d and m*s to f16 (buf_q4_1.rs:276)
    assert abs(got - want) <= tol * scale + 1e-6, (got, want)


@pytest.mark.parametrize("t", [oc.Q8_0, oc.Q4_0])
def test_avx2_order_close_to_scalar(t):
    rng = np.random.default_rng(7)
    k = 4096
    w = random_blocks(t, k // 32, rng)
    act = oc.quantize(oc.Q8_0, rng.standard_normal(k).astype(np.float32))
    a = float(oc.vec_dot(t, w, act, k, 0)); b = float(oc.vec_dot(t, w, act, k, oc.ORDER_AVX2))
    assert abs(a - b) <= 1e-5 * max(1.0, abs(a))


def test_q4k_bugcompat_differs_only_on_wrap():
    # B7: i16 wrap of bsum*min.  With tiny activations no product exceeds i16 and both agree.
    rng = np.random.default_rng(3)
    k = 1024
    w = random_blocks(oc.Q4_K, k // 256, rng)
    x = rng.standard_normal(k).astype(np.float32)
    act = oc.quantize(oc.Q8_K, x)
    bs = act.reshape(-1, 292)[:, 260:].view(np.int16)
    a, b = oc.vec_dot(oc.Q4_K, w, act, k, 0), oc.vec_dot(oc.Q4_K, w, act, k, oc.BUGCOMPAT)
    if (np.abs(bs.astype(np.int32)) * 63 < 32768).all():
        assert a == b


def test_gemv_row_split_matches_single_thread():
    # matmul_vec.rs:41-77: thread split must not change results
    rng = np.random.default_rng(11)
    m, k = 50, 288
    w = random_blocks(oc.Q8_0, m * k // 32, rng)
    x = rng.standard_normal(k).astype(np.float32)
    one = oc.gemv(oc.Q8_0, w, m, k, x, threads=1)
    for th in (2, 3, 8):
        np.testing.assert_array_equal(one, oc.gemv(oc.Q8_0, w, m, k, x, threads=th))
