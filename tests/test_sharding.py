"""Host logic of the sharded decode path (crabml_b200/sharding.py): the shard plan and the GGUF byte slicing.
CPU only; the oracle's dequantiser is the checker."""
import numpy as np
import pytest

from crabml_b200 import sharding as S
from oracle import oracle as oc
from oracle.synth import synth_weight


def test_deal_covers_everything_once():
    for total in (1, 7, 43, 344, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [S.deal(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (a0, ac), (b0, _) in zip(spans, spans[1:]):
                assert a0 + ac == b0
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_llama2_7b_plans():
    # Q8_0: 11008 / 8 = 1376 = 43 blocks of 32 each
    plans = [S.make_plan(32, 32, 4096, 11008, 32000, oc.Q8_0, r, 8) for r in range(8)]
    assert all(p.hidden_local == 1376 and p.q_rows[1] == 512 and p.vocab_rows[1] == 4000 for p in plans)
    # K-quants: 43 super-blocks of 256 -> 6,6,6,5,5,5,5,5
    plans = [S.make_plan(32, 32, 4096, 11008, 32000, oc.Q4_K, r, 8) for r in range(8)]
    assert [p.hidden_local // 256 for p in plans] == [6, 6, 6, 5, 5, 5, 5, 5]
    assert sum(p.hidden_local for p in plans) == 11008
    assert S.make_plan(32, 32, 4096, 11008, 32000, oc.Q8_0, 0, 1).hidden == (0, 11008)


def test_plan_rejects_what_cannot_be_sharded():
    with pytest.raises(ValueError):          # tinyllamas: 3 heads x 48 = 144 columns of wo, not a multiple of 32
        S.make_plan(6, 6, 288, 768, 32000, oc.Q8_0, 0, 2)
    with pytest.raises(ValueError):          # heads do not divide
        S.make_plan(6, 6, 288, 768, 32000, oc.Q8_0, 0, 4)
    with pytest.raises(ValueError):          # Mistral GQA with the F32 cache: kv head = h % n_kv is not local (quirk B13)
        S.make_plan(32, 8, 4096, 14336, 32000, oc.Q8_0, 0, 2, f16_kv=False)
    S.make_plan(32, 8, 4096, 14336, 32000, oc.Q8_0, 0, 2, f16_kv=True)
    with pytest.raises(ValueError):
        S.make_plan(32, 32, 4096, 11008, 32000, oc.Q8_0, 8, 8)


@pytest.mark.parametrize("t", [oc.Q8_0, oc.Q4_0, oc.Q4_1, oc.Q5_0, oc.Q4_K, oc.Q6_K])
def test_byte_slices_are_the_matrix_slices(t):
    be = oc.block_elems(t)
    rows, cols = 12, be * 6
    raw = synth_weight(t, rows, cols, 11, 3, 0.01)
    full = oc.dequantize(t, raw, rows * cols).reshape(rows, cols)
    r = S.slice_rows(raw, rows, cols, t, 4, 5)
    np.testing.assert_array_equal(oc.dequantize(t, r, 5 * cols).reshape(5, cols), full[4:9])
    c = S.slice_cols(raw, rows, cols, t, 2 * be, 3 * be)
    np.testing.assert_array_equal(oc.dequantize(t, c, rows * 3 * be).reshape(rows, 3 * be), full[:, 2 * be:5 * be])
    with pytest.raises(ValueError):
        S.slice_cols(raw, rows, cols, t, be // 2, be)


def test_column_split_matvec_partials_sum_to_the_full_row():
    """The identity the allreduce relies on, on the oracle: sum over ranks of (column shard) . (activation shard) == full
    dot up to the f32 grouping of the partial sums (quantisation is per 32-block, so shards see the same blocks)."""
    t, rows, cols, world = oc.Q8_0, 16, 32 * 12, 4
    raw = synth_weight(t, rows, cols, 5, 1, 0.01)
    x = np.random.default_rng(0).standard_normal(cols).astype(np.float32)
    full = oc.gemv(t, raw, rows, cols, x)
    acc = np.zeros(rows, np.float32)
    for r in range(world):
        c0, cn = S.deal(cols // 32, world, r)
        part = S.slice_cols(raw, rows, cols, t, c0 * 32, cn * 32)
        acc = acc + oc.gemv(t, part, rows, cn * 32, x[c0 * 32:(c0 + cn) * 32])
    np.testing.assert_allclose(acc, full, rtol=0, atol=2e-6 * float(np.abs(full).max() + 1))
