"""Pins the CPU oracle against the reference's own known-answer tests.

Every test names the reference test it restates (file:line in crabml/crabml @0151f893,
relative to crabml-core/src/cpu/).  Values are the literals asserted there.
"""
import numpy as np
import pytest

from oracle import oracle as oc
from oracle.tensor_ref import OracleDevice, OracleTensor, TensorError, TensorStrider


def f16b(x):
    return list(np.array([x], np.float16).view(np.uint8))


def test_block_sizes():
    # size_of assertions: buf_q4_0.rs:263, buf_q4_1.rs:289, buf_q5_0.rs:188, buf_q5_1.rs:186,
    # buf_q8_1.rs:139, buf_q8_k.rs:236, buf_q6_k.rs:250; Appendix A of SURVEY.md
    want = {oc.Q8_0: 34, oc.Q4_0: 18, oc.Q4_1: 20, oc.Q5_0: 22, oc.Q5_1: 24, oc.Q8_1: 36,
            oc.Q2_K: 84, oc.Q3_K: 110, oc.Q4_K: 144, oc.Q5_K: 176, oc.Q6_K: 210, oc.Q8_K: 292}
    for t, n in want.items():
        assert oc.block_bytes(t) == n, oc.TYPE_NAMES[t]
        assert oc.block_elems(t) == (32 if t in (oc.Q8_0, oc.Q4_0, oc.Q4_1, oc.Q5_0, oc.Q5_1, oc.Q8_1) else 256)


def test_rhs_dtype_pairing():
    # buf/api.rs:142-159
    for t in (oc.Q8_0, oc.Q4_0, oc.Q5_0): assert oc.rhs_type(t) == oc.Q8_0
    for t in (oc.Q4_1, oc.Q5_1, oc.Q8_1): assert oc.rhs_type(t) == oc.Q8_1
    for t in (oc.Q2_K, oc.Q3_K, oc.Q4_K, oc.Q5_K, oc.Q6_K, oc.Q8_K): assert oc.rhs_type(t) == oc.Q8_K


def test_q8_0_block_kat():
    # buf_q8_0.rs:293-322
    buf = np.full(68, 1, np.uint8)
    d = f16b(3.0)
    buf[0:2] = d; buf[2] = 2; buf[3] = 3; buf[4] = 4; buf[2 + 31] = 7
    buf[34:36] = d; buf[66] = 9; buf[67] = 9
    got = oc.dequantize(oc.Q8_0, buf, 64)
    want = [6.0, 9.0, 12.0] + [3.0] * 28 + [21.0] + [3.0] * 30 + [27.0, 27.0]
    assert got.tolist() == want


Q80_A = list(range(1, 33))
Q80_B = list(range(32, 0, -1))


def _q80(qs, d):
    return bytes(f16b(d)) + np.array(qs, np.int8).tobytes()


@pytest.mark.parametrize("flags", [0, oc.ORDER_AVX2])
def test_vec_dot_q8_0_q8_0_kat(flags):
    # buf_q8_0.rs:324-389: assert_eq! on both the scalar and the AVX2 CI legs
    a1 = np.frombuffer(_q80(Q80_A, 0.4), np.uint8)
    b1 = np.frombuffer(_q80(Q80_B, 1.3), np.uint8)
    assert oc.vec_dot(oc.Q8_0, a1, b1, 32, flags) == np.float32(3110.453)
    a2 = np.frombuffer(_q80(Q80_A, 0.4) + _q80([-v for v in Q80_A], 0.7), np.uint8)
    b2 = np.frombuffer(_q80(Q80_B, 1.3) + _q80([-v for v in Q80_B], 1.4), np.uint8)
    assert oc.vec_dot(oc.Q8_0, a2, b2, 64, flags) == np.float32(8978.046)


def test_q4_0_block_kat():
    # buf_q4_0.rs:259-298
    buf = np.full(36, 1, np.uint8)
    d = f16b(3.0)
    buf[0:2] = d; buf[2] = 2; buf[3] = 3; buf[4] = 4
    buf[18:20] = d; buf[20] = 2; buf[21] = 3; buf[22] = 4
    got = oc.dequantize(oc.Q4_0, buf, 64)
    one = [-18.0, -15.0, -12.0] + [-21.0] * 13 + [-24.0] * 16
    assert got.tolist() == one + one


def test_q4_1_block_and_quantize_kat():
    # buf_q4_1.rs:286-333: fields; dequantize of the ramp block is exact.  The reference's
    # BlockQ4_1::dequantize interleaves (B10) -- reproduced only under BUGCOMPAT.
    qs = [16, 50, 84, 118, 152, 186, 220, 254] * 2
    blk = np.array(f16b(1.0) + f16b(-8.0) + qs, np.uint8)
    ramp = [float(v) for v in range(-8, 8)] * 2
    assert oc.dequantize(oc.Q4_1, blk, 32, oc.BUGCOMPAT).tolist() == ramp
    # ggml / vec_dot order: low nibbles are elements 0..15, high nibbles 16..31
    lo = [float((q & 15) - 8) for q in qs]; hi = [float((q >> 4) - 8) for q in qs]
    assert oc.dequantize(oc.Q4_1, blk, 32).tolist() == lo + hi


def test_q5_0_kat():
    # buf_q5_0.rs:202-219: ramp quantizes to d=0.5, qs=[0,34,...]; dequantize rounds back to the ramp
    qs = [0, 34, 68, 102, 136, 170, 204, 238] * 2
    # xi = x/d*... the reference test only pins d and qs; rebuild qh from the quantizer rule (:100-140)
    ramp = np.array([float(v) for v in range(-8, 8)] * 2, np.float32)
    d = np.float32(-8.0) / np.float32(-16.0)
    xi = np.minimum((ramp * (np.float32(1.0) / d) + np.float32(16.5)).astype(np.int8), 31).astype(np.uint8)
    qh = 0
    for i in range(16):
        qh |= int((xi[i] & 0x10) >> 4) << i
        qh |= int((xi[i + 16] & 0x10) >> 4) << (i + 16)
    assert [int((xi[i] & 15) | ((xi[i + 16] & 15) << 4)) for i in range(16)] == qs
    blk = np.array(f16b(0.5) + list(np.array([qh], np.uint32).view(np.uint8)) + qs, np.uint8)
    got = oc.dequantize(oc.Q5_0, blk, 32)
    assert np.round(got).tolist() == ramp.tolist()


def test_q5_block_layout_kats():
    # buf_q5_0.rs:176-200, buf_q5_1.rs:174-202: d | (m) | qh[4] | qs[16]
    buf = np.full(22, 1, np.uint8); buf[0:2] = f16b(3.0); buf[2] = 2; buf[3] = 3; buf[4] = 4; buf[17] = 7
    # qh=[2,3,4,1] -> bits; qs[11]=7.  element 0: nibble 1 | bit0 of qh(=0) -> (1-16)*3
    got = oc.dequantize(oc.Q5_0, buf, 32)
    qh = int(np.array([2, 3, 4, 1], np.uint8).view(np.uint32)[0])
    want = []
    qsv = [1] * 16; qsv[11] = 7
    for half in range(2):
        for i in range(16):
            nib = (qsv[i] & 15) if half == 0 else (qsv[i] >> 4)
            bit = (qh >> (i + 16 * half)) & 1
            want.append(float(((nib | (bit << 4)) - 16) * 3.0))
    assert got.tolist() == want
    buf = np.full(24, 1, np.uint8); buf[0:2] = f16b(3.0); buf[2:4] = f16b(1.0); buf[4] = 2; buf[5] = 3; buf[6] = 4; buf[19] = 7
    got = oc.dequantize(oc.Q5_1, buf, 32)
    want = []
    for half in range(2):
        for i in range(16):
            nib = (qsv[i] & 15) if half == 0 else (qsv[i] >> 4)
            bit = (qh >> (i + 16 * half)) & 1
            want.append(float((nib | (bit << 4)) * 3.0 + 1.0))
    assert got.tolist() == want


def test_q8_1_block_kat():
    # buf_q8_1.rs:135-161
    buf = np.full(36, 1, np.uint8)
    buf[0:2] = f16b(3.0); buf[2:4] = f16b(96.0); buf[5] = 2; buf[6] = 3; buf[7] = 4; buf[35] = 7
    got = oc.dequantize(oc.Q8_1, buf, 32)
    assert got.tolist() == [3.0 * q for q in [1, 2, 3, 4] + [1] * 27 + [7]]


def test_q8_k_block_kat():
    # buf_q8_k.rs:232-262: d is the f32 whose bytes are f16(3.0) f16(1.0); bsums i16 pairs of 0x0101
    buf = np.full(292, 1, np.uint8)
    buf[0:2] = f16b(3.0); buf[2:4] = f16b(1.0); buf[4] = 2; buf[5] = 3; buf[6] = 4; buf[19] = 7; buf[283] = 10
    got = oc.dequantize(oc.Q8_K, buf, 256)
    d = np.float32(0.007828236)
    assert np.frombuffer(buf[0:4].tobytes(), np.float32)[0] == d
    assert (got[:16] == d * np.array([2, 3, 4, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 7], np.float32)).all()
    assert buf[260:292].view(np.int16).tolist() == [257] * 11 + [2561] + [257] * 4


def test_q8_k_quantize_kat():
    # buf_q8_k.rs:264-292: ramp -8..7 x16 -> d == 0.0625 and exact round trip
    data = np.array([float(v) for v in range(-8, 8)] * 16, np.float32)
    blk = oc.quantize(oc.Q8_K, data)
    assert np.frombuffer(blk[0:4].tobytes(), np.float32)[0] == np.float32(0.0625)
    assert oc.dequantize(oc.Q8_K, blk, 256).tolist() == data.tolist()
    bs = blk[260:292].view(np.int16)
    assert bs.tolist() == [int(np.sum(np.arange(-8, 8) * 16))] * 16


def test_q6_k_block_kat():
    # buf_q6_k.rs:246-276: d read from bytes 208..210
    buf = np.full(210, 1, np.uint8)
    buf[0:2] = f16b(3.0); buf[2:4] = f16b(1.0); buf[4] = 2; buf[5] = 3; buf[6] = 4; buf[19] = 7; buf[208] = 10
    assert buf[0:8].tolist() == [0, 66, 0, 60, 2, 3, 4, 1]
    d = np.array([int(buf[208]) | (int(buf[209]) << 8)], np.uint16).view(np.float16).astype(np.float32)[0]
    assert d == np.float32(1.5854836e-5)
    got = oc.dequantize(oc.Q6_K, buf, 256)
    # element 0: ql[0]=0 low nibble 0, qh[0]=1 -> (0 | 1<<4) - 32 = -16, scale 1
    assert got[0] == d * np.float32(1.0) * np.float32(-16.0)


def test_get_scale_min_k4_kat():
    # util.rs:352-359: all-ones scales -> sc=63, m=63 for j=0; exercised through Q4_K dequantize
    blk = np.zeros(144, np.uint8)
    blk[0:2] = f16b(1.0); blk[2:4] = f16b(1.0); blk[4:16] = 255; blk[16:] = 0x11
    got = oc.dequantize(oc.Q4_K, blk, 256)
    assert got[0] == 63.0 * 1.0 - 63.0


def test_quantize_q8_0_truncates():
    # buf_q8_0.rs:118-125 (B1): trunc toward zero, d = max/127 in f32, stored f16
    x = np.zeros(32, np.float32); x[0] = 127.0; x[1] = 1.9; x[2] = -1.9; x[3] = 0.99
    blk = oc.quantize(oc.Q8_0, x)
    assert blk[0:2].view(np.float16)[0] == np.float16(1.0)
    assert blk[2:6].view(np.int8).tolist() == [127, 1, -1, 0]
    z = oc.quantize(oc.Q8_0, np.zeros(32, np.float32))      # 0/0 -> NaN -> 0
    assert z.tolist() == [0] * 34


def test_quantize_q8_1_zero_block():
    # buf_q8_1.rs:110-114 (B2): NaN.max(-128) = -128 ; d = 0
    z = oc.quantize(oc.Q8_1, np.zeros(32, np.float32))
    assert z[0:2].view(np.float16)[0] == 0 and z[4:].view(np.int8).tolist() == [-128] * 32
    x = np.arange(32, dtype=np.float32)
    blk = oc.quantize(oc.Q8_1, x)
    d = np.float32(31.0) / np.float32(127.0)
    q = (x / d).astype(np.int8)
    assert blk[4:].view(np.int8).tolist() == q.tolist()
    assert blk[2:4].view(np.float16)[0] == np.float16(np.float32(q.astype(np.float32).sum()) * d)


def test_exp_lut_definition():
    # cpu_device.rs:108-115
    lut = oc.exp_lut()
    assert lut[np.array([0.0], np.float16).view(np.uint16)[0]] == np.array([1.0], np.float16).view(np.uint16)[0]
    assert lut[np.array([-np.inf], np.float16).view(np.uint16)[0]] == 0
    assert lut[np.array([1.0], np.float16).view(np.uint16)[0]] == np.array([np.e], np.float16).view(np.uint16)[0]


# ---- op-level KATs (cpu_tensor.rs:455-600) --------------------------------------------------
def test_tensor_view_and_copy_rows():
    # cpu_tensor.rs:455-482
    dev = OracleDevice()
    t = OracleTensor.new([1, 2, 3, 4, 5, 6], [2, 3], dev).reshape([3, 2]).reshape([2, 3])
    assert t.to_vec().tolist() == [1, 2, 3, 4, 5, 6]
    t1 = OracleTensor.new([1, 2, 3, 4], [2, 2], dev)
    t2 = OracleTensor.new([0, 0], [2], dev)
    t2.copy_rows_from(t1, [1]); assert t2.to_vec().tolist() == [3, 4]
    t2.copy_rows_from(t1, [0]); assert t2.to_vec().tolist() == [1, 2]


def test_rope_kat():
    # cpu_tensor.rs:509-527
    dev = OracleDevice()
    t = OracleTensor.new(np.arange(32, dtype=np.float32), [2, 16], dev).rope_inplace(0, 1, 2)
    want = [-0.841471, 0.54030234] + [float(v) for v in range(2, 16)] + [-5.6601696, 22.648676] + [float(v) for v in range(18, 32)]
    np.testing.assert_allclose(t.to_vec(), want, atol=1e-5)


def test_matmul_kat():
    # cpu_tensor.rs:530-541
    dev = OracleDevice()
    w = OracleTensor.new([4.0] * 32, [16, 2], dev)
    out = w.matmul_vec(OracleTensor.new([1.0, 2.0], [2], dev))
    assert out.to_vec().tolist() == [12.0] * 16


def test_matvec_32x8_kat():
    # crabml-wgpu/src/wgpu_tensor.rs:880-895 (exact F32 matvec)
    dev = OracleDevice()
    w = OracleTensor.new([float(i) for i in range(256)], [32, 8], dev)
    out = w.matmul_vec(OracleTensor.new([2.0] * 8, [8], dev)).to_vec()
    assert out.tolist() == [float(sum(range(8 * r, 8 * r + 8)) * 2) for r in range(32)]


def test_softmax_kat():
    # cpu_tensor.rs:544-555 (eps 1e-3 because of the f16 exp LUT)
    dev = OracleDevice()
    t = OracleTensor.new([1, 2, 3, 4, 5, 6], [2, 3], dev).softmax_inplace(1)
    np.testing.assert_allclose(t.to_vec(), [0.09003057, 0.24472848, 0.66524094] * 2, atol=1e-3)
    with pytest.raises(TensorError):
        OracleTensor.new([1, 2, 3, 4, 5, 6], [2, 3], dev).softmax_inplace(0)


def test_silu_kat():
    # cpu_tensor.rs:558-569
    dev = OracleDevice()
    t = OracleTensor.new([1, 2, 3, 4, 5, 6], [6], dev).silu_inplace()
    np.testing.assert_allclose(t.to_vec(), [0.7310586, 1.761594, 2.8577225, 3.928055, 4.9665356, 5.9851646], atol=1e-1)
    np.testing.assert_allclose(t.to_vec(), [0.7310586, 1.761594, 2.8577225, 3.928055, 4.9665356, 5.9851646], rtol=2e-3)


def test_rms_norm_kat():
    # crabml-wgpu/src/wgpu_tensor.rs:852-877: 1..128, eps 1e-5, against the closed form, eps 1e-7 rel
    dev = OracleDevice()
    v = np.arange(1, 129, dtype=np.float32)
    t = OracleTensor.new(v, [128], dev).rms_norm_inplace(1e-5)
    rms = np.sqrt(np.float32((v.astype(np.float64) ** 2).sum() / 128) + np.float32(1e-5))
    np.testing.assert_allclose(t.to_vec(), v / rms, rtol=3e-7)


def test_contiguous_kat():
    # cpu_tensor.rs:572-600
    dev = OracleDevice()
    t2 = OracleTensor.new([1, 2, 3, 4, 5, 6], [2, 3], dev).transpose([1, 0]).contiguous()
    assert t2.to_vec().tolist() == [1, 4, 2, 5, 3, 6] and t2.shape() == [3, 2]
    t1 = OracleTensor.new([1, 2, 3, 4, 5, 6], [1, 2, 3], dev).transpose([2, 1, 0])
    t2 = t1.contiguous()
    assert t2.to_vec().tolist() == [1, 4, 2, 5, 3, 6] and t2.shape() == [3, 2, 1]


def test_strider_kats():
    # strider.rs:238-339
    s = TensorStrider([3, 4])
    assert s.strides == [4, 1] and s.is_contiguous()
    s = TensorStrider([2, 3, 4])
    assert s.strides == [12, 4, 1]
    t = s.transpose([1, 0, 2])
    assert t.shape == [3, 2, 4] and t.strides == [4, 12, 1] and not t.is_contiguous()
    with pytest.raises(TensorError):
        t.reshape([24])
    r = TensorStrider([8, 2, 3200]).resize([8, 1, 3200])
    assert r.strides == [6400, 3200, 1] and r.shape == [8, 1, 3200]
    assert TensorStrider([1, 32, 128]).transpose([1, 0, 2]).is_contiguous() is False


def test_concatenate_kv_layout():
    # llama2.rs:542-554 + concatenate.rs: element (h, pos, z) lands at h*seq_max*hd + pos*hd + z
    dev = OracleDevice()
    cache = OracleTensor.alloc([2, 4, 3], 0, dev).resize(1, 0)
    for pos in range(3):
        k = OracleTensor.new(np.arange(6, dtype=np.float32) + 10 * pos, [1, 2, 3], dev).transpose([1, 0, 2])
        cache.concatenate(k, 1)
    assert cache.shape() == [2, 3, 3]
    buf = cache.buf.reshape(2, 4, 3)
    for pos in range(3):
        assert buf[0, pos].tolist() == [10 * pos + 0, 10 * pos + 1, 10 * pos + 2]
        assert buf[1, pos].tolist() == [10 * pos + 3, 10 * pos + 4, 10 * pos + 5]
