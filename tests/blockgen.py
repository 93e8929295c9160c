"""Seeded generators of VALID random GGUF quant blocks (test helper; numpy only).

Raw bytes are uniform; f16 scale fields are overwritten with finite, well-conditioned
values so that nothing is NaN/Inf and products stay in f32 range.
"""
import numpy as np

from oracle import oracle as oc

# byte offsets of f16 scale fields inside one block, per type
_F16_FIELDS = {
    oc.Q8_0: [0], oc.Q4_0: [0], oc.Q4_1: [0, 2], oc.Q5_0: [0], oc.Q5_1: [0, 2],
    oc.Q2_K: [80, 82], oc.Q3_K: [108], oc.Q4_K: [0, 2], oc.Q5_K: [0, 2], oc.Q6_K: [208],
}


def random_blocks(t, n_blocks, rng, scale=0.01):
    bb = oc.block_bytes(t)
    raw = rng.integers(0, 256, size=(n_blocks, bb), dtype=np.uint8)
    if t == oc.Q8_K:
        d = (scale * (0.5 + rng.random(n_blocks))).astype(np.float32)
        raw[:, 0:4] = d.view(np.uint8).reshape(n_blocks, 4)
        qs = raw[:, 4:260].view(np.int8).astype(np.int32)
        bs = qs.reshape(n_blocks, 16, 16).sum(-1).astype(np.int16)
        raw[:, 260:292] = bs.view(np.uint8).reshape(n_blocks, 32)
        return raw.reshape(-1)
    for off in _F16_FIELDS[t]:
        d = (scale * (0.5 + rng.random(n_blocks)) * rng.choice([-1.0, 1.0], n_blocks)).astype(np.float16)
        raw[:, off:off + 2] = d.view(np.uint8).reshape(n_blocks, 2)
    return raw.reshape(-1)


def random_weight(t, m, k, rng, scale=0.01):
    be = oc.block_elems(t)
    assert k % be == 0
    return random_blocks(t, m * (k // be), rng, scale)
