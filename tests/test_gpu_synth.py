"""The on-device synthetic weight generator and its CPU twin (oracle/synth.py) produce identical GGUF bytes,
so 7B-shape parity checks and the CPU baseline run on exactly the weights the GPU uses."""
import numpy as np
import pytest

from oracle import oracle as oc
from oracle.synth import synth_weight
from tests.gpu_common import make_device

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("t", oc.QUANT_TYPES)
def test_device_synth_matches_cpu_twin(t):
    from crabml_b200 import CudaTensor
    dev = make_device()
    try:
        rows, cols = 13, 1024
        g = CudaTensor.synth([rows, cols], t, dev, 0x5EED, 7 + t, 0.0123)
        want = synth_weight(t, rows, cols, 0x5EED, 7 + t, 0.0123)
        got = g.export_blocks(want.size)
        if t == oc.Q8_K:
            got.reshape(-1, 292)[:, 260:] = want.reshape(-1, 292)[:, 260:]      # weight-side bsums are not stored
        np.testing.assert_array_equal(got, want)
        assert np.isfinite(oc.dequantize(t, want, rows * cols)).all()
    finally:
        dev.close()
