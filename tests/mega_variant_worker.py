"""Worker of tests/test_gpu_runner.py::test_both_persistent_kernels_bit_identical_on_7b_shapes: decodes a few tokens of a 2-layer
Llama-2-7B-shaped synthetic model with the execution mode / megakernel flag word of its environment and saves the logits.
(The flag word is read once per process: CRABML_MEGA_FLAGS, csrc/mega.cu.)"""
import sys

import numpy as np

sys.path.insert(0, ".")
from crabml_b200 import CudaTensorDevice, capi  # noqa: E402
from crabml_b200 import runner as R  # noqa: E402


def main():
    lazy, wt, ct, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    dev = CudaTensorDevice(0, lazy=lazy)
    try:
        conf = R.LlamaConfig(32, 32, 2, 4096, 11008, 4096, 32000, 1e-5, 128)
        w = R.synthetic_weights(dev, conf, wt, ct, seed=11)
        r = R.LlamaRunner(dev, conf, w, 16)
        logits = np.stack([r.forward([t], p).copy() for p, t in enumerate([1, 777, 31999, 5, 6, 9])])
        variant = dev.mega_variant() if lazy else 0
        r.close()
    finally:
        dev.close()
    np.savez(out, logits=logits, variant=np.int64(variant))


if __name__ == "__main__":
    main()
