"""Developer aid: per-tap divergence of the CUDA replay vs the oracle on the tiny model."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle.llama_replay import GGUFModel, Llama2Runner, load_weights  # noqa: E402
from oracle.tensor_ref import OracleDevice, OracleTensor  # noqa: E402
from tests.conftest import find_fixture  # noqa: E402

from crabml_b200 import CudaTensor, CudaTensorDevice  # noqa: E402

gm = GGUFModel(find_fixture(sys.argv[1] if len(sys.argv) > 1 else "tinyllamas-stories-15m-q8_0.gguf"))
gdev, odev = CudaTensorDevice(debug_named_tensors=True), OracleDevice(debug_named_tensors=True)
rg = Llama2Runner(CudaTensor, gm.conf, load_weights(gm, CudaTensor, gdev), gdev, 200)
ro = Llama2Runner(OracleTensor, gm.conf, load_weights(gm, OracleTensor, odev), odev, 200)
for pos, tok in enumerate([1, 365, 2354]):
    lg, lo = rg.forward([tok], pos), ro.forward([tok], pos)
    print(f"pos {pos} logits rel diff {np.abs(lg - lo).max() / np.abs(lo).max():.3e}")
    for l in range(gm.conf.n_layers):
        for name in (f"attn_rmsnorm:{l}:{pos}", f"attn_out:{l}:{pos}", f"ffn_out:{l}:{pos}"):
            a, b = gdev.dump_debug_tensor(name), odev.dump_debug_tensor(name)
            print(f"  {name:24s} max|d| {np.abs(a - b).max():.3e}  max|ref| {np.abs(b).max():.3e}")
