import sys
import numpy as np
sys.path.insert(0, ".")
from oracle.llama_replay import GGUFModel, Llama2Runner, load_weights
from oracle.tensor_ref import OracleDevice, OracleTensor
from tests.conftest import find_fixture
from crabml_b200 import CudaTensorDevice
from crabml_b200 import runner as R
path = find_fixture("tinyllamas-stories-15m-q8_0.gguf")
gm = GGUFModel(path)
odev = OracleDevice()
ro = Llama2Runner(OracleTensor, gm.conf, load_weights(gm, OracleTensor, odev), odev, 64)
for lazy in (False, True):
    dev = CudaTensorDevice(lazy=lazy)
    conf, w, tok = R.load_gguf(path, dev)
    r = R.LlamaRunner(dev, conf, w, 64)
    ro = Llama2Runner(OracleTensor, gm.conf, load_weights(gm, OracleTensor, odev), odev, 64)
    for pos, t in enumerate([1, 365, 2354, 338, 263, 274, 1082]):
        a = r.forward([t], pos).copy(); b = ro.forward([t], pos)
        print("lazy", lazy, "pos", pos, "rel", float(np.abs(a - b).max() / np.abs(b).max()), "argmax", a.argmax(), b.argmax(), dev.lazy_stats())
    r.close(); dev.close()
