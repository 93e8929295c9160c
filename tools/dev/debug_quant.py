import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import oracle as oc
from crabml_b200 import CudaTensor, CudaTensorDevice
dev = CudaTensorDevice()
rng = np.random.default_rng(13)
n = 256 * 12
x = (rng.standard_normal(n) * rng.choice([1e-3, 1.0, 30.0], n)).astype(np.float32)
x[0:256] = 0.0
x[256:512] = np.tile(np.arange(-8, 8, dtype=np.float32), 16)
x[512] = -5.0; x[513] = 5.0
x[512 + 2:768] = 0.25
x[768:800] = 127.0
x[800:832] = [(-1) ** i * (i + 0.5) for i in range(32)]
for rep in range(2):
    gx = CudaTensor.new(x, [n], dev)
    want = oc.quantize(oc.Q8_0, x).reshape(-1, 34)
    got = gx.quantize_activation(oc.Q8_0, want.size).reshape(-1, 34)
    bad = np.argwhere(got != want)
    print("rep", rep, "mismatches", len(bad))
    for r, c in bad[:12]:
        xv = x[r * 32 + c - 2] if c >= 2 else None
        print(" blk", r, "byte", c, "got", got[r, c], "want", want[r, c], "x", xv, "d_want", want[r, :2].view(np.float16)[0], "d_got", got[r, :2].view(np.float16)[0])
