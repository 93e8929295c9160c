import ctypes as C, sys
sys.path.insert(0, ".")
from crabml_b200 import CudaTensorDevice
dev = CudaTensorDevice(0)
us = C.c_float(0)
for n in (64, 512):
    dev.check(dev.lib.cc_test_mega_barrier_floor(dev.handle, n, C.byref(us)))
    print(f"megakernel, {n} empty phases: {us.value:.3f} us per phase (descriptor fetch + grid barrier)")
dev.close()
