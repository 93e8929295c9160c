import ctypes as C, os, subprocess, sys
sys.path.insert(0, ".")
if len(sys.argv) > 1:
    from crabml_b200 import CudaTensorDevice
    dev = CudaTensorDevice(0)
    us = C.c_float(0)
    dev.check(dev.lib.cc_test_mega_barrier_floor(dev.handle, 512, C.byref(us)))
    print(f"flags={os.environ.get('CRABML_MEGA_FLAGS')}: {us.value:.3f} us per empty phase")
    dev.close()
else:
    for fl, name in ((0, "v0 two-level acq_rel atomics"), (8, "v0 without descriptor fetch"), (16, "v1 flat red.release + master, acquire polls"),
                     (32, "v2 flat red.release + master, relaxed polls + fence"), (16 + 8, "v1 without descriptor fetch")):
        env = dict(os.environ, CRABML_MEGA_FLAGS=str(fl))
        out = subprocess.run([sys.executable, __file__, "x"], env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
        print(name, "->", out)
