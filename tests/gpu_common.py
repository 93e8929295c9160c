import numpy as np

from oracle.tensor_ref import OracleTensor


def make_device(**kw):
    from crabml_b200 import CudaTensorDevice
    return CudaTensorDevice(**kw)


def both(values, shape, gdev, odev):
    from crabml_b200 import CudaTensor
    v = np.asarray(values, np.float32)
    return CudaTensor.new(v, shape, gdev), OracleTensor.new(v, shape, odev)
