"""OracleTensor: the reference's `Tensor` trait (crabml-core/src/tensor/api.rs:11-79) as
implemented by `CpuTensor` (crabml-core/src/cpu/cpu_tensor.rs:126-446), restated on numpy +
the C oracle.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).

Method names, argument meaning and error behaviour follow the trait so that the parity tests
read like the reference's own tests and the same replay code (oracle/llama_replay.py) can drive
either this class or the CUDA mirror.
"""
from __future__ import annotations

import numpy as np

from . import oracle as oc


class TensorError(Exception):
    """ErrorKind::TensorError (crabml-core/src/error.rs:24-25)."""


class TensorStrider:
    """crabml-core/src/tensor/strider.rs:5-236."""

    def __init__(self, shape, strides=None):
        self.shape = [int(s) for s in shape]
        self.strides = self._compute(self.shape) if strides is None else [int(s) for s in strides]

    @staticmethod
    def _compute(shape):                                     # strider.rs:216-224
        strides = [1]
        for i in range(len(shape) - 1):
            strides.append(strides[-1] * shape[len(shape) - i - 1])
        return strides[::-1]

    def clone(self):
        return TensorStrider(self.shape, self.strides)

    def dims(self): return len(self.shape)
    def len(self): return int(np.prod(self.shape)) if self.shape else 1

    def resize(self, new_shape):                             # strider.rs:36-51
        if len(new_shape) != len(self.shape):
            raise TensorError(f"invalid new shape {new_shape} for a tensor of shape {self.shape}")
        return TensorStrider(new_shape, self.strides)

    def reshape(self, shape):                                # strider.rs:143-160
        if not self.is_contiguous():
            raise TensorError("not contiguous")
        if int(np.prod(shape)) != self.len():
            raise TensorError(f"invalid shape {shape} for a tensor's origin shape {self.shape}")
        return TensorStrider(shape)

    def transpose(self, dims):                               # strider.rs:162-180
        if len(dims) != len(self.shape):
            raise TensorError(f"invalid dims {dims} for a tensor of shape {self.shape}")
        return TensorStrider([self.shape[d] for d in dims], [self.strides[d] for d in dims])

    def is_contiguous(self):                                 # strider.rs:182-206
        if not self.strides:
            return True
        if self.strides[-1] != 1:
            return False
        last = 1
        for i in reversed(range(len(self.shape))):
            if last != self.strides[i]:
                return False
            last *= self.shape[i]
        return True


class OracleDevice:
    """CpuTensorDevice (cpu_device.rs): options + debug-tensor tap; LUTs live in oracle.py."""

    def __init__(self, debug_named_tensors=False, thread_num=1, flags=0):
        self.debug_named_tensors = debug_named_tensors
        self.thread_num = thread_num
        self.flags = flags
        self.debug_tensors = {}

    def dump_debug_tensor(self, name):                       # cpu_device.rs:96-98
        return self.debug_tensors.get(name)


class OracleTensor:
    """buf: flat numpy array (float32 / uint16 for F16 / uint8 raw blocks for quant types)."""

    def __init__(self, buf, strider, dtype, device, owned, name=None):
        self.buf, self._strider, self._dtype, self.device, self.owned, self.name = buf, strider, dtype, device, owned, name

    # -- constructors --------------------------------------------------------------------
    @classmethod
    def from_cpu(cls, buf, shape, dtype, device):            # api.rs:14-19 (GPU backends copy host bytes)
        raw = np.frombuffer(bytes(buf), np.uint8).copy() if not isinstance(buf, np.ndarray) else np.ascontiguousarray(buf).view(np.uint8).reshape(-1).copy()
        n = int(np.prod(shape))
        need = oc.nbytes_for(dtype, n) if dtype not in (oc.F32, oc.F16) else n * (4 if dtype == oc.F32 else 2)
        if raw.size < need:
            raise TensorError(f"from_cpu: {raw.size} bytes < {need} needed for shape {shape}")
        raw = raw[:need]                                     # gguf.rs:742-747 slices may carry padding (B16)
        if dtype == oc.F32:
            data = raw.view(np.float32)
        elif dtype == oc.F16:
            data = raw.view(np.uint16)
        else:
            data = raw
        return cls(data, TensorStrider(shape), dtype, device, owned=False)

    @classmethod
    def new(cls, values, shape, device):                     # CpuTensor::new, cpu_tensor.rs:29-46
        values = np.asarray(values, np.float32).reshape(-1).copy()
        if values.size != int(np.prod(shape)):
            raise TensorError(f"invalid shape {shape} for data of length {values.size}")
        return cls(values, TensorStrider(shape), oc.F32, device, owned=True)

    @classmethod
    def alloc(cls, shape, dtype, device):                    # cpu_tensor.rs:138-165
        if dtype not in (oc.F32, oc.F16):
            raise TensorError("only f32/f16 is supported")
        n = int(np.prod(shape))
        buf = np.zeros(n, np.float32 if dtype == oc.F32 else np.uint16)
        return cls(buf, TensorStrider(shape), dtype, device, owned=True)

    # -- metadata ------------------------------------------------------------------------
    def dtype(self): return self._dtype
    def shape(self): return list(self._strider.shape)
    def strider(self): return self._strider
    def is_contiguous(self): return self._strider.is_contiguous()
    def _with(self, strider): return OracleTensor(self.buf, strider, self._dtype, self.device, self.owned, None)

    def resize(self, axis, n):                               # cpu_tensor.rs:167-197
        if axis >= len(self.shape()):
            raise TensorError(f"resize: axis {axis} is larger than the current shape {self.shape()}")
        new_shape = self.shape(); new_shape[axis] = n
        if int(np.prod(new_shape)) > self.buf.size:
            raise TensorError(f"resize: new shape {new_shape} is larger than the current shape {self.shape()}")
        return self._with(self._strider.resize(new_shape))

    def with_strider(self, strider): return self._with(strider.clone())
    def reshape(self, shape): return self._with(self._strider.reshape(list(shape)))
    def transpose(self, dims): return self._with(self._strider.transpose(list(dims)))

    def with_name(self, name):                               # cpu_tensor.rs:232-241
        self.name = name
        if self.device.debug_named_tensors:
            self.device.debug_tensors[name] = self.buf.astype(np.float32).copy()  # whole buffer, like iter_f32
        return self

    # -- data movement ---------------------------------------------------------------------
    def _gather(self):
        st = self._strider
        idx = np.zeros(st.shape, np.int64)
        for ax, (n, s) in enumerate(zip(st.shape, st.strides)):
            sh = [1] * len(st.shape); sh[ax] = n
            idx = idx + (np.arange(n, dtype=np.int64) * s).reshape(sh)
        return self.buf[idx.reshape(-1)]

    def to_vec(self):                                        # cpu_tensor.rs:98-107 (test helper)
        assert self._dtype == oc.F32
        return self._gather().copy()

    def contiguous(self):                                    # cpu_tensor.rs:294-304
        if self.is_contiguous():
            return self
        assert self._dtype in (oc.F32, oc.F16) and len(self.shape()) in (2, 3)
        return OracleTensor(self._gather().copy(), TensorStrider(self.shape()), self._dtype, self.device, True)

    def concatenate(self, rhs, axis):                        # cpu_tensor.rs:251-292, concatenate.rs:12-77
        if not self.owned:
            raise TensorError("tensor not owned on concatenate")
        if self._dtype not in (oc.F32, oc.F16) or rhs._dtype not in (oc.F32, oc.F16):
            raise TensorError("only f32/f16 is supported on concatenate")
        if self._dtype == oc.F32 and rhs._dtype == oc.F16:
            raise TensorError("can not concatenate F32 and F16")
        s1, s2 = self._strider, rhs._strider
        for i in range(len(s1.shape)):
            if i != axis and s1.shape[i] != s2.shape[i]:
                raise TensorError(f"shape mismatch on concatenate, want {s1.shape} but got {s2.shape}")
        src = rhs._gather()
        if self._dtype == oc.F16 and rhs._dtype == oc.F32:
            src = oc.f32_to_f16(src)
        idx = np.zeros(s2.shape, np.int64)
        for ax, (n, s) in enumerate(zip(s2.shape, s1.strides)):
            sh = [1] * len(s2.shape); sh[ax] = n
            off = s1.shape[axis] if ax == axis else 0
            idx = idx + ((np.arange(n, dtype=np.int64) + off) * s).reshape(sh)
        if idx.size and idx.max() >= self.buf.size:
            raise TensorError("concatenate: out of pre-allocated storage")
        self.buf[idx.reshape(-1)] = src
        new_shape = list(s1.shape); new_shape[axis] += s2.shape[axis]
        self._strider = s1.resize(new_shape)

    def copy_rows_from(self, src, rows):                     # cpu_tensor.rs:306-331, buf/api.rs:262-321
        if not self.owned: raise TensorError("not owned")
        if not self.is_contiguous(): raise TensorError("dst tensor is not contiguous")
        if not src.is_contiguous(): raise TensorError("src tensor is not contiguous")
        if src._strider.dims() not in (1, 2): raise TensorError("copy_rows_from: src tensor is not 2d or 1d")
        cols = self.shape()[-1]
        for dst_row, src_row in enumerate(rows):
            so, do = src_row * cols, dst_row * cols
            if src._dtype == oc.F32:
                vals = src.buf[so:so + cols]
            elif src._dtype == oc.F16:
                vals = oc.f16_to_f32(src.buf[so:so + cols])
            else:
                be, bb = oc.block_elems(src._dtype), oc.block_bytes(src._dtype)
                assert so % be == 0 and cols % be == 0
                vals = oc.dequantize(src._dtype, src.buf[so // be * bb:(so + cols) // be * bb], cols, self.device.flags & oc.BUGCOMPAT)
            self.buf[do:do + cols] = vals if self._dtype == oc.F32 else oc.f32_to_f16(vals)

    def export(self):                                        # cpu_tensor.rs:339-349
        assert self.is_contiguous() and self._dtype == oc.F32
        return self.buf[:self._strider.len()].copy()

    def dup(self):                                           # cpu_tensor.rs:333-337
        return OracleTensor.new(self.buf.copy(), self.shape(), self.device)

    # -- compute -----------------------------------------------------------------------------
    def _f32_owned(self):
        if self._dtype != oc.F32 or not self.owned:
            raise TensorError(f"not owned f32, but got {oc.TYPE_NAMES[self._dtype]}, owned: {self.owned}")
        return self.buf

    def rope_inplace(self, mode, pos, rope_dims):            # cpu_tensor.rs:431-437, rope.rs:10-45
        assert self.is_contiguous() and self._strider.dims() in (2, 3)
        oc.rope_(self._f32_owned().reshape(self.shape()), mode, pos, rope_dims)
        return self

    def rms_norm_inplace(self, eps):                         # cpu_tensor.rs:439-445
        assert self.is_contiguous() and self._strider.dims() in (1, 2)
        assert self.shape()[-1] % 32 == 0                    # rms_norm.rs:34
        oc.rms_norm_(self._f32_owned().reshape(self.shape()), eps)
        return self

    def softmax_inplace(self, axis):                         # softmax.rs:11-30
        assert self._strider.dims() in (2, 3) and self.is_contiguous()
        if axis != self._strider.dims() - 1:
            raise TensorError(f"only axis={self._strider.dims() - 1} is supported")
        oc.softmax_(self._f32_owned()[:self._strider.len()].reshape(self.shape()))
        return self

    def silu_inplace(self):
        oc.silu_(self._f32_owned()); return self

    def gelu_inplace(self):
        oc.gelu_(self._f32_owned()); return self

    def _binary(self, rhs, fn):                              # arithmetic.rs:5-68
        a, b = self._f32_owned(), rhs.buf
        assert a.size % b.size == 0
        assert self.shape()[-1] == rhs.shape()[-1] or b.size == 1
        assert self.is_contiguous() and rhs.is_contiguous()
        fn(a, np.ascontiguousarray(b, np.float32))
        return self

    def mul_inplace(self, rhs): return self._binary(rhs, oc.mul_)
    def add_inplace(self, rhs): return self._binary(rhs, oc.add_)

    # -- exchange step of the sharded replay (the reference has none; device.comm = object with all_reduce / all_gather
    #    over numpy arrays, e.g. torch.distributed gloo in tests/test_sharded_gloo.py) ------------------------------------
    def all_reduce_sum_inplace(self):
        a = self._f32_owned()
        a[:self._strider.len()] = self.device.comm.all_reduce(a[:self._strider.len()])
        return self

    def all_gather_from(self, piece):
        self._f32_owned()[:self._strider.len()] = self.device.comm.all_gather(piece.buf[:piece._strider.len()])
        return self

    def scale_inplace(self, rhs):                            # cpu_tensor.rs:404-410
        return self._binary(OracleTensor.new([rhs], [1], self.device), oc.mul_)

    def matmul_vec(self, x):                                 # cpu_tensor.rs:371-386, matmul_vec.rs:9-23
        assert self.is_contiguous() and x.is_contiguous()
        assert self.shape()[-1] == x.shape()[-1]
        m, k = self.shape()
        shape_c = [m] if len(x.shape()) == 1 else [x.shape()[0], m]
        xin = x.buf[:x._strider.len()]
        out = oc.gemv(self._dtype, self.buf, m, k, xin, self.device.thread_num, self.device.flags)
        return OracleTensor.new(out, shape_c, self.device)

    def batch_matmul(self, b):                               # cpu_tensor.rs:352-366, batch_matmul.rs:15-45
        s1, s2 = self._strider, b._strider
        assert s1.dims() == 3 and s2.dims() == 3 and s1.is_contiguous()
        assert s2.strides[1] == 1 or s2.strides[2] == 1
        a = self.buf[:s1.len()].reshape(s1.shape)
        c = oc.batch_matmul(a, b.buf, s2.shape, s2.strides, b._dtype == oc.F16)
        return OracleTensor.new(c, [s1.shape[0], s1.shape[1], s2.shape[2]], self.device)
