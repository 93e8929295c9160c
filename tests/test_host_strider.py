"""Host-side metadata logic (the part of the `Tensor` trait that never crosses the C ABI): the Python mirror
(crabml_b200/tensor.py) and the C++ header (csrc/host/cuda_tensor.hpp) against the reference's strider known answers
(crabml-core/src/tensor/strider.rs:242-338) and against the oracle's strider on random view chains.  CPU only."""
import os
import subprocess

import numpy as np
import pytest

from crabml_b200.capi import TensorError
from crabml_b200.tensor import TensorStrider as PS
from oracle.tensor_ref import TensorError as OTensorError
from oracle.tensor_ref import TensorStrider as OS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_strider_known_answers():
    s = PS([3, 4])
    assert s.strides == [4, 1] and s.is_contiguous()
    with pytest.raises(TensorError):
        s.reshape([4, 2])                                  # strider.rs:249-250
    assert s.reshape([2, 6]).strides == [6, 1]             # strider.rs:252-255
    t = PS([2, 3]).transpose([1, 0])                       # strider.rs:288-292
    assert t.shape == [3, 2] and t.strides == [1, 3] and not t.is_contiguous()
    assert t.transpose([1, 0]).strides == [3, 1]           # strider.rs:294-296
    with pytest.raises(TensorError):
        t.reshape([6])
    assert PS([3, 3200]).resize([0, 3200]).strides == [3200, 1]                     # strider.rs:327-330
    r = PS([3, 8, 3200]).resize([3, 0, 3200])                                       # strider.rs:332-336
    assert r.shape == [3, 0, 3200] and r.strides == [3200 * 8, 3200, 1]


def test_python_mirror_equals_oracle_strider_on_random_view_chains():
    rng = np.random.default_rng(7)
    for _ in range(300):
        nd = int(rng.integers(1, 5))
        shape = [int(rng.integers(1, 6)) for _ in range(nd)]
        a, b = PS(shape), OS(shape)
        for _ in range(4):
            op = int(rng.integers(0, 3))
            if op == 0:
                perm = [int(x) for x in rng.permutation(len(a.shape))]
                a, b = a.transpose(perm), b.transpose(perm)
            elif op == 1:
                new = [int(rng.integers(0, s + 1)) for s in a.shape]
                a, b = a.resize(new), b.resize(new)
            else:
                n = a.len()
                new = [n] if n == 0 or rng.integers(0, 2) else [1, n]
                ea = eb = None
                try:
                    a2 = a.reshape(new)
                except TensorError as e:
                    ea = e
                try:
                    b2 = b.reshape(new)
                except (OTensorError, AssertionError) as e:
                    eb = e
                assert (ea is None) == (eb is None), (shape, a.shape, a.strides, new)
                if ea is None:
                    a, b = a2, b2
            assert a.shape == list(b.shape) and a.strides == list(b.strides)
            assert a.is_contiguous() == b.is_contiguous() and a.len() == b.len()


def test_cpp_header_strider(tmp_path):
    lib_dir = os.path.join(ROOT, "crabml_b200", "lib")
    if not os.path.exists(os.path.join(lib_dir, "libcrabml_cuda.so")):
        pytest.skip("library not built")
    exe = str(tmp_path / "strider_check")
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "host", "strider_check.cpp"), "-o", exe,
                    "-L" + lib_dir, "-lcrabml_cuda", "-Wl,-rpath," + lib_dir], check=True, capture_output=True, timeout=300)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=60).stdout.strip().splitlines()
    got = dict(line.split(" ", 1) for line in out)

    def fmt(s):
        return (f"shape {' '.join(map(str, s.shape))} strides {' '.join(map(str, s.strides))} "
                f"contiguous {int(s.is_contiguous())} len {s.len()}")
    want = {
        "new_3x4": fmt(PS([3, 4])), "reshape_4x2": "TensorError", "reshape_2x6": fmt(PS([3, 4]).reshape([2, 6])),
        "transpose_10": fmt(PS([2, 3]).transpose([1, 0])), "transpose_back": fmt(PS([2, 3]).transpose([1, 0]).transpose([1, 0])),
        "reshape_noncontiguous": "TensorError", "resize_0x3200": fmt(PS([3, 3200]).resize([0, 3200])),
        "resize_3x0x3200": fmt(PS([3, 8, 3200]).resize([3, 0, 3200])), "resize_rank": "TensorError", "transpose_rank": "TensorError",
        "kv_resized": fmt(PS([32, 4096, 128]).resize([32, 5, 128])), "kv_T": fmt(PS([32, 4096, 128]).resize([32, 5, 128]).transpose([0, 2, 1])),
        "q_heads": fmt(PS([1, 32, 128]).transpose([1, 0, 2])), "scalar_like": fmt(PS([1])),
    }
    assert got == want
