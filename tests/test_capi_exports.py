"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/crabml_cuda.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes

import pytest

from crabml_b200 import capi


def test_library_exports_every_declared_symbol():
    L = capi.load_library()
    syms = capi.declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"{s} declared in crabml_cuda.h but not exported"


def test_header_has_one_entry_per_trait_method():
    # crabml-core/src/tensor/api.rs:11-79; metadata-only methods stay host-side (strider)
    need = ["cc_tensor_from_cpu", "cc_tensor_alloc", "cc_tensor_dup", "cc_tensor_export_f32", "cc_copy_rows_from",
            "cc_concatenate", "cc_contiguous", "cc_rope_inplace", "cc_rms_norm_inplace", "cc_softmax_inplace",
            "cc_silu_inplace", "cc_gelu_inplace", "cc_mul_inplace", "cc_add_inplace", "cc_scale_inplace",
            "cc_matmul_vec", "cc_batch_matmul", "cc_debug_tensor_tap", "cc_dump_debug_tensor"]
    syms = set(capi.declared_symbols())
    assert not [n for n in need if n not in syms]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from crabml_b200 import CudaTensorDevice, CudaError
    with pytest.raises(CudaError, match="no CPU fallback"):
        CudaTensorDevice()


def test_view_struct_layout():
    assert ctypes.sizeof(capi.cc_view) == 8 + 8 + 32 + 32
    assert ctypes.sizeof(capi.cc_device_options) == 24
