// Host-only check of the C++ TensorStrider (crabml_b200/csrc/host/cuda_tensor.hpp) against the known answers of the
// reference's strider tests (crabml-core/src/tensor/strider.rs:242-338) plus a few transposed/resized cases the decode
// path uses (KV cache views, llama2.rs:527-603).  Prints one line per case; tests/test_host_strider.py compares the
// lines with the Python mirror and the oracle's strider.  No CUDA call is made.
#include <cstdio>
#include <string>
#include <vector>

#include "../../crabml_b200/csrc/host/cuda_tensor.hpp"

using crabml::TensorStrider;

static void show(const char* name, const TensorStrider& s) {
    std::printf("%s shape", name);
    for (auto v : s.shape()) std::printf(" %lld", (long long)v);
    std::printf(" strides");
    for (auto v : s.strides()) std::printf(" %lld", (long long)v);
    std::printf(" contiguous %d len %lld\n", s.is_contiguous() ? 1 : 0, (long long)s.len());
}
template <class F>
static void expect_throw(const char* name, F&& f) {
    try { f(); std::printf("%s no-error\n", name); }
    catch (const crabml::TensorError&) { std::printf("%s TensorError\n", name); }
}

int main() {
    TensorStrider s({3, 4});
    show("new_3x4", s);
    expect_throw("reshape_4x2", [&] { s.reshape({4, 2}); });                 // strider.rs:249-250
    show("reshape_2x6", s.reshape({2, 6}));                                   // strider.rs:252-255
    TensorStrider t = TensorStrider({2, 3}).transpose({1, 0});               // strider.rs:288-292
    show("transpose_10", t);
    show("transpose_back", t.transpose({1, 0}));                              // strider.rs:294-296
    expect_throw("reshape_noncontiguous", [&] { t.reshape({6}); });         // strider.rs:144-146
    show("resize_0x3200", TensorStrider({3, 3200}).resize({0, 3200}));       // strider.rs:327-330
    show("resize_3x0x3200", TensorStrider({3, 8, 3200}).resize({3, 0, 3200}));   // strider.rs:332-336
    expect_throw("resize_rank", [&] { TensorStrider({3, 4}).resize({12}); });
    expect_throw("transpose_rank", [&] { TensorStrider({3, 4}).transpose({0}); });
    // the KV cache walk of llama2.rs:65-86,527-603: [n_kv, seq, hd] resized to length 5, viewed as [n_kv, hd, seq]
    TensorStrider kv = TensorStrider({32, 4096, 128}).resize({32, 5, 128});
    show("kv_resized", kv);
    show("kv_T", kv.transpose({0, 2, 1}));
    show("q_heads", TensorStrider({1, 32, 128}).transpose({1, 0, 2}));
    show("scalar_like", TensorStrider({1}));
    return 0;
}
