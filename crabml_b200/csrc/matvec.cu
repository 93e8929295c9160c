// matvec.cu -- THE hot path: C[b,m] = sum_k W[m,k] * x[b,k] with W in GGUF quant blocks
// (SURVEY §8a rows a1, a2, a6-a11;  reference: primitives/matmul_vec.rs:26-78 -> buf/buf_q*.rs vec_dot_*).
//
// Design (HBM-bound integer/byte work -- no tensor cores on purpose):
//   * one warp per output row; weight bytes are streamed exactly once with 16-byte LDG.128
//     (L1::no_allocate), consecutive lanes read consecutive 16-byte chunks -> 512 B per warp request;
//   * the quantised activation (Q8_0 / Q8_1 / Q8_K, a few KB) is staged once per CTA in shared memory;
//   * in-register unpack (nibble / high-bit / 2-bit fields), dp4a integer dots, one f32 scale
//     multiply per block, warp-shuffle reduction;
//   * integer sub-block sums are exact; only the order of the final f32 accumulation differs
//     from the reference's scalar loops (tolerance 1e-5 rel in tests/test_gpu_matvec.py).
#include "common.cuh"
#include "vecdot.cuh"

#define MV_THREADS 256
#define MV_WARPS (MV_THREADS / 32)

template <class T>
__global__ void __launch_bounds__(MV_THREADS) matvec_kernel(WPlanes W, const void* act, int64_t n_total,
                                                            float* __restrict__ out, int m, int k) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int bi = blockIdx.y;
    T::stage_act(act, n_total, bi, k, smem);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int row = blockIdx.x * MV_WARPS + warp; row < m; row += gridDim.x * MV_WARPS) {
        float acc = T::row_dot(W, row, k, smem, lane);
        acc = warp_sum(acc);
        if (lane == 0) out[(int64_t)bi * m + row] = acc;
    }
}

template <class T>
static int launch(cc_device* dev, const cc_buf* w, const void* act, int64_t n_total, float* out, int64_t m, int64_t k, int64_t b) {
    WPlanes W;
    for (int i = 0; i < CC_MAX_PLANES; i++) W.p[i] = w->plane[i];
    int smem = T::smem_bytes((int)k);
    if (smem > 48 * 1024) {
        CC_REQUIRE(dev, smem <= 200 * 1024, "matmul_vec: k=%lld needs %d bytes of shared memory", (long long)k, smem);
        CC_CUDA(dev, cudaFuncSetAttribute(matvec_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    int64_t ctas = (m + MV_WARPS - 1) / MV_WARPS;
    int64_t cap = (int64_t)dev->sm_count * 8;
    if (ctas > cap) ctas = cap;
    dim3 grid((unsigned)ctas, (unsigned)b);
    matvec_kernel<T><<<grid, MV_THREADS, smem, dev->stream>>>(W, act, n_total, out, (int)m, (int)k);
    CC_LAUNCH_CHECK(dev);
    return CC_OK;
}

int cc_launch_matvec(cc_device* dev, const cc_buf* w, const void* act, const float* x_f32, float* out,
                     int64_t m, int64_t k, int64_t b) {
    if (m == 0 || b == 0) return CC_OK;
    const int64_t n_total = b * k;
    switch (w->dtype) {
    case CC_F32: return launch<TF32>(dev, w, x_f32, n_total, out, m, k, b);
    case CC_F16: return launch<TF16>(dev, w, act, n_total, out, m, k, b);
    case CC_Q8_0: return launch<TQ8_0>(dev, w, act, n_total, out, m, k, b);
    case CC_Q4_0: return launch<TQ4_0>(dev, w, act, n_total, out, m, k, b);
    case CC_Q5_0: return launch<TQ5_0>(dev, w, act, n_total, out, m, k, b);
    case CC_Q4_1: return launch<TQx_1<false>>(dev, w, act, n_total, out, m, k, b);
    case CC_Q5_1: return launch<TQx_1<true>>(dev, w, act, n_total, out, m, k, b);
    case CC_Q2_K: return launch<TQ2_K>(dev, w, act, n_total, out, m, k, b);
    case CC_Q3_K: return launch<TQ3_K>(dev, w, act, n_total, out, m, k, b);
    case CC_Q4_K: return launch<TQ45_K<false>>(dev, w, act, n_total, out, m, k, b);
    case CC_Q5_K: return launch<TQ45_K<true>>(dev, w, act, n_total, out, m, k, b);
    case CC_Q6_K: return launch<TQ6_K>(dev, w, act, n_total, out, m, k, b);
    case CC_Q8_K: return launch<TQ8_K>(dev, w, act, n_total, out, m, k, b);
    }
    return cc_fail(dev, CC_ERR_UNSUPPORTED, "matmul_vec: weight type %d is not supported", w->dtype);
}
