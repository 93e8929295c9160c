// ops.cu -- the small ops around the matvec, one launch per trait call.
// Compiled with -fmad=false / IEEE div+sqrt so that every op reproduces the reference's f32
// arithmetic; only reduction ORDER differs (the canonical 512-thread tree of common.cuh instead of sequential).
#include "common.cuh"

// ---- rms_norm_inplace: primitives/rms_norm.rs:32-47  x /= sqrt(sum(x^2)/n + eps), no weight ------------
// canonical 512-thread order (common.cuh); cols % 4 == 0 is implied by the reference's own assert (cols % 32 == 0)
__global__ void __launch_bounds__(CC_RED_THREADS) rms_norm_kernel(float* x, int64_t cols, float eps) {
    __shared__ float sh[CC_RED_WARPS];
    float* v = x + (int64_t)blockIdx.x * cols;
    float s = 0.0f;
    const int64_t n4 = cols >> 2;
    for (int64_t i = threadIdx.x; i < n4; i += CC_RED_THREADS) { const float* c = v + 4 * i; s += cc_sq4(make_float4(c[0], c[1], c[2], c[3])); }
    if ((cols & 3) && threadIdx.x == 0) for (int64_t i = n4 * 4; i < cols; i++) s += v[i] * v[i];
    s = cc_block_sum_512(s, sh);
    float rms = sqrtf(s / (float)cols + eps);
    for (int64_t i = threadIdx.x; i < cols; i += CC_RED_THREADS) v[i] = v[i] / rms;
}
int cc_launch_rms_norm(cc_device* dev, float* x, int64_t rows, int64_t cols, float eps) {
    if (rows == 0 || cols == 0) return CC_OK;
    rms_norm_kernel<<<(unsigned)rows, CC_RED_THREADS, 0, dev->stream>>>(x, cols, eps);
    CC_LAUNCH_CHECK(dev);
    return CC_OK;
}

// rope_inplace lives in exact.cu (host-evaluated cos/sin table, used by both modes)

// ---- softmax_inplace: primitives/softmax.rs:39-54 with the f16 exp LUT (quirk B4) ---------------------------
__device__ __forceinline__ float exp_cached(float v, const uint16_t* lut) { return h2f_bits(lut[f2h_bits(v)]); }

__global__ void __launch_bounds__(CC_RED_THREADS) softmax_kernel(float* x, int64_t cols, const uint16_t* __restrict__ lut) {
    __shared__ float sh[CC_RED_WARPS];
    float* v = x + (int64_t)blockIdx.x * cols;
    float m = -INFINITY;
    for (int64_t i = threadIdx.x; i < cols; i += CC_RED_THREADS) m = fmaxf(m, v[i]);
    m = cc_block_max_512(m, sh);
    float s = 0.0f;
    for (int64_t i = threadIdx.x; i < cols; i += CC_RED_THREADS) {
        float e = exp_cached(v[i] - m, lut);
        v[i] = e;
        s += e;
    }
    s = cc_block_sum_512(s, sh);          // canonical order (common.cuh): same bits as the fused attention kernels
    for (int64_t i = threadIdx.x; i < cols; i += CC_RED_THREADS) v[i] = v[i] / s;
}
int cc_launch_softmax(cc_device* dev, float* x, int64_t rows, int64_t cols) {
    if (rows == 0 || cols == 0) return CC_OK;
    softmax_kernel<<<(unsigned)rows, CC_RED_THREADS, 0, dev->stream>>>(x, cols, dev->exp_lut);
    CC_LAUNCH_CHECK(dev);
    return CC_OK;
}

// ---- silu / gelu: primitives/silu.rs:6-13, gelu.rs:10-15 -----------------------------------------------------
__global__ void silu_kernel(float* x, int64_t n, const uint16_t* __restrict__ lut) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float v = x[i];
        x[i] = v / (1.0f + exp_cached(-v, lut));
    }
}
__global__ void gelu_kernel(float* x, int64_t n, const uint16_t* __restrict__ lut) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = h2f_bits(lut[f2h_bits(x[i])]);
}
int cc_launch_silu(cc_device* dev, float* x, int64_t n) {
    if (n == 0) return CC_OK;
    silu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, dev->stream>>>(x, n, dev->exp_lut);
    CC_LAUNCH_CHECK(dev);
    return CC_OK;
}
int cc_launch_gelu(cc_device* dev, float* x, int64_t n) {
    if (n == 0) return CC_OK;
    gelu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, dev->stream>>>(x, n, dev->gelu_lut);
    CC_LAUNCH_CHECK(dev);
    return CC_OK;
}

// ---- add / mul with rhs cycling: primitives/arithmetic.rs:5-68 ---------------------------------------------------
__global__ void binary_kernel(float* x, int64_t n, const float* __restrict__ y, int64_t ny, int op) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float b = y[ny == 1 ? 0 : i % ny];
    x[i] = op == 0 ? x[i] + b : x[i] * b;
}
int cc_launch_binary(cc_device* dev, float* x, int64_t n, const float* y, int64_t ny, int op) {
    if (n == 0) return CC_OK;
    binary_kernel<<<(unsigned)((n + 255) / 256), 256, 0, dev->stream>>>(x, n, y, ny, op);
    CC_LAUNCH_CHECK(dev);
    return CC_OK;
}
__global__ void scale_kernel(float* x, int64_t n, float s) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = x[i] * s;
}
int cc_launch_scale(cc_device* dev, float* x, int64_t n, float s) {     // cpu_tensor.rs:404-410
    if (n == 0) return CC_OK;
    scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, dev->stream>>>(x, n, s);
    CC_LAUNCH_CHECK(dev);
    return CC_OK;
}

// ---- strided copy: contiguous (contiguous.rs:6-66) and concatenate (concatenate.rs:12-77,143-204) ----------------
struct CopyDims { int64_t shape[CC_MAX_DIMS], sstr[CC_MAX_DIMS], dstr[CC_MAX_DIMS]; int ndim; };

__global__ void strided_copy_kernel(const void* src, int src_dtype, void* dst, int dst_dtype, CopyDims d, int64_t dst_offset, int64_t total) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int64_t rem = i, so = 0, dof = dst_offset;
    for (int ax = d.ndim - 1; ax >= 0; ax--) {
        int64_t c = rem % d.shape[ax];
        rem /= d.shape[ax];
        so += c * d.sstr[ax];
        dof += c * d.dstr[ax];
    }
    if (src_dtype == CC_F32 && dst_dtype == CC_F32) ((float*)dst)[dof] = ((const float*)src)[so];
    else if (src_dtype == CC_F16 && dst_dtype == CC_F16) ((__half*)dst)[dof] = ((const __half*)src)[so];
    else if (src_dtype == CC_F32 && dst_dtype == CC_F16) ((__half*)dst)[dof] = __float2half_rn(((const float*)src)[so]);   // f16::from_f32
    else ((float*)dst)[dof] = __half2float(((const __half*)src)[so]);
}
int cc_launch_strided_copy(cc_device* dev, const void* src, int src_dtype, const int64_t* sshape, const int64_t* sstrides,
                           void* dst, int dst_dtype, const int64_t* dstrides, int64_t dst_offset, int ndim) {
    CopyDims d;
    d.ndim = ndim;
    int64_t total = 1;
    for (int i = 0; i < ndim; i++) { d.shape[i] = sshape[i]; d.sstr[i] = sstrides[i]; d.dstr[i] = dstrides[i]; total *= sshape[i]; }
    if (total == 0) return CC_OK;
    strided_copy_kernel<<<(unsigned)((total + 255) / 256), 256, 0, dev->stream>>>(src, src_dtype, dst, dst_dtype, d, dst_offset, total);
    CC_LAUNCH_CHECK(dev);
    return CC_OK;
}

// ---- batch_matmul: primitives/batch_matmul.rs:15-131 -----------------------------------------------------------------
// A dense (ab,m,k); B (bb,k,n) strided with stride_k == 1 (K^T view) or stride_n == 1 (V).
// F32 B: kv head = bi % bb (batch_matmul.rs:63);  F16 B: kv head = bi / (ab/bb) (batch_matmul.rs:89-91).
// (a) stride_k == 1: one warp per output, lanes stride over k
template <bool B_F16>
__global__ void bmm_kcontig_kernel(const float* __restrict__ a, const void* __restrict__ b, float* __restrict__ c,
                                   int64_t ab, int64_t bb, int64_t m, int64_t k, int64_t n, int64_t sb0, int64_t sb2) {
    int64_t o = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (o >= ab * m * n) return;
    int64_t ni = o % n, mi = (o / n) % m, bi = o / (m * n);
    const float* pa = a + bi * (m * k) + mi * k;
    float acc = 0.0f;
    if (B_F16) {
        // A -> f16 (quantize_f32_f16), products and accumulation in f32 (vec_dot_f16_f16_fallback, buf_f16.rs:84-97)
        const __half* pb = (const __half*)b + (bi / (ab / bb)) * sb0 + ni * sb2;
        for (int64_t ki = lane; ki < k; ki += 32) acc += __half2float(__float2half_rn(pa[ki])) * __half2float(pb[ki]);
    } else {
        const float* pb = (const float*)b + (bi % bb) * sb0 + ni * sb2;
        for (int64_t ki = lane; ki < k; ki += 32) acc += pa[ki] * pb[ki];
    }
    acc = warp_sum(acc);
    if (lane == 0) c[o] = acc;
}
// (b) stride_n == 1: one thread per output column, sequential over k (F16: f16 accumulation in the SAME order
//     as vec_fma_f16_f16, buf_f16.rs:152-163, so the result is bit-exact)
template <bool B_F16>
__global__ void bmm_ncontig_kernel(const float* __restrict__ a, const void* __restrict__ b, float* __restrict__ c,
                                   int64_t ab, int64_t bb, int64_t m, int64_t k, int64_t n, int64_t sb0, int64_t sb1) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= ab * m * n) return;
    int64_t ni = o % n, mi = (o / n) % m, bi = o / (m * n);
    const float* pa = a + bi * (m * k) + mi * k;
    if (B_F16) {
        const __half* pb = (const __half*)b + (bi / (ab / bb)) * sb0 + ni;
        __half acc = __float2half_rn(0.0f);
        for (int64_t ki = 0; ki < k; ki++) acc = __hadd(acc, __hmul(pb[ki * sb1], __float2half_rn(pa[ki])));
        c[o] = __half2float(acc);
    } else {
        const float* pb = (const float*)b + (bi % bb) * sb0 + ni;
        float acc = 0.0f;
        for (int64_t ki = 0; ki < k; ki++) acc += pa[ki] * pb[ki * sb1];     // same order as batch_matmul.rs:60-68
        c[o] = acc;
    }
}
int cc_launch_batch_matmul(cc_device* dev, const float* a, const void* b, int b_dtype, float* c, int64_t ab, int64_t bb,
                           int64_t m, int64_t k, int64_t n, int64_t sb0, int64_t sb1, int64_t sb2) {
    int64_t outs = ab * m * n;
    if (outs == 0) return CC_OK;
    bool f16 = b_dtype == CC_F16;
    if (sb1 == 1) {
        unsigned grid = (unsigned)((outs * 32 + 255) / 256);
        if (f16) bmm_kcontig_kernel<true><<<grid, 256, 0, dev->stream>>>(a, b, c, ab, bb, m, k, n, sb0, sb2);
        else bmm_kcontig_kernel<false><<<grid, 256, 0, dev->stream>>>(a, b, c, ab, bb, m, k, n, sb0, sb2);
    } else {
        unsigned grid = (unsigned)((outs + 127) / 128);
        if (f16) bmm_ncontig_kernel<true><<<grid, 128, 0, dev->stream>>>(a, b, c, ab, bb, m, k, n, sb0, sb1);
        else bmm_ncontig_kernel<false><<<grid, 128, 0, dev->stream>>>(a, b, c, ab, bb, m, k, n, sb0, sb1);
    }
    CC_LAUNCH_CHECK(dev);
    return CC_OK;
}

// ---- greedy sampling on the device: sampler.rs:109-116 (`max_by` keeps the LAST maximum; NaN never wins) -------------------------
// one CTA; thread t scans t, t + 512, ...; ties resolve to the larger index at every level, so the result is the last maximum
__device__ __forceinline__ void argmax_merge(float& bv, long long& bi, float ov, long long oi) {
    if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi > bi))) { bv = ov; bi = oi; }
}
__global__ void __launch_bounds__(CC_RED_THREADS) argmax_kernel(const float* __restrict__ x, long long n, long long* slot, long long* hist,
                                                               const long long* hist_index_dev, long long hist_index) {
    __shared__ float sv[CC_RED_WARPS];
    __shared__ long long si[CC_RED_WARPS];
    float bv = 0.0f; long long bi = -1;
    for (long long i = threadIdx.x; i < n; i += CC_RED_THREADS) { const float v = x[i]; if (bi < 0 || !(v < bv)) { bv = v; bi = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        argmax_merge(bv, bi, ov, oi);
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < CC_RED_WARPS; w++) argmax_merge(bv, bi, sv[w], si[w]);
        if (bi < 0) bi = 0;
        *slot = bi;
        const long long h = hist_index_dev ? *hist_index_dev : hist_index;
        if (hist && h >= 0 && h < CC_HISTORY_CAP) hist[h] = bi;
    }
}
int cc_launch_argmax(cc_device* dev, const float* x, int64_t n, int64_t* slot, int64_t* hist, const int64_t* hist_index_dev, int64_t hist_index) {
    argmax_kernel<<<1, CC_RED_THREADS, 0, dev->stream>>>(x, (long long)n, (long long*)slot, (long long*)hist, (const long long*)hist_index_dev, (long long)hist_index);
    CC_LAUNCH_CHECK(dev);
    return CC_OK;
}
