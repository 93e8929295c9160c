// matvec_stream.cu -- the streaming matvec for the decode hot path (b = 1), second design.
//
// What the first design (matvec.cu, kept for b > 1 and the rarer types) got wrong -- profiles/r01a: latency bound
// at 38 % of DRAM peak (4 loads in flight per lane, activation staged before the first weight load, 1.33 waves,
// a separate quantize launch per matmul_vec).  This kernel:
//   * persistent grid: 2 CTAs x 8 warps per SM, rows dealt round-robin to warps;
//   * the weight stream of a warp is a flat sequence of "groups" (32 quant blocks = 1024 weights) crossing row
//     boundaries, software-pipelined through a ring of D register slots: the load of group t+D is issued when
//     group t is consumed, so every lane keeps D x 32 B (Q8_0) / D x 16 B (Q4_0) in flight at all times
//     (8 KB per warp, 128 KB per SM);
//   * the first D groups are requested BEFORE the activation prologue, which then runs under their latency;
//   * the prologue is fused: [optional rms_norm * weight] -> Q8_0 quantisation of x into shared memory
//     (buf_q8_0.rs:87-134 arithmetic, bit-exact) -- no separate quantize / rms_norm / mul launches;
//   * up to 3 matrices that share the activation (wq,wk,wv / gate,up) run as one launch;
//   * epilogues: store | + residual (llama2.rs:266,636) | silu(gate) * up (llama2.rs:620-630).
// Q8_0 device layout: inside each group of 32 blocks the 16-byte first halves of all blocks precede the second
// halves, so lane l reads block 32g+l with two fully coalesced LDG.128 (512 B per warp request).
#include "common.cuh"

#define MS_THREADS 256
#define MS_WARPS 8
#define MS_CTAS_PER_SM 2

struct StreamMats {
    const uint8_t* qs[3];
    const uint16_t* d[3];
    float* out[3];
    int m[3];
    int n;
};
struct StreamArgs {
    StreamMats mats;
    const float* x;          // f32 activation [k]
    int k;
    int prologue;            // 0 plain, 1 rms_norm(eps) * norm_w first
    const float* norm_w;
    float eps;
    int epilogue;            // 0 store, 1 add residual, 2 silu(mat0 row) * (mat1 row)
    const float* residual;
    const uint16_t* exp_lut;
};

__device__ __forceinline__ int dp16(const int4& w, const int4& a) {
    return __dp4a(w.x, a.x, __dp4a(w.y, a.y, __dp4a(w.z, a.z, __dp4a(w.w, a.w, 0))));
}

template <int TYPE> struct Slot;
template <> struct Slot<CC_Q8_0> { int4 a, b; uint16_t s; };
template <> struct Slot<CC_Q4_0> { int4 a; uint16_t s; };

template <int TYPE>
__device__ __forceinline__ void load_slot(Slot<TYPE>& sl, const StreamMats& M, int row, int g, int nb, int lane, bool valid) {
    // row -> (matrix, local row)
    int mi = 0, r = row;
    if (M.n > 1 && r >= M.m[0]) { r -= M.m[0]; mi = 1; if (M.n > 2 && r >= M.m[1]) { r -= M.m[1]; mi = 2; } }
    const int nbg = min(32, nb - 32 * g);
    const bool on = valid && lane < nbg;
    if constexpr (TYPE == CC_Q8_0) {
        const uint8_t* p = M.qs[mi] + (size_t)r * nb * 32 + (size_t)g * 1024 + lane * 16;
        Slot<CC_Q8_0>& s8 = sl;
        if (on) { s8.a = ld_stream_16(p); s8.b = ld_stream_16(p + 16 * nbg); s8.s = M.d[mi][(size_t)r * nb + 32 * g + lane]; }
        else { s8.a = make_int4(0, 0, 0, 0); s8.b = s8.a; s8.s = 0; }
    } else {
        const uint8_t* p = M.qs[mi] + (size_t)r * nb * 16 + (size_t)g * 512 + lane * 16;
        Slot<CC_Q4_0>& s4 = sl;
        if (on) { s4.a = ld_stream_16(p); s4.s = M.d[mi][(size_t)r * nb + 32 * g + lane]; }
        else { s4.a = make_int4(0, 0, 0, 0); s4.s = 0; }
    }
}

// shared memory: activation quants [k] int8 | scales [nb] f32 | (Q4_0) block sums [nb] i32
template <int TYPE>
__device__ __forceinline__ float consume_slot(const Slot<TYPE>& sl, int g, int nb, int lane, const int4* aq, const float* ad, const int* as) {
    const int b = 32 * g + lane;
    if (b >= nb) return 0.0f;
    if constexpr (TYPE == CC_Q8_0) {
        const Slot<CC_Q8_0>& s8 = sl;
        int sumi = dp16(s8.a, aq[2 * b]) + dp16(s8.b, aq[2 * b + 1]);
        return (float)sumi * h2f_bits(s8.s) * ad[b];                   // buf_q8_0.rs:283 per-block term
    } else {
        const Slot<CC_Q4_0>& s4 = sl;
        int4 lo = make_int4(s4.a.x & 0x0F0F0F0F, s4.a.y & 0x0F0F0F0F, s4.a.z & 0x0F0F0F0F, s4.a.w & 0x0F0F0F0F);
        int4 hi = make_int4((s4.a.x >> 4) & 0x0F0F0F0F, (s4.a.y >> 4) & 0x0F0F0F0F, (s4.a.z >> 4) & 0x0F0F0F0F, (s4.a.w >> 4) & 0x0F0F0F0F);
        int sumi = dp16(lo, aq[2 * b]) + dp16(hi, aq[2 * b + 1]) - 8 * as[b];   // buf_q4_0.rs:244-249
        return (float)sumi * h2f_bits(s4.s) * ad[b];
    }
}

__device__ __forceinline__ float ms_block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < MS_WARPS; i++) t += sh[i];
    __syncthreads();
    return t;
}

template <int TYPE, int D>
__global__ void __launch_bounds__(MS_THREADS, MS_CTAS_PER_SM) matvec_stream_kernel(StreamArgs A) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ float s_red[MS_WARPS];
    const int k = A.k, nb = k >> 5, GR = (nb + 31) >> 5;
    int8_t* s_q = (int8_t*)smem;
    float* s_d = (float*)(smem + ((k + 15) & ~15));
    int* s_s = (int*)(smem + ((k + 15) & ~15) + ((nb * 4 + 15) & ~15));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * MS_WARPS + warp, TW = gridDim.x * MS_WARPS;
    int m_total = A.mats.m[0] + (A.mats.n > 1 ? A.mats.m[1] : 0) + (A.mats.n > 2 ? A.mats.m[2] : 0);
    const bool pair = A.epilogue == 2;               // rows of mat0 and mat1 are consumed pairwise by the same warp
    if (pair) m_total = A.mats.m[0];
    const int n_rows_w = gw < m_total ? (m_total - gw + TW - 1) / TW : 0;
    const int sub = pair ? 2 : 1;
    const int T = n_rows_w * sub * GR;               // flat group count of this warp

    // ---- 1. request the first D groups of the weight stream -------------------------------------------
    Slot<TYPE> slot[D];
    int li = 0, lsub = 0, lg = 0;                    // load cursor (row iteration, sub-row, group)
#pragma unroll
    for (int s = 0; s < D; s++) {
        const int row = gw + li * TW + (pair && lsub ? A.mats.m[0] : 0);
        load_slot<TYPE>(slot[s], A.mats, row, lg, nb, lane, s < T);
        if (++lg == GR) { lg = 0; if (++lsub == sub) { lsub = 0; li++; } }
    }

    // ---- 2. prologue under the latency of (1): [rms_norm * w] + Q8_0 quantisation of x into shared memory ----
    // All loads of a pass are issued before any is used (one L2 latency per pass, not one per block).
    constexpr int PB = 8;                            // blocks per warp per pass
    float rms = 0.0f;
    if (A.prologue == 1) {                           // rms_norm.rs:32-47 (sum order differs: tree)
        float ss = 0.0f;
        for (int i0 = 0; i0 < k; i0 += MS_THREADS * PB) {
            float v[PB];
#pragma unroll
            for (int j = 0; j < PB; j++) { int i = i0 + j * MS_THREADS + threadIdx.x; v[j] = i < k ? A.x[i] : 0.0f; }
#pragma unroll
            for (int j = 0; j < PB; j++) ss += v[j] * v[j];
        }
        ss = ms_block_sum(ss, s_red);
        rms = sqrtf(ss / (float)k + A.eps);
    }
    for (int b0 = 0; b0 < nb; b0 += MS_WARPS * PB) { // one warp per 32-element block (buf_q8_0.rs:87-134)
        float v[PB], nw[PB];
#pragma unroll
        for (int j = 0; j < PB; j++) {
            const int b = b0 + j * MS_WARPS + warp;
            v[j] = b < nb ? A.x[b * 32 + lane] : 0.0f;
            if (A.prologue == 1) nw[j] = b < nb ? A.norm_w[b * 32 + lane] : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < PB; j++) {
            const int b = b0 + j * MS_WARPS + warp;
            float x = v[j];
            if (A.prologue == 1) x = (x / rms) * nw[j];          // x/rms then * weight (llama2.rs:231-232)
            float amax = warp_max(fabsf(x));
            float d = amax / 127.0f;
            int q = __float2int_rz(x / d);
            if (b < nb) {
                s_q[b * 32 + lane] = (int8_t)q;
                if constexpr (TYPE == CC_Q4_0) { int s = warp_sum_i(q); if (lane == 0) s_s[b] = s; }
                if (lane == 0) s_d[b] = __half2float(__float2half_rn(d));
            }
        }
    }
    __syncthreads();
    const int4* aq = (const int4*)s_q;

    // ---- 3. stream: consume group t, refill its slot with group t + D ------------------------------------
    float acc = 0.0f, first = 0.0f;
    int ci = 0, csub = 0, cg = 0;
    for (int t0 = 0; t0 < T; t0 += D) {
#pragma unroll
        for (int s = 0; s < D; s++) {
            const int t = t0 + s;
            if (t < T) {
                acc += consume_slot<TYPE>(slot[s], cg, nb, lane, aq, s_d, s_s);
                {   // refill
                    const int row = gw + li * TW + (pair && lsub ? A.mats.m[0] : 0);
                    load_slot<TYPE>(slot[s], A.mats, row, lg, nb, lane, t + D < T);
                    if (++lg == GR) { lg = 0; if (++lsub == sub) { lsub = 0; li++; } }
                }
                if (++cg == GR) {                    // a (sub-)row is complete
                    cg = 0;
                    float r = warp_sum(acc);
                    acc = 0.0f;
                    const int row = gw + ci * TW;
                    if (pair) {
                        if (csub == 0) { first = r; csub = 1; }
                        else {                       // silu.rs:6-13 then mul (llama2.rs:625-630)
                            csub = 0; ci++;
                            if (lane == 0) {
                                float g = first;
                                float nexp = h2f_bits(A.exp_lut[f2h_bits(-g)]);
                                A.mats.out[0][row] = (g / (1.0f + nexp)) * r;
                            }
                        }
                    } else {
                        ci++;
                        if (lane == 0) {
                            int mi = 0, rr = row;
                            if (A.mats.n > 1 && rr >= A.mats.m[0]) { rr -= A.mats.m[0]; mi = 1; if (A.mats.n > 2 && rr >= A.mats.m[1]) { rr -= A.mats.m[1]; mi = 2; } }
                            if (A.epilogue == 1) r = r + A.residual[rr];
                            A.mats.out[mi][rr] = r;
                        }
                    }
                }
            }
        }
    }
}

static size_t stream_smem_bytes(int type, int k) {
    size_t nb = k / 32;
    return ((k + 15) & ~15) + ((nb * 4 + 15) & ~15) + (type == CC_Q4_0 ? ((nb * 4 + 15) & ~15) : 0);
}

bool cc_stream_supported(int type, int64_t k) { return (type == CC_Q8_0 || type == CC_Q4_0) && k % 32 == 0 && k <= 65536; }

int cc_launch_matvec_stream(cc_device* dev, int type, const StreamArgs& A) {
    size_t smem = stream_smem_bytes(type, A.k);
    int grid = dev->sm_count * MS_CTAS_PER_SM;
    int64_t m_total = (int64_t)A.mats.m[0] + (A.mats.n > 1 && A.epilogue != 2 ? A.mats.m[1] : 0) + (A.mats.n > 2 ? A.mats.m[2] : 0);
    int64_t need = (m_total + MS_WARPS - 1) / MS_WARPS;
    if (need < grid) grid = (int)(need > 0 ? need : 1);
    if (type == CC_Q8_0) {
        if (smem > 48 * 1024) CC_CUDA(dev, cudaFuncSetAttribute(matvec_stream_kernel<CC_Q8_0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        matvec_stream_kernel<CC_Q8_0, 8><<<grid, MS_THREADS, smem, dev->stream>>>(A);
    } else {
        if (smem > 48 * 1024) CC_CUDA(dev, cudaFuncSetAttribute(matvec_stream_kernel<CC_Q4_0, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        matvec_stream_kernel<CC_Q4_0, 16><<<grid, MS_THREADS, smem, dev->stream>>>(A);
    }
    CC_LAUNCH_CHECK(dev);
    return CC_OK;
}

// plain matmul_vec through the streaming kernel (eager trait call): one launch, quantisation fused
int cc_launch_matvec_stream_plain(cc_device* dev, const cc_buf* w, const float* x, float* out, int64_t m, int64_t k) {
    StreamArgs A = {};
    A.mats.n = 1;
    A.mats.qs[0] = w->plane[0];
    A.mats.d[0] = (const uint16_t*)w->plane[1];
    A.mats.out[0] = out;
    A.mats.m[0] = (int)m;
    A.x = x;
    A.k = (int)k;
    A.exp_lut = dev->exp_lut;
    return cc_launch_matvec_stream(dev, w->dtype, A);
}
