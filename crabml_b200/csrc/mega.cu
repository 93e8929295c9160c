// mega.cu -- one persistent kernel per token ("megakernel") for the fused decode path.
//
// Why: profiles/r01c -- on B200 a kernel boundary costs ~4-5 us for a full-GPU streaming kernel (drain, launch latency,
// ramp-up) and ~2-3 us for a tiny one; a fused Llama-2-7B token still has ~260 of them (~1 ms), as much as the 1.05 ms the
// weights need at HBM speed.  Here the whole token is ONE launch of one 512-thread CTA per SM that walks a table of
// phases (built by lazy.cu from the recorded trait calls) separated by grid-wide barriers (2.15 us each, measured):
//     MATVEC  streaming matvec over 1-3 matrices + epilogue, optionally with a fused prologue ([dup] + rms_norm * w + Q8_0
//             quantisation of the input row, recomputed by every CTA) and, on the sharded path, the exchange with the other GPUs
//     NORMQ   the same normalise + quantise as a phase of its own (only when the f32 row must be materialised)
//     ATTN    rope + KV append + attention + output quantise    (one CTA per head, K/V chunks through a TMA pipeline)
//     ROWS    copy_rows_from (embedding row dequantisation)
//     REDUCE / GATHER   second half of an exchange when it cannot fold into the next MATVEC prologue
// Measurements and the per-phase time breakdown: profiles/r01f_megakernel_ncu.md.
// Data written by one CTA and read by another in a later phase is always read with ld.global.cg (L2), never through
// the non-coherent L1.  All CTAs execute the same number of barriers.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "dequant.cuh"
#include "quantize_dev.cuh"
#include "vecdot.cuh"

// one CTA of 16 warps per SM: the grid barrier has 148 participants instead of 296 (its cost is what bounds a phase)
#define MK_THREADS 512
#define MK_WARPS 16
#define MK_CTAS_PER_SM 1
#define MK_SEG 4
#define MK_XSTAGE_ROWS 512         // exchange phases: rows of one CTA's contiguous block (2 KB stage in shared memory)

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// Split grid barrier (measured with tools/barrier_floor.py on B200, 148 CTAs: 2.1 us per phase; the two-level
// acq_rel-atomic version cost 3.1 us, relaxed polling + fence 2.6 us):
//   arrive : bar.sync, then ONE thread does a fire-and-forget red.release.gpu.add on a flat monotonic counter
//            (the release publishes the CTA's phase output; that thread never issues prefetch loads);
//   ...      the other warps may already issue the next phase's weight prefetch;
//   wait   : CTA 0 watches the counter reach (gen+1) * nblocks and publishes the generation word; everyone else spins on
//            the generation with ld.acquire; then bar.sync.
// Layout (u32 words on separate 128-byte lines): [0] arrival counter, [32] generation.  Both are monotonic ACROSS launches (u32
// wrap-around included: only equality is tested): a launch starts from the generation the previous one left, so a graph
// replay needs no reset node in front of the kernel.
__device__ __forceinline__ void red_add_release(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
#define MK_BAR_THREAD (MK_THREADS - 1)
#define MK_BAR_ERR 64                        // bar[64]: non-zero once any spin on this GPU has timed out
__device__ __forceinline__ unsigned long long globaltimer_ns() { return cc_globaltimer_ns(); }
// every spin is bounded (CcSpin, common.cuh): a barrier that cannot complete ends in CC_ERR_CUDA "megakernel barrier timeout"
struct MkSpin {
    CcSpin sp;
    __device__ __forceinline__ bool expired(unsigned* bar, unsigned* err_host, unsigned code) { return sp.expired(&bar[MK_BAR_ERR], err_host, code); }
};
// xgpu: this CTA stored partial rows into the peers' exchange slots.  Those stores are ordered before the peers' reads by the chain
// (CTA i) red.release.gpu -> (CTA 0) ld.acquire.gpu ... fence.sys + st.release.sys -> (peer) ld.acquire.sys: release/acquire patterns of
// different scopes compose (PTX memory model: causality order is transitive over morally strong synchronisation), so a system-scope
// fence in EVERY CTA is not required (default: off; the 2-GPU parity test runs this way); sysfence = true adds it (it waits for this
// CTA's NVLink stores to be acknowledged, ~2.5 us per exchange).
__device__ __forceinline__ void grid_barrier_arrive(unsigned* bar, unsigned nblocks, unsigned gen, bool xgpu = false, bool sysfence = false) {
    __syncthreads();
    if (threadIdx.x == MK_BAR_THREAD) {
        if (xgpu && sysfence) __threadfence_system();
        red_add_release(&bar[0], 1u);
    }
}
// xseq != 0: the barrier doubles as the handshake of exchange number xseq with the other GPUs (protocol: comm.cu).  CTA 0's
// barrier thread, once every local CTA has arrived (all partial rows are stored in the peers' slots), publishes xseq in each
// peer's flag word, waits for every peer's xseq in its own flag words, and only then opens the local barrier.
// poll_counter: (local barriers only) every CTA watches the arrival counter itself -- one L2 round trip less than
// counter -> CTA 0 -> generation word; CTA 0 still publishes the generation (the next launch starts from it).
__device__ __forceinline__ void grid_barrier_wait(unsigned* bar, unsigned nblocks, unsigned gen, const CommDev& comm, unsigned xseq, bool poll_counter,
                                                  int* s_abort, unsigned* err_host) {
    const int lane = threadIdx.x & 31;
    if ((threadIdx.x >> 5) == MK_WARPS - 1) {              // the warp of MK_BAR_THREAD (its lane 31)
        const unsigned target = (gen + 1u) * nblocks;
        bool ok = true;
        if (blockIdx.x == 0) {
            if (lane == 31) { MkSpin sp; while ((int)(ld_acquire_u32(&bar[0]) - target) < 0) if (sp.expired(bar, err_host, 1u)) { ok = false; break; } }   // (poll mode: the others may already be arriving at the next barrier)
            if (xseq) {                                     // kernel-uniform: the whole warp takes this branch together
                __syncwarp();                               // every local CTA has arrived: all partial rows are in the peers' slots
                if (lane < comm.world) {                    // one lane per peer: publish and poll in parallel, not rank after rank
                    // st.release.sys orders everything this warp has observed (lane 31's acquire of the arrival counter, handed over
                    // by the __syncwarp above) before the flag: no separate system fence
                    cc_st_release_sys(comm.flag[lane] + comm.rank * 32, xseq);
                    const unsigned* f = comm.flag[comm.rank] + lane * 32;
                    MkSpin sp;
                    while ((int)(cc_ld_acquire_sys(f) - xseq) < 0) if (sp.expired(bar, err_host, 2u)) { ok = false; break; }
                }
                __syncwarp();
            }
            if (lane == 31) st_release_u32(&bar[32], gen + 1u);
        } else if (lane == 31) {
            MkSpin sp;
            if (poll_counter && !xseq) { while ((int)(ld_acquire_u32(&bar[0]) - target) < 0) if (sp.expired(bar, err_host, 1u)) { ok = false; break; } }   // fast CTAs may already have arrived at the NEXT barrier
            else { while (ld_acquire_u32(&bar[32]) != gen + 1u) if (sp.expired(bar, err_host, 1u)) { ok = false; break; } }
        }
        if (!ok) *s_abort = 1;
    }
    __syncthreads();
}

__device__ __forceinline__ float ldcg_f(const float* p) { return __ldcg(p); }

// ---- NORMQ phase (fused.cu normq_kernel, grid-wide) -----------------------------------------------------------------
__device__ void phase_normq(const MkPhase& ph, float* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = ph.n;
    float* x = ph.x;
    float rms = 1.0f;
    // write_back: the normalised row must be materialised in place -> one CTA does the whole row (nobody else may
    // still be summing x while it is overwritten)
    if (ph.write_back && blockIdx.x != 0) return;
    // this warp's blocks are requested first, so their latency overlaps the row pass below (one L2 round trip in total)
    const int nb0 = n >> 5;
    const int gw0 = ph.write_back ? warp : blockIdx.x * MK_WARPS + warp, tw0 = ph.write_back ? MK_WARPS : gridDim.x * MK_WARPS;
    float pre[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { const int b = gw0 + j * tw0; pre[j] = b < nb0 ? ldcg_f(x + b * 32 + lane) : 0.0f; }
    if (ph.norm_w) {
        float ss = 0.0f;
        const float4* x4 = (const float4*)x;
        const int n4 = n >> 2;
        for (int i0 = 0; i0 < n4; i0 += MK_THREADS * 4) {
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { int i = i0 + j * MK_THREADS + threadIdx.x; v[j] = i < n4 ? __ldcg(x4 + i) : make_float4(0, 0, 0, 0); }
#pragma unroll
            for (int j = 0; j < 4; j++) ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
        }
        ss = warp_sum(ss);
        if (lane == 0) s_red[warp] = ss;
        __syncthreads();
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < MK_WARPS; w++) t += s_red[w];
        rms = sqrtf(t / (float)n + ph.eps);
        __syncthreads();
    }
    ActQ8_0 act = ph.act;
    const int nb = n >> 5;
    const int gw = ph.write_back ? warp : blockIdx.x * MK_WARPS + warp, tw = ph.write_back ? MK_WARPS : gridDim.x * MK_WARPS;
    auto do_block = [&](int b, float v) {
        if (ph.orig) ph.orig[b * 32 + lane] = v;
        if (ph.norm_w) { v = (v / rms) * ph.norm_w[b * 32 + lane]; if (ph.write_back) x[b * 32 + lane] = v; }
        float amax = warp_max(fabsf(v));
        float d = amax / 127.0f;
        int q = __float2int_rz(v / d);
        act.qs[b * 32 + lane] = (int8_t)q;
        int s = warp_sum_i(q);
        if (lane == 0) { act.d[b] = __half2float(__float2half_rn(d)); act.isum[b] = s; }
    };
#pragma unroll
    for (int j = 0; j < 4; j++) { const int b = gw + j * tw; if (b < nb) do_block(b, pre[j]); }
    for (int b = gw + 4 * tw; b < nb; b += tw) do_block(b, ldcg_f(x + b * 32 + lane));
}

// ---- MATVEC phase: the body of matvec_stream_kernel (see matvec_stream.cu for the design notes) -------------------------
__device__ __forceinline__ int mk_dp16(const int4& w, const int4& a) {
    return __dp4a(w.x, a.x, __dp4a(w.y, a.y, __dp4a(w.z, a.z, __dp4a(w.w, a.w, 0))));
}
typedef KSeg MkSeg;                                                  // (Q4_0 leaves b unused)
static_assert(MK_SEG == 4, "KSeg holds 4 groups");
struct MkRowPtr { const uint8_t* q; const uint16_t* d; };

template <int TYPE>
__device__ __forceinline__ void mk_seg_load(MkSeg& S, const MkRowPtr& p, int seg, int nb, int GR, int last_half_off, int lane, bool valid) {
    constexpr int GB = TYPE == CC_Q8_0 ? 1024 : 512;
    const uint8_t* q = p.q + (size_t)seg * (MK_SEG * GB);
    const uint16_t* d = p.d + seg * (MK_SEG * 32);
#pragma unroll
    for (int g = 0; g < MK_SEG; g++) {
        const int gi = seg * MK_SEG + g;
        const bool on = valid && (gi * 32 + lane < nb);
        if constexpr (TYPE == CC_Q8_0) {
            const int hoff = gi == GR - 1 ? last_half_off : 512;
            if (on) { S.a[g] = ld_stream_16(q + g * GB); S.b[g] = ld_stream_16(q + g * GB + hoff); S.s[g] = d[g * 32]; }
            else { S.a[g] = make_int4(0, 0, 0, 0); S.b[g] = S.a[g]; S.s[g] = 0; }
        } else {
            if (on) { S.a[g] = ld_stream_16(q + g * GB); S.s[g] = d[g * 32]; }
            else { S.a[g] = make_int4(0, 0, 0, 0); S.s[g] = 0; }
        }
    }
}
template <int TYPE>
__device__ __forceinline__ float mk_seg_dot(const MkSeg& S, int seg, const int4* aq_l, const float* ad_l, const int* as_l) {
    float acc = 0.0f;
    const int4* aq = aq_l + seg * (MK_SEG * 64);
    const float* ad = ad_l + seg * (MK_SEG * 32);
#pragma unroll
    for (int g = 0; g < MK_SEG; g++) {
        if constexpr (TYPE == CC_Q8_0) {
            int sumi = mk_dp16(S.a[g], aq[g * 64]) + mk_dp16(S.b[g], aq[g * 64 + 1]);
            acc += (float)sumi * h2f_bits(S.s[g]) * ad[g * 32];
        } else {
            const int4 w = S.a[g];
            int4 lo = make_int4(w.x & 0x0F0F0F0F, w.y & 0x0F0F0F0F, w.z & 0x0F0F0F0F, w.w & 0x0F0F0F0F);
            int4 hi = make_int4((w.x >> 4) & 0x0F0F0F0F, (w.y >> 4) & 0x0F0F0F0F, (w.z >> 4) & 0x0F0F0F0F, (w.w >> 4) & 0x0F0F0F0F);
            int sumi = mk_dp16(lo, aq[g * 64]) + mk_dp16(hi, aq[g * 64 + 1]) - 8 * as_l[(seg * MK_SEG + g) * 32];
            acc += (float)sumi * h2f_bits(S.s[g]) * ad[g * 32];
        }
    }
    return acc;
}

__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nMK_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra MK_WAIT_%=;\n}\n" ::"r"(mbar), "r"(parity) : "memory");
}
// look-ahead arguments of a coming MATVEC phase, fetched one word per thread at phase start (64 threads per slot)
struct MkNext { StreamArgs mv; int wtype; int norm_n; const float* norm_w; };
static_assert(sizeof(StreamArgs) % 4 == 0 && sizeof(StreamArgs) / 4 + 4 <= 64, "MkNext fetch layout: 64 threads per look-ahead slot");
struct MkPipe { MkSeg buf0, buf1; };      // register stages of the weight stream, live across phases and barriers

// (Tried and removed in round 2, profiles/r02a_*, r02b_*: an L2 look-ahead of each warp's coming rows -- cp.async.bulk.prefetch.L2 as
// well as per-lane prefetch.global.L2 -- made the token 7-10 % SLOWER: a bulk prefetch request occupies its issuing thread for
// ~9 us per 4 KB, line prefetches cost issue slots at the phase boundary, and the phase bodies already stream at HBM speed.)

// geometry of one MATVEC phase for this warp
struct MkGeo {
    int nb, GR, NSEG, U, last_half_off, gw, TW, rpc;
    bool pair;
};
__device__ __forceinline__ MkGeo mk_geo(const StreamArgs& A) {
    MkGeo g;
    const int warp = threadIdx.x >> 5;
    g.nb = A.k >> 5; g.GR = (g.nb + 31) >> 5; g.NSEG = (g.GR + MK_SEG - 1) / MK_SEG;
    // warp-major numbering: when rows do not divide by the warp count, every SM gets the same mix of k- and (k+1)-row warps
    // (CTA-major numbering left the last SMs with half the work of the first ones)
    g.gw = warp * gridDim.x + blockIdx.x; g.TW = gridDim.x * MK_WARPS;
    g.pair = A.epilogue == 2;
    const StreamMats& M = A.mats;
    const int m_cat = g.pair ? M.m[0] : M.m[0] + (M.n > 1 ? M.m[1] : 0) + (M.n > 2 ? M.m[2] : 0);
    int n_rows = g.gw < m_cat ? (m_cat - g.gw + g.TW - 1) / g.TW : 0;
    g.rpc = 0;
    if (A.epilogue == 3) {
        // exchange phases: every CTA owns ONE contiguous block of rows (rpc rows, a multiple of 4), warp w takes the rows w, w + 16, ...
        // of the block -- so the CTA's partial results are one contiguous run of floats and go to each peer as a single coalesced store
        g.rpc = (((m_cat + (int)gridDim.x - 1) / (int)gridDim.x) + 3) & ~3;
        const int first = (int)blockIdx.x * g.rpc;
        const int cnt = min(g.rpc, max(0, m_cat - first));
        g.gw = first + warp; g.TW = MK_WARPS;
        n_rows = warp < cnt ? (cnt - warp + MK_WARPS - 1) / MK_WARPS : 0;
    }
    g.U = (g.pair ? 2 * n_rows : n_rows) * g.NSEG;
    g.last_half_off = 16 * (g.nb - 32 * (g.GR - 1));
    return g;
}
template <int TYPE>
__device__ __forceinline__ MkRowPtr mk_vrow_ptr(const StreamMats& M, const MkGeo& g, int i, int lane) {
    constexpr int BB = TYPE == CC_Q8_0 ? 32 : 16;
    int mat = 0, r;
    if (g.pair) { mat = i & 1; r = g.gw + (i >> 1) * g.TW; }
    else {
        r = g.gw + i * g.TW;
        if (M.n > 1 && r >= M.m[0]) { r -= M.m[0]; mat = 1; if (M.n > 2 && r >= M.m[1]) { r -= M.m[1]; mat = 2; } }
    }
    const uint8_t* q0 = mat == 0 ? M.qs[0] : mat == 1 ? M.qs[1] : M.qs[2];
    const uint16_t* d0 = mat == 0 ? M.d[0] : mat == 1 ? M.d[1] : M.d[2];
    MkRowPtr p;
    p.q = q0 + (size_t)r * g.nb * BB + lane * 16;
    p.d = d0 + (size_t)r * g.nb + lane;
    return p;
}

// Issue the loads of this warp's first two segments of a MATVEC phase (register stages).  Weights are immutable, so this may run
// long before the phase itself -- across barriers and small phases -- keeping HBM busy while the grid synchronises.
template <int TYPE>
__device__ __forceinline__ void matvec_prefetch(const StreamArgs& A, MkPipe& P) {
    const int lane = threadIdx.x & 31;
    const MkGeo g = mk_geo(A);
    int l_i = 0, l_seg = 0;
    MkRowPtr l_ptr = mk_vrow_ptr<TYPE>(A.mats, g, 0, lane);
    auto advance_load = [&]() { if (++l_seg == g.NSEG) { l_seg = 0; l_ptr = mk_vrow_ptr<TYPE>(A.mats, g, ++l_i, lane); } };
    mk_seg_load<TYPE>(P.buf0, l_ptr, l_seg, g.nb, g.GR, g.last_half_off, lane, g.U > 0); advance_load();
    mk_seg_load<TYPE>(P.buf1, l_ptr, l_seg, g.nb, g.GR, g.last_half_off, lane, g.U > 1);
}

// precondition: the pipe holds this warp's segments 0, 1 (matvec_prefetch).  s_w: staging area of the norm weights at the top
// of dynamic shared memory; w_staged: they were already requested there (cp.async, before the barrier) by the look-ahead.
template <int TYPE>
__device__ void phase_matvec(const MkPhase& ph, uint8_t* smem, float* s_w, bool w_staged, bool x_staged, const uint16_t* exp_lut, MkPipe& P,
                             const CommDev& comm, unsigned xseq, unsigned long long* stamp1, const MkNext* early_next, int next_w) {
    const StreamArgs& A = ph.mv;
    const int k = A.k;
    const MkGeo g = mk_geo(A);
    const int nb = g.nb, GR = g.GR, NSEG = g.NSEG, U = g.U, gw = g.gw, TW = g.TW;
    const bool pair = g.pair;
    const int nbp = NSEG * MK_SEG * 32;
    int8_t* s_q = (int8_t*)smem;
    float* s_d = (float*)(smem + (size_t)nbp * 32);
    int* s_s = (int*)(smem + (size_t)nbp * 32 + (size_t)nbp * 4);
    const int lane = threadIdx.x & 31;
    const StreamMats& M = A.mats;
    // load cursor: points at segment 2
    int l_i = 0, l_seg = 0;
    MkRowPtr l_ptr = mk_vrow_ptr<TYPE>(M, g, 0, lane);
    auto advance_load = [&]() { if (++l_seg == NSEG) { l_seg = 0; l_ptr = mk_vrow_ptr<TYPE>(M, g, ++l_i, lane); } };
    if (++l_seg == NSEG) { l_seg = 0; l_ptr = mk_vrow_ptr<TYPE>(M, g, ++l_i, lane); }       // -> segment 1 (already requested)
    if (++l_seg == NSEG) { l_seg = 0; l_ptr = mk_vrow_ptr<TYPE>(M, g, ++l_i, lane); }       // -> segment 2
    MkSeg& buf0 = P.buf0;
    MkSeg& buf1 = P.buf1;
    if (ph.x) {
        // Fused prologue: [rms_norm * w] + Q8_0 quantisation of x, computed by EVERY CTA straight into its shared memory
        // (redundant across SMs, ~1.5 us of issue time) -- cheaper than a separate NORMQ phase, which costs a grid barrier
        // (~2.1 us) plus its own latency chain.  The weight segments requested by matvec_prefetch are in flight meanwhile.
        const int n = k;
        const int warp = threadIdx.x >> 5;
        float* s_red = (float*)(smem + (size_t)nbp * 40);                // scratch behind the activation arrays (256 B), then the 2 KB exchange stage
        float* s_x = s_red + 64 + 512;                                    // f32 copy of x
        {   // one L2 round trip: every 16-byte chunk of x (and of the norm weights) requested at once
            const int n4 = n >> 2;
            const unsigned sx = (unsigned)__cvta_generic_to_shared(s_x), sw = (unsigned)__cvta_generic_to_shared(s_w);
            if (ph.norm_w && !w_staged)
                for (int i = threadIdx.x; i < n4; i += MK_THREADS)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sw + i * 16), "l"(ph.norm_w + i * 4) : "memory");
            if (ph.red_n) {
                // second half of the exchange that ended the previous phase (comm.cu): x = sum over ranks of the partial rows
                // in rank order (+ residual), rebuilt by every CTA from this GPU's window -- no separate REDUCE phase
                const float* base = comm.data[comm.rank] + (size_t)(xseq & 1u) * CC_COMM_MAX_RANKS * CC_COMM_MAX_ELEMS;
                for (int i = threadIdx.x; i < n4; i += MK_THREADS) {
                    float4 acc4 = __ldcg((const float4*)base + i);
                    for (int p = 1; p < comm.world; p++) {
                        const float4 t4 = __ldcg((const float4*)(base + (size_t)p * CC_COMM_MAX_ELEMS) + i);
                        acc4.x += t4.x; acc4.y += t4.y; acc4.z += t4.z; acc4.w += t4.w;
                    }
                    if (ph.red_res) { const float4 r4 = __ldcg((const float4*)ph.red_res + i); acc4.x += r4.x; acc4.y += r4.y; acc4.z += r4.z; acc4.w += r4.w; }
                    ((float4*)s_x)[i] = acc4;
                }
            } else if (!x_staged) {
                for (int i = threadIdx.x; i < n4; i += MK_THREADS)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sx + i * 16), "l"(ph.x + i * 4) : "memory");
            }
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            if (stamp1) stamp1[3] = globaltimer_ns();          // x (and the norm weights) are in shared memory
        }
        float rms = 1.0f;
        if (ph.norm_w) {
            float ss = 0.0f;
            const float4* x4 = (const float4*)s_x;
            for (int i = threadIdx.x; i < (n >> 2); i += MK_THREADS) { float4 v = x4[i]; ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
            ss = warp_sum(ss);
            if (lane == 0) s_red[warp] = ss;
            __syncthreads();
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < MK_WARPS; w++) t += s_red[w];
            rms = sqrtf(t / (float)n + ph.eps);
        }
        if (stamp1) stamp1[4] = globaltimer_ns();              // rms known
        if (ph.orig && blockIdx.x == 0)                              // Tensor::dup of the un-normalised row (llama2.rs:227,607)
            for (int i = threadIdx.x; i < (n >> 2); i += MK_THREADS) ((float4*)ph.orig)[i] = ((const float4*)s_x)[i];
        // quantise: 4 consecutive elements per thread, 8 threads per 32-block, 64 blocks per pass (same arithmetic per element as
        // quantize.cu: d = amax / 127, q = trunc(x / d), stored scale = f32(f16(d)))
        const int sub = threadIdx.x & 7;
        for (int b = threadIdx.x >> 3; b < nbp; b += MK_THREADS / 8) {        // nbp % 128 == 0: uniform trip count per warp
            const bool live = b < nb;
            float4 v = live ? ((const float4*)s_x)[b * 8 + sub] : make_float4(0, 0, 0, 0);
            if (ph.norm_w && live) {
                const float4 w4 = ((const float4*)s_w)[b * 8 + sub];
                v.x = (v.x / rms) * w4.x; v.y = (v.y / rms) * w4.y; v.z = (v.z / rms) * w4.z; v.w = (v.w / rms) * w4.w;
            }
            float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            const float d = amax / 127.0f;
            const int q0 = live ? __float2int_rz(v.x / d) : 0, q1 = live ? __float2int_rz(v.y / d) : 0;
            const int q2 = live ? __float2int_rz(v.z / d) : 0, q3 = live ? __float2int_rz(v.w / d) : 0;
            ((int*)s_q)[b * 8 + sub] = (q0 & 255) | ((q1 & 255) << 8) | ((q2 & 255) << 16) | (q3 << 24);
            if constexpr (TYPE == CC_Q4_0) {
                int sq = q0 + q1 + q2 + q3;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                if (sub == 0) s_s[b] = sq;
            }
            if (sub == 0) s_d[b] = live ? __half2float(__float2half_rn(d)) : 0.0f;
        }
    } else {   // stage the quantised activation (written by other CTAs in the previous phase: L2 loads)
        const uint8_t* act = (const uint8_t*)A.act;
        const int4* gq = (const int4*)act;
        int4* sq4 = (int4*)s_q;
        const float* gd = (const float*)(act + ((k + 15) & ~15));
        const int* gs = (const int*)(act + ((k + 15) & ~15) + ((nb * 4 + 15) & ~15));
        if (nbp <= MK_THREADS) {           // every load of the thread is issued before its first store (one L2 round trip, not three)
            const int i0 = threadIdx.x, i1 = threadIdx.x + MK_THREADS;
            const int4 z4 = make_int4(0, 0, 0, 0);
            const int4 qa = i0 < nb * 2 ? __ldcg(gq + i0) : z4, qb = i1 < nb * 2 ? __ldcg(gq + i1) : z4;
            const float dv = i0 < nb ? __ldcg(gd + i0) : 0.0f;
            int sv = 0;
            if constexpr (TYPE == CC_Q4_0) sv = i0 < nb ? __ldcg(gs + i0) : 0;
            if (i0 < nbp * 2) sq4[i0] = qa;
            if (i1 < nbp * 2) sq4[i1] = qb;
            if (i0 < nbp) { s_d[i0] = dv; if constexpr (TYPE == CC_Q4_0) s_s[i0] = sv; }
        } else {
            for (int i = threadIdx.x; i < nbp * 2; i += MK_THREADS) sq4[i] = i < nb * 2 ? __ldcg(gq + i) : make_int4(0, 0, 0, 0);
            for (int i = threadIdx.x; i < nbp; i += MK_THREADS) {
                s_d[i] = i < nb ? __ldcg(gd + i) : 0.0f;
                if constexpr (TYPE == CC_Q4_0) s_s[i] = i < nb ? __ldcg(gs + i) : 0;
            }
        }
    }
    if (early_next && threadIdx.x < 128) {      // look-ahead arguments (loaded at phase start) become visible with the barrier below
        const int slot = threadIdx.x >> 6, t = threadIdx.x & 63;
        if (t < (int)(sizeof(StreamArgs) / 4) + 4) ((int*)&early_next[slot])[t] = next_w;
    }
    __syncthreads();
    if (stamp1) *stamp1 = globaltimer_ns();
    const int4* aq_l = (const int4*)s_q + 2 * lane;
    const float* ad_l = s_d + lane;
    const int* as_l = s_s + lane;
    float* s_part = (float*)(smem + (size_t)nbp * 40 + 256);             // exchange stage: this CTA's block of partial rows (<= MK_XSTAGE_ROWS floats)
    float acc = 0.0f, first = 0.0f;
    int c_i = 0, c_seg = 0;
    // Epilogues that need a value from memory (the residual, or the exp LUT entry of silu) are finished ONE ROW LATER: the load
    // is issued when the row's dot is known and consumed after the next row, so the warp never stalls an L2 round trip with its
    // weight stream idle (in-order issue).  lane 0 only.
    float pend_a = 0.0f, pend_b = 0.0f, pend_res = 0.0f;
    unsigned short pend_lut = 0;
    int pend_row = -1;
    auto flush_pending = [&]() {
        if (lane == 0 && pend_row >= 0) {
            if (pair) M.out[0][pend_row] = (pend_a / (1.0f + h2f_bits(pend_lut))) * pend_b;
            else M.out[0][pend_row] = pend_a + pend_res;
        }
        pend_row = -1;
    };
    auto finish_segment = [&]() {
        if (++c_seg < NSEG) return;
        c_seg = 0;
        float r = warp_sum(acc);
        acc = 0.0f;
        const int i = c_i++;
        if (pair) {
            if ((i & 1) == 0) { first = r; return; }
            flush_pending();
            if (lane == 0) {
                pend_a = first; pend_b = r; pend_row = gw + (i >> 1) * TW;
                pend_lut = exp_lut[f2h_bits(-first)];
            }
            return;
        }
        if (A.epilogue == 1) {                 // single matrix: out = dot + residual (llama2.rs:266,636)
            flush_pending();
            if (lane == 0) { pend_a = r; pend_row = gw + i * TW; pend_res = ldcg_f(A.residual + pend_row); }
            return;
        }
        if (lane == 0) {
            int mat = 0, rr = gw + i * TW;
            if (M.n > 1 && rr >= M.m[0]) { rr -= M.m[0]; mat = 1; if (M.n > 2 && rr >= M.m[1]) { rr -= M.m[1]; mat = 2; } }
            if (A.epilogue == 3) {         // partial row -> this CTA's stage; sent to the peers as one run when the phase body is done
                s_part[rr - (int)blockIdx.x * g.rpc] = r;
                return;
            }
            float* o = mat == 0 ? M.out[0] : mat == 1 ? M.out[1] : M.out[2];
            o[rr] = r;
        }
    };
    for (int u = 0; u < U; u += 2) {          // two segments (8 KB of Q8_0) in flight per warp at all times
        acc += mk_seg_dot<TYPE>(buf0, c_seg, aq_l, ad_l, as_l);
        finish_segment();
        mk_seg_load<TYPE>(buf0, l_ptr, l_seg, nb, GR, g.last_half_off, lane, u + 2 < U);
        advance_load();
        if (u + 1 >= U) break;
        acc += mk_seg_dot<TYPE>(buf1, c_seg, aq_l, ad_l, as_l);
        finish_segment();
        mk_seg_load<TYPE>(buf1, l_ptr, l_seg, nb, GR, g.last_half_off, lane, u + 3 < U);
        advance_load();
    }
    flush_pending();
    if (A.epilogue == 3) {
        // the CTA's block of partial rows -> slot[rank] of every GPU's exchange window: warp p serves peer p with ONE coalesced NVLink
        // store of 16 bytes per lane (28 rows = 112 contiguous bytes at 7B shapes) instead of one 4-byte store per row and peer
        __syncthreads();
        const int warp = threadIdx.x >> 5;
        const int first_row = (int)blockIdx.x * g.rpc;
        const int m_all = M.m[0];
        const int cnt = min(g.rpc, max(0, m_all - first_row));
        if (warp < comm.world) {
            const size_t off = ((size_t)((xseq + 1u) & 1u) * CC_COMM_MAX_RANKS + comm.rank) * CC_COMM_MAX_ELEMS + first_row;
            for (int c4 = lane * 4; c4 < cnt; c4 += 128) *(float4*)(comm.data[warp] + off + c4) = *(const float4*)(s_part + c4);
        }
    }
    // this warp is done: its register stages are free, so it requests its first segments of the next MATVEC phase right away instead
    // of idling until the slowest warp of the CTA reaches the barrier (the tail of a phase becomes prefetch time)
    if (early_next) { if (early_next[0].wtype == CC_Q8_0) matvec_prefetch<CC_Q8_0>(early_next[0].mv, P); else if (early_next[0].wtype == CC_Q4_0) matvec_prefetch<CC_Q4_0>(early_next[0].mv, P); }
}

// ---- generic MATVEC phase: K-quant weights (Q2_K .. Q6_K, Q8_K) against the Q8_K-quantised activation ---------------------------
// Same shape as phase_matvec -- fused prologue ([rms_norm * w] + activation quantisation, recomputed by every CTA), rows dealt
// warp-major, the same epilogues -- but the row dot is the type's T::row_dot of vecdot.cuh (what the eager matvec_kernel runs, hence
// the same bits) and the weights are not pipelined through registers across phase boundaries.
// shared memory: qs [k] | d [k/256] | bsums [k/16] (TKBase) | reduction scratch | f32 x
__device__ __forceinline__ int mk_generic_sx_offset(int k) { return ((TKBase::smem_bytes(k) + 15) & ~15) + 256; }
template <class T>
__device__ void phase_matvec_generic(const MkPhase& ph, uint8_t* smem, float* s_w, bool w_staged, bool x_staged, const uint16_t* exp_lut, MkPipe& P, unsigned long long* stamp1) {
    const StreamArgs& A = ph.mv;
    const StreamMats& M = A.mats;
    const int k = A.k;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int8_t* s_q = (int8_t*)smem;
    float* s_d = (float*)(smem + al16i(k));
    int16_t* s_bs = (int16_t*)(smem + al16i(k) + al16i(k / 256 * 4));
    float* s_red = (float*)(smem + mk_generic_sx_offset(k) - 256);
    float* s_x = (float*)(smem + mk_generic_sx_offset(k));
    {   // one L2 round trip: the f32 row (unless requested right after the barrier) and the norm weights (unless staged before it)
        const int n4 = k >> 2;
        const unsigned sx = (unsigned)__cvta_generic_to_shared(s_x), sw = (unsigned)__cvta_generic_to_shared(s_w);
        if (ph.norm_w && !w_staged)
            for (int i = threadIdx.x; i < n4; i += MK_THREADS) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sw + i * 16), "l"(ph.norm_w + i * 4) : "memory");
        if (!x_staged)
            for (int i = threadIdx.x; i < n4; i += MK_THREADS) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sx + i * 16), "l"(ph.x + i * 4) : "memory");
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (stamp1) stamp1[3] = globaltimer_ns();
    }
    float rms = 1.0f;
    if (ph.norm_w) {                                                    // canonical order (common.cuh)
        float ss = 0.0f;
        const float4* x4 = (const float4*)s_x;
        for (int i = threadIdx.x; i < (k >> 2); i += MK_THREADS) ss += cc_sq4(x4[i]);
        rms = sqrtf(cc_block_sum_512(ss, s_red) / (float)k + ph.eps);
    }
    if (stamp1) stamp1[4] = globaltimer_ns();
    if (ph.orig && blockIdx.x == 0)                                     // Tensor::dup of the un-normalised row (llama2.rs:227,607)
        for (int i = threadIdx.x; i < (k >> 2); i += MK_THREADS) ((float4*)ph.orig)[i] = ((const float4*)s_x)[i];
    for (int sb = warp; sb < (k >> 8); sb += MK_WARPS) {                // one warp per 256-element super-block (buf_q8_k.rs:84-131)
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            v[i] = s_x[sb * 256 + lane * 8 + i];
            if (ph.norm_w) v[i] = (v[i] / rms) * s_w[sb * 256 + lane * 8 + i];
        }
        cc_quant_q8k_sblock(v, lane, s_q + sb * 256, s_d + sb, s_bs + sb * 16);
    }
    __syncthreads();
    if (stamp1) *stamp1 = globaltimer_ns();
    // rows: the same warp-major dealing and epilogues as phase_matvec
    const int gw = warp * gridDim.x + blockIdx.x, TW = gridDim.x * MK_WARPS;
    const bool pair = A.epilogue == 2;
    const int m_cat = pair ? M.m[0] : M.m[0] + (M.n > 1 ? M.m[1] : 0) + (M.n > 2 ? M.m[2] : 0);
    const int n_rows = gw < m_cat ? (m_cat - gw + TW - 1) / TW : 0;
    const int n_vrows = pair ? 2 * n_rows : n_rows;
    float first = 0.0f, pend_a = 0.0f, pend_b = 0.0f, pend_res = 0.0f;
    unsigned short pend_lut = 0;
    int pend_row = -1;
    auto flush_pending = [&]() {
        if (lane == 0 && pend_row >= 0) {
            if (pair) M.out[0][pend_row] = (pend_a / (1.0f + h2f_bits(pend_lut))) * pend_b;
            else M.out[0][pend_row] = pend_a + pend_res;
        }
        pend_row = -1;
    };
    auto locate = [&](int i, int& mat) -> int {                          // i-th virtual row of this warp -> (matrix, row)
        int r;
        mat = 0;
        if (pair) { mat = i & 1; r = gw + (i >> 1) * TW; }
        else {
            r = gw + i * TW;
            if (M.n > 1 && r >= M.m[0]) { r -= M.m[0]; mat = 1; if (M.n > 2 && r >= M.m[1]) { r -= M.m[1]; mat = 2; } }
        }
        return r;
    };
    auto planes = [&](int mat) -> WPlanes {
        WPlanes W;
        W.p[0] = mat == 0 ? M.qs[0] : mat == 1 ? M.qs[1] : M.qs[2];
        W.p[1] = (const uint8_t*)(mat == 0 ? M.d[0] : mat == 1 ? M.d[1] : M.d[2]);
        W.p[2] = mat == 0 ? M.p2[0] : mat == 1 ? M.p2[1] : M.p2[2];
        W.p[3] = mat == 0 ? M.p3[0] : mat == 1 ? M.p3[1] : M.p3[2];
        return W;
    };
    auto emit = [&](int i, float v) {                                    // row i of this warp is reduced: epilogue
        int mat;
        const int r = locate(i, mat);
        if (pair) {
            if ((i & 1) == 0) { first = v; return; }
            flush_pending();
            if (lane == 0) { pend_a = first; pend_b = v; pend_row = r; pend_lut = exp_lut[f2h_bits(-first)]; }
        } else if (A.epilogue == 1) {
            flush_pending();
            if (lane == 0) { pend_a = v; pend_row = r; pend_res = ldcg_f(A.residual + r); }
        } else if (lane == 0) {
            float* o = mat == 0 ? M.out[0] : mat == 1 ? M.out[1] : M.out[2];
            o[r] = v;
        }
    };
    if constexpr (T::kSegmented) {
        // rows cut into segments of 16 super-blocks; the loads of segment u + 2 are issued before segment u + 1 is consumed (two
        // segments = 16-24 LDG.128 per lane in flight), across row boundaries
        const int NSEG = ((k >> 8) + 15) >> 4;
        const int U = n_vrows * NSEG;
        KSeg& S0 = P.buf0;                                                   // the weight pipe's registers (no streaming look-ahead is pending: caller)
        KSeg& S1 = P.buf1;
        int x0[4] = {0, 0, 0, 0}, x1[4] = {0, 0, 0, 0};
        int l_i = 0, l_seg = 0;
        auto load = [&](KSeg& S, int (&x)[4], bool valid) {
            if (valid) { int mat; const int r = locate(l_i, mat); T::seg_load(S, x, planes(mat), r, k, l_seg, lane); }
            if (++l_seg == NSEG) { l_seg = 0; l_i++; }
        };
        load(S0, x0, U > 0);
        load(S1, x1, U > 1);
        float acc = 0.0f;
        int c_i = 0, c_seg = 0;
        auto finish = [&]() { if (++c_seg < NSEG) return; c_seg = 0; const float v = warp_sum(acc); acc = 0.0f; emit(c_i++, v); };
        for (int u = 0; u < U; u += 2) {
            acc = T::seg_dot(S0, x0, k, c_seg, smem, lane, acc);
            finish();
            load(S0, x0, u + 2 < U);
            if (u + 1 >= U) break;
            acc = T::seg_dot(S1, x1, k, c_seg, smem, lane, acc);
            finish();
            load(S1, x1, u + 3 < U);
        }
    } else {
        for (int i = 0; i < n_vrows; i++) {
            int mat;
            const int r = locate(i, mat);
            emit(i, warp_sum(T::row_dot(planes(mat), r, k, smem, lane)));
        }
    }
    flush_pending();
}

// ---- ATTN phase: arithmetic of fused.cu attn_decode_kernel, heads dealt to CTAs.  The K (then V) rows of the head are
// staged in shared memory in chunks of AT_CH positions with ALL loads of a chunk in flight at once: at decode the cost of
// this phase is HBM/L2 latency, not bandwidth, so round trips are what matters.
#define AT_CH 64
#define AT_NBUF 3
template <bool KV_F16>
__device__ void phase_attn(const MkPhase& ph, float* sm, float* s_red, const uint8_t* dyn, const uint16_t* exp_lut, unsigned abar0, unsigned& apar) {
    const AttnArgs& a = ph.at;
    const int n_heads = a.n_heads, n_kv = a.n_kv, hd = a.hd, rope_dim = a.rope_dim;
    const int64_t seq_stride = a.seq_stride;
    const int64_t* dynv = (const int64_t*)(dyn + ph.dyn_off);
    const float* rope_tab = (const float*)(dyn + ph.rope_off);
    const int kv_len = (int)dynv[1], L = kv_len + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* s_q = sm; float* s_k = sm + hd; float* s_v = sm + 2 * hd; float* s_p = sm + 3 * hd;
    // AT_NBUF chunk buffers of AT_CH cache rows each (raw bytes: f32 or f16), filled by TMA bulk copies -- the rows of one kv head
    // are contiguous -- in a K-chunks-then-V-chunks job sequence with AT_NBUF jobs in flight; one mbarrier per buffer.
    uint8_t* s_buf = (uint8_t*)(sm + 3 * hd + ((a.max_len + 8 + 3) & ~3));
    const unsigned s_buf_smem = (unsigned)__cvta_generic_to_shared(s_buf);
    constexpr int ELT = KV_F16 ? 2 : 4;
    const unsigned buf_bytes = (unsigned)(AT_CH * hd * ELT);
    const int pairs = rope_dim >> 1;
    const int hd4 = hd >> 2;
    const int NC = (kv_len + AT_CH - 1) / AT_CH;                  // chunks per pass; jobs 0..NC-1 = K chunks, NC..2NC-1 = V chunks
    auto issue_job = [&](int g, int j) {                            // one elected thread
        const int c = j < NC ? j : j - NC;
        const int p0 = c * AT_CH, cnt = min(AT_CH, kv_len - p0);
        const uint8_t* src = (const uint8_t*)(j < NC ? a.kcache : a.vcache) + ((int64_t)g * seq_stride + (int64_t)p0 * hd) * ELT;
        const unsigned bytes = (unsigned)(cnt * hd * ELT), bar = abar0 + 8u * (unsigned)(j % AT_NBUF), dst = s_buf_smem + (unsigned)(j % AT_NBUF) * buf_bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
    };
    auto wait_job = [&](int j) {                                    // all threads, in job order
        const int bsel = j % AT_NBUF;
        mbar_wait(abar0 + 8u * (unsigned)bsel, (apar >> bsel) & 1u);
        apar ^= 1u << bsel;
    };
    auto ld_kv = [&](const uint8_t* buf, int idx) -> float { return KV_F16 ? __half2float(((const __half*)buf)[idx]) : ((const float*)buf)[idx]; };
    for (int h = blockIdx.x; h < n_heads; h += gridDim.x) {
        const int g = KV_F16 ? h / (n_heads / n_kv) : h % n_kv;
        // the first AT_NBUF jobs are requested up front; every later job is issued as soon as its buffer has been consumed
        if (threadIdx.x == 0) for (int j = 0; j < min(AT_NBUF, 2 * NC); j++) issue_job(g, j);
        for (int i = threadIdx.x; i < hd; i += MK_THREADS) {
            float qv, kvv;
            if (i < rope_dim) {
                const int j = i >> 1;
                const float c = rope_tab[j], s = rope_tab[pairs + j];
                const float q0 = ldcg_f(a.q + h * hd + 2 * j), q1 = ldcg_f(a.q + h * hd + 2 * j + 1);
                const float k0 = ldcg_f(a.k + g * hd + 2 * j), k1 = ldcg_f(a.k + g * hd + 2 * j + 1);
                qv = (i & 1) ? q0 * s + q1 * c : q0 * c - q1 * s;
                kvv = (i & 1) ? k0 * s + k1 * c : k0 * c - k1 * s;
            } else {
                qv = ldcg_f(a.q + h * hd + i);
                kvv = ldcg_f(a.k + g * hd + i);
            }
            s_q[i] = qv * a.scale;
            s_k[i] = kvv;
            s_v[i] = ldcg_f(a.v + g * hd + i);
        }
        __syncthreads();
        const bool owner = KV_F16 ? (h % (n_heads / n_kv) == 0) : (h < n_kv);
        if (owner) {
            for (int i = threadIdx.x; i < hd; i += MK_THREADS) {
                const int64_t off = (int64_t)g * seq_stride + (int64_t)kv_len * hd + i;
                if (KV_F16) { ((__half*)a.kcache)[off] = __float2half_rn(s_k[i]); ((__half*)a.vcache)[off] = __float2half_rn(s_v[i]); }
                else { ((float*)a.kcache)[off] = s_k[i]; ((float*)a.vcache)[off] = s_v[i]; }
            }
        }
        // scores, chunk by chunk; per-lane summation order i = lane, lane+32, ... as in fused.cu
        for (int j = 0; j < NC; j++) {
            const int p0 = j * AT_CH, cnt = min(AT_CH, kv_len - p0);
            const uint8_t* kb = s_buf + (size_t)(j % AT_NBUF) * buf_bytes;
            wait_job(j);
            for (int s = warp; s < cnt; s += MK_WARPS) {
                float acc = 0.0f;
                for (int i = lane; i < hd; i += 32) acc += (KV_F16 ? __half2float(__float2half_rn(s_q[i])) : s_q[i]) * ld_kv(kb, s * hd + i);
                acc = warp_sum(acc);
                if (lane == 0) s_p[p0 + s] = acc;
            }
            __syncthreads();                                       // buffer consumed by every warp -> refill it
            if (threadIdx.x == 0 && j + AT_NBUF < 2 * NC) issue_job(g, j + AT_NBUF);
        }
        if (warp == 0) {                                           // this token's own position
            float acc = 0.0f;
            for (int i = lane; i < hd; i += 32) {
                if (KV_F16) acc += __half2float(__float2half_rn(s_q[i])) * __half2float(__float2half_rn(s_k[i]));
                else acc += s_q[i] * s_k[i];
            }
            acc = warp_sum(acc);
            if (lane == 0) s_p[kv_len] = acc;
        }
        __syncthreads();
        float m = -INFINITY;
        for (int s = threadIdx.x; s < L; s += MK_THREADS) m = fmaxf(m, s_p[s]);
        m = warp_max(m);
        if (lane == 0) s_red[warp] = m;
        __syncthreads();
        m = s_red[0];
#pragma unroll
        for (int w = 1; w < MK_WARPS; w++) m = fmaxf(m, s_red[w]);
        __syncthreads();
        float sum = 0.0f;
        for (int s = threadIdx.x; s < L; s += MK_THREADS) {
            float e = h2f_bits(exp_lut[f2h_bits(s_p[s] - m)]);
            s_p[s] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        if (lane == 0) s_red[warp] = sum;
        __syncthreads();
        sum = 0.0f;
#pragma unroll
        for (int w = 0; w < MK_WARPS; w++) sum += s_red[w];
        for (int s = threadIdx.x; s < L; s += MK_THREADS) s_p[s] = s_p[s] / sum;
        __syncthreads();
        // out[d] = sum_s p[s] * V[s][d], sequential over s (batch_matmul.rs:60-68 order); F16: f16 accumulation (buf_f16.rs:152-163)
        float accf = 0.0f;
        __half acch = __float2half_rn(0.0f);
        const int d = threadIdx.x;
        for (int j = NC; j < 2 * NC; j++) {
            const int p0 = (j - NC) * AT_CH, cnt = min(AT_CH, kv_len - p0);
            const uint8_t* vb = s_buf + (size_t)(j % AT_NBUF) * buf_bytes;
            wait_job(j);
            if (d < hd) {
                for (int s = 0; s < cnt; s++) {
                    if (KV_F16) acch = __hadd(acch, __hmul(((const __half*)vb)[s * hd + d], __float2half_rn(s_p[p0 + s])));
                    else accf += s_p[p0 + s] * ((const float*)vb)[s * hd + d];
                }
            }
            __syncthreads();
            if (threadIdx.x == 0 && j + AT_NBUF < 2 * NC) issue_job(g, j + AT_NBUF);
        }
        float* s_o = s_k;
        __syncthreads();
        if (d < hd) {
            float o;
            if (KV_F16) { acch = __hadd(acch, __hmul(__float2half_rn(s_v[d]), __float2half_rn(s_p[kv_len]))); o = __half2float(acch); }
            else { accf += s_p[kv_len] * s_v[d]; o = accf; }
            a.out[h * hd + d] = o;
            s_o[d] = o;
        }
        __syncthreads();
        if (a.act_scratch) {
            ActQ8_0 act = ph.act;
            for (int b = warp; b < (hd >> 5); b += MK_WARPS) {
                float v = s_o[b * 32 + lane];
                float amax = warp_max(fabsf(v));
                float dd = amax / 127.0f;
                int qq = __float2int_rz(v / dd);
                const int gb = h * (hd >> 5) + b;
                act.qs[gb * 32 + lane] = (int8_t)qq;
                int ss = warp_sum_i(qq);
                if (lane == 0) { act.d[gb] = __half2float(__float2half_rn(dd)); act.isum[gb] = ss; }
            }
        }
        __syncthreads();
    }
}

// ---- ROWS phase: copy_rows_from with the row indices in dyn (embedding lookup / row pick) -----------------------------------
__device__ void phase_rows(const MkPhase& ph, const uint8_t* dyn) {
    const int64_t* rows = ph.rows_dev ? (const int64_t*)ph.rows_dev : (const int64_t*)(dyn + ph.dyn_off);      // a device slot: read through L2 (written by a previous launch)
    const int64_t total = (int64_t)ph.n_rows * ph.cols;
    for (int64_t i = (int64_t)blockIdx.x * MK_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * MK_THREADS) {
        const int64_t r = i / ph.cols, c = i - r * ph.cols;
        const int64_t e = rows[r] * ph.cols + c;
        float v;
        if (ph.src_dtype == CC_F32) v = __ldcg((const float*)ph.planes.p[0] + e);
        else if (ph.src_dtype == CC_F16) v = __half2float(((const __half*)ph.planes.p[0])[e]);
        else v = dequant_elem(ph.src_dtype, ph.planes, e);
        if (ph.dst_dtype == CC_F32) ((float*)ph.dst)[i] = v; else ((__half*)ph.dst)[i] = __float2half_rn(v);
    }
}

// ---- ARGMAX phase: greedy sampling on the device (ops.cu argmax_kernel, the LAST maximum: sampler.rs:109-116); CTA 0 only -----------------
__device__ void phase_argmax(const MkPhase& ph, const uint8_t* dyn, float* s_red) {
    if (blockIdx.x != 0) return;
    __shared__ long long s_idx[MK_WARPS];
    const float* x = ph.x;
    float bv = 0.0f; long long bi = -1;
    for (long long i = threadIdx.x; i < ph.n; i += MK_THREADS) { const float v = ldcg_f(x + i); if (bi < 0 || !(v < bv)) { bv = v; bi = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi > bi))) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = bv; s_idx[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < MK_WARPS; w++) { const float ov = s_red[w]; const long long oi = s_idx[w]; if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi > bi))) { bv = ov; bi = oi; } }
        if (bi < 0) bi = 0;
        *ph.slot_dev = bi;
        const long long h = *(const long long*)(dyn + ph.dyn_off);
        if (h >= 0 && h < CC_HISTORY_CAP) ph.hist_dev[h] = bi;
    }
    __syncthreads();
}

// ---- REDUCE / GATHER phases: second half of an exchange (the first half is epilogue 3 of the MATVEC phase + the handshake
// carried by its barrier).  REDUCE: dst = sum over ranks of the partial rows, rank order, (+ residual)   GATHER: dst = slices
__device__ void phase_reduce(const MkPhase& ph, const CommDev& comm, unsigned xseq, bool gather) {
    const float* base = comm.data[comm.rank] + (size_t)(xseq & 1u) * CC_COMM_MAX_RANKS * CC_COMM_MAX_ELEMS;
    const int n4 = ph.red_n >> 2;
    if (gather) {
        for (int i = blockIdx.x * MK_THREADS + threadIdx.x; i < n4 * comm.world; i += gridDim.x * MK_THREADS) {
            const int p = i / n4, j = i - p * n4;
            ((float4*)(ph.red_dst + (size_t)p * ph.red_n))[j] = __ldcg((const float4*)(base + (size_t)p * CC_COMM_MAX_ELEMS) + j);
        }
        return;
    }
    for (int i = blockIdx.x * MK_THREADS + threadIdx.x; i < n4; i += gridDim.x * MK_THREADS) {
        float4 a = __ldcg((const float4*)base + i);
        for (int p = 1; p < comm.world; p++) {
            const float4 b = __ldcg((const float4*)(base + (size_t)p * CC_COMM_MAX_ELEMS) + i);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        if (ph.red_res) { const float4 r = __ldcg((const float4*)ph.red_res + i); a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w; }
        ((float4*)ph.red_dst)[i] = a;
    }
}


#define MK_PROF_SLOTS 8      // developer profiling: u64 stamps per phase (CTA 0 / thread 0): 0 start, 1 activation ready, 2 rows done, 3 arrived, 4 x staged, 5 rms known
// look-ahead arguments of the next two MATVEC phases, fetched one word per thread at phase start

// flags: 1 look-ahead weight prefetch | 4 norm weights staged before the barrier | 8 every CTA polls the arrival counter |
//        32 per-warp early look-ahead | 64 x of the next fused prologue requested right after the barrier
#define MK_F_LOOK 1
#define MK_F_WSTAGE 4
#define MK_F_POLLCNT 8
#define MK_F_SYSFENCE 256      // exchange phases: a system-scope fence in EVERY CTA before its arrival (not needed, see grid_barrier_arrive;
                               // profiles/r02j: 2351 us vs 2265 us per token at N = 2)
#define MK_F_TESTSTALL 128     // test hook: the last CTA leaves before barrier 2 -> every other CTA must time out, not hang
#define MK_F_XEARLY 64         // the f32 row of the next fused prologue is requested (cp.async) right after the barrier opens
#define MK_F_EARLY 32          // a warp requests its first segments of the next MATVEC phase as soon as IT has finished its rows
#define MK_TYPE_CALL(T, CALL_Q8, CALL_Q4) do { if ((T) == CC_Q8_0) { CALL_Q8; } else { CALL_Q4; } } while (0)

// GEN: the phase table contains generic (K-quant) MATVEC phases.  The streaming-only instantiation carries none of their code, so
// its register allocation (the weight pipe lives in registers across phases) is not disturbed by them.
template <bool GEN>
__global__ void __launch_bounds__(MK_THREADS, MK_CTAS_PER_SM) mega_kernel(const MkPhase* __restrict__ phases, int n_phases, const uint8_t* dyn,
                                                                         unsigned* bar, const uint16_t* exp_lut, unsigned long long* prof, int flags, int wtop_off,
                                                                         unsigned* err_host, const CommDev comm) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ float s_red[MK_WARPS];
    __shared__ MkPhase s_phs[2];             // phase descriptors, double-buffered: p+1 is fetched while p runs
    __shared__ MkNext s_next[2];             // arguments of the next two MATVEC phases (look-ahead prefetch)
    __shared__ int s_abort;
    __shared__ __align__(8) unsigned long long s_abar[AT_NBUF];                // attention chunk buffers (TMA completion)
    unsigned apar = 0u;                      // per-buffer wait parity of the attention chunk pipeline
    const unsigned abar0 = (unsigned)__cvta_generic_to_shared(&s_abar[0]);
    if (threadIdx.x == 0) { for (int i = 0; i < AT_NBUF; i++) mbar_init(abar0 + 8u * i, 1u); s_abort = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    MkPipe pipe;                             // weight prefetch registers, live across phases and barriers
    uint8_t* work = smem;                    // per-phase working area (activation arrays, attention tiles)
    float* s_w = (float*)(smem + wtop_off);  // norm weights of the next fused prologue (top of dynamic shared memory)
    int prefetched = -1;                     // phase index whose first segments sit in the pipe
    int wstaged = -1;                        // phase index whose norm weights were requested into s_w
    int xstaged = -1;                        // phase index whose f32 input row was requested into its prologue's staging area
    unsigned gen = 0;                        // barriers completed; starts from the value left by the last launch
    if (threadIdx.x == MK_BAR_THREAD) gen = ld_acquire_u32(&bar[32]);
    unsigned xseq = comm.world > 0 ? *comm.seq : 0u;     // exchanges finished so far on this rank (comm.cu)
    {
        const int* src = (const int*)phases;
        int* dst = (int*)&s_phs[0];
        for (int i = threadIdx.x; i < (int)(sizeof(MkPhase) / 4); i += MK_THREADS) dst[i] = src[i];
    }
    for (int p = 0; p < n_phases; p++) {
        // developer profiling, 4 stamps per phase from CTA 0 / thread 0: start, activation ready (MATVEC), rows done, arrived + prefetch issued
        const bool stamp = prof && blockIdx.x == 0 && threadIdx.x == 0;
        if (stamp) { prof[p * MK_PROF_SLOTS] = globaltimer_ns(); prof[p * MK_PROF_SLOTS + 1] = 0; prof[p * MK_PROF_SLOTS + 4] = 0; prof[p * MK_PROF_SLOTS + 5] = 0; }
        __syncthreads();                     // descriptor p is in shared memory (stored one phase ago)
        const MkPhase& s_ph = s_phs[p & 1];
        // Descriptor p+1 and the arguments of the next MATVEC phases (look-ahead prefetch) are LOADED now, into one register each,
        // and STORED to shared memory after the phase body: a load followed directly by its st.shared would block the thread for
        // an L2 round trip (in-order issue) before it could issue the phase's own loads.
        static_assert(sizeof(MkPhase) / 4 <= MK_THREADS, "descriptor does not fit one word per thread");
        const int nx = s_ph.next_matvec, nx2 = s_ph.next_matvec2;
        const bool look = (flags & MK_F_LOOK) && nx > p && nx < n_phases && prefetched != nx && p + 1 < n_phases;     // (the next MATVEC may turn out generic: checked below)
        int desc_w = 0, next_w = 0;
        if (p + 1 < n_phases && threadIdx.x < sizeof(MkPhase) / 4) desc_w = ((const int*)(phases + p + 1))[threadIdx.x];
        if (look && threadIdx.x < 128) {
            const int slot = threadIdx.x >> 6, t = threadIdx.x & 63;
            const int q = slot == 0 ? nx : nx2;
            if (q > p && q < n_phases) {
                const MkPhase* ph = phases + q;
                constexpr int NW = (int)(sizeof(StreamArgs) / 4);
                if (t < NW) next_w = ((const int*)&ph->mv)[t];
                else if (t == NW) next_w = ph->wtype;
                else if (t == NW + 1) next_w = ph->x && ph->norm_w ? ph->n : 0;
                else if (t == NW + 2) next_w = ((const int*)&ph->norm_w)[0];
                else if (t == NW + 3) next_w = ((const int*)&ph->norm_w)[1];
            } else if (t == (int)(sizeof(StreamArgs) / 4)) next_w = -1;     // no such phase
        }
        bool early = false;                  // the MATVEC phase issued the look-ahead itself, warp by warp
        switch (s_ph.type) {
        case MK_NORMQ: phase_normq(s_ph, s_red); break;
        case MK_MATVEC:
            if (GEN && s_ph.act_type == CC_Q8_K) {   // K-quant weights: generic phase, no register look-ahead
                unsigned long long* st1 = stamp ? prof + p * MK_PROF_SLOTS + 1 : nullptr;
                switch (s_ph.wtype) {
                case CC_Q2_K: phase_matvec_generic<TQ2_K>(s_ph, work, s_w, wstaged == p, xstaged == p, exp_lut, pipe, st1); break;
                case CC_Q3_K: phase_matvec_generic<TQ3_K>(s_ph, work, s_w, wstaged == p, xstaged == p, exp_lut, pipe, st1); break;
                case CC_Q4_K: phase_matvec_generic<TQ45_K<false>>(s_ph, work, s_w, wstaged == p, xstaged == p, exp_lut, pipe, st1); break;
                case CC_Q5_K: phase_matvec_generic<TQ45_K<true>>(s_ph, work, s_w, wstaged == p, xstaged == p, exp_lut, pipe, st1); break;
                case CC_Q6_K: phase_matvec_generic<TQ6_K>(s_ph, work, s_w, wstaged == p, xstaged == p, exp_lut, pipe, st1); break;
                default: phase_matvec_generic<TQ8_K>(s_ph, work, s_w, wstaged == p, xstaged == p, exp_lut, pipe, st1); break;
                }
                prefetched = -1;              // the generic phase used the pipe's registers: a pending streaming look-ahead (mixed models) is gone
                break;
            }
            if (prefetched != p) MK_TYPE_CALL(s_ph.wtype, matvec_prefetch<CC_Q8_0>(s_ph.mv, pipe), matvec_prefetch<CC_Q4_0>(s_ph.mv, pipe));
            early = look && (flags & MK_F_EARLY);
            MK_TYPE_CALL(s_ph.wtype,
                         phase_matvec<CC_Q8_0>(s_ph, work, s_w, wstaged == p, xstaged == p, exp_lut, pipe, comm, xseq, stamp ? prof + p * MK_PROF_SLOTS + 1 : nullptr,
                                               early ? s_next : nullptr, next_w),
                         phase_matvec<CC_Q4_0>(s_ph, work, s_w, wstaged == p, xstaged == p, exp_lut, pipe, comm, xseq, stamp ? prof + p * MK_PROF_SLOTS + 1 : nullptr,
                                               early ? s_next : nullptr, next_w));
            break;
        case MK_ATTN:
            if (s_ph.at.kv_f16) phase_attn<true>(s_ph, (float*)work, s_red, dyn, exp_lut, abar0, apar); else phase_attn<false>(s_ph, (float*)work, s_red, dyn, exp_lut, abar0, apar);
            break;
        case MK_ROWS: phase_rows(s_ph, dyn); break;
        case MK_REDUCE: phase_reduce(s_ph, comm, xseq, false); break;
        case MK_GATHER: phase_reduce(s_ph, comm, xseq, true); break;
        case MK_ARGMAX: phase_argmax(s_ph, dyn, s_red); break;
        }
        if (stamp) prof[p * MK_PROF_SLOTS + 2] = globaltimer_ns();
        if (p + 1 < n_phases && threadIdx.x < sizeof(MkPhase) / 4) ((int*)&s_phs[(p + 1) & 1])[threadIdx.x] = desc_w;
        if (look && !early && threadIdx.x < 128) {
            const int slot = threadIdx.x >> 6, t = threadIdx.x & 63;
            if (t < (int)(sizeof(StreamArgs) / 4) + 4) ((int*)&s_next[slot])[t] = next_w;
        }
        // look-ahead: request the first two weight segments of the next MATVEC phase (and, into L2, the rows behind them and the first
        // rows of the phase after it) before waiting at the barrier, so HBM keeps streaming through the barrier, the prologue and any
        // small (NORMQ / ATTN / ROWS) phases in between
        const bool more = p + 1 < n_phases;
        const bool xg = s_ph.xgpu != 0;
        // test hook (tests/test_gpu_robustness.py): one CTA deserts before the third barrier, as if it had never become resident
        if ((flags & MK_F_TESTSTALL) && p == 2 && blockIdx.x == gridDim.x - 1) return;
        if (more) grid_barrier_arrive(bar, gridDim.x, gen, xg, (flags & MK_F_SYSFENCE) != 0);       // its bar.sync also publishes s_next (written just above)
        if (look) {
            const bool next_stream = s_next[0].wtype == CC_Q8_0 || s_next[0].wtype == CC_Q4_0;
            if (!early && next_stream) MK_TYPE_CALL(s_next[0].wtype, matvec_prefetch<CC_Q8_0>(s_next[0].mv, pipe), matvec_prefetch<CC_Q4_0>(s_next[0].mv, pipe));
            if (next_stream) prefetched = nx;
            if ((flags & MK_F_WSTAGE) && s_next[0].norm_n > 0) {     // immutable norm weights of the next fused prologue: one L2 trip less after the barrier
                const unsigned sw = (unsigned)__cvta_generic_to_shared(s_w);
                const float* nw = s_next[0].norm_w;
                for (int i = threadIdx.x; i < (s_next[0].norm_n >> 2); i += MK_THREADS)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sw + i * 16), "l"(nw + i * 4) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                wstaged = nx;
            }
        }
        if (stamp) prof[p * MK_PROF_SLOTS + 3] = globaltimer_ns();
        if (more) {
            grid_barrier_wait(bar, gridDim.x, gen, comm, xg ? xseq + 1u : 0u, (flags & MK_F_POLLCNT) != 0, &s_abort, err_host);
            gen++; if (xg) xseq++;
            if (s_abort) break;              // a barrier timed out (a CTA never became resident, or a peer GPU died): bail out, host reports
            // the barrier is open: the row the next fused prologue normalises is complete -- request it before anything else (descriptor
            // bookkeeping, geometry, look-ahead loads) so that its L2 round trip overlaps them
            const MkPhase& nph = s_phs[(p + 1) & 1];
            if ((flags & MK_F_XEARLY) && nph.type == MK_MATVEC && nph.x && !nph.red_n) {
                const int nb = nph.mv.k >> 5, nbp = ((((nb + 31) >> 5) + MK_SEG - 1) / MK_SEG) * MK_SEG * 32;
                const size_t xoff = nph.act_type == CC_Q8_K ? (size_t)mk_generic_sx_offset(nph.mv.k) : (size_t)nbp * 40 + 256 + 2048;
                const unsigned sx = (unsigned)__cvta_generic_to_shared(work + xoff);      // = s_x of the phase's prologue
                const float* xg = nph.x;
                for (int i = threadIdx.x; i < (nph.mv.k >> 2); i += MK_THREADS)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sx + i * 16), "l"(xg + i * 4) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                xstaged = p + 1;
            }
        }
    }
    if (comm.world > 0 && blockIdx.x == 0 && threadIdx.x == 0) *comm.seq = xseq;
    if (prof && blockIdx.x == 0 && threadIdx.x == 0) prof[n_phases * MK_PROF_SLOTS] = globaltimer_ns();
}

// working shared memory of one phase (the staging area of the norm weights comes on top, see cc_launch_mega)
size_t cc_mega_smem_for_phase(const MkPhase& ph) {
    if (ph.type == MK_MATVEC && ph.act_type == CC_Q8_K) return (size_t)(((TKBase::smem_bytes(ph.mv.k) + 15) & ~15) + 256) + (size_t)ph.mv.k * 4;
    if (ph.type == MK_MATVEC) {
        const size_t k = (size_t)ph.mv.k, nb = k / 32, GR = (nb + 31) / 32, NSEG = (GR + MK_SEG - 1) / MK_SEG, nbp = NSEG * MK_SEG * 32;
        // quants | scales | block sums | prologue: reduction scratch, f32 x
        return nbp * 40 + 256 + 2048 + (ph.x ? k * 4 : 0);
    }
    if (ph.type == MK_ATTN) return (size_t)(3 * ph.at.hd + ((ph.at.max_len + 8 + 3) & ~3) + AT_NBUF * AT_CH * ph.at.hd) * 4 + 64;
    return 1024;
}

// developer hook: a table of `n` empty phases -> the pure per-phase floor (descriptor fetch + grid barrier)
extern "C" CC_API int cc_test_mega_barrier_floor(cc_device* dev, int n, float* us_per_phase) {
    if (!dev || n < 2 || !us_per_phase) return CC_ERR_ARG;
    std::vector<MkPhase> tab((size_t)n);
    for (auto& p : tab) { memset(&p, 0, sizeof(p)); p.type = 99; p.next_matvec = -1; p.next_matvec2 = -1; }
    MkPhase* d_tab = nullptr; unsigned* d_bar = nullptr;
    CC_CUDA(dev, cudaMalloc(&d_tab, tab.size() * sizeof(MkPhase)));
    CC_CUDA(dev, cudaMalloc(&d_bar, 4096));
    CC_CUDA(dev, cudaMemcpy(d_tab, tab.data(), tab.size() * sizeof(MkPhase), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CC_CUDA(dev, cudaMemsetAsync(d_bar, 0, 4096, dev->stream));
        cudaEventRecord(e0, dev->stream);
        int rc = cc_launch_mega(dev, d_tab, n, nullptr, d_bar, 1024, 0, nullptr, nullptr, false);
        if (rc) return rc;
        cudaEventRecord(e1, dev->stream);
        CC_CUDA(dev, cudaEventSynchronize(e1));
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    *us_per_phase = best * 1e3f / (float)n;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_tab); cudaFree(d_bar);
    return CC_OK;
}

bool cc_mega_generic_supported(int type, int64_t k) {
    return (type == CC_Q2_K || type == CC_Q3_K || type == CC_Q4_K || type == CC_Q5_K || type == CC_Q6_K || type == CC_Q8_K) && k % 256 == 0 && k <= 32768;
}

// developer A/B switches: CRABML_MEGA_FLAGS replaces the default flag word (see MK_F_* and the L2 budget byte)
#define MK_DEFAULT_FLAGS (MK_F_LOOK | MK_F_WSTAGE | MK_F_POLLCNT | MK_F_XEARLY)      // profiles/r02c: 2407 us vs 2586 us (0x5) on the same box
int cc_mega_flags() {
    static const int f = getenv("CRABML_MEGA_FLAGS") ? (int)strtol(getenv("CRABML_MEGA_FLAGS"), nullptr, 0) : MK_DEFAULT_FLAGS;
    return f;
}

int cc_launch_mega(cc_device* dev, const MkPhase* phases_dev, int n_phases, const uint8_t* dyn_dev, unsigned* bar_dev, size_t smem_work, size_t smem_wstage,
                   unsigned long long* prof, const CommDev* comm, bool generic) {
    const int flags = cc_mega_flags();
    int max_ctas_per_sm = 0;
    const size_t wtop = (smem_work + 15) & ~(size_t)15;
    const size_t smem = wtop + smem_wstage;
    CC_REQUIRE(dev, smem <= 227 * 1024, "megakernel: a phase needs %zu bytes of shared memory", smem);
    auto kern = generic ? mega_kernel<true> : mega_kernel<false>;
    if (smem > 48 * 1024) CC_CUDA(dev, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CC_CUDA(dev, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_ctas_per_sm, kern, MK_THREADS, smem));
    CC_REQUIRE(dev, max_ctas_per_sm >= 1, "megakernel does not fit on an SM");
    int per_sm = max_ctas_per_sm < MK_CTAS_PER_SM ? max_ctas_per_sm : MK_CTAS_PER_SM;
    int grid = dev->sm_count * per_sm;          // all CTAs co-resident: required by the grid barrier
    CommDev cd;
    memset(&cd, 0, sizeof(cd));
    if (comm) cd = *comm;
    // The grid barrier needs every CTA resident at once.  On a GPU this process owns, a plain launch of sm_count CTAs (1 per SM)
    // is co-resident by construction.  With another tenant on the same GPU (a second process, MPS) a partially scheduled grid
    // cannot finish a barrier: every spin in the kernel is bounded (MkSpin) and ends in CC_ERR_CUDA "megakernel barrier timeout"
    // instead of a hang; CRABML_MEGA_COOP=1 adds the cooperative launch attribute (all-or-nothing placement).  That is opt-in
    // because a cooperative kernel node costs ~1.3 ms per graph launch on this driver (274 vs 416 tok/s, same call).
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(MK_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = dev->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = getenv("CRABML_MEGA_COOP") ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const uint16_t* lut = dev->exp_lut;
    CC_CUDA(dev, cudaLaunchKernelEx(&cfg, kern, phases_dev, n_phases, dyn_dev, bar_dev, lut, prof, flags, (int)wtop, dev->err_host, (const CommDev)cd));
    CC_LAUNCH_CHECK(dev);
    return CC_OK;
}
