// vecdot.cuh -- per-row dot products of every GGUF weight type against the on-the-fly quantised activation (SURVEY 8a rows a6-a11;
// reference: buf/buf_q*.rs vec_dot_*).  One warp per output row: `T::row_dot` returns this lane's partial, the caller finishes with
// the xor-butterfly warp_sum.  Shared by the eager kernel (matvec.cu) and the generic MATVEC phase of the megakernel (mega.cu), so
// both produce the same bits.  T::stage_act copies the activation scratch (quantize.cu layout) into shared memory.
#pragma once
#include "common.cuh"


__device__ __forceinline__ int dp4(int a, int b, int c) { return __dp4a(a, b, c); }
__device__ __forceinline__ int dot16(const int4& w, const int4& a) {
    return dp4(w.x, a.x, dp4(w.y, a.y, dp4(w.z, a.z, dp4(w.w, a.w, 0))));
}
__device__ __forceinline__ int4 and4(const int4& v, int m) { return make_int4(v.x & m, v.y & m, v.z & m, v.w & m); }
__device__ __forceinline__ int4 shr4(const int4& v, int s, int m) {
    return make_int4((int)(((unsigned)v.x >> s) & m), (int)(((unsigned)v.y >> s) & m),
                     (int)(((unsigned)v.z >> s) & m), (int)(((unsigned)v.w >> s) & m));
}
__device__ __forceinline__ int4 or4(const int4& a, const int4& b) { return make_int4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w); }
__device__ __forceinline__ int4 shl4(const int4& v, int s) { return make_int4(v.x << s, v.y << s, v.z << s, v.w << s); }
// spread 4 bits to bit 4 of 4 bytes
__device__ __forceinline__ int spread4(unsigned bits) { return (int)((((bits & 0xF) * 0x00204081u) & 0x01010101u) << 4); }
__device__ __forceinline__ int byte_of(unsigned w0, unsigned w1, int idx) { return (int)(((idx < 4 ? w0 : w1) >> (8 * (idx & 3))) & 0xFF); }

// cooperative global -> shared copy of `bytes` (multiple of 4) bytes
__device__ __forceinline__ void stage(void* dst, const void* src, int bytes) {
    if (((uintptr_t)src & 15) != 0) {          // batched activations: row bi may start 4-byte aligned only
        for (int i = threadIdx.x; i < (bytes >> 2); i += blockDim.x) ((int*)dst)[i] = ((const int*)src)[i];
        return;
    }
    int n16 = bytes >> 4;
    const int4* s = (const int4*)src;
    int4* d = (int4*)dst;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) d[i] = s[i];
    int rem = (bytes & 15) >> 2;
    if ((int)threadIdx.x < rem) ((int*)dst)[n16 * 4 + threadIdx.x] = ((const int*)src)[n16 * 4 + threadIdx.x];
}
__host__ __device__ inline int al16i(int v) { return (v + 15) & ~15; }

// ------------------------------------------------------------------------------------------------
struct WPlanes { const uint8_t* p[CC_MAX_PLANES]; };

// Each Traits provides:
//   smem_bytes(k); stage_act(scratch, bi, k, smem); row_dot(W, row, k, smem, lane) -> per-lane partial
// ------------------------------------------------------------------------------------------------

// ---- Q8_0 x Q8_0 : buf_q8_0.rs:136-286 -------------------------------------------------------------
struct TQ8_0 {
    static __host__ __device__ int smem_bytes(int k) { return al16i(k) + al16i(k / 32 * 4); }
    static __device__ void stage_act(const void* scratch, int64_t n_total, int bi, int k, uint8_t* sm) {
        const uint8_t* p = (const uint8_t*)scratch;
        const int8_t* qs = (const int8_t*)p + (int64_t)bi * k;
        const float* d = (const float*)(p + ((n_total + 15) & ~15ll)) + (int64_t)bi * (k / 32);
        stage(sm, qs, k);
        stage(sm + al16i(k), d, k / 32 * 4);
    }
    static __device__ float row_dot(const WPlanes& W, int64_t row, int k, const uint8_t* sm, int lane) {
        // device layout: groups of 32 blocks, first halves then second halves (matvec_stream.cu)
        const int nb = k >> 5, nchunk = k >> 4;
        const int4* wq = (const int4*)W.p[0] + row * nchunk;
        const uint16_t* wd = (const uint16_t*)W.p[1] + row * CC_D_STRIDE(nb);
        const int4* aq = (const int4*)sm;
        const float* ad = (const float*)(sm + al16i(k));
        float acc = 0.0f;
        for (int c = lane; c < nchunk; c += 32) {
            const int g = c >> 6, idx = c & 63;
            const int nbg = min(32, nb - 32 * g);
            const int half = idx >= nbg ? 1 : 0;
            const int blk = 32 * g + idx - half * nbg;
            int4 w0 = ld_stream_16(wq + c);
            acc += (float)dot16(w0, aq[2 * blk + half]) * h2f_bits(wd[blk]) * ad[blk];
        }
        return acc;
    }
};

// ---- Q4_0 x Q8_0 : buf_q4_0.rs:126-253 -------------------------------------------------------------
// shared: qs [k] | d [k/32] f32 | isum [k/32] i32
struct TQ4_0 {
    static __host__ __device__ int smem_bytes(int k) { return al16i(k) + 2 * al16i(k / 32 * 4); }
    static __device__ void stage_act(const void* scratch, int64_t n_total, int bi, int k, uint8_t* sm) {
        const uint8_t* p = (const uint8_t*)scratch;
        int64_t o1 = (n_total + 15) & ~15ll, o2 = o1 + ((n_total / 32 * 4 + 15) & ~15ll);
        stage(sm, p + (int64_t)bi * k, k);
        stage(sm + al16i(k), p + o1 + (int64_t)bi * (k / 32) * 4, k / 32 * 4);
        stage(sm + al16i(k) + al16i(k / 32 * 4), p + o2 + (int64_t)bi * (k / 32) * 4, k / 32 * 4);
    }
    static __device__ __forceinline__ float block(const int4& w, const int4* aq, const float* ad, const int* as, int b, uint16_t hd) {
        int4 lo = and4(w, 0x0F0F0F0F), hi = shr4(w, 4, 0x0F0F0F0F);
        int s = dot16(lo, aq[2 * b]) + dot16(hi, aq[2 * b + 1]) - 8 * as[b];     // sum (q-8)*a
        return (float)s * h2f_bits(hd) * ad[b];
    }
    static __device__ float row_dot(const WPlanes& W, int64_t row, int k, const uint8_t* sm, int lane) {
        const int nb = k >> 5;
        const int4* wq = (const int4*)W.p[0] + row * nb;
        const uint16_t* wd = (const uint16_t*)W.p[1] + row * CC_D_STRIDE(nb);
        const int4* aq = (const int4*)sm;
        const float* ad = (const float*)(sm + al16i(k));
        const int* as = (const int*)(sm + al16i(k) + al16i(nb * 4));
        float acc = 0.0f;
        int b = lane;
        for (; b + 96 < nb; b += 128) {
            int4 w0 = ld_stream_16(wq + b), w1 = ld_stream_16(wq + b + 32), w2 = ld_stream_16(wq + b + 64), w3 = ld_stream_16(wq + b + 96);
            uint16_t h0 = wd[b], h1 = wd[b + 32], h2 = wd[b + 64], h3 = wd[b + 96];
            acc += block(w0, aq, ad, as, b, h0);
            acc += block(w1, aq, ad, as, b + 32, h1);
            acc += block(w2, aq, ad, as, b + 64, h2);
            acc += block(w3, aq, ad, as, b + 96, h3);
        }
        for (; b < nb; b += 32) acc += block(ld_stream_16(wq + b), aq, ad, as, b, wd[b]);
        return acc;
    }
};

// ---- Q5_0 x Q8_0 : buf_q5_0.rs:145-163 --------------------------------------------------------------
struct TQ5_0 : TQ4_0 {
    static __device__ float row_dot(const WPlanes& W, int64_t row, int k, const uint8_t* sm, int lane) {
        const int nb = k >> 5;
        const int4* wq = (const int4*)W.p[0] + row * nb;
        const uint16_t* wd = (const uint16_t*)W.p[1] + row * nb;
        const uint32_t* wh = (const uint32_t*)W.p[2] + row * nb;
        const int4* aq = (const int4*)sm;
        const float* ad = (const float*)(sm + al16i(k));
        const int* as = (const int*)(sm + al16i(k) + al16i(nb * 4));
        float acc = 0.0f;
        for (int b = lane; b < nb; b += 32) {
            int4 w = ld_stream_16(wq + b);
            uint32_t qh = wh[b];
            int4 lo = and4(w, 0x0F0F0F0F), hi = shr4(w, 4, 0x0F0F0F0F);
            lo = or4(lo, make_int4(spread4(qh), spread4(qh >> 4), spread4(qh >> 8), spread4(qh >> 12)));
            hi = or4(hi, make_int4(spread4(qh >> 16), spread4(qh >> 20), spread4(qh >> 24), spread4(qh >> 28)));
            int s = dot16(lo, aq[2 * b]) + dot16(hi, aq[2 * b + 1]) - 16 * as[b];
            acc += (float)s * h2f_bits(wd[b]) * ad[b];
        }
        return acc;
    }
};

// ---- Q4_1 / Q5_1 x Q8_1 : buf_q4_1.rs:266-280 (scalar path, B11), buf_q5_1.rs:142-161 ------------------
// shared: qs [k] | (d,s) half2 [k/32]
template <bool FIVE>
struct TQx_1 {
    static __host__ __device__ int smem_bytes(int k) { return al16i(k) + al16i(k / 32 * 4); }
    static __device__ void stage_act(const void* scratch, int64_t n_total, int bi, int k, uint8_t* sm) {
        const uint8_t* p = (const uint8_t*)scratch;
        int64_t o1 = (n_total + 15) & ~15ll;
        stage(sm, p + (int64_t)bi * k, k);
        stage(sm + al16i(k), p + o1 + (int64_t)bi * (k / 32) * 4, k / 32 * 4);
    }
    static __device__ float row_dot(const WPlanes& W, int64_t row, int k, const uint8_t* sm, int lane) {
        const int nb = k >> 5;
        const int4* wq = (const int4*)W.p[0] + row * nb;
        const uint32_t* wdm = (const uint32_t*)W.p[1] + row * nb;     // (d, m) f16 pair
        const uint32_t* wh = FIVE ? (const uint32_t*)W.p[2] + row * nb : nullptr;
        const int4* aq = (const int4*)sm;
        const uint32_t* ads = (const uint32_t*)(sm + al16i(k));     // (d, s) f16 pair
        float acc = 0.0f;
        for (int b = lane; b < nb; b += 32) {
            int4 w = ld_stream_16(wq + b);
            int4 lo = and4(w, 0x0F0F0F0F), hi = shr4(w, 4, 0x0F0F0F0F);
            if (FIVE) {
                uint32_t qh = wh[b];
                lo = or4(lo, make_int4(spread4(qh), spread4(qh >> 4), spread4(qh >> 8), spread4(qh >> 12)));
                hi = or4(hi, make_int4(spread4(qh >> 16), spread4(qh >> 20), spread4(qh >> 24), spread4(qh >> 28)));
            }
            int s = dot16(lo, aq[2 * b]) + dot16(hi, aq[2 * b + 1]);
            uint32_t dm = wdm[b], ds = ads[b];
            // half::f16 `a * b` = f16(f32(a) * f32(b))
            float dd = h2f_bits(f2h_bits(h2f_bits(dm & 0xFFFF) * h2f_bits(ds & 0xFFFF)));
            float ms = h2f_bits(f2h_bits(h2f_bits(dm >> 16) * h2f_bits(ds >> 16)));
            if (FIVE) acc += (float)s * dd + ms;      // buf_q5_1.rs:157
            else      acc += dd * (float)s + ms;      // buf_q4_1.rs:276
        }
        return acc;
    }
};

// ---- K-quants x Q8_K -----------------------------------------------------------------------------------
// shared: qs [k] | d [k/256] f32 | bsums [k/16] i16
struct KAct {
    const int4* q; const float* d; const int16_t* bs;
    __device__ KAct(const uint8_t* sm, int k) : q((const int4*)sm), d((const float*)(sm + al16i(k))), bs((const int16_t*)(sm + al16i(k) + al16i(k / 256 * 4))) {}
};
struct TKBase {
    static constexpr bool kSegmented = false;      // overridden by the types that offer seg_load / seg_dot (Q4_K, Q6_K)
    static __device__ __forceinline__ void seg_load(KSeg&, int (&)[4], const WPlanes&, int64_t, int, int, int) {}
    static __device__ __forceinline__ float seg_dot(const KSeg&, const int (&)[4], int, int, const uint8_t*, int, float acc) { return acc; }
    static __host__ __device__ int smem_bytes(int k) { return al16i(k) + al16i(k / 256 * 4) + al16i(k / 16 * 2); }
    static __device__ void stage_act(const void* scratch, int64_t n_total, int bi, int k, uint8_t* sm) {
        const uint8_t* p = (const uint8_t*)scratch;
        int64_t o1 = (n_total + 15) & ~15ll, o2 = o1 + ((n_total / 256 * 4 + 15) & ~15ll);
        stage(sm, p + (int64_t)bi * k, k);
        stage(sm + al16i(k), p + o1 + (int64_t)bi * (k / 256) * 4, k / 256 * 4);
        stage(sm + al16i(k) + al16i(k / 256 * 4), p + o2 + (int64_t)bi * (k / 16) * 2, k / 16 * 2);
    }
};

// unpack of the 12 scale bytes into 8 scales (s0,s1) and 8 mins (m0,m1): buf_q4_k.rs:219-234
__device__ __forceinline__ void k4_unpack(unsigned u0, unsigned u1, unsigned u2, unsigned& s0, unsigned& s1, unsigned& m0, unsigned& m1) {
    const unsigned K1 = 0x3f3f3f3fu, K2 = 0x0f0f0f0fu, K3 = 0x03030303u;
    m1 = ((u2 >> 4) & K2) | (((u1 >> 6) & K3) << 4);
    m0 = u1 & K1;
    s1 = (u2 & K2) | (((u0 >> 6) & K3) << 4);
    s0 = u0 & K1;
}

// Q4_K: buf_q4_k.rs:192-277.  8 lanes per super-block, lane j owns the 16-byte qs chunk j.
template <bool FIVE>
struct TQ45_K : TKBase {
    static __device__ float row_dot(const WPlanes& W, int64_t row, int k, const uint8_t* sm, int lane) {
        const int nsb = k >> 8, BB = FIVE ? 176 : 144, QOFF = FIVE ? 48 : 16;
        const uint8_t* wrow = W.p[0] + row * (int64_t)nsb * BB;
        KAct A(sm, k);
        const int j = lane & 7, g = j >> 1, l0 = (j & 1) * 16;
        float acc = 0.0f;
        for (int sb = lane >> 3; sb < nsb; sb += 4) {
            const uint8_t* blk = wrow + (int64_t)sb * BB;
            int4 hdr = ld_stream_16(blk);
            int4 q = ld_stream_16(blk + QOFF + 16 * j);
            unsigned s0, s1, m0, m1;
            k4_unpack((unsigned)hdr.y, (unsigned)hdr.z, (unsigned)hdr.w, s0, s1, m0, m1);
            int4 lo = and4(q, 0x0F0F0F0F), hi = shr4(q, 4, 0x0F0F0F0F);
            if (FIVE) {                                                  // buf_q5_k.rs:240-258: bit 2g / 2g+1 of qh[l]
                int4 qh = ld_stream_16(blk + 16 + l0);
                lo = or4(lo, shl4(shr4(qh, 2 * g, 0x01010101), 4));
                hi = or4(hi, shl4(shr4(qh, 2 * g + 1, 0x01010101), 4));
            }
            const int abase = sb * 16 + g * 4 + (j & 1);               // int4 index of act elems 256 sb + 64 g + l0
            int sum_lo = dot16(lo, A.q[abase]), sum_hi = dot16(hi, A.q[abase + 2]);
            int isum = byte_of(s0, s1, 2 * g) * sum_lo + byte_of(s0, s1, 2 * g + 1) * sum_hi;
            // mins: lane j covers min index j  ->  (bsums[2j] + bsums[2j+1]) * mins[j]   (i32: B7 fixed)
            int msum = ((int)A.bs[sb * 16 + 2 * j] + (int)A.bs[sb * 16 + 2 * j + 1]) * byte_of(m0, m1, j);
            float da = A.d[sb];
            float d = h2f_bits((uint16_t)(hdr.x & 0xFFFF)) * da, dmin = h2f_bits((uint16_t)((unsigned)hdr.x >> 16)) * da;
            acc += d * (float)isum - dmin * (float)msum;
        }
        return acc;
    }
    // ---- the same dot cut into SEGMENTS of 16 super-blocks (4 iterations of the loop above) whose 8 16-byte loads are issued as one
    // batch into a KSeg, so a caller can keep the next segment in flight while it computes the current one (megakernel generic
    // phase).  Per lane the terms are added in the same order as in row_dot -> identical bits.  x: 4 extra words per segment (Q6_K).
    static constexpr bool kSegmented = !FIVE;          // Q5_K would need 4 more 16-byte registers per segment (qh): plain loop
    static __device__ __forceinline__ void seg_load(KSeg& S, int (&x)[4], const WPlanes& W, int64_t row, int k, int seg, int lane) {
        const int nsb = k >> 8, BB = 144, QOFF = 16;
        const uint8_t* wrow = W.p[0] + row * (int64_t)nsb * BB;
        const int j = lane & 7;
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int sb = seg * 16 + it * 4 + (lane >> 3);
            if (sb < nsb) {
                const uint8_t* blk = wrow + (int64_t)sb * BB;
                S.a[it] = ld_stream_16(blk);
                S.b[it] = ld_stream_16(blk + QOFF + 16 * j);
            }
        }
    }
    static __device__ __forceinline__ float seg_dot(const KSeg& S, const int (&x)[4], int k, int seg, const uint8_t* sm, int lane, float acc) {
        const int nsb = k >> 8;
        KAct A(sm, k);
        const int j = lane & 7, g = j >> 1;
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int sb = seg * 16 + it * 4 + (lane >> 3);
            if (sb < nsb) {
                const int4 hdr = S.a[it], q = S.b[it];
                unsigned s0, s1, m0, m1;
                k4_unpack((unsigned)hdr.y, (unsigned)hdr.z, (unsigned)hdr.w, s0, s1, m0, m1);
                int4 lo = and4(q, 0x0F0F0F0F), hi = shr4(q, 4, 0x0F0F0F0F);
                const int abase = sb * 16 + g * 4 + (j & 1);
                int sum_lo = dot16(lo, A.q[abase]), sum_hi = dot16(hi, A.q[abase + 2]);
                int isum = byte_of(s0, s1, 2 * g) * sum_lo + byte_of(s0, s1, 2 * g + 1) * sum_hi;
                int msum = ((int)A.bs[sb * 16 + 2 * j] + (int)A.bs[sb * 16 + 2 * j + 1]) * byte_of(m0, m1, j);
                float da = A.d[sb];
                float d = h2f_bits((uint16_t)(hdr.x & 0xFFFF)) * da, dmin = h2f_bits((uint16_t)((unsigned)hdr.x >> 16)) * da;
                acc += d * (float)isum - dmin * (float)msum;
            }
        }
        return acc;
    }
};

// Q6_K: buf_q6_k.rs:183-235.  8 lanes per super-block, lane j owns ql chunk j.
struct TQ6_K : TKBase {
    static __device__ float row_dot(const WPlanes& W, int64_t row, int k, const uint8_t* sm, int lane) {
        const int nsb = k >> 8;
        const uint8_t* ql = W.p[0] + row * (int64_t)nsb * 128;
        const uint8_t* qh = W.p[1] + row * (int64_t)nsb * 64;
        const int8_t* sc = (const int8_t*)W.p[2] + row * (int64_t)nsb * 16;
        const uint16_t* wd = (const uint16_t*)W.p[3] + row * (int64_t)nsb;
        KAct A(sm, k);
        const int j = lane & 7, n = j >> 2, jj = j & 3, r = jj & 1, p = jj >> 1;
        float acc = 0.0f;
        for (int sb = lane >> 3; sb < nsb; sb += 4) {
            int4 l = ld_stream_16(ql + (int64_t)sb * 128 + 16 * j);
            int4 h = ld_stream_16(qh + (int64_t)sb * 64 + 32 * n + 16 * r);
            int4 lo = or4(and4(l, 0x0F0F0F0F), shl4(shr4(h, 2 * p, 0x03030303), 4));
            int4 hi = or4(shr4(l, 4, 0x0F0F0F0F), shl4(shr4(h, 2 * p + 4, 0x03030303), 4));
            const int e_lo = 128 * n + 32 * p + 16 * r;                 // element offset inside the super-block
            const int a_lo = sb * 16 + (e_lo >> 4), a_hi = a_lo + 4;
            int s_lo = dot16(lo, A.q[a_lo]) - 32 * (int)A.bs[a_lo];     // sum (q-32)*a
            int s_hi = dot16(hi, A.q[a_hi]) - 32 * (int)A.bs[a_hi];
            int isum = (int)sc[sb * 16 + (e_lo >> 4)] * s_lo + (int)sc[sb * 16 + (e_lo >> 4) + 4] * s_hi;
            acc += (float)isum * (h2f_bits(wd[sb]) * A.d[sb]);
        }
        return acc;
    }
    // segments of 16 super-blocks, loads batched (see TQ45_K): ql chunk, qh chunk, the two sub-block scales and d per iteration
    static constexpr bool kSegmented = true;
    static __device__ __forceinline__ void seg_load(KSeg& S, int (&x)[4], const WPlanes& W, int64_t row, int k, int seg, int lane) {
        const int nsb = k >> 8;
        const uint8_t* ql = W.p[0] + row * (int64_t)nsb * 128;
        const uint8_t* qh = W.p[1] + row * (int64_t)nsb * 64;
        const int8_t* sc = (const int8_t*)W.p[2] + row * (int64_t)nsb * 16;
        const uint16_t* wd = (const uint16_t*)W.p[3] + row * (int64_t)nsb;
        const int j = lane & 7, n = j >> 2, jj = j & 3, r = jj & 1, p = jj >> 1;
        const int e_lo = 128 * n + 32 * p + 16 * r;
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int sb = seg * 16 + it * 4 + (lane >> 3);
            if (sb < nsb) {
                S.a[it] = ld_stream_16(ql + (int64_t)sb * 128 + 16 * j);
                S.b[it] = ld_stream_16(qh + (int64_t)sb * 64 + 32 * n + 16 * r);
                x[it] = ((int)sc[sb * 16 + (e_lo >> 4)] & 0xFFFF) | ((int)sc[sb * 16 + (e_lo >> 4) + 4] << 16);     // the two sub-block scales
                S.s[it] = wd[sb];
            }
        }
    }
    static __device__ __forceinline__ float seg_dot(const KSeg& S, const int (&x)[4], int k, int seg, const uint8_t* sm, int lane, float acc) {
        const int nsb = k >> 8;
        KAct A(sm, k);
        const int j = lane & 7, n = j >> 2, jj = j & 3, r = jj & 1, p = jj >> 1;
        const int e_lo = 128 * n + 32 * p + 16 * r;
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int sb = seg * 16 + it * 4 + (lane >> 3);
            if (sb < nsb) {
                const int4 l = S.a[it], h = S.b[it];
                int4 lo = or4(and4(l, 0x0F0F0F0F), shl4(shr4(h, 2 * p, 0x03030303), 4));
                int4 hi = or4(shr4(l, 4, 0x0F0F0F0F), shl4(shr4(h, 2 * p + 4, 0x03030303), 4));
                const int a_lo = sb * 16 + (e_lo >> 4), a_hi = a_lo + 4;
                int s_lo = dot16(lo, A.q[a_lo]) - 32 * (int)A.bs[a_lo];
                int s_hi = dot16(hi, A.q[a_hi]) - 32 * (int)A.bs[a_hi];
                int isum = (int)(int16_t)(x[it] & 0xFFFF) * s_lo + (x[it] >> 16) * s_hi;
                acc += (float)isum * (h2f_bits(S.s[it]) * A.d[sb]);
            }
        }
        return acc;
    }
};

// Q2_K: buf_q2_k.rs:214-257.  4 lanes per super-block, lane j owns qs chunk j (16 B), all four 2-bit planes.
struct TQ2_K : TKBase {
    static __device__ float row_dot(const WPlanes& W, int64_t row, int k, const uint8_t* sm, int lane) {
        const int nsb = k >> 8;
        const uint8_t* qs = W.p[0] + row * (int64_t)nsb * 64;
        const uint8_t* sc = W.p[1] + row * (int64_t)nsb * 16;
        const uint32_t* dd = (const uint32_t*)W.p[2] + row * (int64_t)nsb;
        KAct A(sm, k);
        const int j = lane & 3, n = j >> 1, r = j & 1;
        float acc = 0.0f;
        for (int sb = lane >> 2; sb < nsb; sb += 8) {
            int4 q = ld_stream_16(qs + (int64_t)sb * 64 + 16 * j);
            int isum = 0, msum = 0;
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const int is = 8 * n + 2 * s + r;                       // 16-element group index
                int4 v = shr4(q, 2 * s, 0x03030303);
                int scb = sc[sb * 16 + is];
                isum += (scb & 0xF) * dot16(v, A.q[sb * 16 + is]);
                msum += (int)A.bs[sb * 16 + is] * (scb >> 4);
            }
            uint32_t d2 = dd[sb];
            float da = A.d[sb];
            acc += (da * h2f_bits(d2 & 0xFFFF)) * (float)isum - (da * h2f_bits(d2 >> 16)) * (float)msum;
        }
        return acc;
    }
};

// Q3_K: buf_q3_k.rs:240-328.  4 lanes per super-block.
struct TQ3_K : TKBase {
    static __device__ float row_dot(const WPlanes& W, int64_t row, int k, const uint8_t* sm, int lane) {
        const int nsb = k >> 8;
        const uint8_t* qs = W.p[0] + row * (int64_t)nsb * 64;
        const uint8_t* hm = W.p[1] + row * (int64_t)nsb * 32;
        const uint32_t* s12 = (const uint32_t*)(W.p[2] + row * (int64_t)nsb * 12);
        const uint16_t* wd = (const uint16_t*)W.p[3] + row * (int64_t)nsb;
        KAct A(sm, k);
        const int j = lane & 3, n = j >> 1, r = j & 1;
        float acc = 0.0f;
        for (int sb = lane >> 2; sb < nsb; sb += 8) {
            int4 q = ld_stream_16(qs + (int64_t)sb * 64 + 16 * j);
            int4 h = ld_stream_16(hm + (int64_t)sb * 32 + 16 * r);
            // 16 six-bit scales (buf_q3_k.rs:286-296)
            unsigned a0 = s12[sb * 3], a1 = s12[sb * 3 + 1], tmp = s12[sb * 3 + 2];
            const unsigned K1 = 0x03030303u, K2 = 0x0f0f0f0fu;
            unsigned x0 = (a0 & K2) | ((tmp & K1) << 4), x1 = (a1 & K2) | (((tmp >> 2) & K1) << 4);
            unsigned x2 = ((a0 >> 4) & K2) | (((tmp >> 4) & K1) << 4), x3 = ((a1 >> 4) & K2) | (((tmp >> 6) & K1) << 4);
            int isum = 0;
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const int is = 8 * n + 2 * s + r;
                // value = q2 - (hbit ? 0 : 4) = (q2 | hbit<<2) - 4
                int4 v = or4(shr4(q, 2 * s, 0x03030303), shl4(shr4(h, 4 * n + s, 0x01010101), 2));
                int dot = dot16(v, A.q[sb * 16 + is]) - 4 * (int)A.bs[sb * 16 + is];
                unsigned word = is < 4 ? x0 : is < 8 ? x1 : is < 12 ? x2 : x3;
                int scale = (int)(int8_t)((word >> (8 * (is & 3))) & 0xFF) - 32;
                isum += scale * dot;
            }
            acc += (h2f_bits(wd[sb]) * A.d[sb]) * (float)isum;
        }
        return acc;
    }
};

// Q8_K x Q8_K: buf_q8_k.rs:133-224
struct TQ8_K : TKBase {
    static __device__ float row_dot(const WPlanes& W, int64_t row, int k, const uint8_t* sm, int lane) {
        const int nchunk = k >> 4;
        const int4* wq = (const int4*)W.p[0] + row * nchunk;
        const float* wd = (const float*)W.p[1] + row * (k >> 8);
        KAct A(sm, k);
        float acc = 0.0f;
        for (int c = lane; c < nchunk; c += 32) acc += (float)dot16(ld_stream_16(wq + c), A.q[c]) * wd[c >> 4] * A.d[c >> 4];
        return acc;
    }
};

// ---- F32 / F16 weights (fixtures only): buf_f32.rs:19-27, buf_f16.rs:84-97 -------------------------------
struct TF32 {
    static __host__ __device__ int smem_bytes(int k) { return al16i(k * 4); }
    static __device__ void stage_act(const void* x, int64_t, int bi, int k, uint8_t* sm) { stage(sm, (const float*)x + (int64_t)bi * k, k * 4); }
    static __device__ float row_dot(const WPlanes& W, int64_t row, int k, const uint8_t* sm, int lane) {
        const float* w = (const float*)W.p[0] + row * (int64_t)k;
        const float* a = (const float*)sm;
        float acc = 0.0f;
        for (int i = lane; i < k; i += 32) acc += w[i] * a[i];
        return acc;
    }
};
struct TF16 {
    static __host__ __device__ int smem_bytes(int k) { return al16i(k * 2); }
    static __device__ void stage_act(const void* x, int64_t, int bi, int k, uint8_t* sm) {
        const __half* src = (const __half*)x + (int64_t)bi * k;
        __half* d = (__half*)sm;
        for (int i = threadIdx.x; i < k; i += blockDim.x) d[i] = src[i];
    }
    static __device__ float row_dot(const WPlanes& W, int64_t row, int k, const uint8_t* sm, int lane) {
        const __half* w = (const __half*)W.p[0] + row * (int64_t)k;
        const __half* a = (const __half*)sm;
        float acc = 0.0f;
        for (int i = lane; i < k; i += 32) acc += __half2float(w[i]) * __half2float(a[i]);
        return acc;
    }
};

