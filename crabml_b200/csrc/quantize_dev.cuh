// quantize_dev.cuh -- warp-level activation quantisers shared by the eager kernels (quantize.cu) and the fused prologues of the
// megakernel (mega.cu): one definition, hence the same bits in every execution mode.
#pragma once
#include "common.cuh"

// Q8_K (buf_q8_k.rs:84-131): one warp quantises one 256-element super-block; lane l holds the elements 8l .. 8l+7 in v.
//   scale = -128 / x[first argmax |x|],  q = min(round_half_away(scale * x), 127),  d = 1 / scale,  bsums[j] = sum of 16 quants
// qs: the super-block's 256 quants, d: its scale, bsums: its 16 sums (global or shared memory).
__device__ __forceinline__ void cc_quant_q8k_sblock(const float (&v)[8], int lane, int8_t* qs, float* d_out, int16_t* bsums) {
    // first occurrence of the maximum |x| (strict `>` in buf_q8_k.rs:92-98)
    float best_abs = 0.0f, best_val = 0.0f;
    int best_idx = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float av = fabsf(v[i]);
        if (av > best_abs) { best_abs = av; best_val = v[i]; best_idx = lane * 8 + i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float oa = __shfl_xor_sync(0xffffffffu, best_abs, o);
        float ov = __shfl_xor_sync(0xffffffffu, best_val, o);
        int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
        if (oa > best_abs || (oa == best_abs && oi < best_idx)) { best_abs = oa; best_val = ov; best_idx = oi; }
    }
    int8_t q[8];
    int s0 = 0;
    float d = 0.0f;
    if (best_abs == 0.0f) {
#pragma unroll
        for (int i = 0; i < 8; i++) q[i] = 0;
    } else {
        float scale = -128.0f / best_val;
        d = 1.0f / scale;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float r = fminf(roundf(scale * v[i]), 127.0f);      // f32::round = half away from zero (B3)
            int qi = __float2int_rz(r);
            qi = max(qi, -128);
            q[i] = (int8_t)qi;
            s0 += qi;
        }
    }
    *reinterpret_cast<int2*>(qs + lane * 8) = *reinterpret_cast<int2*>(q);
    int s1 = __shfl_down_sync(0xffffffffu, s0, 1);
    if ((lane & 1) == 0) bsums[lane >> 1] = (int16_t)(s0 + s1);
    if (lane == 0) *d_out = d;
}
