// dequant.cuh -- element-wise dequantisation from the device planes (shared by repack.cu and mega.cu; header-only because
// the library is built without relocatable device code).
#pragma once
#include "common.cuh"

// Q8_0 qs plane: byte j of block b (row-relative) -> offset inside the row.  Groups of 32 blocks; within a
// group of nbg blocks all first halves (16 B) come first, then all second halves (matvec_stream.cu).
__host__ __device__ inline int64_t q8_0_row_offset(int b, int j, int nb) {
    int g = b >> 5, l = b & 31;
    int nbg = nb - 32 * g < 32 ? nb - 32 * g : 32;
    return (int64_t)g * 1024 + (j >> 4) * 16 * nbg + 16 * l + (j & 15);
}

// one thread per (block, plane byte): trivially parallel, load-time only

// ---------------------------------------------------------------------------------------------------
// Element-wise dequantisation from the device planes: BlockQ*::dequantize (cited per type).
// Association of the float products follows the reference so results are BIT-EXACT
// (file compiled with -fmad=false).
// ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ void get_scale_min_k4(int j, const uint8_t* q, uint8_t* d, uint8_t* m) {   // util.rs:18-27
    if (j < 4) {
        *d = q[j] & 63;
        *m = q[j + 4] & 63;
    } else {
        *d = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4);
        *m = (q[j + 4] >> 4) | ((q[j] >> 6) << 4);
    }
}

// value of element `e` (0-based within the whole tensor, row-major)
__device__ inline float dequant_elem(int t, const DeqPlanes& P, int64_t e) {
    switch (t) {
    case CC_Q8_0: {                                                        // buf_q8_0.rs:18-23
        int64_t b = e >> 5;
        const int nb = (int)(P.cols >> 5);
        int64_t row = e / P.cols;
        float d = h2f_bits(((const uint16_t*)P.p[1])[row * CC_D_STRIDE(nb) + (b - row * nb)]);
        int64_t off = row * P.cols + q8_0_row_offset((int)(b - row * nb), (int)(e & 31), nb);
        return (float)((const int8_t*)P.p[0])[off] * d;
    }
    case CC_Q4_0: {                                                        // buf_q4_0.rs:18-27
        int64_t b = e >> 5; int i = (int)(e & 31);
        uint8_t q = P.p[0][b * 16 + (i & 15)];
        int x = (i < 16 ? (q & 0x0F) : (q >> 4)) - 8;
        const int64_t nb = P.cols >> 5, row = e / P.cols;
        return (float)x * h2f_bits(((const uint16_t*)P.p[1])[row * CC_D_STRIDE(nb) + (b - row * nb)]);
    }
    case CC_Q4_1: {                                                        // vec_dot order, buf_q4_1.rs:266-280 (B10)
        int64_t b = e >> 5; int i = (int)(e & 31);
        uint8_t q = P.p[0][b * 16 + (i & 15)];
        float x = (float)(i < 16 ? (q & 0x0F) : (q >> 4));
        const uint16_t* dm = (const uint16_t*)P.p[1] + b * 2;
        return x * h2f_bits(dm[0]) + h2f_bits(dm[1]);
    }
    case CC_Q5_0: {                                                        // buf_q5_0.rs:22-37
        int64_t b = e >> 5; int i = (int)(e & 31);
        uint8_t q = P.p[0][b * 16 + (i & 15)];
        uint32_t qh = ((const uint32_t*)P.p[2])[b];
        int x = (int)((i < 16 ? (q & 0x0F) : (q >> 4)) | (((qh >> i) & 1) << 4)) - 16;
        return (float)x * h2f_bits(((const uint16_t*)P.p[1])[b]);
    }
    case CC_Q5_1: {                                                        // buf_q5_1.rs:20-37
        int64_t b = e >> 5; int i = (int)(e & 31);
        uint8_t q = P.p[0][b * 16 + (i & 15)];
        uint32_t qh = ((const uint32_t*)P.p[2])[b];
        float x = (float)((i < 16 ? (q & 0x0F) : (q >> 4)) | (((qh >> i) & 1) << 4));
        const uint16_t* dm = (const uint16_t*)P.p[1] + b * 2;
        return x * h2f_bits(dm[0]) + h2f_bits(dm[1]);
    }
    case CC_Q2_K: {                                                        // buf_q2_k.rs:35-69
        int64_t b = e >> 8; int i = (int)(e & 255);
        int half = i >> 7, r = i & 127, j = r >> 5, l = r & 31;            // 128-half, shift group j, byte l
        uint8_t sc = P.p[1][b * 16 + half * 8 + j * 2 + (l >> 4)];
        const uint16_t* dd = (const uint16_t*)P.p[2] + b * 2;
        float dl = h2f_bits(dd[0]) * (float)(sc & 0xF), ml = h2f_bits(dd[1]) * (float)(sc >> 4);
        uint8_t q = P.p[0][b * 64 + half * 32 + l];
        return dl * (float)((q >> (2 * j)) & 3) - ml;
    }
    case CC_Q3_K: {                                                        // buf_q3_k.rs:37-84
        int64_t b = e >> 8; int i = (int)(e & 255);
        int half = i >> 7, r = i & 127, j = r >> 5, l = r & 31;
        const uint8_t* s12 = P.p[2] + b * 12;
        int is = half * 8 + j * 2 + (l >> 4);
        // 6-bit scale `is` out of the 12 packed bytes (kmask shuffle of buf_q3_k.rs:44-58)
        int lo = is < 8 ? (s12[is] & 0xF) : (s12[is - 8] >> 4);
        int hi = (s12[8 + (is & 3)] >> (2 * (is >> 2))) & 3;
        int scale = (int)(int8_t)(lo | (hi << 4)) - 32;
        float dl = h2f_bits(((const uint16_t*)P.p[3])[b]) * (float)scale;
        uint8_t m = (uint8_t)(1u << (half * 4 + j));
        int mm = (P.p[1][b * 32 + l] & m) ? 0 : 4;
        uint8_t q = P.p[0][b * 64 + half * 32 + l];
        return dl * (float)((int)((q >> (2 * j)) & 3) - mm);
    }
    case CC_Q4_K: {                                                        // buf_q4_k.rs:24-48
        int64_t b = e >> 8; int i = (int)(e & 255);
        const uint8_t* blk = P.p[0] + b * 144;
        int c = i >> 6, r = i & 63, l = r & 31, hi = r >> 5;
        uint8_t sc, m;
        get_scale_min_k4(2 * c + hi, blk + 4, &sc, &m);
        float d = h2f_bits(*(const uint16_t*)blk) * (float)sc, mn = h2f_bits(*(const uint16_t*)(blk + 2)) * (float)m;
        uint8_t q = blk[16 + 32 * c + l];
        return d * (float)(hi ? (q >> 4) : (q & 0xF)) - mn;
    }
    case CC_Q5_K: {                                                        // buf_q5_k.rs:23-59 on the ggml field order
        int64_t b = e >> 8; int i = (int)(e & 255);
        const uint8_t* blk = P.p[0] + b * 176;
        int c = i >> 6, r = i & 63, l = r & 31, hi = r >> 5;
        uint8_t sc, m;
        get_scale_min_k4(2 * c + hi, blk + 4, &sc, &m);
        float d = h2f_bits(*(const uint16_t*)blk) * (float)sc, mn = h2f_bits(*(const uint16_t*)(blk + 2)) * (float)m;
        uint8_t q = blk[48 + 32 * c + l];
        uint8_t qh = blk[16 + l];
        float v = (float)(hi ? (q >> 4) : (q & 0xF)) + ((qh & (1u << (2 * c + hi))) ? 16.0f : 0.0f);
        return d * v - mn;
    }
    case CC_Q6_K: {                                                        // buf_q6_k.rs:21-47
        int64_t b = e >> 8; int i = (int)(e & 255);
        int half = i >> 7, r = i & 127, g = r >> 5, l = r & 31;            // g: q1..q4
        const uint8_t* ql = P.p[0] + b * 128 + 64 * half;
        uint8_t qh = P.p[1][b * 64 + 32 * half + l];
        int8_t sc = ((const int8_t*)P.p[2])[b * 16 + 8 * half + (l >> 4) + 2 * g];
        uint8_t lo = (g & 1) ? ql[l + 32] : ql[l];
        int nib = (g >= 2) ? (lo >> 4) : (lo & 0xF);
        int q = (int)(int8_t)(nib | (((qh >> (2 * g)) & 3) << 4)) - 32;
        float d = h2f_bits(((const uint16_t*)P.p[3])[b]);
        return d * (float)sc * (float)q;
    }
    case CC_Q8_K: {                                                        // buf_q8_k.rs:15-20
        int64_t b = e >> 8;
        return ((const float*)P.p[1])[b] * (float)((const int8_t*)P.p[0])[e];
    }
    }
    return 0.0f;
}

