/*
 * crabml_oracle.c -- CPU restatement of crabml's quantized decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see crabml_oracle.h).  Not shipped, not linked by
 * the CUDA product.  Parity is PINNED against the reference's own KATs and
 * golden generations (tests/test_oracle_*.py).
 *
 * Citations are file:line in crabml/crabml @0151f893, relative to
 * crabml-core/src/cpu/ unless a longer path is given.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -mavx2 -mfma -mf16c).
 * -ffp-contract=off matters: rustc never fuses a*b+c, so neither may we.
 */
#define _GNU_SOURCE
#include "crabml_oracle.h"

#include <immintrin.h>
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define QK_K 256

/* ---------------------------------------------------------------- f16 -- */
/* `half` crate conversions are IEEE round-to-nearest-even; so is _Float16. */
static inline float h2f(uint16_t h) {
    _Float16 v;
    memcpy(&v, &h, 2);
    return (float)v;
}
static inline uint16_t f2h(float f) {
    _Float16 v = (_Float16)f;
    uint16_t h;
    memcpy(&h, &v, 2);
    return h;
}
/* half::f16 `a * b`: widen, multiply in f32, round back to f16 */
static inline uint16_t hmul(uint16_t a, uint16_t b) { return f2h(h2f(a) * h2f(b)); }
static inline uint16_t hadd(uint16_t a, uint16_t b) { return f2h(h2f(a) + h2f(b)); }

void oc_f16_to_f32(const uint16_t* src, float* dst, size_t n) {
    for (size_t i = 0; i < n; i++) dst[i] = h2f(src[i]);
}
void oc_f32_to_f16(const float* src, uint16_t* dst, size_t n) {
    for (size_t i = 0; i < n; i++) dst[i] = f2h(src[i]);
}

/* cpu_device.rs:108-115: LUT[bits] = f16(exp(f32(f16 bits))) */
void oc_exp_lut(uint16_t* lut) {
    for (uint32_t x = 0; x < 65536; x++) lut[x] = f2h(expf(h2f((uint16_t)x)));
}
/* gelu.rs:17-21 + cpu_device.rs:117-124 */
static inline float gelu_single(float x) {
    const float COEF_A = 0.044715f;
    const float SQRT_2_OVER_PI = (float)0.7978845608028654;
    return 0.5f * x * (1.0f + tanhf(SQRT_2_OVER_PI * x * (1.0f + COEF_A * x * x)));
}
void oc_gelu_lut(uint16_t* lut) {
    for (uint32_t x = 0; x < 65536; x++) lut[x] = f2h(gelu_single(h2f((uint16_t)x)));
}
/* buf_f32.rs:29-35 */
static inline float exp_cached(float x, const uint16_t* lut) { return h2f(lut[f2h(x)]); }

/* ------------------------------------------------------- block structs -- */
#pragma pack(push, 1)
typedef struct { uint16_t d; int8_t qs[32]; } blk_q8_0;                       /* buf_q8_0.rs:8-13 */
typedef struct { uint16_t d; uint8_t qs[16]; } blk_q4_0;                      /* buf_q4_0.rs:10-15 */
typedef struct { uint16_t d, m; uint8_t qs[16]; } blk_q4_1;                   /* buf_q4_1.rs:10-16 */
typedef struct { uint16_t d; uint8_t qh[4]; uint8_t qs[16]; } blk_q5_0;       /* buf_q5_0.rs:13-19 */
typedef struct { uint16_t d, m; uint8_t qh[4]; uint8_t qs[16]; } blk_q5_1;    /* buf_q5_1.rs:10-17 */
typedef struct { uint16_t d, s; int8_t qs[32]; } blk_q8_1;                    /* buf_q8_1.rs:73-79 */
typedef struct { uint8_t scales[16]; uint8_t qs[64]; uint16_t d, dmin; } blk_q2_k;      /* buf_q2_k.rs:17-28 */
typedef struct { uint8_t hmask[32]; uint8_t qs[64]; uint8_t scales[12]; uint16_t d; } blk_q3_k; /* buf_q3_k.rs:19-30 */
typedef struct { uint16_t d, dmin; uint8_t scales[12]; uint8_t qs[128]; } blk_q4_k;     /* buf_q4_k.rs:14-21 */
/* ggml / GGUF on-disk order (B9: the reference struct buf_q5_k.rs:15-21 differs) */
typedef struct { uint16_t d, dmin; uint8_t scales[12]; uint8_t qh[32]; uint8_t qs[128]; } blk_q5_k;
typedef struct { uint8_t qs[128]; uint8_t qh[32]; uint8_t scales[12]; uint16_t d, dmin; } blk_q5_k_ref;
typedef struct { uint8_t ql[128]; uint8_t qh[64]; int8_t scales[16]; uint16_t d; } blk_q6_k;    /* buf_q6_k.rs:11-18 */
typedef struct { float d; int8_t qs[256]; int16_t bsums[16]; } blk_q8_k;                /* buf_q8_k.rs:6-12 */
#pragma pack(pop)

int oc_block_elems(int t) {
    switch (t) {
    case OC_F32: case OC_F16: return 1;
    case OC_Q4_0: case OC_Q4_1: case OC_Q5_0: case OC_Q5_1: case OC_Q8_0: case OC_Q8_1: return 32;
    case OC_Q2_K: case OC_Q3_K: case OC_Q4_K: case OC_Q5_K: case OC_Q6_K: case OC_Q8_K: return 256;
    }
    return 0;
}
size_t oc_block_bytes(int t) {
    switch (t) {
    case OC_F32: return 4; case OC_F16: return 2;
    case OC_Q4_0: return sizeof(blk_q4_0); case OC_Q4_1: return sizeof(blk_q4_1);
    case OC_Q5_0: return sizeof(blk_q5_0); case OC_Q5_1: return sizeof(blk_q5_1);
    case OC_Q8_0: return sizeof(blk_q8_0); case OC_Q8_1: return sizeof(blk_q8_1);
    case OC_Q2_K: return sizeof(blk_q2_k); case OC_Q3_K: return sizeof(blk_q3_k);
    case OC_Q4_K: return sizeof(blk_q4_k); case OC_Q5_K: return sizeof(blk_q5_k);
    case OC_Q6_K: return sizeof(blk_q6_k); case OC_Q8_K: return sizeof(blk_q8_k);
    }
    return 0;
}
/* buf/api.rs:142-159 */
int oc_vec_dot_rhs_type(int t) {
    switch (t) {
    case OC_F32: return OC_F32; case OC_F16: return OC_F16;
    case OC_Q8_0: case OC_Q4_0: case OC_Q5_0: return OC_Q8_0;
    case OC_Q8_1: case OC_Q4_1: case OC_Q5_1: return OC_Q8_1;
    case OC_Q2_K: case OC_Q3_K: case OC_Q4_K: case OC_Q5_K: case OC_Q6_K: case OC_Q8_K: return OC_Q8_K;
    }
    return -1;
}

/* util.rs:18-27 */
static inline void get_scale_min_k4(int j, const uint8_t* q, uint8_t* d, uint8_t* m) {
    if (j < 4) {
        *d = q[j] & 63;
        *m = q[j + 4] & 63;
    } else {
        *d = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4);
        *m = (q[j + 4] >> 4) | ((q[j] >> 6) << 4);
    }
}
static inline uint32_t rd_u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

/* --------------------------------------------------------- dequantize -- */
static void deq_q8_0(const blk_q8_0* b, float* o) {            /* buf_q8_0.rs:18-23 */
    float d = h2f(b->d);
    for (int i = 0; i < 32; i++) o[i] = (float)b->qs[i] * d;
}
static void deq_q4_0(const blk_q4_0* b, float* o) {            /* buf_q4_0.rs:18-27 */
    float d = h2f(b->d);
    for (int i = 0; i < 16; i++) {
        int x0 = (b->qs[i] & 0x0F) - 8, x1 = (b->qs[i] >> 4) - 8;
        o[i] = (float)x0 * d;
        o[i + 16] = (float)x1 * d;
    }
}
static void deq_q4_1(const blk_q4_1* b, float* o, int bug) {
    float d = h2f(b->d), m = h2f(b->m);
    for (int i = 0; i < 16; i++) {
        float x0 = (float)(b->qs[i] & 0x0F), x1 = (float)(b->qs[i] >> 4);
        if (bug) {             /* buf_q4_1.rs:19-30: interleaved (2i, 2i+1) -- B10 */
            o[i * 2] = x0 * d + m;
            o[i * 2 + 1] = x1 * d + m;
        } else {               /* element order of vec_dot (buf_q4_1.rs:266-280) == ggml */
            o[i] = x0 * d + m;
            o[i + 16] = x1 * d + m;
        }
    }
}
static void deq_q5_0(const blk_q5_0* b, float* o) {            /* buf_q5_0.rs:22-37 */
    float d = h2f(b->d);
    uint32_t qh = rd_u32(b->qh);
    for (int i = 0; i < 16; i++) {
        uint8_t xh0 = (uint8_t)(((qh >> i) << 4) & 0x10);
        uint8_t xh1 = (uint8_t)((qh >> (i + 12)) & 0x10);
        int x0 = (int)((b->qs[i] & 0x0F) | xh0) - 16;
        int x1 = (int)((b->qs[i] >> 4) | xh1) - 16;
        o[i] = (float)x0 * d;
        o[i + 16] = (float)x1 * d;
    }
}
static void deq_q5_1(const blk_q5_1* b, float* o) {            /* buf_q5_1.rs:20-37 */
    float d = h2f(b->d), m = h2f(b->m);
    uint32_t qh = rd_u32(b->qh);
    for (int i = 0; i < 16; i++) {
        uint8_t xh0 = (uint8_t)(((qh >> i) << 4) & 0x10);
        uint8_t xh1 = (uint8_t)((qh >> (i + 12)) & 0x10);
        uint8_t x0 = (b->qs[i] & 0x0F) | xh0, x1 = (b->qs[i] >> 4) | xh1;
        o[i] = (float)x0 * d + m;
        o[i + 16] = (float)x1 * d + m;
    }
}
static void deq_q8_1(const blk_q8_1* b, float* o) {            /* buf_q8_1.rs:81-88 */
    float d = h2f(b->d);
    for (int i = 0; i < 32; i++) o[i] = (float)b->qs[i] * d;
}
static void deq_q2_k(const blk_q2_k* b, float* o) {            /* buf_q2_k.rs:35-69 */
    float d = h2f(b->d), min = h2f(b->dmin);
    int is = 0, oi = 0;
    for (int n = 0; n < QK_K; n += 128) {
        const uint8_t* qs = b->qs + (n / 128) * 32;
        int shift = 0;
        for (int j = 0; j < 4; j++) {
            uint8_t sc = b->scales[is++];
            float dl = d * (float)(sc & 0xF), ml = min * (float)(sc >> 4);
            for (int l = 0; l < 16; l++) o[oi++] = dl * (float)((qs[l] >> shift) & 3) - ml;
            sc = b->scales[is++];
            dl = d * (float)(sc & 0xF); ml = min * (float)(sc >> 4);
            for (int l = 16; l < 32; l++) o[oi++] = dl * (float)((qs[l] >> shift) & 3) - ml;
            shift += 2;
        }
    }
}
static void q3k_scales(const uint8_t* sc12, int8_t out[16]) {  /* buf_q3_k.rs:44-58 */
    const uint32_t KMASK_1 = 0x03030303u, KMASK_2 = 0x0f0f0f0fu;
    uint32_t aux[4] = {0, 0, 0, 0};
    memcpy(aux, sc12, 12);
    uint32_t tmp = aux[2];
    aux[2] = ((aux[0] >> 4) & KMASK_2) | (((tmp >> 4) & KMASK_1) << 4);
    aux[3] = ((aux[1] >> 4) & KMASK_2) | (((tmp >> 6) & KMASK_1) << 4);
    aux[0] = (aux[0] & KMASK_2) | (((tmp) & KMASK_1) << 4);
    aux[1] = (aux[1] & KMASK_2) | (((tmp >> 2) & KMASK_1) << 4);
    memcpy(out, aux, 16);
}
static void deq_q3_k(const blk_q3_k* b, float* o) {            /* buf_q3_k.rs:37-84 */
    float d_all = h2f(b->d);
    int8_t scales[16];
    q3k_scales(b->scales, scales);
    uint8_t m = 1;
    int qs_i = 0, oi = 0, is = 0;
    for (int n = 0; n < QK_K; n += 128) {
        int shift = 0;
        for (int j = 0; j < 4; j++) {
            float dl = d_all * (float)(scales[is++] - 32);
            for (int l = 0; l < 16; l++) {
                int mm = (b->hmask[l] & m) ? 0 : 4;
                o[oi++] = dl * (float)((int)((b->qs[l + qs_i] >> shift) & 3) - mm);
            }
            dl = d_all * (float)(scales[is++] - 32);
            for (int l = 0; l < 16; l++) {
                int mm = (b->hmask[l + 16] & m) ? 0 : 4;
                o[oi++] = dl * (float)((int)((b->qs[l + qs_i + 16] >> shift) & 3) - mm);
            }
            shift += 2;
            m <<= 1;
        }
        qs_i += 32;
    }
}
static void deq_q4_k(const blk_q4_k* b, float* o) {            /* buf_q4_k.rs:24-48 */
    float d = h2f(b->d), min = h2f(b->dmin);
    int is = 0;
    for (int c = 0; c < 4; c++) {
        const uint8_t* q = b->qs + 32 * c;
        float* oc = o + 64 * c;
        uint8_t sc, m;
        get_scale_min_k4(is, b->scales, &sc, &m);
        float d1 = d * (float)sc, m1 = min * (float)m;
        get_scale_min_k4(is + 1, b->scales, &sc, &m);
        float d2 = d * (float)sc, m2 = min * (float)m;
        for (int l = 0; l < 32; l++) {
            oc[l] = d1 * (float)(q[l] & 0xF) - m1;
            oc[l + 32] = d2 * (float)(q[l] >> 4) - m2;
        }
        is += 2;
    }
}
static void deq_q5_k_fields(const uint8_t* qs, const uint8_t* qh, const uint8_t* scales,
                            uint16_t hd, uint16_t hdmin, float* o) {  /* buf_q5_k.rs:23-59 */
    float d = h2f(hd), min = h2f(hdmin);
    uint8_t u1 = 1, u2 = 2;
    int is = 0;
    for (int c = 0; c < 4; c++) {
        const uint8_t* q = qs + 32 * c;
        float* oc = o + 64 * c;
        uint8_t sc, m;
        get_scale_min_k4(is, scales, &sc, &m);
        float d1 = d * (float)sc, m1 = min * (float)m;
        get_scale_min_k4(is + 1, scales, &sc, &m);
        float d2 = d * (float)sc, m2 = min * (float)m;
        for (int l = 0; l < 32; l++) {
            oc[l] = d1 * ((float)(q[l] & 0xF) + ((qh[l] & u1) ? 16.0f : 0.0f)) - m1;
            oc[l + 32] = d2 * ((float)(q[l] >> 4) + ((qh[l] & u2) ? 16.0f : 0.0f)) - m2;
        }
        is += 2;
        u1 <<= 2;
        u2 <<= 2;
    }
}
static void deq_q6_k(const blk_q6_k* b, float* o) {            /* buf_q6_k.rs:21-47 */
    float d = h2f(b->d);
    for (int idx = 0; idx < 2; idx++) {
        float* oc = o + 128 * idx;
        const int8_t* sc = b->scales + 8 * idx;
        const uint8_t* ql = b->ql + 64 * idx;
        const uint8_t* qh = b->qh + 32 * idx;
        for (int l = 0; l < 32; l++) {
            int is = l / 16;
            int8_t q1 = (int8_t)((ql[l] & 0xF) | ((qh[l] & 3) << 4)) - 32;
            int8_t q2 = (int8_t)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
            int8_t q3 = (int8_t)((ql[l] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
            int8_t q4 = (int8_t)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
            oc[l] = d * (float)sc[is] * (float)q1;
            oc[l + 32] = d * (float)sc[is + 2] * (float)q2;
            oc[l + 64] = d * (float)sc[is + 4] * (float)q3;
            oc[l + 96] = d * (float)sc[is + 6] * (float)q4;
        }
    }
}
static void deq_q8_k(const blk_q8_k* b, float* o) {            /* buf_q8_k.rs:15-20 */
    for (int i = 0; i < 256; i++) o[i] = b->d * (float)b->qs[i];
}

int oc_dequantize(int type, const void* blocks, size_t n, float* out, int flags) {
    int be = oc_block_elems(type);
    if (be == 0 || n % (size_t)be) return -1;
    size_t nb = n / (size_t)be;
    size_t bb = oc_block_bytes(type);
    const uint8_t* p = (const uint8_t*)blocks;
    int bug = flags & OC_BUGCOMPAT;
    for (size_t i = 0; i < nb; i++, p += bb, out += be) {
        switch (type) {
        case OC_F32: memcpy(out, p, 4); break;
        case OC_F16: { uint16_t h; memcpy(&h, p, 2); *out = h2f(h); } break;
        case OC_Q8_0: deq_q8_0((const blk_q8_0*)p, out); break;
        case OC_Q4_0: deq_q4_0((const blk_q4_0*)p, out); break;
        case OC_Q4_1: deq_q4_1((const blk_q4_1*)p, out, bug); break;
        case OC_Q5_0: deq_q5_0((const blk_q5_0*)p, out); break;
        case OC_Q5_1: deq_q5_1((const blk_q5_1*)p, out); break;
        case OC_Q8_1: deq_q8_1((const blk_q8_1*)p, out); break;
        case OC_Q2_K: deq_q2_k((const blk_q2_k*)p, out); break;
        case OC_Q3_K: deq_q3_k((const blk_q3_k*)p, out); break;
        case OC_Q4_K: deq_q4_k((const blk_q4_k*)p, out); break;
        case OC_Q5_K:
            if (bug) { const blk_q5_k_ref* b = (const blk_q5_k_ref*)p; deq_q5_k_fields(b->qs, b->qh, b->scales, b->d, b->dmin, out); }
            else     { const blk_q5_k* b = (const blk_q5_k*)p;         deq_q5_k_fields(b->qs, b->qh, b->scales, b->d, b->dmin, out); }
            break;
        case OC_Q6_K: deq_q6_k((const blk_q6_k*)p, out); break;
        case OC_Q8_K: deq_q8_k((const blk_q8_k*)p, out); break;
        default: return -1;
        }
    }
    return 0;
}

/* ----------------------------------------------- activation quantizers -- */
/* Rust `f32 as i32` / simd cast: truncate toward zero, saturate, NaN -> 0 */
static inline int32_t cast_f32_i32(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}
static inline int8_t cast_f32_i8(float v) {      /* Rust `f32 as i8` */
    if (v != v) return 0;
    if (v >= 127.0f) return 127;
    if (v <= -128.0f) return -128;
    return (int8_t)(int32_t)v;
}
/* Rust f32::max / f32::min: if one operand is NaN return the other */
static inline float rmaxf(float a, float b) { return (a != a) ? b : (b != b) ? a : (a > b ? a : b); }
static inline float rminf(float a, float b) { return (a != a) ? b : (b != b) ? a : (a < b ? a : b); }

/* buf_q8_0.rs:87-134: d = max|x| / 127 (f32), q = trunc(x / d) (B1), d stored f16 */
static void quantize_q8_0(const float* x, size_t n, blk_q8_0* out) {
    for (size_t i = 0; i < n; i += 32, out++) {
        float max = 0.0f;
        for (int j = 0; j < 32; j++) {
            /* simd_max/reduce_max: NaN handling as in maxnum; inputs are finite here */
            float a = fabsf(x[i + j]);
            max = rmaxf(max, a);
        }
        float d = max / 127.0f;
        for (int j = 0; j < 32; j++) {
            float v = x[i + j] / d;
            out->qs[j] = (int8_t)cast_f32_i32(v);   /* i32 -> i8 wraps; |v| <= 127 here */
        }
        out->d = f2h(d);
    }
}
/* buf_q8_1.rs:90-129 */
static void quantize_q8_1(const float* x, size_t n, blk_q8_1* out) {
    for (size_t i = 0; i < n; i += 32, out++) {
        float max_abs = 0.0f;
        for (int j = 0; j < 32; j++) {
            float a = fabsf(x[i + j]);
            if (a > max_abs) max_abs = a;
        }
        float d = max_abs / 127.0f;
        float s = 0.0f;
        for (int j = 0; j < 32; j++) {
            float sv = x[i + j] / d;
            int8_t q = cast_f32_i8(rminf(rmaxf(sv, -128.0f), 127.0f));
            out->qs[j] = q;
            s += (float)q;
        }
        s *= d;
        out->d = f2h(d);
        out->s = f2h(s);
    }
}
/* buf_q8_k.rs:84-131; `round()` = half away from zero (B3) */
static void quantize_q8_k(const float* x, size_t n, blk_q8_k* out) {
    for (size_t i = 0; i < n; i += 256, out++) {
        float max_abs = 0.0f, max_val = 0.0f;
        for (int j = 0; j < 256; j++) {
            float a = fabsf(x[i + j]);
            if (a > max_abs) { max_abs = a; max_val = x[i + j]; }
        }
        memset(out->bsums, 0, sizeof(out->bsums));
        if (max_abs == 0.0f) {
            out->d = 0.0f;
            memset(out->qs, 0, 256);
            continue;
        }
        float scale = -128.0f / max_val;
        out->d = 1.0f / scale;
        for (int j = 0; j < 256; j++) {
            float v = roundf(scale * x[i + j]);
            out->qs[j] = cast_f32_i8(rminf(v, 127.0f));
        }
        for (int g = 0; g < 16; g++) {
            int32_t sum = 0;
            for (int j = 0; j < 16; j++) sum += out->qs[g * 16 + j];
            out->bsums[g] = (int16_t)sum;
        }
    }
}
int oc_quantize(int act_type, const float* x, size_t n, void* out) {
    switch (act_type) {
    case OC_Q8_0: if (n % 32) return -1; quantize_q8_0(x, n, (blk_q8_0*)out); return 0;
    case OC_Q8_1: if (n % 32) return -1; quantize_q8_1(x, n, (blk_q8_1*)out); return 0;
    case OC_Q8_K: if (n % 256) return -1; quantize_q8_k(x, n, (blk_q8_k*)out); return 0;
    case OC_F32: memcpy(out, x, n * 4); return 0;
    case OC_F16: oc_f32_to_f16(x, (uint16_t*)out, n); return 0;   /* buf/api.rs:198 */
    }
    return -1;
}

/* -------------------------------------------------------------- vec_dot -- */
/* archutil/x86_64.rs:4-53 restated with the same intrinsics */
static inline __m256 sum_i16_pairs_float(__m128i xh, __m128i xl) {
    __m128i ones = _mm_set1_epi16(1);
    __m128i sl = _mm_madd_epi16(ones, xl);
    __m128i sh = _mm_madd_epi16(ones, xh);
    return _mm256_cvtepi32_ps(_mm256_set_m128i(sh, sl));
}
static inline __m256 mul_sum_us8_pairs_float(__m256i ax, __m256i sy) {
    __m128i axl = _mm256_castsi256_si128(ax), axh = _mm256_extractf128_si256(ax, 1);
    __m128i syl = _mm256_castsi256_si128(sy), syh = _mm256_extractf128_si256(sy, 1);
    __m128i dotl = _mm_maddubs_epi16(axl, syl), doth = _mm_maddubs_epi16(axh, syh);
    return sum_i16_pairs_float(doth, dotl);
}
static inline __m256 mul_sum_i8_pairs_float(__m256i x, __m256i y) {
    __m256i ax = _mm256_sign_epi8(x, x);
    __m256i sy = _mm256_sign_epi8(y, x);
    return mul_sum_us8_pairs_float(ax, sy);
}
static inline float hsum_float_8(__m256 x) {
    __m128 res = _mm256_extractf128_ps(x, 1);
    res = _mm_add_ps(res, _mm256_castps256_ps128(x));
    res = _mm_add_ps(res, _mm_movehl_ps(res, res));
    res = _mm_add_ss(res, _mm_movehdup_ps(res));
    return _mm_cvtss_f32(res);
}
static inline __m256i bytes_from_nibbles_32(const uint8_t* p) {
    __m128i tmp = _mm_loadu_si128((const __m128i*)p);
    __m256i bytes = _mm256_set_m128i(_mm_srli_epi16(tmp, 4), tmp);
    return _mm256_and_si256(_mm256_set1_epi8(0xF), bytes);
}

/* buf_q8_0.rs:275-286 */
static float dot_q8_0_scalar(const blk_q8_0* a, const blk_q8_0* b, size_t nb) {
    float sumf = 0.0f;
    for (size_t i = 0; i < nb; i++) {
        int32_t sumi = 0;
        for (int j = 0; j < 32; j++) sumi += (int32_t)a[i].qs[j] * (int32_t)b[i].qs[j];
        sumf += (float)sumi * h2f(a[i].d) * h2f(b[i].d);
    }
    return sumf;
}
/* buf_q8_0.rs:228-272 */
static float dot_q8_0_avx2(const blk_q8_0* a, const blk_q8_0* b, size_t nb) {
    __m256 acc0 = _mm256_setzero_ps(), acc1 = _mm256_setzero_ps();
    size_t i = 0;
    for (; i + 1 < nb; i += 2) {
        __m256 d0 = _mm256_set1_ps(h2f(a[i].d) * h2f(b[i].d));
        __m256 d1 = _mm256_set1_ps(h2f(a[i + 1].d) * h2f(b[i + 1].d));
        __m256i qa0 = _mm256_loadu_si256((const __m256i*)a[i].qs);
        __m256i qb0 = _mm256_loadu_si256((const __m256i*)b[i].qs);
        __m256i qa1 = _mm256_loadu_si256((const __m256i*)a[i + 1].qs);
        __m256i qb1 = _mm256_loadu_si256((const __m256i*)b[i + 1].qs);
        acc0 = _mm256_fmadd_ps(d0, mul_sum_i8_pairs_float(qa0, qb0), acc0);
        acc1 = _mm256_fmadd_ps(d1, mul_sum_i8_pairs_float(qa1, qb1), acc1);
    }
    if (nb % 2 == 1) {
        __m256 d = _mm256_set1_ps(h2f(a[nb - 1].d) * h2f(b[nb - 1].d));
        __m256i qa = _mm256_loadu_si256((const __m256i*)a[nb - 1].qs);
        __m256i qb = _mm256_loadu_si256((const __m256i*)b[nb - 1].qs);
        acc0 = _mm256_fmadd_ps(d, mul_sum_i8_pairs_float(qa, qb), acc0);
    }
    return hsum_float_8(_mm256_add_ps(acc0, acc1));
}
/* buf_q4_0.rs:240-253 */
static float dot_q4_0_scalar(const blk_q4_0* a, const blk_q8_0* b, size_t nb) {
    float sumf = 0.0f;
    for (size_t i = 0; i < nb; i++) {
        int32_t sumi = 0;
        for (int j = 0; j < 16; j++) {
            int v0 = (a[i].qs[j] & 0x0F) - 8, v1 = (a[i].qs[j] >> 4) - 8;
            sumi += v0 * b[i].qs[j] + v1 * b[i].qs[j + 16];
        }
        sumf += (float)sumi * h2f(a[i].d) * h2f(b[i].d);
    }
    return sumf;
}
/* buf_q4_0.rs:215-238, including the blocks%32 gate (B12) */
static float dot_q4_0_avx2(const blk_q4_0* a, const blk_q8_0* b, size_t nb) {
    if (nb % 32 != 0) return dot_q4_0_scalar(a, b, nb);
    __m256 acc = _mm256_setzero_ps();
    for (size_t i = 0; i < nb; i++) {
        __m256 d = _mm256_set1_ps(h2f(a[i].d) * h2f(b[i].d));
        __m256i bx = bytes_from_nibbles_32(a[i].qs);
        bx = _mm256_sub_epi8(bx, _mm256_set1_epi8(8));
        __m256i by = _mm256_loadu_si256((const __m256i*)b[i].qs);
        acc = _mm256_fmadd_ps(d, mul_sum_i8_pairs_float(bx, by), acc);
    }
    return hsum_float_8(acc);
}
/* buf_q4_1.rs:266-280 (scalar; the AVX2 kernel is wrong, B11) */
static float dot_q4_1(const blk_q4_1* a, const blk_q8_1* b, size_t nb) {
    float sumf = 0.0f;
    for (size_t i = 0; i < nb; i++) {
        int32_t sumi = 0;
        for (int j = 0; j < 16; j++) {
            int v0 = a[i].qs[j] & 0x0F, v1 = (a[i].qs[j] >> 4) & 0x0F;
            sumi += v0 * b[i].qs[j] + v1 * b[i].qs[j + 16];
        }
        sumf += h2f(hmul(a[i].d, b[i].d)) * (float)sumi + h2f(hmul(a[i].m, b[i].s));
    }
    return sumf;
}
/* buf_q5_0.rs:145-163 */
static float dot_q5_0(const blk_q5_0* a, const blk_q8_0* b, size_t nb) {
    float sumf = 0.0f;
    for (size_t i = 0; i < nb; i++) {
        uint32_t qh = rd_u32(a[i].qh);
        int32_t sumi = 0;
        for (int j = 0; j < 16; j++) {
            uint32_t xh0 = ((qh & (1u << j)) >> j) << 4;
            uint32_t xh1 = (qh & (1u << (j + 16))) >> (j + 12);
            int32_t x0 = (int32_t)((a[i].qs[j] & 0x0F) | xh0) - 16;
            int32_t x1 = (int32_t)((a[i].qs[j] >> 4) | xh1) - 16;
            sumi += x0 * b[i].qs[j] + x1 * b[i].qs[j + 16];
        }
        sumf += (float)sumi * h2f(a[i].d) * h2f(b[i].d);
    }
    return sumf;
}
/* buf_q5_1.rs:142-161 */
static float dot_q5_1(const blk_q5_1* a, const blk_q8_1* b, size_t nb) {
    float sumf = 0.0f;
    for (size_t i = 0; i < nb; i++) {
        uint32_t qh = rd_u32(a[i].qh);
        int32_t sumi = 0;
        for (int j = 0; j < 16; j++) {
            uint32_t xh0 = ((qh >> j) << 4) & 0x10;
            uint32_t xh1 = (qh >> (j + 12)) & 0x10;
            int32_t x0 = (int32_t)((a[i].qs[j] & 0xF) | xh0);
            int32_t x1 = (int32_t)((a[i].qs[j] >> 4) | xh1);
            sumi += x0 * b[i].qs[j] + x1 * b[i].qs[j + 16];
        }
        sumf += (float)sumi * h2f(hmul(a[i].d, b[i].d)) + h2f(hmul(a[i].m, b[i].s));
    }
    return sumf;
}
/* buf_q8_1.rs has no vec_dot body worth restating (never a weight type in GGUF);
 * api.rs:238 routes (Q8_1,Q8_1) to a plain int dot -- restated as d*d*sumi. */
static float dot_q8_1(const blk_q8_1* a, const blk_q8_1* b, size_t nb) {
    float sumf = 0.0f;
    for (size_t i = 0; i < nb; i++) {
        int32_t sumi = 0;
        for (int j = 0; j < 32; j++) sumi += (int32_t)a[i].qs[j] * (int32_t)b[i].qs[j];
        sumf += (float)sumi * h2f(a[i].d) * h2f(b[i].d);
    }
    return sumf;
}
/* buf_q2_k.rs:214-257.  summs: i16 in the reference (B7) -> i32 unless bugcompat */
static float dot_q2_k(const blk_q2_k* a, const blk_q8_k* b, size_t nb, int bug) {
    float sumf = 0.0f;
    for (size_t i = 0; i < nb; i++) {
        int32_t summs = 0;
        int16_t summs16 = 0;
        for (int j = 0; j < 16; j++) {
            summs += (int32_t)b[i].bsums[j] * (int32_t)(a[i].scales[j] >> 4);
            summs16 = (int16_t)(summs16 + (int16_t)(b[i].bsums[j] * (int16_t)(a[i].scales[j] >> 4)));
        }
        float dall = b[i].d * h2f(a[i].d), dmin = b[i].d * h2f(a[i].dmin);
        int32_t isum = 0;
        int is = 0, q8_i = 0, q2_i = 0;
        for (int n = 0; n < QK_K / 128; n++) {
            int shift = 0;
            for (int j = 0; j < 4; j++) {
                int32_t d = a[i].scales[is++] & 0xF, isuml = 0;
                for (int l = 0; l < 16; l++) isuml += (int32_t)b[i].qs[q8_i + l] * (int32_t)((a[i].qs[q2_i + l] >> shift) & 3);
                isum += d * isuml;
                d = a[i].scales[is++] & 0xF; isuml = 0;
                for (int l = 16; l < 32; l++) isuml += (int32_t)b[i].qs[q8_i + l] * (int32_t)((a[i].qs[q2_i + l] >> shift) & 3);
                isum += d * isuml;
                shift += 2;
                q8_i += 32;
            }
            q2_i += 32;
        }
        sumf += dall * (float)isum - dmin * (float)(bug ? (int32_t)summs16 : summs);
    }
    return sumf;
}
/* buf_q3_k.rs:240-328 */
static float dot_q3_k(const blk_q3_k* a, const blk_q8_k* b, size_t nb) {
    float sums[8] = {0};
    int8_t aux8[QK_K];
    for (size_t i = 0; i < nb; i++) {
        int a8 = 0, q3 = 0;
        uint8_t m = 1;
        for (int n = 0; n < QK_K; n += 128) {
            for (int sh = 0; sh < 8; sh += 2) {
                for (int l = 0; l < 32; l++) aux8[a8 + l] = (int8_t)((a[i].qs[q3 + l] >> sh) & 3);
                for (int l = 0; l < 32; l++) aux8[a8 + l] -= (a[i].hmask[l] & m) ? 0 : 4;
                a8 += 32;
                m <<= 1;
            }
            q3 += 32;
        }
        int32_t aux32[8] = {0};
        int8_t scales[16];
        q3k_scales(a[i].scales, scales);
        int q8_i = 0;
        a8 = 0;
        for (int j = 0; j < 16; j++) {
            int32_t sc = (int32_t)scales[j] - 32;
            for (int h = 0; h < 2; h++) {
                for (int l = 0; l < 8; l++) aux32[l] += sc * (int32_t)(int16_t)(b[i].qs[q8_i + l] * aux8[a8 + l]);
                q8_i += 8;
                a8 += 8;
            }
        }
        float d = h2f(a[i].d) * b[i].d;
        for (int l = 0; l < 8; l++) sums[l] += d * (float)aux32[l];
    }
    float s = sums[0];
    for (int l = 1; l < 8; l++) s += sums[l];
    return s;
}
/* unpack of the 12 scale bytes into 8 scales + 8 mins: buf_q4_k.rs:219-234 */
static void k4_scales_mins(const uint8_t* sc12, uint8_t scales[8], uint8_t mins[8]) {
    const uint32_t KMASK1 = 0x3f3f3f3fu, KMASK2 = 0x0f0f0f0fu, KMASK3 = 0x03030303u;
    uint32_t utmp[4];
    memcpy(utmp, sc12, 12);
    utmp[3] = ((utmp[2] >> 4) & KMASK2) | (((utmp[1] >> 6) & KMASK3) << 4);
    uint32_t uaux = utmp[1] & KMASK1;
    utmp[1] = (utmp[2] & KMASK2) | (((utmp[0] >> 6) & KMASK3) << 4);
    utmp[2] = uaux;
    utmp[0] &= KMASK1;
    memcpy(scales, &utmp[0], 8);
    memcpy(mins, &utmp[2], 8);
}
/* shared tail of buf_q4_k.rs:236-270 / buf_q5_k.rs:278-312: f32 per-lane accumulation (B8) */
static void k45_accumulate(const int8_t* aux8, const blk_q8_k* b, const uint8_t* scales, const uint8_t* mins,
                           float d, float dmin, float sums[8], float* sumf, int bug) {
    float aux32[8] = {0};
    int64_t sumi = 0;
    for (int j = 0; j < 16; j++) {
        if (bug) sumi += (int16_t)(b->bsums[j] * (int16_t)mins[j / 2]);   /* i16 wrap, B7 */
        else     sumi += (int32_t)b->bsums[j] * (int32_t)mins[j / 2];
    }
    for (int is = 0; is < 8; is++) {
        float scale = (float)scales[is];
        const int8_t* a8 = aux8 + 32 * is;
        const int8_t* q8 = b->qs + 32 * is;
        for (int g = 0; g < 4; g++)
            for (int l = 0; l < 8; l++) {
                int16_t p = (int16_t)(q8[l + 8 * g] * a8[l + 8 * g]);
                aux32[l] += scale * (float)p;
            }
    }
    for (int l = 0; l < 8; l++) sums[l] += d * aux32[l];
    *sumf -= dmin * (float)sumi;
}
/* buf_q4_k.rs:192-277 */
static float dot_q4_k(const blk_q4_k* a, const blk_q8_k* b, size_t nb, int bug) {
    float sums[8] = {0}, sumf = 0.0f;
    int8_t aux8[256];
    for (size_t i = 0; i < nb; i++) {
        for (int c = 0; c < 4; c++)
            for (int l = 0; l < 32; l++) {
                aux8[64 * c + l] = (int8_t)(a[i].qs[32 * c + l] & 0xF);
                aux8[64 * c + l + 32] = (int8_t)(a[i].qs[32 * c + l] >> 4);
            }
        uint8_t scales[8], mins[8];
        k4_scales_mins(a[i].scales, scales, mins);
        k45_accumulate(aux8, &b[i], scales, mins, h2f(a[i].d) * b[i].d, h2f(a[i].dmin) * b[i].d, sums, &sumf, bug);
    }
    for (int l = 0; l < 8; l++) sumf += sums[l];
    return sumf;
}
/* buf_q5_k.rs:223-319 on the ggml field order (B9) */
static float dot_q5_k(const void* av, const blk_q8_k* b, size_t nb, int bug) {
    float sums[8] = {0}, sumf = 0.0f;
    int8_t aux8[256];
    for (size_t i = 0; i < nb; i++) {
        const uint8_t *qs, *qh, *sc;
        uint16_t hd, hdmin;
        if (bug) { const blk_q5_k_ref* a = (const blk_q5_k_ref*)av + i; qs = a->qs; qh = a->qh; sc = a->scales; hd = a->d; hdmin = a->dmin; }
        else     { const blk_q5_k* a = (const blk_q5_k*)av + i;         qs = a->qs; qh = a->qh; sc = a->scales; hd = a->d; hdmin = a->dmin; }
        uint8_t m = 1;
        for (int c = 0; c < 4; c++) {
            for (int l = 0; l < 32; l++) aux8[64 * c + l] = (int8_t)((qs[32 * c + l] & 0xF) + ((qh[l] & m) ? 16 : 0));
            m <<= 1;
            for (int l = 0; l < 32; l++) aux8[64 * c + l + 32] = (int8_t)((qs[32 * c + l] >> 4) + ((qh[l] & m) ? 16 : 0));
            m <<= 1;
        }
        uint8_t scales[8], mins[8];
        k4_scales_mins(sc, scales, mins);
        k45_accumulate(aux8, &b[i], scales, mins, h2f(hd) * b[i].d, h2f(hdmin) * b[i].d, sums, &sumf, bug);
    }
    for (int l = 0; l < 8; l++) sumf += sums[l];
    return sumf;
}
/* buf_q6_k.rs:183-235 */
static float dot_q6_k(const blk_q6_k* a, const blk_q8_k* b, size_t nb) {
    float sums[8] = {0};
    int8_t aux8[256];
    for (size_t i = 0; i < nb; i++) {
        float aux32[8] = {0};
        for (int j = 0; j < 256; j += 128) {
            int8_t* x = aux8 + j;
            const uint8_t* q4 = a[i].ql + j / 2;
            const uint8_t* qh = a[i].qh + j / 4;
            for (int l = 0; l < 32; l++) {
                x[l] = (int8_t)((int)((q4[l] & 0xF) | ((qh[l] & 3) << 4)) - 32);
                x[l + 32] = (int8_t)((int)((q4[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32);
                x[l + 64] = (int8_t)((int)((q4[l] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32);
                x[l + 96] = (int8_t)((int)((q4[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32);
            }
        }
        for (int j = 0; j < 16; j++) {
            float scale = (float)a[i].scales[j];
            for (int h = 0; h < 2; h++)
                for (int l = 0; l < 8; l++) {
                    int16_t p = (int16_t)(b[i].qs[16 * j + 8 * h + l] * aux8[16 * j + 8 * h + l]);
                    aux32[l] += scale * (float)p;
                }
        }
        float d = h2f(a[i].d) * b[i].d;
        for (int l = 0; l < 8; l++) sums[l] += aux32[l] * d;
    }
    float s = 0.0f;
    for (int l = 0; l < 8; l++) s += sums[l];
    return s;
}
/* buf_q8_k.rs:133-224 (scalar fallback form): sumf += d_a*d_b*sumi */
static float dot_q8_k(const blk_q8_k* a, const blk_q8_k* b, size_t nb) {
    float sumf = 0.0f;
    for (size_t i = 0; i < nb; i++) {
        int32_t sumi = 0;
        for (int j = 0; j < 256; j++) sumi += (int32_t)a[i].qs[j] * (int32_t)b[i].qs[j];
        sumf += (float)sumi * a[i].d * b[i].d;   /* buf_q8_k.rs:213-224 */
    }
    return sumf;
}
/* buf_f32.rs:19-27, buf_f16.rs:84-97 */
static float dot_f32(const float* a, const float* b, size_t n) {
    float s = 0.0f;
    for (size_t i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}
static float dot_f16(const uint16_t* a, const uint16_t* b, size_t n) {
    float s = 0.0f;
    for (size_t i = 0; i < n; i++) s += h2f(a[i]) * h2f(b[i]);
    return s;
}

float oc_vec_dot(int w_type, const void* w, const void* act, size_t n, int flags) {
    int be = oc_block_elems(w_type);
    if (be == 0 || n % (size_t)be) return NAN;
    size_t nb = n / (size_t)be;
    int avx = flags & OC_ORDER_AVX2, bug = flags & OC_BUGCOMPAT;
    switch (w_type) {
    case OC_F32: return dot_f32((const float*)w, (const float*)act, n);
    case OC_F16: return dot_f16((const uint16_t*)w, (const uint16_t*)act, n);
    case OC_Q8_0: return avx ? dot_q8_0_avx2(w, act, nb) : dot_q8_0_scalar(w, act, nb);
    case OC_Q4_0: return avx ? dot_q4_0_avx2(w, act, nb) : dot_q4_0_scalar(w, act, nb);
    case OC_Q4_1: return dot_q4_1(w, act, nb);
    case OC_Q5_0: return dot_q5_0(w, act, nb);
    case OC_Q5_1: return dot_q5_1(w, act, nb);
    case OC_Q8_1: return dot_q8_1(w, act, nb);
    case OC_Q2_K: return dot_q2_k(w, act, nb, bug);
    case OC_Q3_K: return dot_q3_k(w, act, nb);
    case OC_Q4_K: return dot_q4_k(w, act, nb, bug);
    case OC_Q5_K: return dot_q5_k(w, act, nb, bug);
    case OC_Q6_K: return dot_q6_k(w, act, nb);
    case OC_Q8_K: return dot_q8_k(w, act, nb);
    }
    return NAN;
}

/* ----------------------------------------------------------------- gemv -- */
typedef struct {
    int w_type, flags;
    const uint8_t* w;
    const uint8_t* act;
    size_t m, k, row_bytes, act_row_bytes;
    float* out;
    size_t begin, end;   /* span of C handled by this worker */
} gemv_job;

/* primitives/matmul_vec.rs:57-76: each worker walks its span in chunks of 16 */
static void* gemv_worker(void* p) {
    gemv_job* j = (gemv_job*)p;
    for (size_t e = j->begin; e < j->end; e++) {
        size_t mi = e % j->m, bi = e / j->m;
        j->out[e] = oc_vec_dot(j->w_type, j->w + mi * j->row_bytes, j->act + bi * j->act_row_bytes, j->k, j->flags);
    }
    return NULL;
}

/* thread_pool.rs:8-88: N persistent workers; `scoped` runs the first thunk on the caller, hands the others to
 * the workers and busy-waits on a counter.  Restated with a generation counter + spin (then yield) waits. */
#define OC_MAX_WORKERS 256
static struct {
    pthread_t tid[OC_MAX_WORKERS];
    int n_workers;
    gemv_job* jobs;              /* jobs[1..n_jobs-1] go to workers 0..n_jobs-2 */
    int n_jobs;
    volatile unsigned long gen;  /* bumped by the caller to publish a batch */
    volatile int pending;        /* jobs not yet finished */
    pthread_mutex_t mu;
} g_pool = {.n_workers = 0, .mu = PTHREAD_MUTEX_INITIALIZER};

static void* pool_worker(void* arg) {
    long id = (long)arg;
    unsigned long seen = 0;
    for (;;) {
        unsigned spins = 0;
        while (__atomic_load_n(&g_pool.gen, __ATOMIC_ACQUIRE) == seen) {
            if (++spins > 20000) { sched_yield(); if (spins > 40000) { usleep(200); } }
        }
        seen = __atomic_load_n(&g_pool.gen, __ATOMIC_ACQUIRE);
        if (id + 1 < g_pool.n_jobs) {
            gemv_worker(&g_pool.jobs[id + 1]);
            __atomic_fetch_sub(&g_pool.pending, 1, __ATOMIC_ACQ_REL);
        }
    }
    return NULL;
}
static void pool_ensure(int workers) {
    if (workers > OC_MAX_WORKERS) workers = OC_MAX_WORKERS;
    while (g_pool.n_workers < workers) {
        long id = g_pool.n_workers;
        pthread_create(&g_pool.tid[id], NULL, pool_worker, (void*)id);
        pthread_detach(g_pool.tid[id]);
        g_pool.n_workers++;
    }
}

int oc_gemv_q(int w_type, const void* w, size_t m, size_t k, const void* act, size_t b,
              float* out, int threads, int flags) {
    int be = oc_block_elems(w_type);
    if (be == 0 || k % (size_t)be) return -1;
    int at = oc_vec_dot_rhs_type(w_type);
    size_t len = m * b;
    if (threads < 1) threads = 1;
    /* matmul_vec.rs:45: work_len = len / thread_num; chunks_mut(work_len) gives
     * ceil(len/work_len) spans (thread_num or thread_num+1, B14) */
    size_t work_len = len / (size_t)threads;
    if (work_len == 0) { work_len = len; }
    size_t nspans = (len + work_len - 1) / work_len;
    gemv_job* jobs = (gemv_job*)calloc(nspans, sizeof(gemv_job));
    for (size_t s = 0; s < nspans; s++) {
        jobs[s] = (gemv_job){w_type, flags, (const uint8_t*)w, (const uint8_t*)act, m, k,
                             (k / (size_t)be) * oc_block_bytes(w_type),
                             (k / (size_t)oc_block_elems(at)) * oc_block_bytes(at),
                             out, s * work_len, (s + 1) * work_len < len ? (s + 1) * work_len : len};
    }
    /* thread_pool.rs:38-70: first thunk inline, others on the pool's workers, busy-wait join */
    if (nspans == 1 || nspans - 1 > OC_MAX_WORKERS) {
        for (size_t s = 0; s < nspans; s++) gemv_worker(&jobs[s]);
    } else {
        pthread_mutex_lock(&g_pool.mu);
        pool_ensure((int)nspans - 1);
        g_pool.jobs = jobs;
        g_pool.n_jobs = (int)nspans;
        __atomic_store_n(&g_pool.pending, (int)nspans - 1, __ATOMIC_RELEASE);
        __atomic_fetch_add(&g_pool.gen, 1, __ATOMIC_ACQ_REL);
        gemv_worker(&jobs[0]);
        while (__atomic_load_n(&g_pool.pending, __ATOMIC_ACQUIRE) > 0) { /* spin: thread_pool.rs:66-69 */ }
        g_pool.n_jobs = 0;
        pthread_mutex_unlock(&g_pool.mu);
    }
    free(jobs);
    return 0;
}

int oc_gemv(int w_type, const void* w, size_t m, size_t k, const float* x, size_t b,
            float* out, int threads, int flags) {
    int at = oc_vec_dot_rhs_type(w_type);
    if (at < 0) return -1;
    int abe = oc_block_elems(at);
    if (k % (size_t)abe) return -1;
    size_t act_bytes = (b * k / (size_t)abe) * oc_block_bytes(at);
    void* act = malloc(act_bytes ? act_bytes : 1);
    /* matmul_vec.rs:37-40: quantize the whole (b,k) activation once */
    int rc = oc_quantize(at, x, b * k, act);
    if (rc == 0) rc = oc_gemv_q(w_type, w, m, k, act, b, out, threads, flags);
    free(act);
    return rc;
}

/* ----------------------------------------------------------- primitives -- */
/* rms_norm.rs:32-47: sum over 32-lane chunks (reduce_sum of each chunk added to a scalar) */
void oc_rms_norm(float* x, size_t rows, size_t cols, float eps) {
    for (size_t r = 0; r < rows; r++) {
        float* v = x + r * cols;
        float sum = 0.0f;
        for (size_t c = 0; c + 32 <= cols; c += 32) {
            /* f32x32::reduce_sum = simd_reduce_add_ordered: left-to-right; the chunk
             * total is then added to the running scalar (order noise only, B18) */
            float cs = 0.0f;
            for (int l = 0; l < 32; l++) cs += v[c + l] * v[c + l];
            sum += cs;
        }
        float rms = sqrtf(sum / (float)cols + eps);
        for (size_t c = 0; c < cols; c++) v[c] /= rms;
    }
}

/* rope.rs:47-80 */
void oc_rope(float* x, size_t n_batch, size_t batch_stride, size_t head_dim, int mode, size_t pos, size_t rope_dim) {
    for (size_t bi = 0; bi < n_batch; bi++) {
        size_t seq_pos = pos + bi;                               /* rope.rs:36 */
        float* row = x + bi * batch_stride;
        size_t n_heads = batch_stride / head_dim;
        if (mode == 0) {                                          /* Llama */
            float theta_scale = powf(10000.0f, -2.0f / (float)head_dim);
            for (size_t h = 0; h < n_heads; h++) {
                float* c = row + h * head_dim;
                float theta = (float)seq_pos;
                for (size_t i = 0; i < rope_dim; i += 2) {
                    float ct = cosf(theta), st = sinf(theta);
                    theta *= theta_scale;
                    float q0 = c[i], q1 = c[i + 1];
                    c[i] = q0 * ct - q1 * st;
                    c[i + 1] = q0 * st + q1 * ct;
                }
            }
        } else {                                                  /* Neox */
            for (size_t h = 0; h < n_heads; h++) {
                float* c = row + h * head_dim;
                for (size_t i = 0; i < rope_dim / 2; i++) {
                    float fe = 2.0f * (float)i / (float)head_dim;
                    float timescale = powf(10000.0f, fe);
                    float theta = (float)seq_pos / timescale;
                    float ct = cosf(theta), st = sinf(theta);
                    float q0 = c[i], q1 = c[i + head_dim / 2];
                    c[i] = q0 * ct - q1 * st;
                    c[i + head_dim / 2] = q0 * st + q1 * ct;
                }
            }
        }
    }
}

/* softmax.rs:39-54 */
void oc_softmax(float* x, size_t rows, size_t cols, const uint16_t* lut) {
    for (size_t r = 0; r < rows; r++) {
        float* v = x + r * cols;
        float max = -INFINITY;
        for (size_t c = 0; c < cols; c++) max = rmaxf(v[c], max);
        float sum = 0.0f;
        for (size_t c = 0; c < cols; c++) {
            v[c] = exp_cached(v[c] - max, lut);
            sum += v[c];
        }
        for (size_t c = 0; c < cols; c++) v[c] /= sum;
    }
}
/* silu.rs:6-13 */
void oc_silu(float* x, size_t n, const uint16_t* lut) {
    for (size_t i = 0; i < n; i++) {
        float nexp = exp_cached(-x[i], lut);
        x[i] /= 1.0f + nexp;
    }
}
/* gelu.rs:10-15 */
void oc_gelu(float* x, size_t n, const uint16_t* lut) {
    for (size_t i = 0; i < n; i++) x[i] = h2f(lut[f2h(x[i])]);
}
/* arithmetic.rs:5-34: rhs of length 1 is a scalar, else cycled in chunks of 4 */
void oc_add(float* x, size_t n, const float* y, size_t ny) {
    if (ny == 1) { for (size_t i = 0; i < n; i++) x[i] += y[0]; return; }
    for (size_t i = 0; i + 4 <= n; i += 4)
        for (int l = 0; l < 4; l++) x[i + l] += y[(i % (ny - ny % 4)) + l];
}
/* arithmetic.rs:36-68 */
void oc_mul(float* x, size_t n, const float* y, size_t ny) {
    if (ny == 1) { for (size_t i = 0; i < n; i++) x[i] *= y[0]; return; }
    for (size_t i = 0; i + 4 <= n; i += 4)
        for (int l = 0; l < 4; l++) x[i + l] *= y[(i % (ny - ny % 4)) + l];
}

/* batch_matmul.rs:47-71: C += A*B, k innermost, C zero-initialised by alloc; kv head = bi % b_batch */
void oc_batch_matmul_f32(const float* a, const float* b, float* c,
                         size_t a_batch, size_t b_batch, size_t m, size_t k, size_t n,
                         size_t sb0, size_t sb1, size_t sb2) {
    for (size_t bi = 0; bi < a_batch; bi++)
        for (size_t mi = 0; mi < m; mi++)
            for (size_t ni = 0; ni < n; ni++) {
                float acc = 0.0f;
                for (size_t ki = 0; ki < k; ki++)
                    acc += a[bi * (m * k) + mi * k + ki] * b[(bi % b_batch) * sb0 + ki * sb1 + ni * sb2];
                c[bi * (m * n) + mi * n + ni] = acc;
            }
}
/* batch_matmul.rs:73-131: A -> f16; K path = f32-accumulated f16 dot, V path = f16 FMA accumulation;
 * kv head = bi / (a_batch / b_batch) */
void oc_batch_matmul_f16(const float* a, const uint16_t* b, float* c,
                         size_t a_batch, size_t b_batch, size_t m, size_t k, size_t n,
                         size_t sb0, size_t sb1, size_t sb2) {
    size_t bc = a_batch / b_batch;
    uint16_t* ah = (uint16_t*)malloc(a_batch * m * k * 2 + 2);
    oc_f32_to_f16(a, ah, a_batch * m * k);
    if (sb1 == 1) {
        for (size_t i = 0; i < a_batch * m * n; i++) {
            size_t ni = i % n, mi = (i / n) % m, bi = i / (m * n);
            const uint16_t* pa = ah + bi * (m * k) + mi * k;
            const uint16_t* pb = b + (bi / bc) * sb0 + ni * sb2;
            c[i] = dot_f16(pa, pb, k);
        }
    } else {
        uint16_t* tmp = (uint16_t*)calloc(a_batch * m * n + 1, 2);
        for (size_t bi = 0; bi < a_batch; bi++)
            for (size_t mi = 0; mi < m; mi++)
                for (size_t ki = 0; ki < k; ki++) {
                    uint16_t av = ah[bi * (m * k) + mi * k + ki];
                    const uint16_t* pb = b + (bi / bc) * sb0 + ki * sb1;
                    uint16_t* pc = tmp + bi * (m * n) + mi * n;
                    for (size_t ni = 0; ni < n; ni++) pc[ni] = hadd(pc[ni], hmul(pb[ni], av));  /* buf_f16.rs:152-163 */
                }
        oc_f16_to_f32(tmp, c, a_batch * m * n);
        free(tmp);
    }
    free(ah);
}

/* ------------------------------------------------------- synthetic weights -- */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
int oc_synth_blocks(int t, size_t nblocks, uint64_t seed, uint64_t tid, float scale, void* outv) {
    int n_f16 = 0, off[2] = {0, 0}, is_min[2] = {0, 0}, d32 = -1;
    switch (t) {
    case OC_Q8_0: case OC_Q4_0: case OC_Q5_0: n_f16 = 1; break;
    case OC_Q4_1: case OC_Q5_1: case OC_Q4_K: case OC_Q5_K: n_f16 = 2; off[1] = 2; is_min[1] = 1; break;
    case OC_Q2_K: n_f16 = 2; off[0] = 80; off[1] = 82; is_min[1] = 1; break;
    case OC_Q3_K: n_f16 = 1; off[0] = 108; break;
    case OC_Q6_K: n_f16 = 1; off[0] = 208; break;
    case OC_Q8_K: d32 = 0; break;
    default: return -1;
    }
    uint8_t* out = (uint8_t*)outv;
    size_t bb = oc_block_bytes(t), total = nblocks * bb;
    uint64_t key = splitmix64(seed ^ splitmix64(tid));
    for (size_t w = 0; w * 8 < total; w++) {
        uint64_t r = splitmix64(key ^ (uint64_t)w);
        for (int j = 0; j < 8 && w * 8 + j < total; j++) out[w * 8 + j] = (uint8_t)(r >> (8 * j));
    }
    for (size_t b = 0; b < nblocks; b++) {
        uint64_t r = splitmix64(key ^ 0xD1B54A32D192ED03ull ^ (uint64_t)b);
        for (int f = 0; f < n_f16; f++) {
            float u = (float)((r >> (16 * f)) & 0xFFFF) * (1.0f / 65536.0f);
            float v = scale * (0.75f + 0.5f * u);
            if (is_min[f]) v *= 0.25f;
            uint16_t h = f2h(v);
            memcpy(out + b * bb + off[f], &h, 2);
        }
        if (d32 >= 0) {
            float u = (float)(r & 0xFFFF) * (1.0f / 65536.0f);
            float v = scale * (0.75f + 0.5f * u);
            memcpy(out + b * bb + d32, &v, 4);
        }
    }
    return 0;
}

int oc_hw_threads(void) { return (int)sysconf(_SC_NPROCESSORS_ONLN); }
