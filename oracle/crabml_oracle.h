/*
 * crabml_oracle.h -- CPU restatement of crabml's quantized decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is the parity checker for the CUDA
 * backend in crabml_b200/.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path
 * (crabml_b200/) never links or calls anything in oracle/.
 *
 * Every function cites the reference file:line (relative to the crabml/crabml
 * checkout @0151f893) whose arithmetic it restates.  Parity is PINNED: the
 * restatement is checked against the reference's own known-answer tests
 * (tests/test_oracle_kats.py) and its golden generations on the tinyllamas
 * fixtures (tests/test_oracle_golden_text.py).
 */
#ifndef CRABML_ORACLE_H
#define CRABML_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* GGML type ids, crabml-core/src/gguf.rs:86-108 */
enum oc_type {
    OC_F32 = 0, OC_F16 = 1, OC_Q4_0 = 2, OC_Q4_1 = 3, OC_Q5_0 = 6, OC_Q5_1 = 7,
    OC_Q8_0 = 8, OC_Q8_1 = 9, OC_Q2_K = 10, OC_Q3_K = 11, OC_Q4_K = 12,
    OC_Q5_K = 13, OC_Q6_K = 14, OC_Q8_K = 15
};

/* flags */
#define OC_BUGCOMPAT      1  /* reproduce reference bugs B7/B9/B10 */
#define OC_ORDER_AVX2     2  /* f32 accumulation order of the reference's AVX2 kernels */

int    oc_block_elems(int type);
size_t oc_block_bytes(int type);
int    oc_vec_dot_rhs_type(int type);              /* buf/api.rs:142-159 */

void   oc_f16_to_f32(const uint16_t* src, float* dst, size_t n);
void   oc_f32_to_f16(const float* src, uint16_t* dst, size_t n);
void   oc_exp_lut(uint16_t* lut65536);             /* cpu_device.rs:108-115 */
void   oc_gelu_lut(uint16_t* lut65536);            /* cpu_device.rs:117-124, gelu.rs:17-21 */

/* BlockQ*::dequantize -- n elements (multiple of the block size) */
int    oc_dequantize(int type, const void* blocks, size_t n, float* out, int flags);

/* activation quantizers: buf_q8_0.rs:87-134, buf_q8_1.rs:90-129, buf_q8_k.rs:84-131 */
int    oc_quantize(int act_type, const float* x, size_t n, void* out);

/* vec_dot_*: one row of n elements against an activation quantized to the partner type */
float  oc_vec_dot(int w_type, const void* w, const void* act, size_t n, int flags);

/* gemv_dense_2d_2d, primitives/matmul_vec.rs:26-78.  out[b*m + i] */
int    oc_gemv(int w_type, const void* w, size_t m, size_t k, const float* x, size_t b,
               float* out, int threads, int flags);
/* same but with a pre-quantized activation (for timing the dot alone) */
int    oc_gemv_q(int w_type, const void* w, size_t m, size_t k, const void* act, size_t b,
                 float* out, int threads, int flags);

/* primitives */
void   oc_rms_norm(float* x, size_t rows, size_t cols, float eps);           /* rms_norm.rs:9-47 */
void   oc_rope(float* x, size_t n_batch, size_t batch_stride, size_t head_dim,
               int mode, size_t pos, size_t rope_dim);                      /* rope.rs:10-80 */
void   oc_softmax(float* x, size_t rows, size_t cols, const uint16_t* exp_lut); /* softmax.rs:11-57 */
void   oc_silu(float* x, size_t n, const uint16_t* exp_lut);                 /* silu.rs:6-13 */
void   oc_gelu(float* x, size_t n, const uint16_t* gelu_lut);                /* gelu.rs:10-15 */
void   oc_add(float* x, size_t n, const float* y, size_t ny);                /* arithmetic.rs:5-34 */
void   oc_mul(float* x, size_t n, const float* y, size_t ny);                /* arithmetic.rs:36-68 */
/* batch_matmul.rs:47-71 (F32 B) and :73-131 (F16 B).  A dense (ab,m,k); B strided. */
void   oc_batch_matmul_f32(const float* a, const float* b, float* c,
                           size_t a_batch, size_t b_batch, size_t m, size_t k, size_t n,
                           size_t sb0, size_t sb1, size_t sb2);
void   oc_batch_matmul_f16(const float* a, const uint16_t* b, float* c,
                           size_t a_batch, size_t b_batch, size_t m, size_t k, size_t n,
                           size_t sb0, size_t sb1, size_t sb2);
/* CPU twin of the device synthetic-weight generator (crabml_b200/csrc/repack.cu); not reference code */
int    oc_synth_blocks(int type, size_t nblocks, uint64_t seed, uint64_t tensor_id, float scale, void* out);
/* hardware info used by bench.py */
int    oc_hw_threads(void);

#ifdef __cplusplus
}
#endif
#endif
