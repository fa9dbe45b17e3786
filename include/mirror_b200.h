/*
 * mirror_b200 C ABI — the drop-in boundary of the B200-native MIRROR pre-training path.
 *
 * The reference (TianyiFranklinWang/MIRROR) is pure Python/PyTorch: every device
 * operation of its hot path is a torch library call.  This header
 * declares the entry points that replace those call sites; each comment cites the
 * reference lines (relative to the reference checkout) the entry point stands in for.
 *
 * Conventions:
 *  - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *    (PyTorch's caching allocator); the library never allocates, frees or retains them.
 *  - every call enqueues work on the caller's `stream` and returns immediately;
 *    no entry point synchronises the device or uses the legacy default stream.
 *  - return value 0 = success, otherwise a negative library code or a cudaError_t;
 *    `mirror_last_error()` returns a thread-local message.  No exceptions, no exit().
 *  - re-entrant and thread-safe: backward entry points are called from PyTorch's
 *    autograd thread concurrently with forward calls from the main thread.
 *  - "bf16" buffers are raw 16-bit bfloat16; "f32" are IEEE float.
 */
#ifndef MIRROR_B200_H
#define MIRROR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mirror_stream_t; /* cudaStream_t */

const char* mirror_last_error(void);
int mirror_abi_version(void);
/* 1 when the current device is compute capability 10.x (tcgen05/TMEM present). */
int mirror_device_supported(void);

/* ------------------------------------------------------------------------------------------------
 * Batched GEMM on 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA-fed, persistent).
 *
 *   for every (b2,b1):  acc[M,N] = sum_k A[m,k] * B[n,k]            (bf16 x bf16 -> fp32)
 *   v = alpha*acc + bias[n];  v = act(v);  v = dropout(v);
 *   v += gamma * R[m,n] + beta * out_f32_old[m,n];   out_f32 = v;  out_bf16 = bf16(v)
 *
 * Replaces every nn.Linear / einsum / `@` of the hot path:
 *   models/mirror.py:346,654 (_fc1), :594-605 heads, :70-74,:100 RNA qkv/proj, timm Mlp
 *   (:138,:217,:815), :823-827 style heads/prototypes; nystrom_attention to_qkv / sim1-3 /
 *   pinv iterations / attn@v products / to_out (call site models/mirror.py:299-312);
 *   losses/mirror_loss.py:39-40 logits; and all of their autograd backward GEMMs.
 *
 * A is logically [M,K], B is logically [N,K] (C = A * B^T).  Each operand is either
 * "K-major" (K contiguous, `ld` = elements between consecutive m / n) or "MN-major"
 * (m / n contiguous, `ld` = elements between consecutive k), so all of NN/NT/TN/TT are covered.
 * Alignment: base pointers 16 B; ld and batch strides multiples of 8 elements.
 * ---------------------------------------------------------------------------------------------- */
enum { MIRROR_ACT_NONE = 0, MIRROR_ACT_RELU = 1, MIRROR_ACT_GELU = 2 };

typedef struct {
  const void* a;
  const void* b;
  int32_t a_mn_major, b_mn_major;
  int64_t lda, ldb;
  int64_t a_bs1, a_bs2, b_bs1, b_bs2; /* batch strides in elements */
  int32_t M, N, K, batch1, batch2;
  float alpha;
  const float* bias; /* [N] or NULL */
  int32_t act;
  float drop_p; /* 0 = off.  keep(idx) = hash(drop_seed, idx) >= p, idx = ((b2*batch1+b1)*M+m)*N+n */
  uint64_t drop_seed;
  const void* res; /* residual R, NULL = none */
  int32_t res_is_bf16;
  float gamma;
  int64_t ldr, r_bs1, r_bs2;
  float beta;
  float* out_f32; /* may be NULL */
  int64_t ldc32, c32_bs1, c32_bs2;
  void* out_bf16; /* may be NULL */
  int64_t ldc16, c16_bs1, c16_bs2;
  int32_t split_k; /* >1: K is split over CTAs and fp32 partials are atomically added into out_f32
                      (caller pre-zeroes it); only alpha is applied */
} mirror_gemm_args;

int mirror_gemm_bf16(const mirror_gemm_args* args, mirror_stream_t stream);
/* Same contract on CUDA cores (one thread per output element).  Test/diagnostic tool used by the
 * GPU unit tests to cross-check the tensor-core kernel at sizes where a host reference is slow. */
int mirror_gemm_bf16_simt(const mirror_gemm_args* args, mirror_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MIRROR_B200_H */
