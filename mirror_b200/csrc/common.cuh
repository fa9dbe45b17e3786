// Shared host/device helpers for the mirror_b200 kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/mirror_b200.h"

namespace mb {

typedef __nv_bfloat16 bf16;

// ---- error reporting (thread-local, SURVEY.md §8b "Error convention") ----
void set_error(const char* fmt, ...);
enum { MB_ERR_ARG = -1, MB_ERR_ALIGN = -2, MB_ERR_DRIVER = -3, MB_ERR_UNSUPPORTED = -4 };

#define MB_CHECK_ARG(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      mb::set_error(__VA_ARGS__);    \
      return mb::MB_ERR_ARG;         \
    }                                \
  } while (0)

#define MB_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      mb::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return (int)e__;                                                             \
    }                                                                              \
  } while (0)

#define MB_LAUNCH_CHECK() MB_CUDA(cudaGetLastError())

// per-device caches (a process may drive several devices: attributes set on one device do not carry over)
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
inline int num_sms() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (!n[dev]) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}
// Programmatic dependent launch (PDL), opt-in (MIRROR_B200_PDL=1).  A step is ~630 dependent launches; kernels that call
// pdl_wait() before their first global-memory access may be launched with the programmatic-serialization attribute: their CTAs
// are scheduled while the previous grid drains (barrier init, TMEM allocation, descriptor prefetch overlap its tail), and
// griddepcontrol.wait holds them until that grid has completed and flushed its memory.  Measured on the full step (round 2):
// 49.98 ms with PDL on the GEMM and flash launches against 49.18 ms without (same box) -- the persistent 1-CTA-per-SM kernels
// cannot co-reside with their predecessor, so nothing overlaps and the early scheduling only adds contention.  Default: off.
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MIRROR_B200_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// `static DeviceOnce once; if (once.first()) cudaFuncSetAttribute(...)`: true the first time per device (benign race:
// the attribute set is idempotent)
struct DeviceOnce {
  bool done[kMaxDevices] = {};
  bool first() {
    const int dev = current_device();
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

// ---- stateless counter-based uniform in [0,1): keep-mask of the fused dropout.  32-bit "lowbias32" finaliser over the
// element index mixed with the 64-bit seed: ~10 integer ops per element (the epilogues evaluate it for every output).
__host__ // 2^x as ONE MUFU.EX2 (flush-to-zero; exp2f without .ftz costs a range check and two extra multiplies per element)
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float hash_u01(uint64_t seed, uint64_t idx) {
  uint32_t h = (uint32_t)idx + (uint32_t)seed * 0x9E3779B1u + (uint32_t)(idx >> 32) * 0x85EBCA77u;
  h ^= (uint32_t)(seed >> 32);
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

// Graph-safe dropout: a CUDA graph bakes the by-value seed of every captured launch, so a replayed step would repeat its
// masks.  mirror_set_dropout_epoch installs a DEVICE counter; while it is set, every dropout launch mixes *epoch into its
// seed at run time (the owner of the counter bumps it once per replay).  NULL (the default) = seeds are used as passed.
const unsigned long long* drop_epoch_ptr();
__device__ __forceinline__ uint64_t epoch_seed(uint64_t seed, const unsigned long long* epoch) {
  return epoch ? seed ^ (*epoch * 0x9E3779B97F4A7C15ull) : seed;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum / max for blockDim.x <= 1024 (multiple of 32); `sh` holds >= 32 floats
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) v = warp_sum(v);
  if (threadIdx.x == 0) sh[0] = v;
  __syncthreads();
  return sh[0];
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : -INFINITY;
  if (w == 0) v = warp_max(v);
  if (threadIdx.x == 0) sh[0] = v;
  __syncthreads();
  return sh[0];
}

// ---- GEMM epilogue shared by the tcgen05 kernel and the SIMT cross-check kernel ----
struct Epi {
  int M, N, batch1;
  float alpha, diag;
  const float* bias;
  int act;
  float drop_p, drop_scale;
  uint64_t drop_seed;
  const unsigned long long* drop_epoch;  // optional device counter mixed into drop_seed (graph replay), see epoch_seed
  const void* res;
  const bf16* res2;  // second residual, bf16, same element strides as res
  int res_is_bf16;
  float gamma, gamma2;
  long long ldr, r_bs1, r_bs2;
  int res_row_div;  // residual row = output row / res_row_div (>= 1)
  float beta;
  float* o32;
  long long ldc32, c32_bs1, c32_bs2;
  bf16* o16;
  long long ldc16, c16_bs1, c16_bs2;
  int atomic;  // split-K partial: atomically add alpha*acc into o32
  int mode;    // MIRROR_GEMM_*: fused row-softmax epilogues
  float* stats;
  int nparts;  // partial slots per row (two per N tile)
};

__device__ __forceinline__ void epi_store_scalar(const Epi& e, float acc, int b1, int b2, int row, int col) {
  if (e.atomic) {
    atomicAdd(e.o32 + b2 * e.c32_bs2 + b1 * e.c32_bs1 + (long long)row * e.ldc32 + col, e.alpha * acc);
    return;
  }
  float v = e.alpha * acc;
  if (row == col) v += e.diag;
  if (e.bias) v += e.bias[col];
  if (e.act == MIRROR_ACT_RELU) v = fmaxf(v, 0.f);
  else if (e.act == MIRROR_ACT_GELU) v = gelu_erf(v);
  if (e.drop_p > 0.f) {
    const uint64_t idx = ((uint64_t)(b2 * e.batch1 + b1) * e.M + row) * e.N + col;
    v = hash_u01(epoch_seed(e.drop_seed, e.drop_epoch), idx) >= e.drop_p ? v * e.drop_scale : 0.f;
  }
  if (e.res) {
    const long long off = b2 * e.r_bs2 + b1 * e.r_bs1 + (long long)(row / e.res_row_div) * e.ldr + col;
    v += e.gamma * (e.res_is_bf16 ? __bfloat162float(((const bf16*)e.res)[off]) : ((const float*)e.res)[off]);
    if (e.res2) v += e.gamma2 * __bfloat162float(e.res2[off]);
  }
  if (e.o32) {
    float* p = e.o32 + b2 * e.c32_bs2 + b1 * e.c32_bs1 + (long long)row * e.ldc32 + col;
    if (e.beta != 0.f) v += e.beta * *p;
    *p = v;
  }
  if (e.o16) e.o16[b2 * e.c16_bs2 + b1 * e.c16_bs1 + (long long)row * e.ldc16 + col] = __float2bfloat16(v);
}

}  // namespace mb
