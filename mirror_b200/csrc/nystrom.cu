// Nyström-attention specific HBM-bound kernels (call site models/mirror.py:299-312; algorithm of the
// un-vendored nystrom_attention package):
//  - 33-tap depthwise "value residual" convolution along the token axis, forward / data-grad / weight-grad
//  - Moore-Penrose iteration set-up: global maxima of the row/column abs-sums of attn2, z0 = attn2^T / (c*r),
//    and the backward of that set-up (including the gradient that flows through the two global maxima).
#include "common.cuh"

namespace mb {
namespace {

constexpr int TAPS = 33;

// Depthwise 33-tap FIR along the token axis.  Each thread owns one bf16 channel pair and a strip of TT consecutive
// tokens: the TT+32 inputs of the strip are loaded once into registers and reused by all 33 taps (3 loads per
// output pair instead of 33), threads of a warp cover 64 consecutive channels (128-byte coalesced rows).
//   FWD: out16[b,t,c] = sum_j w[h(c),j] * v[b,t+j-16,c]      (v = value slot of qkv [B,n,3E], column 2E+c)
//   BWD: dv16[b,t,c]   = sum_j w[h(c),j] * dout[b,t-j+16,c]        (data gradient: the same FIR with flipped taps)
constexpr int TT = 16;

template <bool BWD>
__global__ void __launch_bounds__(128)
res_conv_kernel(const bf16* __restrict__ src, long long src_ld, int src_col0, const float* __restrict__ w, int n, int E, int d,
                bf16* __restrict__ dst16, long long dst_ld, int dst_col0) {
  __shared__ float sw[8 * TAPS];
  for (int i = threadIdx.x; i < 8 * TAPS; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int c = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (c >= E) return;
  const int t0 = blockIdx.y * TT;
  const long long b = blockIdx.z;
  const float* wh = sw + (c / d) * TAPS;
  float2 in[TT + TAPS - 1];
#pragma unroll
  for (int i = 0; i < TT + TAPS - 1; ++i) {
    const int tt = t0 + i - 16;
    in[i] = (tt >= 0 && tt < n)
                ? __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(src + (b * n + tt) * src_ld + src_col0 + c))
                : make_float2(0.f, 0.f);
  }
  float2 acc[TT];
#pragma unroll
  for (int i = 0; i < TT; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < TAPS; ++j) {
    const float wj = BWD ? wh[TAPS - 1 - j] : wh[j];  // BWD: out[t] = sum_j w[j] in[t-j+16] = sum_j' w[32-j'] in[t+j'-16]
#pragma unroll
    for (int i = 0; i < TT; ++i) {
      acc[i].x += wj * in[i + j].x;
      acc[i].y += wj * in[i + j].y;
    }
  }
#pragma unroll
  for (int i = 0; i < TT; ++i) {
    const int t = t0 + i;
    if (t >= n) break;
    *reinterpret_cast<__nv_bfloat162*>(dst16 + (b * n + t) * dst_ld + dst_col0 + c) = __floats2bfloat162_rn(acc[i].x, acc[i].y);
  }
}

// dw[h,j] += sum_{b,t,c in head h} dout[b,t,c] * v[b,t+j-16,c].  Same strip decomposition, but a thread keeps its 33
// partial sums in registers across ALL token strips of its chunk (grid.y chunks per slide); only then are they reduced
// (warp shuffle -> shared atomics -> one global atomic per (head, tap) and CTA).  One global atomic per strip and tap
// (the first version) serialised 7 M atomics on 264 addresses and took 2.8 ms.
__global__ void __launch_bounds__(128)
res_conv_wgrad_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ qkv, int n, int E, int d, float* __restrict__ dw,
                      int strips_per_block) {
  __shared__ float sacc[8 * TAPS];
  for (int i = threadIdx.x; i < 8 * TAPS; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int c = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const long long b = blockIdx.z;
  // warp-uniform fast path: all 32 lanes active and inside one head -> shuffle-reduce, one shared atomic per tap
  const int c_first = 2 * (blockIdx.x * blockDim.x + (threadIdx.x & ~31));
  const bool warp_one_head = (c_first + 62 < E) && (c_first / d == (c_first + 62) / d);
  if (c < E) {
    float acc[TAPS];
#pragma unroll
    for (int j = 0; j < TAPS; ++j) acc[j] = 0.f;
    const int s0 = blockIdx.y * strips_per_block;
    for (int s = s0; s < s0 + strips_per_block; ++s) {
      const int t0 = s * TT;
      if (t0 >= n) break;
      float2 vin[TT + TAPS - 1];
#pragma unroll
      for (int i = 0; i < TT + TAPS - 1; ++i) {
        const int tt = t0 + i - 16;
        vin[i] = (tt >= 0 && tt < n)
                     ? __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(qkv + (b * n + tt) * 3LL * E + 2 * E + c))
                     : make_float2(0.f, 0.f);
      }
      float2 g[TT];
#pragma unroll
      for (int i = 0; i < TT; ++i) {
        const int t = t0 + i;
        g[i] = t < n ? __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dout + (b * n + t) * (long long)E + c))
                     : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < TAPS; ++j) {
#pragma unroll
        for (int i = 0; i < TT; ++i) acc[j] += g[i].x * vin[i + j].x + g[i].y * vin[i + j].y;
      }
    }
    const int h = c / d;
#pragma unroll
    for (int j = 0; j < TAPS; ++j) {
      float a = acc[j];
      if (warp_one_head) {
        a = warp_sum(a);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sacc[h * TAPS + j], a);
      } else {
        atomicAdd(&sacc[h * TAPS + j], a);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * TAPS; i += blockDim.x)
    if (sacc[i] != 0.f) atomicAdd(dw + i, sacc[i]);
}

// ---- pinv set-up.  a2: [BH, m, m] f32 (row softmax).  scal: [0]=max row abs-sum, [1]=max col abs-sum (as ordered
// uint64 keys: float bits << 32 | ~index so that atomicMax also yields the FIRST arg-max).
__device__ __forceinline__ unsigned long long pack_key(float v, unsigned idx) {
  return ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}
__global__ void pinv_scale_kernel(const float* __restrict__ a2, int m, unsigned long long* __restrict__ keys) {
  const int bh = blockIdx.x;
  const float* a = a2 + (long long)bh * m * m;
  unsigned long long best_r = 0, best_c = 0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {  // thread i: row-sum of row i (strided reads, L1-resident tile)
    float rs = 0.f, cs = 0.f;
    for (int j = 0; j < m; ++j) {
      rs += fabsf(a[(long long)i * m + j]);
      cs += fabsf(a[(long long)j * m + i]);
    }
    const unsigned long long kr = pack_key(rs, (unsigned)(bh * m + i)), kc = pack_key(cs, (unsigned)(bh * m + i));
    best_r = kr > best_r ? kr : best_r;
    best_c = kc > best_c ? kc : best_c;
  }
  atomicMax(keys, best_r);
  atomicMax(keys + 1, best_c);
}
__device__ __forceinline__ float key_val(unsigned long long k) { return __uint_as_float((unsigned)(k >> 32)); }
__device__ __forceinline__ unsigned key_idx(unsigned long long k) { return 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFu); }

// z0[bh,i,j] = a2[bh,j,i] / (c*r)  via a 32x32 shared tile
__global__ void pinv_init_kernel(const float* __restrict__ a2, int m, const unsigned long long* __restrict__ keys,
                                 float* __restrict__ z32, bf16* __restrict__ z16) {
  __shared__ float tile[32][33];
  const float inv = 1.f / (key_val(keys[0]) * key_val(keys[1]));
  const long long base = (long long)blockIdx.z * m * m;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r, j = j0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < m && j < m) ? a2[base + (long long)i * m + j] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int j = j0 + r, i = i0 + threadIdx.x;  // output row j, column i
    if (i < m && j < m) {
      const float v = tile[threadIdx.x][r] * inv;
      if (z32) z32[base + (long long)j * m + i] = v;
      if (z16) z16[base + (long long)j * m + i] = __float2bfloat16(v);
    }
  }
}

// backward of z0 = x^T / D, D = c*r:
//   s = <g_z0, z0>  (whole tensor)      dD = -s / D      dc = dD*r   dr = dD*c
//   gx[bh,i,j] (+)= g_z0[bh,j,i]/D + [row (bh,i) is the arg-max row]*dc + [col (bh,j) is the arg-max col]*dr
__global__ void dot_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float* __restrict__ out) {
  __shared__ float sh[32];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += a[i] * b[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
__global__ void pinv_init_bwd_kernel(const float* __restrict__ gz0, int m, const unsigned long long* __restrict__ keys,
                                     const float* __restrict__ dotp, float* __restrict__ gx, int accumulate) {
  __shared__ float tile[32][33];
  const float c = key_val(keys[0]), r = key_val(keys[1]);
  const float D = c * r, inv = 1.f / D;
  const float dD = -dotp[0] * inv;
  const float dc = dD * r, dr = dD * c;
  const unsigned arg_row = key_idx(keys[0]), arg_col = key_idx(keys[1]);
  const int bh = blockIdx.z;
  const long long base = (long long)bh * m * m;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;  // output tile rows i0.., cols j0..
  for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
    const int j = j0 + rr, i = i0 + threadIdx.x;  // read g_z0[j, i]
    tile[rr][threadIdx.x] = (i < m && j < m) ? gz0[base + (long long)j * m + i] : 0.f;
  }
  __syncthreads();
  for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
    const int i = i0 + rr, j = j0 + threadIdx.x;
    if (i < m && j < m) {
      float v = tile[threadIdx.x][rr] * inv;
      if ((unsigned)(bh * m + i) == arg_row) v += dc;
      if ((unsigned)(bh * m + j) == arg_col) v += dr;
      float* p = gx + base + (long long)i * m + j;
      *p = accumulate ? *p + v : v;
    }
  }
}

}  // namespace
}  // namespace mb

using namespace mb;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

static int ew_grid(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

extern "C" int mirror_res_conv_fwd(const void* qkv, const float* w, int32_t B, int32_t n, int32_t E, void* out_bf16,
                                   mirror_stream_t stream) {
  MB_CHECK_ARG(qkv && w && out_bf16 && B > 0 && n > 0 && E % 16 == 0, "res_conv_fwd: bad args");
  dim3 grid((E / 2 + 127) / 128, (n + TT - 1) / TT, B);
  res_conv_kernel<false><<<grid, 128, 0, STREAM>>>(reinterpret_cast<const bf16*>(qkv), 3LL * E, 2 * E, w, n, E, E / 8,
                                                  reinterpret_cast<bf16*>(out_bf16), E, 0);
  MB_LAUNCH_CHECK();
  return 0;
}
extern "C" int mirror_res_conv_bwd(const void* dout_bf16, const void* qkv, const float* w, int32_t B, int32_t n, int32_t E,
                                   void* dv_bf16, float* dw, mirror_stream_t stream) {
  MB_CHECK_ARG(dout_bf16 && qkv && w && dv_bf16 && dw && B > 0 && n > 0 && E % 16 == 0, "res_conv_bwd: bad args");
  dim3 grid((E / 2 + 127) / 128, (n + TT - 1) / TT, B);
  res_conv_kernel<true><<<grid, 128, 0, STREAM>>>(reinterpret_cast<const bf16*>(dout_bf16), E, 0, w, n, E, E / 8,
                                                 reinterpret_cast<bf16*>(dv_bf16), E, 0);
  MB_LAUNCH_CHECK();
  const int strips = (n + TT - 1) / TT;
  int chunks = (2 * num_sms() + 3 * B - 1) / (3 * B);  // ~2 waves of CTAs in total
  if (chunks < 1) chunks = 1;
  if (chunks > strips) chunks = strips;
  const int spb = (strips + chunks - 1) / chunks;
  res_conv_wgrad_kernel<<<dim3(grid.x, (strips + spb - 1) / spb, B), 128, 0, STREAM>>>(
      reinterpret_cast<const bf16*>(dout_bf16), reinterpret_cast<const bf16*>(qkv), n, E, E / 8, dw, spb);
  MB_LAUNCH_CHECK();
  return 0;
}

/* scratch: 4 x 8 bytes: [0],[1] ordered max keys, [2] float dot accumulator (zeroed here / by pinv_init_bwd) */
extern "C" int mirror_pinv_init(const float* a2, int32_t BH, int32_t m, void* scratch32, float* z_f32, void* z_bf16,
                                mirror_stream_t stream) {
  MB_CHECK_ARG(a2 && scratch32 && (z_f32 || z_bf16) && BH > 0 && m > 0, "pinv_init: bad args");
  MB_CUDA(cudaMemsetAsync(scratch32, 0, 32, STREAM));
  pinv_scale_kernel<<<BH, 128, 0, STREAM>>>(a2, m, reinterpret_cast<unsigned long long*>(scratch32));
  MB_LAUNCH_CHECK();
  dim3 grid((m + 31) / 32, (m + 31) / 32, BH), block(32, 8);
  pinv_init_kernel<<<grid, block, 0, STREAM>>>(a2, m, reinterpret_cast<const unsigned long long*>(scratch32), z_f32,
                                               reinterpret_cast<bf16*>(z_bf16));
  MB_LAUNCH_CHECK();
  return 0;
}
extern "C" int mirror_pinv_init_bwd(const float* gz0, const float* z0_f32, int32_t BH, int32_t m, void* scratch32, float* gx,
                                    int32_t accumulate, mirror_stream_t stream) {
  MB_CHECK_ARG(gz0 && z0_f32 && scratch32 && gx && BH > 0 && m > 0, "pinv_init_bwd: bad args");
  float* dotp = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch32) + 16);
  MB_CUDA(cudaMemsetAsync(dotp, 0, 4, STREAM));
  const long long n = (long long)BH * m * m;
  dot_kernel<<<ew_grid(n, 256 * 4), 256, 0, STREAM>>>(gz0, z0_f32, n, dotp);
  MB_LAUNCH_CHECK();
  dim3 grid((m + 31) / 32, (m + 31) / 32, BH), block(32, 8);
  pinv_init_bwd_kernel<<<grid, block, 0, STREAM>>>(gz0, m, reinterpret_cast<const unsigned long long*>(scratch32), dotp, gx,
                                                   accumulate);
  MB_LAUNCH_CHECK();
  return 0;
}
