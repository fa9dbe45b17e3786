// Nyström-attention specific HBM-bound kernels (call site models/mirror.py:299-312; algorithm of the
// un-vendored nystrom_attention package):
//  - 33-tap depthwise "value residual" convolution along the token axis, forward / data-grad / weight-grad
//  - Moore-Penrose iteration set-up: global maxima of the row/column abs-sums of attn2, z0 = attn2^T / (c*r),
//    and the backward of that set-up (including the gradient that flows through the two global maxima).
#include "common.cuh"

namespace mb {
namespace {

constexpr int TAPS = 33;

// Depthwise 33-tap FIR along the token axis.  Each thread owns one bf16 channel pair (a warp covers 64 consecutive
// channels = 128-byte rows) and walks a CHUNK of consecutive TT-token strips with a sliding register window: the 48
// inputs a strip needs are the previous strip's last 32 plus 16 new ones, and those 16 are fetched (still packed) while
// the 33 x 16 x 2 FMAs of the current strip issue -- the load latency hides inside the warp instead of needing occupancy
// (the first version loaded all 48 inputs per strip up front: 0.7 IPC per SM, 0.72 ms per call at the benchmark shape).
//   FWD: out16[b,t,c] = sum_j w[h(c),j] * v[b,t+j-16,c]      (v = value slot of qkv [B,n,3E], column 2E+c)
//   BWD: dv16[b,t,c]   = sum_j w[h(c),j] * dout[b,t-j+16,c]        (data gradient: the same FIR with flipped taps)
constexpr int TT = 16;
constexpr int HALO = TAPS - 1;  // 32

__device__ __forceinline__ uint32_t ld_pair(const bf16* base, long long ld, long long b, int n, int tt) {
  return (tt >= 0 && tt < n) ? *reinterpret_cast<const uint32_t*>(base + (b * n + tt) * ld) : 0u;
}
__device__ __forceinline__ float2 unpack_pair(uint32_t u) {  // bf16 -> f32 is a shift
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

template <bool BWD>
__global__ void __launch_bounds__(128)
res_conv_kernel(const bf16* __restrict__ src, long long src_ld, int src_col0, const float* __restrict__ w, int n, int E, int d,
                bf16* __restrict__ dst16, long long dst_ld, int dst_col0, int strips_per_chunk) {
  __shared__ float sw[8 * TAPS];
  for (int i = threadIdx.x; i < 8 * TAPS; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int c = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (c >= E) return;
  const long long b = blockIdx.z;
  const float* wh = sw + (c / d) * TAPS;
  const bf16* sp = src + src_col0 + c;
  const int strips = (n + TT - 1) / TT;
  const int s0 = blockIdx.y * strips_per_chunk, s1 = min(strips, s0 + strips_per_chunk);
  if (s0 >= s1) return;
  float2 win[TT + HALO];  // tokens t0-16 .. t0+31 of the current strip
#pragma unroll
  for (int i = 0; i < HALO; ++i) win[i] = unpack_pair(ld_pair(sp, src_ld, b, n, s0 * TT + i - 16));
  uint32_t nxt[TT];
#pragma unroll
  for (int i = 0; i < TT; ++i) nxt[i] = ld_pair(sp, src_ld, b, n, s0 * TT + HALO + i - 16);
  for (int s = s0; s < s1; ++s) {
    const int t0 = s * TT;
#pragma unroll
    for (int i = 0; i < TT; ++i) win[HALO + i] = unpack_pair(nxt[i]);
    if (s + 1 < s1) {  // the next strip's new inputs travel while this strip computes
#pragma unroll
      for (int i = 0; i < TT; ++i) nxt[i] = ld_pair(sp, src_ld, b, n, t0 + TT + HALO + i - 16);
    }
    float2 acc[TT];
#pragma unroll
    for (int i = 0; i < TT; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < TAPS; ++j) {
      const float wj = BWD ? wh[TAPS - 1 - j] : wh[j];  // BWD: out[t] = sum_j w[j] in[t-j+16] = sum_j' w[32-j'] in[t+j'-16]
#pragma unroll
      for (int i = 0; i < TT; ++i) {
        acc[i].x += wj * win[i + j].x;
        acc[i].y += wj * win[i + j].y;
      }
    }
#pragma unroll
    for (int i = 0; i < TT; ++i) {
      const int t = t0 + i;
      if (t < n) *reinterpret_cast<__nv_bfloat162*>(dst16 + (b * n + t) * dst_ld + dst_col0 + c) = __floats2bfloat162_rn(acc[i].x, acc[i].y);
    }
#pragma unroll
    for (int i = 0; i < HALO; ++i) win[i] = win[i + TT];
  }
}

// dw[h,j] += sum_{b,t,c in head h} dout[b,t,c] * v[b,t+j-16,c].  The same sliding window over v plus the strip's 16
// gradients, both prefetched one strip ahead; a thread keeps its 33 partial sums in registers across ALL strips of its
// chunk and only then reduces them (warp shuffle -> shared atomics -> one global atomic per (head, tap) and CTA).
__global__ void __launch_bounds__(128)
res_conv_wgrad_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ qkv, int n, int E, int d, float* __restrict__ dw,
                      int strips_per_chunk) {
  __shared__ float sacc[8 * TAPS];
  for (int i = threadIdx.x; i < 8 * TAPS; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int c = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const long long b = blockIdx.z;
  // warp-uniform fast path: all 32 lanes active and inside one head -> shuffle-reduce, one shared atomic per tap
  const int c_first = 2 * (blockIdx.x * blockDim.x + (threadIdx.x & ~31));
  const bool warp_one_head = (c_first + 62 < E) && (c_first / d == (c_first + 62) / d);
  const int strips = (n + TT - 1) / TT;
  const int s0 = blockIdx.y * strips_per_chunk, s1 = min(strips, s0 + strips_per_chunk);
  if (c < E && s0 < s1) {
    const bf16* vp = qkv + 2 * E + c;
    const bf16* gp = dout + c;
    float acc[TAPS];
#pragma unroll
    for (int j = 0; j < TAPS; ++j) acc[j] = 0.f;
    float2 win[TT + HALO];
#pragma unroll
    for (int i = 0; i < HALO; ++i) win[i] = unpack_pair(ld_pair(vp, 3LL * E, b, n, s0 * TT + i - 16));
    uint32_t nxt[TT], gn[TT];
#pragma unroll
    for (int i = 0; i < TT; ++i) {
      nxt[i] = ld_pair(vp, 3LL * E, b, n, s0 * TT + HALO + i - 16);
      gn[i] = ld_pair(gp, E, b, n, s0 * TT + i);
    }
    for (int s = s0; s < s1; ++s) {
      const int t0 = s * TT;
      float2 g[TT];
#pragma unroll
      for (int i = 0; i < TT; ++i) {
        win[HALO + i] = unpack_pair(nxt[i]);
        g[i] = unpack_pair(gn[i]);
      }
      if (s + 1 < s1) {
#pragma unroll
        for (int i = 0; i < TT; ++i) {
          nxt[i] = ld_pair(vp, 3LL * E, b, n, t0 + TT + HALO + i - 16);
          gn[i] = ld_pair(gp, E, b, n, t0 + TT + i);
        }
      }
#pragma unroll
      for (int j = 0; j < TAPS; ++j) {
#pragma unroll
        for (int i = 0; i < TT; ++i) acc[j] += g[i].x * win[i + j].x + g[i].y * win[i + j].y;
      }
#pragma unroll
      for (int i = 0; i < HALO; ++i) win[i] = win[i + TT];
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

// The same weight gradient on the tensor cores (head_dim % 16 == 0).  For one (slide, head) and a block of 64 tokens,
// P = D V_win^T (D = dO block [64 x d], V_win = the 96 value rows t0-16 .. t0+79) holds every product the 33 lags need:
// dw[j] = sum_t P[t, t + j] (column index relative to the window).  Each warp owns 16 rows and only the 48-column band
// of P around its diagonal: 6 x (d/16) mma.sync.m16n8k16 per 16 tokens instead of 16 x d x 33 FMAs.  An accumulator
// fragment element always sits on the same lag (column - row is fixed per thread, n-block and element), so the sums over
// all token blocks of the chunk stay in 24 registers per thread and are binned by lag once at the end.
constexpr int WG_TB = 64, WG_WIN = WG_TB + 32, WG_MAXD = 128;
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(128)
res_conv_wgrad_mma_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ qkv, int n, int E, int d, float* __restrict__ dw,
                          int blocks_per_chunk) {
  extern __shared__ __align__(16) unsigned char wg_smem[];
  const int pitch = d + 8;  // bf16 elements per staged row: 16-byte aligned rows, conflict-free fragment loads for d = 96
  bf16* Ds = reinterpret_cast<bf16*>(wg_smem);
  bf16* Vs = Ds + WG_TB * pitch;
  __shared__ float sacc[TAPS];
  if (threadIdx.x < TAPS) sacc[threadIdx.x] = 0.f;
  const int h = blockIdx.y;
  const long long b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int nblocks = (n + WG_TB - 1) / WG_TB;
  const int blk0 = blockIdx.x * blocks_per_chunk, blk1 = min(nblocks, blk0 + blocks_per_chunk);
  const int c16 = d / 8;  // 16-byte pieces per row
  const bf16* dbase = dout + b * n * (long long)E + h * d;
  const bf16* vbase = qkv + b * n * 3LL * E + 2 * E + h * d;
  float acc[6][4];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
  const int r0 = warp * 16;
  for (int blk = blk0; blk < blk1; ++blk) {
    const int t0 = blk * WG_TB;
    __syncthreads();  // the previous block's fragments have been read
    for (int i = threadIdx.x; i < WG_TB * c16; i += blockDim.x) {
      const int r = i / c16, p8 = (i - r * c16) * 8;
      const int t = t0 + r;
      const uint4 v = t < n ? *reinterpret_cast<const uint4*>(dbase + (long long)t * E + p8) : make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(Ds + r * pitch + p8) = v;
    }
    for (int i = threadIdx.x; i < WG_WIN * c16; i += blockDim.x) {
      const int r = i / c16, p8 = (i - r * c16) * 8;
      const int t = t0 - 16 + r;
      const uint4 v = (t >= 0 && t < n) ? *reinterpret_cast<const uint4*>(vbase + (long long)t * 3 * E + p8) : make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(Vs + r * pitch + p8) = v;
    }
    __syncthreads();
    for (int ks = 0; ks < d / 16; ++ks) {
      const bf16* ap = Ds + (r0 + g) * pitch + ks * 16 + 2 * tg;
      const uint32_t a0 = *reinterpret_cast<const uint32_t*>(ap), a1 = *reinterpret_cast<const uint32_t*>(ap + 8 * pitch);
      const uint32_t a2 = *reinterpret_cast<const uint32_t*>(ap + 8), a3 = *reinterpret_cast<const uint32_t*>(ap + 8 * pitch + 8);
#pragma unroll
      for (int nb = 0; nb < 6; ++nb) {  // window row of column (nb*8 + g) of this warp's band: r0 + nb*8 + g
        const bf16* bp = Vs + (r0 + nb * 8 + g) * pitch + ks * 16 + 2 * tg;
        mma_bf16_16816(acc[nb], a0, a1, a2, a3, *reinterpret_cast<const uint32_t*>(bp), *reinterpret_cast<const uint32_t*>(bp + 8));
      }
    }
  }
  __syncthreads();  // sacc zeroed by all before anyone adds
#pragma unroll
  for (int nb = 0; nb < 6; ++nb)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = nb * 8 + 2 * tg + (e & 1) - g - ((e & 2) ? 8 : 0);  // lag index = column - row inside the band
      if (j >= 0 && j < TAPS) atomicAdd(&sacc[j], acc[nb][e]);
    }
  __syncthreads();
  if (threadIdx.x < TAPS && sacc[threadIdx.x] != 0.f) atomicAdd(dw + h * TAPS + threadIdx.x, sacc[threadIdx.x]);
}

// The FIR itself on the tensor cores (head_dim % 16 == 0): out[t, c] = sum_k T[t, k] * x[t0-16+k, c] with the 16 x 48 band
// Toeplitz matrix T[r, k] = w[k - r] (data gradient: flipped taps) as the A operand -- constant per head, so its
// fragments (split into bf16 hi + lo to keep the fp32 taps) live in registers -- and the staged token window as B
// (row-major tokens x channels, read with ldmatrix.trans).  6 mma.sync.m16n8k16 per 16 tokens x 8 channels.
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(smem_addr));
}
template <bool BWD>
__global__ void __launch_bounds__(128)
res_conv_mma_kernel(const bf16* __restrict__ src, long long src_ld, int src_col0, const float* __restrict__ w, int n, int E, int d,
                    bf16* __restrict__ dst16, int blocks_per_chunk) {
  extern __shared__ __align__(16) unsigned char wg_smem[];
  const int pitch = d + 8;
  bf16* Xs = reinterpret_cast<bf16*>(wg_smem);  // [96][pitch]: tokens t0-16 .. t0+79
  bf16* Os = Xs + WG_WIN * pitch;               // [64][pitch]: results of this block before the 16-byte stores
  const int h = blockIdx.y;
  const long long b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int c16 = d / 8;
  const bf16* xbase = src + b * n * src_ld + src_col0 + h * d;
  bf16* obase = dst16 + b * n * (long long)E + h * d;
  // A fragments: a[ks][0] = (row g, k = 16 ks + 2tg, +1), [1] = (row g+8, same k), [2] = (row g, k+8, +9), [3] = (row g+8, k+8, +9)
  uint32_t ahi[3][4], alo[3][4];
  const float* wh = w + h * TAPS;
#pragma unroll
  for (int ks = 0; ks < 3; ++ks)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = g + ((q & 1) ? 8 : 0), k = ks * 16 + 2 * tg + ((q & 2) ? 8 : 0);
      float t[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = k + e - r;  // tap index of T[r, k + e]
        t[e] = (j >= 0 && j < TAPS) ? wh[BWD ? TAPS - 1 - j : j] : 0.f;
      }
      const __nv_bfloat162 hi = __floats2bfloat162_rn(t[0], t[1]);
      const float2 hf = __bfloat1622float2(hi);
      const __nv_bfloat162 lo = __floats2bfloat162_rn(t[0] - hf.x, t[1] - hf.y);
      ahi[ks][q] = *reinterpret_cast<const uint32_t*>(&hi);
      alo[ks][q] = *reinterpret_cast<const uint32_t*>(&lo);
    }
  const int nblocks = (n + WG_TB - 1) / WG_TB;
  const int blk0 = blockIdx.x * blocks_per_chunk, blk1 = min(nblocks, blk0 + blocks_per_chunk);
  const int r0 = warp * 16;
  const uint32_t xs_addr = (uint32_t)__cvta_generic_to_shared(Xs);
  for (int blk = blk0; blk < blk1; ++blk) {
    const int t0 = blk * WG_TB;
    __syncthreads();  // previous block: fragments read, results stored
    for (int i = threadIdx.x; i < WG_WIN * c16; i += blockDim.x) {
      const int r = i / c16, p8 = (i - r * c16) * 8;
      const int t = t0 - 16 + r;
      const uint4 v = (t >= 0 && t < n) ? *reinterpret_cast<const uint4*>(xbase + (long long)t * src_ld + p8) : make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(Xs + r * pitch + p8) = v;
    }
    __syncthreads();
    for (int nb = 0; nb < d / 8; ++nb) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
        // B fragment: tokens (window rows) r0 + 16 ks + 0..15, channels nb*8 .. +7; lanes 0-15 supply the row addresses
        uint32_t b0, b1;
        ldmatrix_x2_trans(b0, b1, xs_addr + (uint32_t)(((r0 + ks * 16 + (lane & 15)) * pitch + nb * 8) * 2));
        mma_bf16_16816(acc, ahi[ks][0], ahi[ks][1], ahi[ks][2], ahi[ks][3], b0, b1);
        mma_bf16_16816(acc, alo[ks][0], alo[ks][1], alo[ks][2], alo[ks][3], b0, b1);
      }
      *reinterpret_cast<__nv_bfloat162*>(Os + (r0 + g) * pitch + nb * 8 + 2 * tg) = __floats2bfloat162_rn(acc[0], acc[1]);
      *reinterpret_cast<__nv_bfloat162*>(Os + (r0 + g + 8) * pitch + nb * 8 + 2 * tg) = __floats2bfloat162_rn(acc[2], acc[3]);
    }
    __syncwarp();
    for (int i = lane; i < 16 * c16; i += 32) {  // this warp's 16 rows, 16 bytes per lane and step
      const int r = i / c16, p8 = (i - r * c16) * 8;
      const int t = t0 + r0 + r;
      if (t < n) *reinterpret_cast<uint4*>(obase + (long long)t * E + p8) = *reinterpret_cast<const uint4*>(Os + (r0 + r) * pitch + p8);
    }
  }
}

// ---- pinv set-up.  a2: [BH, m, m] f32 (row softmax).  scal: [0]=max row abs-sum, [1]=max col abs-sum (as ordered
// uint64 keys: float bits << 32 | ~index so that atomicMax also yields the FIRST arg-max).
__device__ __forceinline__ unsigned long long pack_key(float v, unsigned idx) {
  return ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}
// One CTA per matrix.  Column sums: thread j walks down column j (a warp reads 128 contiguous bytes per row, rows in
// index order so the sum order is that of a sequential loop).  Row sums: one warp per row, lanes strided, shuffle tree.
// (The first version let every thread walk its own ROW: 32 sectors per load instruction, 0.35 ms per call.)
__global__ void __launch_bounds__(512)
pinv_scale_kernel(const float* __restrict__ a2, int m, unsigned long long* __restrict__ keys) {
  __shared__ unsigned long long sk[2];
  if (threadIdx.x < 2) sk[threadIdx.x] = 0ull;
  __syncthreads();
  const int bh = blockIdx.x;
  const float* a = a2 + (long long)bh * m * m;
  unsigned long long best_r = 0, best_c = 0;
  for (int j = threadIdx.x; j < m; j += blockDim.x) {
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;  // four loads in flight; rows still enter each partial in index order
    int i = 0;
    for (; i + 4 <= m; i += 4) {
      c0 += fabsf(a[(long long)i * m + j]);
      c1 += fabsf(a[(long long)(i + 1) * m + j]);
      c2 += fabsf(a[(long long)(i + 2) * m + j]);
      c3 += fabsf(a[(long long)(i + 3) * m + j]);
    }
    for (; i < m; ++i) c0 += fabsf(a[(long long)i * m + j]);
    const unsigned long long kc = pack_key((c0 + c1) + (c2 + c3), (unsigned)(bh * m + j));
    best_c = kc > best_c ? kc : best_c;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = warp; i < m; i += nw) {
    float rs = 0.f;
    for (int j = lane; j < m; j += 32) rs += fabsf(a[(long long)i * m + j]);
    rs = warp_sum(rs);
    const unsigned long long kr = pack_key(rs, (unsigned)(bh * m + i));
    best_r = kr > best_r ? kr : best_r;
  }
  atomicMax(&sk[0], best_r);
  atomicMax(&sk[1], best_c);
  __syncthreads();
  if (threadIdx.x < 2) atomicMax(keys + threadIdx.x, sk[threadIdx.x]);
}
__device__ __forceinline__ float key_val(unsigned long long k) { return __uint_as_float((unsigned)(k >> 32)); }
__device__ __forceinline__ unsigned key_idx(unsigned long long k) { return 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFu); }

// z0[bh,i,j] = a2[bh,j,i] / (c*r)  via a 64x64 shared tile (256 threads = 64 columns x 4 row phases, 16 elements each)
constexpr int TP = 64;
__global__ void __launch_bounds__(256)
pinv_init_kernel(const float* __restrict__ a2, int m, const unsigned long long* __restrict__ keys, float* __restrict__ z32,
                 bf16* __restrict__ z16) {
  __shared__ float tile[TP][TP + 1];
  const float inv = 1.f / (key_val(keys[0]) * key_val(keys[1]));
  const long long base = (long long)blockIdx.z * m * m;
  const int i0 = blockIdx.y * TP, j0 = blockIdx.x * TP;
  const int tx = threadIdx.x & (TP - 1), ty = threadIdx.x / TP;
#pragma unroll 4
  for (int r = ty; r < TP; r += 4) {
    const int i = i0 + r, j = j0 + tx;
    tile[r][tx] = (i < m && j < m) ? a2[base + (long long)i * m + j] : 0.f;
  }
  __syncthreads();
#pragma unroll 4
  for (int r = ty; r < TP; r += 4) {
    const int j = j0 + r, i = i0 + tx;  // output row j, column i
    if (i < m && j < m) {
      const float v = tile[tx][r] * inv;
      if (z32) z32[base + (long long)j * m + i] = v;
      if (z16) z16[base + (long long)j * m + i] = __float2bfloat16(v);
    }
  }
}

// backward of z0 = x^T / D, D = c*r:
//   s = <g_z0, z0>  (whole tensor)      dD = -s / D      dc = dD*r   dr = dD*c
//   gx[bh,i,j] (+)= g_z0[bh,j,i]/D + [row (bh,i) is the arg-max row]*dc + [col (bh,j) is the arg-max col]*dr
// <a, b> with a in f32 and b = the bf16 copy of z0 that the forward keeps anyway (an fp32 z0 was written only for this sum)
__global__ void dot_kernel(const float* __restrict__ a, const bf16* __restrict__ b, long long n, float* __restrict__ out) {
  __shared__ float sh[32];
  float s = 0.f;
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(a)[i];
    const uint2 u = reinterpret_cast<const uint2*>(b)[i];
    const float2 y0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 y1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    s += x.x * y0.x + x.y * y0.y + x.z * y1.x + x.w * y1.y;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) s += a[i] * __bfloat162float(b[i]);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
__global__ void __launch_bounds__(256)
pinv_init_bwd_kernel(const float* __restrict__ gz0, int m, const unsigned long long* __restrict__ keys,
                     const float* __restrict__ dotp, float* __restrict__ gx, int accumulate) {
  __shared__ float tile[TP][TP + 1];
  const float c = key_val(keys[0]), r = key_val(keys[1]);
  const float D = c * r, inv = 1.f / D;
  const float dD = -dotp[0] * inv;
  const float dc = dD * r, dr = dD * c;
  const unsigned arg_row = key_idx(keys[0]), arg_col = key_idx(keys[1]);
  const int bh = blockIdx.z;
  const long long base = (long long)bh * m * m;
  const int i0 = blockIdx.y * TP, j0 = blockIdx.x * TP;  // output tile rows i0.., cols j0..
  const int tx = threadIdx.x & (TP - 1), ty = threadIdx.x / TP;
#pragma unroll 4
  for (int rr = ty; rr < TP; rr += 4) {
    const int j = j0 + rr, i = i0 + tx;  // read g_z0[j, i]
    tile[rr][tx] = (i < m && j < m) ? gz0[base + (long long)j * m + i] : 0.f;
  }
  __syncthreads();
#pragma unroll 4
  for (int rr = ty; rr < TP; rr += 4) {
    const int i = i0 + rr, j = j0 + tx;
    if (i < m && j < m) {
      float v = tile[tx][rr] * inv;
      if ((unsigned)(bh * m + i) == arg_row) v += dc;
      if ((unsigned)(bh * m + j) == arg_col) v += dr;
      float* p = gx + base + (long long)i * m + j;
      *p = accumulate ? *p + v : v;
    }
  }
}

// pinv_init_bwd followed by the row-softmax backward of attn2, in one pass:
//   g[i,j] = ga2[i,j] + gz0[j,i]/D + [row i is the arg-max row]*dc + [col j is the arg-max col]*dr
//   ds[i,j] = scale * p[i,j] * (g[i,j] - sum_j p[i,j] g[i,j])                       (p = bf16 attn2)
// One CTA = 32 rows of one matrix: the m x 32 block of gz0 those rows need is staged (transposed use) in shared memory,
// then each warp finishes 4 rows with the row kept in registers (m <= 512).  Saves the read-modify-write of ga2 and its
// re-read by a separate softmax kernel (1.5 GB -> 0.9 GB per layer at the benchmark shape) -- but measured SLOWER than the
// two kernels it replaces (0.43 vs 0.34 ms per layer: the 128-byte staging reads), so the caller uses it only on request.
constexpr int kFuseMaxChunks = 16;
__global__ void __launch_bounds__(256)
pinv_init_softmax_bwd_kernel(const float* __restrict__ ga2, const float* __restrict__ gz0, const bf16* __restrict__ p16, int m,
                             const unsigned long long* __restrict__ keys, const float* __restrict__ dotp, float scale,
                             bf16* __restrict__ ds16) {
  extern __shared__ float tile[];  // [m][33]
  const float c = key_val(keys[0]), r = key_val(keys[1]);
  const float D = c * r, inv = 1.f / D;
  const float dD = -dotp[0] * inv;
  const float dc = dD * r, dr = dD * c;
  const unsigned arg_row = key_idx(keys[0]), arg_col = key_idx(keys[1]);
  const int bh = blockIdx.y, i0 = blockIdx.x * 32;
  const long long base = (long long)bh * m * m;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j0 = warp * 8; j0 < m; j0 += 64) {  // eight independent 128-byte row segments in flight per warp
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (j0 + u < m && i0 + lane < m) ? gz0[base + (long long)(j0 + u) * m + i0 + lane] : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (j0 + u < m) tile[(j0 + u) * 33 + lane] = v[u];
  }
  __syncthreads();
  const int nch = m / 32;
  for (int ii = warp * 4; ii < warp * 4 + 4; ++ii) {
    const int i = i0 + ii;
    if (i >= m) break;
    const float rowadd = ((unsigned)(bh * m + i) == arg_row) ? dc : 0.f;
    float g[kFuseMaxChunks], p[kFuseMaxChunks];
    float dot = 0.f;
#pragma unroll
    for (int cc = 0; cc < kFuseMaxChunks; ++cc) {
      if (cc < nch) {
        const int j = cc * 32 + lane;
        g[cc] = ga2[base + (long long)i * m + j] + tile[j * 33 + ii] * inv + rowadd + (((unsigned)(bh * m + j) == arg_col) ? dr : 0.f);
        p[cc] = __bfloat162float(p16[base + (long long)i * m + j]);
        dot += p[cc] * g[cc];
      }
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int cc = 0; cc < kFuseMaxChunks; ++cc) {
      if (cc < nch) ds16[base + (long long)i * m + cc * 32 + lane] = __float2bfloat16(scale * p[cc] * (g[cc] - dot));
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

// strips a thread walks: long chunks amortise the 32-token halo and keep the prefetch pipeline running (measured at the
// benchmark shape: 16 strips 0.255 ms vs 4 strips 0.271 ms forward; weight-grad 0.56 vs 1.05 ms), but the grid should still
// hold ~6 CTAs per SM.  MIRROR_B200_CONV_SPC overrides (tuning).
static int conv_strips_per_chunk(int B, int n, int E, int max_spc) {
  static const int env = [] { const char* v = getenv("MIRROR_B200_CONV_SPC"); return v && *v ? atoi(v) : 0; }();
  const int strips = (n + TT - 1) / TT;
  const long long cols = (E / 2 + 127) / 128;
  const long long want_chunks = (6LL * num_sms() + cols * B - 1) / (cols * B);
  int spc = env > 0 ? env : (int)(strips / want_chunks);
  if (env <= 0) spc = spc < 4 ? 4 : (spc > max_spc ? max_spc : spc);
  return spc > strips ? strips : spc;
}

// tensor-core versions of the three FIR kernels (head_dim a multiple of 16), verified by the GPU parity tests.  Measured at the
// benchmark shape with 3 token blocks per CTA: FIR / data gradient 0.254 -> 0.197 ms, weight gradient 0.52 -> 0.26 ms.
// MIRROR_B200_AB_NO_CONV_MMA=1 switches back to the register-window kernels (A/B).
static bool conv_mma_ok(int E) {
  static const int off = [] { const char* v = getenv("MIRROR_B200_AB_NO_CONV_MMA"); return v && *v == '1'; }();
  const int d = E / 8;
  return !off && d % 16 == 0 && d <= WG_MAXD;
}
static bool conv_wgrad_mma_ok(int E) { return conv_mma_ok(E); }
static size_t conv_mma_smem(int E) { return (size_t)(WG_TB + WG_WIN) * (E / 8 + 8) * sizeof(bf16); }
static dim3 conv_mma_grid(int B, int n, int* blocks_per_chunk) {  // (token chunks, heads, slides)
  // 64-token blocks per CTA: few, so that the grid has many more CTAs than fit at once and they overlap each other's loads
  // (18 per CTA: 0.294 ms, 6: 0.214, 3: 0.197, 2: 0.199 forward).  MIRROR_B200_CONV_MMA_BPC overrides (tuning).
  static const int env = [] { const char* v = getenv("MIRROR_B200_CONV_MMA_BPC"); return v && *v ? atoi(v) : 0; }();
  const int nblocks = (n + WG_TB - 1) / WG_TB;
  int bpc = env > 0 ? env : 3;
  if (bpc > nblocks) bpc = nblocks;
  (void)B;
  *blocks_per_chunk = bpc;
  return dim3((unsigned)((nblocks + bpc - 1) / bpc), 8, (unsigned)B);
}

extern "C" int mirror_res_conv_fwd(const void* qkv, const float* w, int32_t B, int32_t n, int32_t E, void* out_bf16,
                                   mirror_stream_t stream) {
  MB_CHECK_ARG(qkv && w && out_bf16 && B > 0 && n > 0 && E % 16 == 0, "res_conv_fwd: bad args");
  if (conv_mma_ok(E)) {
    int bpc;
    const dim3 mg = conv_mma_grid(B, n, &bpc);
    res_conv_mma_kernel<false><<<mg, 128, conv_mma_smem(E), STREAM>>>(reinterpret_cast<const bf16*>(qkv), 3LL * E, 2 * E, w, n, E, E / 8,
                                                                     reinterpret_cast<bf16*>(out_bf16), bpc);
    MB_LAUNCH_CHECK();
    return 0;
  }
  const int spc = conv_strips_per_chunk(B, n, E, 16);
  dim3 grid((E / 2 + 127) / 128, ((n + TT - 1) / TT + spc - 1) / spc, B);
  res_conv_kernel<false><<<grid, 128, 0, STREAM>>>(reinterpret_cast<const bf16*>(qkv), 3LL * E, 2 * E, w, n, E, E / 8,
                                                  reinterpret_cast<bf16*>(out_bf16), E, 0, spc);
  MB_LAUNCH_CHECK();
  return 0;
}
extern "C" int mirror_res_conv_bwd(const void* dout_bf16, const void* qkv, const float* w, int32_t B, int32_t n, int32_t E,
                                   void* dv_bf16, float* dw, mirror_stream_t stream) {
  MB_CHECK_ARG(dout_bf16 && qkv && w && dv_bf16 && dw && B > 0 && n > 0 && E % 16 == 0, "res_conv_bwd: bad args");
  const int spc = conv_strips_per_chunk(B, n, E, 16);
  dim3 grid((E / 2 + 127) / 128, ((n + TT - 1) / TT + spc - 1) / spc, B);
  if (conv_mma_ok(E)) {
    int bpc;
    const dim3 mg = conv_mma_grid(B, n, &bpc);
    res_conv_mma_kernel<true><<<mg, 128, conv_mma_smem(E), STREAM>>>(reinterpret_cast<const bf16*>(dout_bf16), E, 0, w, n, E, E / 8,
                                                                    reinterpret_cast<bf16*>(dv_bf16), bpc);
  } else {
    res_conv_kernel<true><<<grid, 128, 0, STREAM>>>(reinterpret_cast<const bf16*>(dout_bf16), E, 0, w, n, E, E / 8,
                                                   reinterpret_cast<bf16*>(dv_bf16), E, 0, spc);
  }
  MB_LAUNCH_CHECK();
  const int d = E / 8;
  if (conv_wgrad_mma_ok(E)) {  // tensor-core weight gradient
    int bpc;
    const dim3 mg = conv_mma_grid(B, n, &bpc);
    res_conv_wgrad_mma_kernel<<<mg, 128, conv_mma_smem(E), STREAM>>>(reinterpret_cast<const bf16*>(dout_bf16),
                                                                    reinterpret_cast<const bf16*>(qkv), n, E, d, dw, bpc);
  } else {
    const int wspc = conv_strips_per_chunk(B, n, E, 24);
    res_conv_wgrad_kernel<<<dim3(grid.x, ((n + TT - 1) / TT + wspc - 1) / wspc, B), 128, 0, STREAM>>>(
        reinterpret_cast<const bf16*>(dout_bf16), reinterpret_cast<const bf16*>(qkv), n, E, d, dw, wspc);
  }
  MB_LAUNCH_CHECK();
  return 0;
}

/* ds_bf16 = softmax_bwd(attn2, ga2 + pinv_init_bwd terms) without touching ga2; needs m % 32 == 0 and m <= 512 */
extern "C" int mirror_pinv_init_softmax_bwd(const float* ga2, const float* gz0, const void* z0_bf16, const void* a2_bf16, int32_t BH,
                                            int32_t m, void* scratch32, float scale, void* ds_bf16, mirror_stream_t stream) {
  MB_CHECK_ARG(ga2 && gz0 && z0_bf16 && a2_bf16 && scratch32 && ds_bf16 && BH > 0 && m > 0 && m % 32 == 0 && m <= 32 * kFuseMaxChunks,
               "pinv_init_softmax_bwd: bad args (m must be a multiple of 32, at most 512)");
  float* dotp = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch32) + 16);
  MB_CUDA(cudaMemsetAsync(dotp, 0, 4, STREAM));
  const long long n = (long long)BH * m * m;
  dot_kernel<<<ew_grid(n, 256 * 8), 256, 0, STREAM>>>(gz0, reinterpret_cast<const bf16*>(z0_bf16), n, dotp);
  MB_LAUNCH_CHECK();
  const size_t smem = (size_t)m * 33 * sizeof(float);
  static DeviceOnce once;
  if (once.first()) {
    MB_CUDA(cudaFuncSetAttribute(pinv_init_softmax_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * kFuseMaxChunks * 33 * 4));
  }
  pinv_init_softmax_bwd_kernel<<<dim3(m / 32, BH), 256, smem, STREAM>>>(ga2, gz0, reinterpret_cast<const bf16*>(a2_bf16), m,
                                                                       reinterpret_cast<const unsigned long long*>(scratch32), dotp,
                                                                       scale, reinterpret_cast<bf16*>(ds_bf16));
  MB_LAUNCH_CHECK();
  return 0;
}

/* scratch: 4 x 8 bytes: [0],[1] ordered max keys, [2] float dot accumulator (zeroed here / by pinv_init_bwd) */
extern "C" int mirror_pinv_init(const float* a2, int32_t BH, int32_t m, void* scratch32, float* z_f32, void* z_bf16,
                                mirror_stream_t stream) {
  MB_CHECK_ARG(a2 && scratch32 && (z_f32 || z_bf16) && BH > 0 && m > 0, "pinv_init: bad args");
  MB_CUDA(cudaMemsetAsync(scratch32, 0, 32, STREAM));
  pinv_scale_kernel<<<BH, 512, 0, STREAM>>>(a2, m, reinterpret_cast<unsigned long long*>(scratch32));
  MB_LAUNCH_CHECK();
  dim3 grid((m + TP - 1) / TP, (m + TP - 1) / TP, BH);
  pinv_init_kernel<<<grid, 256, 0, STREAM>>>(a2, m, reinterpret_cast<const unsigned long long*>(scratch32), z_f32,
                                               reinterpret_cast<bf16*>(z_bf16));
  MB_LAUNCH_CHECK();
  return 0;
}
extern "C" int mirror_pinv_init_bwd(const float* gz0, const void* z0_bf16, int32_t BH, int32_t m, void* scratch32, float* gx,
                                    int32_t accumulate, mirror_stream_t stream) {
  MB_CHECK_ARG(gz0 && z0_bf16 && scratch32 && gx && BH > 0 && m > 0, "pinv_init_bwd: bad args");
  float* dotp = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch32) + 16);
  MB_CUDA(cudaMemsetAsync(dotp, 0, 4, STREAM));
  const long long n = (long long)BH * m * m;
  dot_kernel<<<ew_grid(n, 256 * 8), 256, 0, STREAM>>>(gz0, reinterpret_cast<const bf16*>(z0_bf16), n, dotp);
  MB_LAUNCH_CHECK();
  dim3 grid((m + TP - 1) / TP, (m + TP - 1) / TP, BH);
  pinv_init_bwd_kernel<<<grid, 256, 0, STREAM>>>(gz0, m, reinterpret_cast<const unsigned long long*>(scratch32), dotp, gx,
                                                   accumulate);
  MB_LAUNCH_CHECK();
  return 0;
}
