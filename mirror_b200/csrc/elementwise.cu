// HBM-bound glue kernels of the MIRROR step: casts, activation fwd/bwd with the fused dropout
// mask, token assembly (cls + wrap-around padding), mask/pos-embed, rank-based random masking,
// landmark means and their backward, column sums, reparameterisation.  All are coalesced,
// vectorised where the layout allows, and sized as grid-stride loops over a multiple of the SM count.
#include "common.cuh"

namespace mb {
namespace {

inline int grid_for(long long n, int block, int per_thread = 1) {
  long long g = (n + (long long)block * per_thread - 1) / ((long long)block * per_thread);
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// dst[r, 0:cols_out] = bf16(src[r, 0:cols]) zero padded
__global__ void cast_pad_kernel(const float* __restrict__ src, long long rows, int cols, long long lds,
                                bf16* __restrict__ dst, int cols_out, long long ldd) {
  const long long total = rows * cols_out;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols_out;
    const int c = (int)(i - r * cols_out);
    dst[r * ldd + c] = __float2bfloat16(c < cols ? src[r * lds + c] : 0.f);
  }
}
// contiguous fast path, 8 elements per thread
__global__ void cast_vec_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = src[2 * i], b = src[2 * i + 1];
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
    h[0] = __floats2bfloat162_rn(a.x, a.y);
    h[1] = __floats2bfloat162_rn(a.z, a.w);
    h[2] = __floats2bfloat162_rn(b.x, b.y);
    h[3] = __floats2bfloat162_rn(b.z, b.w);
    dst[i] = u;
  }
}

// bf16 "split-3" operand for near-fp32 products on bf16 tensor cores: x = hi + lo (+ O(2^-17)).
// order 0: (hi, lo, hi)   order 1: (hi, hi, lo)   so that  <A0, A1> over the tripled contraction = hi.hi + lo.hi + hi.lo
// The three blocks are laid side by side (stack_rows = 0: dst is [rows_out, 3*cols_out]) or on top of each other
// (stack_rows = 1: dst is [3*rows_out, cols_out]); rows/cols beyond the source extent are zero.
__global__ void cast_split3_kernel(const float* __restrict__ src, long long rows, int cols, long long lds,
                                   bf16* __restrict__ dst, long long rows_out, int cols_out, int stack_rows, int order) {
  const long long total = rows_out * cols_out;
  const long long ldd = stack_rows ? cols_out : 3LL * cols_out;
  const long long blk = stack_rows ? rows_out * (long long)cols_out : cols_out;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols_out;
    const int c = (int)(i - r * cols_out);
    const float x = (r < rows && c < cols) ? src[r * lds + c] : 0.f;
    const bf16 hi = __float2bfloat16(x);
    const bf16 lo = __float2bfloat16(x - __bfloat162float(hi));
    bf16* d = dst + r * ldd + c;
    d[0] = hi;
    d[blk] = order == 0 ? lo : hi;
    d[2 * blk] = order == 0 ? hi : lo;
  }
}

// 8 columns per thread (cols, cols_out, lds multiples of 8, 16-byte aligned bases): two float4 loads, three 16-byte stores
__global__ void cast_split3_vec_kernel(const float* __restrict__ src, long long rows, int cols, long long lds,
                                       bf16* __restrict__ dst, long long rows_out, int cols_out, int stack_rows, int order) {
  const int c8n = cols_out / 8;
  const long long total = rows_out * c8n;
  const long long ldd = stack_rows ? cols_out : 3LL * cols_out;
  const long long blk = stack_rows ? rows_out * (long long)cols_out : cols_out;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c8n;
    const int c = (int)(i - r * c8n) * 8;
    float x[8];
    if (r < rows && c < cols) {  // cols % 8 == 0: a group of 8 is inside or outside as a whole
      const float4 a = *reinterpret_cast<const float4*>(src + r * lds + c), b = *reinterpret_cast<const float4*>(src + r * lds + c + 4);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = 0.f;
    }
    uint4 hi, lo;
    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&hi);
    __nv_bfloat162* ll = reinterpret_cast<__nv_bfloat162*>(&lo);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      hh[k] = __floats2bfloat162_rn(x[2 * k], x[2 * k + 1]);
      const float2 hf = __bfloat1622float2(hh[k]);
      ll[k] = __floats2bfloat162_rn(x[2 * k] - hf.x, x[2 * k + 1] - hf.y);
    }
    bf16* d = dst + r * ldd + c;
    *reinterpret_cast<uint4*>(d) = hi;
    *reinterpret_cast<uint4*>(d + blk) = order == 0 ? lo : hi;
    *reinterpret_cast<uint4*>(d + 2 * blk) = order == 0 ? hi : lo;
  }
}

// strided 2-D copy (gathers e.g. the cls rows of [B,N+1,E] into a dense [B,E] block)
__global__ void copy_rows_kernel(const float* __restrict__ src, long long lds, long long rows, int cols, float* __restrict__ dst,
                                 long long ldd) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    dst[r * ldd + c] = src[r * lds + c];
  }
}

__global__ void axpy_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n, float alpha) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] += alpha * src[i];
}

// dst[r, :] = float(src[idx[r], :]): the dataset's per-slide resampling (datasets/dataset_pretrain.py:157-161) done on the device
// over the packed patch features of the batch; src is fp32 (src_bf16 = 0) or bf16 (= 1: half the H2D bytes of the feature
// store).  One warp per output row, 16-byte lanes; HBM-bound (reads + writes rows * cols elements once).
__global__ void __launch_bounds__(256)
gather_rows_kernel(const void* __restrict__ src, int src_bf16, long long lds, const long long* __restrict__ idx, long long rows, int cols,
                   long long n_src, float* __restrict__ dst) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long r = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * wpb) {
    long long s = idx[r];
    s = s < 0 ? 0 : (s >= n_src ? n_src - 1 : s);  // clamp: a bad index must not read out of bounds
    float* d = dst + r * cols;
    if (src_bf16) {
      const bf16* sp = reinterpret_cast<const bf16*>(src) + s * lds;
      if ((cols & 7) == 0 && (lds & 7) == 0) {
        for (int c = lane * 8; c < cols; c += 256) {
          const uint4 v = *reinterpret_cast<const uint4*>(sp + c);
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
          float4 lo, hi;
          lo.x = __uint_as_float(w[0] << 16); lo.y = __uint_as_float(w[0] & 0xffff0000u);
          lo.z = __uint_as_float(w[1] << 16); lo.w = __uint_as_float(w[1] & 0xffff0000u);
          hi.x = __uint_as_float(w[2] << 16); hi.y = __uint_as_float(w[2] & 0xffff0000u);
          hi.z = __uint_as_float(w[3] << 16); hi.w = __uint_as_float(w[3] & 0xffff0000u);
          *reinterpret_cast<float4*>(d + c) = lo;
          *reinterpret_cast<float4*>(d + c + 4) = hi;
        }
      } else {
        for (int c = lane; c < cols; c += 32) d[c] = __bfloat162float(sp[c]);
      }
    } else {
      const float* sp = reinterpret_cast<const float*>(src) + s * lds;
      if ((cols & 3) == 0 && (lds & 3) == 0) {
        for (int c = lane * 4; c < cols; c += 128) *reinterpret_cast<float4*>(d + c) = *reinterpret_cast<const float4*>(sp + c);
      } else {
        for (int c = lane; c < cols; c += 32) d[c] = sp[c];
      }
    }
  }
}

// out = dropout(act(pre)) as bf16 (and optionally f32)
__global__ void act_fwd_kernel(const float* __restrict__ pre, long long n, int act, float drop_p, float drop_scale,
                               uint64_t seed, const unsigned long long* epoch, bf16* __restrict__ o16, float* __restrict__ o32) {
  seed = epoch_seed(seed, epoch);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = pre[i];
    if (act == MIRROR_ACT_RELU) v = fmaxf(v, 0.f);
    else if (act == MIRROR_ACT_GELU) v = gelu_erf(v);
    if (drop_p > 0.f) v = hash_u01(seed, (uint64_t)i) >= drop_p ? v * drop_scale : 0.f;
    if (o16) o16[i] = __float2bfloat16(v);
    if (o32) o32[i] = v;
  }
}
// dx = dy * dropmask * act'(pre) on [B,T,C] views with independent batch / row strides (padded layouts).
__global__ void act_bwd_kernel(const float* __restrict__ dy, long long bs_dy, long long ld_dy, const float* __restrict__ pre,
                               long long bs_pre, long long ld_pre, int B, int T, int C, int act, float drop_p, float drop_scale,
                               uint64_t seed, const unsigned long long* epoch, bf16* __restrict__ o16, long long bs16, long long ld16,
                               float* __restrict__ o32, long long bs32, long long ld32) {
  seed = epoch_seed(seed, epoch);
  const long long total = (long long)B * T * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long bt = i / C;
    const int t = (int)(bt % T);
    const long long b = bt / T;
    float g = dy[b * bs_dy + t * ld_dy + c];
    if (drop_p > 0.f) g = hash_u01(seed, (uint64_t)i) >= drop_p ? g * drop_scale : 0.f;
    if (act != MIRROR_ACT_NONE) {
      const float x = pre[b * bs_pre + t * ld_pre + c];
      if (act == MIRROR_ACT_RELU) g = x > 0.f ? g : 0.f;
      else g *= gelu_erf_grad(x);
    }
    if (o16) o16[b * bs16 + t * ld16 + c] = __float2bfloat16(g);
    if (o32) o32[b * bs32 + t * ld32 + c] = g;
  }
}

// The same for C % 4 == 0 and 16-byte aligned rows: one warp per (b,t) row, float4 lanes -- no 64-bit div/mod per element
// (the scalar kernel moved 0.6 GB in 0.4 ms on the dropout backward of a token matrix).
__global__ void __launch_bounds__(256)
act_bwd_rows_kernel(const float* __restrict__ dy, long long bs_dy, long long ld_dy, const float* __restrict__ pre, long long bs_pre,
                    long long ld_pre, int B, int T, int C, int act, float drop_p, float drop_scale, uint64_t seed,
                    const unsigned long long* epoch, bf16* __restrict__ o16, long long bs16, long long ld16, float* __restrict__ o32,
                    long long bs32, long long ld32) {
  seed = epoch_seed(seed, epoch);
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const long long rows = (long long)B * T;
  for (long long r = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * wpb) {
    const long long b = r / T;
    const int t = (int)(r - b * T);
    const float* gy = dy + b * bs_dy + t * ld_dy;
    const float* px = pre ? pre + b * bs_pre + t * ld_pre : nullptr;
    for (int c = lane * 4; c < C; c += 128) {
      float4 g = *reinterpret_cast<const float4*>(gy + c);
      if (drop_p > 0.f) {
        const uint64_t i = (uint64_t)r * C + c;  // dense (b,t,c) index, as in the forward
        g.x = hash_u01(seed, i) >= drop_p ? g.x * drop_scale : 0.f;
        g.y = hash_u01(seed, i + 1) >= drop_p ? g.y * drop_scale : 0.f;
        g.z = hash_u01(seed, i + 2) >= drop_p ? g.z * drop_scale : 0.f;
        g.w = hash_u01(seed, i + 3) >= drop_p ? g.w * drop_scale : 0.f;
      }
      if (act != MIRROR_ACT_NONE) {
        const float4 x = *reinterpret_cast<const float4*>(px + c);
        if (act == MIRROR_ACT_RELU) {
          g.x = x.x > 0.f ? g.x : 0.f; g.y = x.y > 0.f ? g.y : 0.f; g.z = x.z > 0.f ? g.z : 0.f; g.w = x.w > 0.f ? g.w : 0.f;
        } else {
          g.x *= gelu_erf_grad(x.x); g.y *= gelu_erf_grad(x.y); g.z *= gelu_erf_grad(x.z); g.w *= gelu_erf_grad(x.w);
        }
      }
      if (o16) {
        uint2 u;
        *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(g.x, g.y);
        *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(g.z, g.w);
        *reinterpret_cast<uint2*>(o16 + b * bs16 + t * ld16 + c) = u;
      }
      if (o32) *reinterpret_cast<float4*>(o32 + b * bs32 + t * ld32 + c) = g;
    }
  }
}

// h[b,0,:] = cls ; h[b,1+N+j,:] = h[b,1+j,:] for j < add      (models/mirror.py:656-665)
__global__ void assemble_fwd_kernel(float* __restrict__ h, const float* __restrict__ cls, int B, int N, int add, int E) {
  const int S = 1 + N + add;
  const long long total = (long long)B * (1 + add) * E;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % E);
    const int j = (int)((i / E) % (1 + add));
    const int b = (int)(i / ((long long)E * (1 + add)));
    float* hb = h + (long long)b * S * E;
    if (j == 0) hb[e] = cls[e];
    else hb[(long long)(N + j) * E + e] = hb[(long long)j * E + e];
  }
}
// dpre[b,j,:] = relu'(h[b,1+j,:]) * (dh[b,1+j,:] + (j<add ? dh[b,1+N+j,:] : 0));  dcls[e] += sum_b dh[b,0,e]
__global__ void wsi_embed_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ h, int B, int N, int add, int E,
                                     bf16* __restrict__ dpre, float* __restrict__ dcls) {
  const int S = 1 + N + add, E2 = E / 2;
  const long long total = (long long)B * N * E2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = 2 * (int)(i % E2);
    const long long bj = i / E2;
    const int j = (int)(bj % N);
    const long long b = bj / N;
    const long long o = (b * S + 1 + j) * E + e;
    float2 g = *reinterpret_cast<const float2*>(dh + o);
    if (j < add) {
      const float2 w = *reinterpret_cast<const float2*>(dh + (b * S + 1 + N + j) * E + e);
      g.x += w.x;
      g.y += w.y;
    }
    const float2 y = *reinterpret_cast<const float2*>(h + o);
    *reinterpret_cast<__nv_bfloat162*>(dpre + bj * E + e) = __floats2bfloat162_rn(y.x > 0.f ? g.x : 0.f, y.y > 0.f ? g.y : 0.f);
  }
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dh[(long long)b * S * E + e];
    dcls[e] += s;
  }
}

// mask[b,j] = rank_j >= keep, rank_j = #{i : noise_i < noise_j or (== and i < j)}  (argsort of argsort,
// models/mirror.py:516-531, 630-647)
__global__ void rank_mask_kernel(const float* __restrict__ noise, int N, int keep, float* __restrict__ mask) {
  extern __shared__ float sn[];
  const int b = blockIdx.x;
  const float* nb = noise + (long long)b * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) sn[i] = nb[i];
  __syncthreads();
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    const float v = sn[j];
    int rank = 0;
    for (int i = 0; i < N; ++i) {
      const float u = sn[i];
      rank += (u < v) || (u == v && i < j);
    }
    mask[(long long)b * N + j] = rank >= keep ? 1.f : 0.f;
  }
}

// The same mask in O(N) per slide for long bags (N = 16 384: the counting kernel above needs N^2 = 2.7e8 compares per slide):
// radix select of the keep-th smallest noise value (four 8-bit passes over order-preserving integer keys, histogram in shared
// memory), then one pass that keeps everything below it and the first `quota` elements EQUAL to it in index order -- exactly
// the stable double argsort of the reference, ties included.
__device__ __forceinline__ unsigned int order_key(float x) {
  const unsigned int u = __float_as_uint(x + 0.f);  // -0 -> +0: equal as floats, so equal as keys
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending float order == ascending unsigned order
}
__global__ void __launch_bounds__(1024)
rank_mask_select_kernel(const float* __restrict__ noise, int N, int keep, float* __restrict__ mask) {
  extern __shared__ unsigned int sk[];  // N keys (later: the 0/1 results)
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_k, s_less;
  __shared__ unsigned int wsum[32];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const float* nb = noise + (long long)b * N;
  float* mb_ = mask + (long long)b * N;
  if (keep <= 0 || keep >= N) {
    for (int i = tid; i < N; i += nt) mb_[i] = keep <= 0 ? 1.f : 0.f;
    return;
  }
  for (int i = tid; i < N; i += nt) sk[i] = order_key(nb[i]);
  if (tid == 0) { s_prefix = 0u; s_k = (unsigned int)(keep - 1); s_less = 0u; }
  __syncthreads();
  unsigned int pmask = 0u;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += nt) hist[i] = 0u;
    __syncthreads();
    const unsigned int prefix = s_prefix;
    for (int i = tid; i < N; i += nt) {
      const unsigned int key = sk[i];
      if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {  // 256 bins: a serial walk is ~1 us per pass
      unsigned int k = s_k, cum = 0u;
      int bin = 0;
      for (; bin < 256; ++bin) {
        if (cum + hist[bin] > k) break;
        cum += hist[bin];
      }
      s_prefix = prefix | ((unsigned int)bin << shift);
      s_k = k - cum;
      s_less += cum;
    }
    pmask |= 0xFFu << shift;
    __syncthreads();
  }
  const unsigned int vstar = s_prefix;                       // key of the element of rank keep - 1
  const unsigned int quota = (unsigned int)keep - s_less;    // how many elements equal to it are kept (>= 1)
  // order of the equal elements by index: contiguous chunk per thread, exclusive block scan of the per-chunk counts
  const int chunk = (N + nt - 1) / nt;
  const int i0 = min(N, tid * chunk), i1 = min(N, i0 + chunk);
  unsigned int cnt = 0u;
  for (int i = i0; i < i1; ++i) cnt += sk[i] == vstar;
  unsigned int incl = cnt;
  const int lane = tid & 31, w = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  if (w == 0) {
    unsigned int v = lane < (nt >> 5) ? wsum[lane] : 0u, iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, iv, o);
      if (lane >= o) iv += t;
    }
    wsum[lane] = iv - v;  // exclusive prefix of the warp sums
  }
  __syncthreads();
  unsigned int seen = wsum[w] + incl - cnt;  // equal elements before this thread's chunk
  for (int i = i0; i < i1; ++i) {
    const unsigned int key = sk[i];
    unsigned int m;
    if (key == vstar) m = seen++ >= quota;
    else m = key > vstar;
    sk[i] = m;
  }
  __syncthreads();
  for (int i = tid; i < N; i += nt) mb_[i] = sk[i] ? 1.f : 0.f;
}

// r[b,t,e] = (t >= first && mask[b,t-first] ? tok[e*tok_stride] : r[b,t,e]) + pos[t,e]
__global__ void mask_pos_fwd_kernel(float* __restrict__ r, const float* __restrict__ mask, const float* __restrict__ tok,
                                    int tok_stride, const float* __restrict__ pos, int B, int T, int E, int first) {
  const long long total = (long long)B * T * E;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % E);
    const int t = (int)((i / E) % T);
    const int b = (int)(i / ((long long)E * T));
    const bool m = t >= first && mask[(long long)b * (T - first) + (t - first)] != 0.f;
    r[i] = (m ? tok[(long long)e * tok_stride] : r[i]) + pos[(long long)t * E + e];
  }
}
// E % 4 == 0, tok_stride == 1: one warp per token row, float4 lanes (masked rows never read r)
__global__ void __launch_bounds__(256)
mask_pos_fwd_rows_kernel(float* __restrict__ r, const float* __restrict__ mask, const float* __restrict__ tok,
                         const float* __restrict__ pos, int B, int T, int E, int first) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const long long rows = (long long)B * T;
  for (long long row = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    const long long b = row / T;
    const int t = (int)(row - b * T);
    const bool m = t >= first && mask[b * (T - first) + (t - first)] != 0.f;
    float* rr = r + row * E;
    const float* pp = pos + (long long)t * E;
    for (int c = lane * 4; c < E; c += 128) {
      const float4 a = m ? *reinterpret_cast<const float4*>(tok + c) : *reinterpret_cast<const float4*>(rr + c);
      const float4 q = *reinterpret_cast<const float4*>(pp + c);
      *reinterpret_cast<float4*>(rr + c) = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
    }
  }
}
// dr = masked ? 0 : dy (out of place: with distinct restrict pointers the loads of the slide loop are batched; the in-place
// version serialised every load behind the previous store);  dpos[t,e] += sum_b dy;  dtok[e*tok_stride] += sum over masked slots
__global__ void mask_pos_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ mask, float* __restrict__ dr,
                                    float* __restrict__ dtok, int tok_stride, float* __restrict__ dpos, int B, int T, int E, int first) {
  const long long total = (long long)T * E;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % E);
    const int t = (int)(i / E);
    float sp = 0.f, st = 0.f;
#pragma unroll 8
    for (int b = 0; b < B; ++b) {
      const long long o = ((long long)b * T + t) * E + e;
      const float g = dy[o];
      const bool masked = t >= first && mask[(long long)b * (T - first) + (t - first)] != 0.f;
      sp += g;
      st += masked ? g : 0.f;
      dr[o] = masked ? 0.f : g;
    }
    dpos[i] += sp;
    if (st != 0.f) atomicAdd(dtok + (long long)e * tok_stride, st);
  }
}

// lm[b,j,c] = mean_{s<seg} qkv[b, j*seg+s, c]  for c < 2E (q and k slots); 8 channels (16 bytes) per thread
__global__ void landmark_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ lm, int B, int n, int m, int seg,
                                    int E) {
  const int C8 = 2 * E / 8;
  const long long total = (long long)B * m * C8;
  const float inv = 1.f / seg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int j = (int)((i / C8) % m);
    const int b = (int)(i / ((long long)C8 * m));
    const uint4* src = reinterpret_cast<const uint4*>(qkv + ((long long)b * n + (long long)j * seg) * 3 * E) + c8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < seg; ++s) {
      const uint4 u = src[(long long)s * 3 * E / 8];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __bfloat1622float2(h[t]);
        acc[2 * t] += f.x;
        acc[2 * t + 1] += f.y;
      }
    }
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int t = 0; t < 4; ++t) oh[t] = __floats2bfloat162_rn(acc[2 * t] * inv, acc[2 * t + 1] * inv);
    reinterpret_cast<uint4*>(lm + ((long long)b * m + j) * 2 * E)[c8] = o;
  }
}
// out[c] += sum_r x[r, c]   (bias gradients).  Block = 32 x 8 threads over a [rows_chunk, 32-col] tile.
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long long rows, int cols, long long ld, float* __restrict__ out,
                              int rows_per_block) {
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s = 0.f;
  if (c < cols)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) s += (float)x[r * ld + c];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// z = mu + exp(0.5*logvar)*eps   (models/mirror.py:830-833)
__global__ void reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ eps,
                                   long long n, bf16* __restrict__ z16, float* __restrict__ z32) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float z = mu[i] + __expf(0.5f * lv[i]) * eps[i];
    if (z16) z16[i] = __float2bfloat16(z);
    if (z32) z32[i] = z;
  }
}
// dmu += dz ; dlv += dz * eps * 0.5 * exp(0.5*logvar)
__global__ void reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ lv, const float* __restrict__ eps,
                                   long long n, float* __restrict__ dmu, float* __restrict__ dlv) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = dz[i];
    dmu[i] += g;
    dlv[i] += g * eps[i] * 0.5f * __expf(0.5f * lv[i]);
  }
}

// out[b,0,:] = full[b,0,:] + cls[b,:];  out[b,t,:] = full[b,t,:] + tok[b,t-1,:] (t >= 1); absent terms are zero.
// The gradient of "one token matrix read as a whole, as its cls row and as its patch rows" in ONE pass (float4 lanes).
__global__ void token_fanout_bwd_kernel(const float* __restrict__ full, const float* __restrict__ cls, const float* __restrict__ tok,
                                        long long tok_bs, long long tok_ld, int B, int T, int E, float* __restrict__ out) {
  const int e4 = E / 4;
  const long long n = (long long)B * T * e4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % e4) * 4;
    const long long r = i / e4;
    const int t = (int)(r % T);
    const long long b = r / T;
    float4 v = full ? *reinterpret_cast<const float4*>(full + r * E + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float* add = t == 0 ? (cls ? cls + b * E + c : nullptr) : (tok ? tok + b * tok_bs + (long long)(t - 1) * tok_ld + c : nullptr);
    if (add) {
      const float4 a = *reinterpret_cast<const float4*>(add);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    *reinterpret_cast<float4*>(out + r * E + c) = v;
  }
}

// out[r] = <a[r,:], b[r,:] - sub[r,:]> (sub optional) for contiguous bf16 rows; 8 lanes per row, 8 elements per lane and step
__global__ void rowdot_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, const bf16* __restrict__ minus, long long rows,
                                   int cols, float* __restrict__ out) {
  const int sub = threadIdx.x & 7;
  const unsigned gmask = 0xFFu << (threadIdx.x & 24);  // the 8 lanes that share a row leave the loop together
  for (long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3; r < rows; r += ((long long)gridDim.x * blockDim.x) >> 3) {
    float acc = 0.f;
    for (int c = sub * 8; c < cols; c += 64) {
      const uint4 x = *reinterpret_cast<const uint4*>(a + r * cols + c);
      const uint4 y = *reinterpret_cast<const uint4*>(b + r * cols + c);
      const uint4 z = minus ? *reinterpret_cast<const uint4*>(minus + r * cols + c) : make_uint4(0u, 0u, 0u, 0u);
      const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&x);
      const __nv_bfloat162* yh = reinterpret_cast<const __nv_bfloat162*>(&y);
      const __nv_bfloat162* zh = reinterpret_cast<const __nv_bfloat162*>(&z);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 p = __bfloat1622float2(xh[t]), q = __bfloat1622float2(yh[t]), s = __bfloat1622float2(zh[t]);
        acc += p.x * (q.x - s.x) + p.y * (q.y - s.y);
      }
    }
    acc += __shfl_xor_sync(gmask, acc, 1);
    acc += __shfl_xor_sync(gmask, acc, 2);
    acc += __shfl_xor_sync(gmask, acc, 4);
    if (sub == 0) out[r] = acc;
  }
}

}  // namespace
}  // namespace mb

using namespace mb;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int mirror_cast_f32_bf16(const float* src, int64_t rows, int32_t cols, int64_t lds, void* dst, int32_t cols_out,
                                    int64_t ldd, mirror_stream_t stream) {
  MB_CHECK_ARG(src && dst && rows >= 0 && cols > 0 && cols_out >= cols && ldd >= cols_out && lds >= cols, "cast: bad args");
  if (rows == 0) return 0;
  const long long n = rows * (long long)cols;
  if (cols == cols_out && lds == cols && ldd == cols && n % 8 == 0 && ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0) {
    cast_vec_kernel<<<grid_for(n / 8, 256), 256, 0, STREAM>>>(reinterpret_cast<const float4*>(src),
                                                              reinterpret_cast<uint4*>(dst), n / 8);
  } else {
    cast_pad_kernel<<<grid_for(rows * (long long)cols_out, 256), 256, 0, STREAM>>>(src, rows, cols, lds,
                                                                                 reinterpret_cast<bf16*>(dst), cols_out, ldd);
  }
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_cast_split3(const float* src, int64_t rows, int32_t cols, int64_t lds, void* dst, int64_t rows_out,
                                  int32_t cols_out, int32_t stack_rows, int32_t order, mirror_stream_t stream) {
  MB_CHECK_ARG(src && dst && rows > 0 && cols > 0 && cols_out >= cols && rows_out >= rows && lds >= cols && (order == 0 || order == 1),
               "cast_split3: bad args");
  const bool al = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  if (cols % 8 == 0 && cols_out % 8 == 0 && lds % 8 == 0 && al) {
    cast_split3_vec_kernel<<<grid_for(rows_out * (long long)(cols_out / 8), 256), 256, 0, STREAM>>>(
        src, rows, cols, lds, reinterpret_cast<bf16*>(dst), rows_out, cols_out, stack_rows, order);
  } else {
    cast_split3_kernel<<<grid_for(rows_out * (long long)cols_out, 256), 256, 0, STREAM>>>(
        src, rows, cols, lds, reinterpret_cast<bf16*>(dst), rows_out, cols_out, stack_rows, order);
  }
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_copy_rows_f32(const float* src, int64_t lds, int64_t rows, int32_t cols, float* dst, int64_t ldd,
                                    mirror_stream_t stream) {
  MB_CHECK_ARG(src && dst && rows > 0 && cols > 0, "copy_rows: bad args");
  copy_rows_kernel<<<grid_for(rows * (long long)cols, 256), 256, 0, STREAM>>>(src, lds, rows, cols, dst, ldd);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_rowdot_bf16(const void* a, const void* b, const void* sub, int64_t rows, int32_t cols, float* out,
                                  mirror_stream_t stream) {
  MB_CHECK_ARG(a && b && out && rows > 0 && cols > 0 && cols % 8 == 0, "rowdot_bf16: bad args (cols must be a multiple of 8)");
  const long long threads = ((rows * 8 + 255) / 256) * 256;  // whole 8-lane groups: the shuffles need every lane of a warp alive
  rowdot_bf16_kernel<<<grid_for(threads, 256), 256, 0, STREAM>>>(reinterpret_cast<const bf16*>(a), reinterpret_cast<const bf16*>(b), reinterpret_cast<const bf16*>(sub), rows, cols, out);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_axpy_f32(float* dst, const float* src, int64_t n, float alpha, mirror_stream_t stream) {
  MB_CHECK_ARG(dst && src && n >= 0, "axpy: bad args");
  if (n == 0) return 0;
  axpy_kernel<<<grid_for(n, 256), 256, 0, STREAM>>>(dst, src, n, alpha);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_token_fanout_bwd(const float* d_full, const float* d_cls, const float* d_tok, int64_t tok_bs, int64_t tok_ld,
                                       int32_t B, int32_t T, int32_t E, float* out, mirror_stream_t stream) {
  MB_CHECK_ARG(out && B > 0 && T > 0 && E > 0 && E % 4 == 0 && (!d_tok || (tok_bs % 4 == 0 && tok_ld % 4 == 0)),
               "token_fanout_bwd: bad args (E and the strides must be multiples of 4)");
  token_fanout_bwd_kernel<<<grid_for((long long)B * T * (E / 4), 256), 256, 0, STREAM>>>(d_full, d_cls, d_tok, tok_bs, tok_ld, B, T, E, out);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_act_fwd(const float* pre, int64_t n, int32_t act, float drop_p, uint64_t seed, void* out_bf16,
                              float* out_f32, mirror_stream_t stream) {
  MB_CHECK_ARG(pre && n > 0 && (out_bf16 || out_f32) && drop_p >= 0.f && drop_p < 1.f, "act_fwd: bad args");
  act_fwd_kernel<<<grid_for(n, 256), 256, 0, STREAM>>>(pre, n, act, drop_p, 1.f / (1.f - drop_p), seed, drop_epoch_ptr(),
                                                       reinterpret_cast<bf16*>(out_bf16), out_f32);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_act_bwd(const float* dy, int64_t bs_dy, int64_t ld_dy, const float* pre, int64_t bs_pre, int64_t ld_pre,
                              int32_t B, int32_t T, int32_t C, int32_t act, float drop_p, uint64_t seed, void* out_bf16,
                              int64_t bs16, int64_t ld16, float* out_f32, int64_t bs32, int64_t ld32, mirror_stream_t stream) {
  MB_CHECK_ARG(dy && B > 0 && T > 0 && C > 0 && (out_bf16 || out_f32) && (act == 0 || pre) && drop_p >= 0.f && drop_p < 1.f,
               "act_bwd: bad args");
  auto al = [](const void* ptr, long long bs, long long ld, int bytes) {
    return !ptr || (((reinterpret_cast<uintptr_t>(ptr) * 1ULL) % (4 * bytes)) == 0 && bs % 4 == 0 && ld % 4 == 0);
  };
  if (C % 4 == 0 && C >= 128 && al(dy, bs_dy, ld_dy, 4) && al(pre, bs_pre, ld_pre, 4) && al(out_bf16, bs16, ld16, 2) &&
      al(out_f32, bs32, ld32, 4)) {
    act_bwd_rows_kernel<<<grid_for((long long)B * T, 8), 256, 0, STREAM>>>(
        dy, bs_dy, ld_dy, pre, bs_pre, ld_pre, B, T, C, act, drop_p, 1.f / (1.f - drop_p), seed, drop_epoch_ptr(),
        reinterpret_cast<bf16*>(out_bf16), bs16, ld16, out_f32, bs32, ld32);
  } else {
    act_bwd_kernel<<<grid_for((long long)B * T * C, 256), 256, 0, STREAM>>>(
        dy, bs_dy, ld_dy, pre, bs_pre, ld_pre, B, T, C, act, drop_p, 1.f / (1.f - drop_p), seed, drop_epoch_ptr(),
        reinterpret_cast<bf16*>(out_bf16), bs16, ld16, out_f32, bs32, ld32);
  }
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_wsi_assemble_fwd(float* h, const float* cls, int32_t B, int32_t N, int32_t add, int32_t E,
                                       mirror_stream_t stream) {
  MB_CHECK_ARG(h && cls && B > 0 && N > 0 && add >= 0 && add <= N && E > 0, "assemble_fwd: bad args");
  assemble_fwd_kernel<<<grid_for((long long)B * (1 + add) * E, 256), 256, 0, STREAM>>>(h, cls, B, N, add, E);
  MB_LAUNCH_CHECK();
  return 0;
}
extern "C" int mirror_wsi_embed_bwd(const float* dh, const float* h, int32_t B, int32_t N, int32_t add, int32_t E, void* dpre_bf16,
                                    float* dcls, mirror_stream_t stream) {
  MB_CHECK_ARG(dh && h && dpre_bf16 && dcls && B > 0 && N > 0 && add >= 0 && add <= N && E % 2 == 0, "wsi_embed_bwd: bad args");
  wsi_embed_bwd_kernel<<<grid_for((long long)B * N * E / 2, 256), 256, 0, STREAM>>>(dh, h, B, N, add, E,
                                                                                 reinterpret_cast<bf16*>(dpre_bf16), dcls);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_rank_mask(const float* noise, int32_t B, int32_t N, int32_t keep, float* mask, mirror_stream_t stream) {
  MB_CHECK_ARG(noise && mask && B > 0 && N > 0 && keep >= 0 && keep <= N, "rank_mask: bad args");
  const size_t smem = (size_t)N * sizeof(float);
  MB_CHECK_ARG(smem <= 200 * 1024, "rank_mask: N=%d too large for the shared-memory row", N);
  static DeviceOnce once;
  if (once.first()) {
    MB_CUDA(cudaFuncSetAttribute(rank_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  if (N >= 1024) {  // long rows: O(N) radix select; short rows: the N^2 counting kernel is faster than four histogram passes
    static DeviceOnce once2;
    if (once2.first()) MB_CUDA(cudaFuncSetAttribute(rank_mask_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    rank_mask_select_kernel<<<B, 1024, smem, STREAM>>>(noise, N, keep, mask);
  } else {
    rank_mask_kernel<<<B, 512, smem, STREAM>>>(noise, N, keep, mask);
  }
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_mask_pos_fwd(float* r, const float* mask, const float* tok, int32_t tok_stride, const float* pos,
                                   int32_t B, int32_t T, int32_t E, int32_t first, mirror_stream_t stream) {
  MB_CHECK_ARG(r && mask && tok && pos && B > 0 && T > first && E > 0 && first >= 0, "mask_pos_fwd: bad args");
  const bool al16 = ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(tok) | reinterpret_cast<uintptr_t>(pos)) & 15) == 0;
  if (E % 4 == 0 && E >= 128 && tok_stride == 1 && al16) {
    mask_pos_fwd_rows_kernel<<<grid_for((long long)B * T, 8), 256, 0, STREAM>>>(r, mask, tok, pos, B, T, E, first);
  } else {
    mask_pos_fwd_kernel<<<grid_for((long long)B * T * E, 256), 256, 0, STREAM>>>(r, mask, tok, tok_stride, pos, B, T, E, first);
  }
  MB_LAUNCH_CHECK();
  return 0;
}
extern "C" int mirror_mask_pos_bwd(const float* dy, const float* mask, float* dr, float* dtok, int32_t tok_stride, float* dpos,
                                   int32_t B, int32_t T, int32_t E, int32_t first, mirror_stream_t stream) {
  MB_CHECK_ARG(dy && mask && dr && dr != dy && dtok && dpos && B > 0 && T > first && E > 0 && first >= 0,
               "mask_pos_bwd: bad args (dr must not alias dy)");
  mask_pos_bwd_kernel<<<grid_for((long long)T * E, 128), 128, 0, STREAM>>>(dy, mask, dr, dtok, tok_stride, dpos, B, T, E, first);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_landmark_fwd(const void* qkv, void* lm, int32_t B, int32_t n, int32_t m, int32_t seg, int32_t E,
                                   mirror_stream_t stream) {
  MB_CHECK_ARG(qkv && lm && B > 0 && m > 0 && seg > 0 && n == m * seg && E % 8 == 0, "landmark_fwd: bad args (n=%d m=%d seg=%d)",
               n, m, seg);
  landmark_fwd_kernel<<<grid_for((long long)B * m * (E / 4), 256), 256, 0, STREAM>>>(reinterpret_cast<const bf16*>(qkv),
                                                                             reinterpret_cast<bf16*>(lm), B, n, m, seg, E);
  MB_LAUNCH_CHECK();
  return 0;
}
extern "C" int mirror_colsum(const void* x, int32_t is_bf16, int64_t rows, int32_t cols, int64_t ld, float* out,
                             mirror_stream_t stream) {
  MB_CHECK_ARG(x && out && rows > 0 && cols > 0 && ld >= cols, "colsum: bad args");
  const int gx = (cols + 31) / 32;
  long long gy = (long long)num_sms() * 8 / gx;
  if (gy < 1) gy = 1;
  long long rpb = (rows + gy - 1) / gy;
  if (rpb < 64) rpb = 64;
  gy = (rows + rpb - 1) / rpb;
  dim3 grid(gx, (unsigned)gy), block(32, 8);
  if (is_bf16) colsum_kernel<bf16><<<grid, block, 0, STREAM>>>(reinterpret_cast<const bf16*>(x), rows, cols, ld, out, (int)rpb);
  else colsum_kernel<float><<<grid, block, 0, STREAM>>>(reinterpret_cast<const float*>(x), rows, cols, ld, out, (int)rpb);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_reparam_fwd(const float* mu, const float* logvar, const float* eps, int64_t n, void* z_bf16, float* z_f32,
                                  mirror_stream_t stream) {
  MB_CHECK_ARG(mu && logvar && eps && (z_bf16 || z_f32) && n > 0, "reparam_fwd: bad args");
  reparam_fwd_kernel<<<grid_for(n, 256), 256, 0, STREAM>>>(mu, logvar, eps, n, reinterpret_cast<bf16*>(z_bf16), z_f32);
  MB_LAUNCH_CHECK();
  return 0;
}
extern "C" int mirror_reparam_bwd(const float* dz, const float* logvar, const float* eps, int64_t n, float* dmu, float* dlogvar,
                                  mirror_stream_t stream) {
  MB_CHECK_ARG(dz && logvar && eps && dmu && dlogvar && n > 0, "reparam_bwd: bad args");
  reparam_bwd_kernel<<<grid_for(n, 256), 256, 0, STREAM>>>(dz, logvar, eps, n, dmu, dlogvar);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_gather_rows(const void* src, int32_t src_bf16, int64_t lds, int64_t n_src, const int64_t* idx, int64_t rows,
                                  int32_t cols, float* dst, mirror_stream_t stream) {
  MB_CHECK_ARG(src && idx && dst && rows > 0 && cols > 0 && n_src > 0 && lds >= cols, "gather_rows: bad args");
  MB_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
               "gather_rows: buffers must be 16-byte aligned");
  gather_rows_kernel<<<grid_for(rows, 8), 256, 0, STREAM>>>(src, src_bf16, lds, reinterpret_cast<const long long*>(idx), rows, cols,
                                                           n_src, dst);
  MB_LAUNCH_CHECK();
  return 0;
}
