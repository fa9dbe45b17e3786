// LayerNorm fwd/bwd, row softmax fwd/bwd and row L2-normalise fwd/bwd.  One warp per row,
// fp32 statistics, 128-bit loads; rows are streamed once (fwd) or twice (bwd) through L1.
#include "common.cuh"

namespace mb {
namespace {

constexpr int kWarps = 8;

// four consecutive gradient elements, from an f32 row or (dy16) from a bf16 row: the token-sized backward GEMM that produces
// d(LN output) writes bf16 through its lean epilogue, and this kernel reads 2 instead of 4 bytes per element
__device__ __forceinline__ float4 load_g4(const float* f32row, const bf16* b16row, int c) {
  if (b16row) {
    const uint2 u = *reinterpret_cast<const uint2*>(b16row + c);
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
  }
  return *reinterpret_cast<const float4*>(f32row + c);
}

// ------------------------------------------------------------------ LayerNorm
// x: [B, X, E] f32, of which the first S rows per slide are normalised (X > S drops the wrap-padding tokens of the WSI
// encoder without a copy).  Outputs are written in a padded row layout [B, n_out, E] at row offset `pad`
// (rows < pad are zero-filled: the Nyström layer front-pads the sequence with zero rows).
__global__ void __launch_bounds__(kWarps * 32)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
              int B, int S, int X, int E, int n_out, int pad, bf16* __restrict__ o16, float* __restrict__ o32,
              float* __restrict__ mean, float* __restrict__ rstd) {
  const int lane = threadIdx.x & 31;
  const long long rows_out = (long long)B * n_out;
  for (long long ro = blockIdx.x * (long long)kWarps + (threadIdx.x >> 5); ro < rows_out; ro += (long long)gridDim.x * kWarps) {
    const int b = (int)(ro / n_out), t = (int)(ro % n_out);
    if (t < pad) {
      for (int c = lane * 4; c < E; c += 128) {
        if (o16) *reinterpret_cast<uint2*>(o16 + ro * E + c) = make_uint2(0u, 0u);
        if (o32) *reinterpret_cast<float4*>(o32 + ro * E + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      continue;
    }
    const int s = t - pad;
    if (s >= S) continue;
    const long long ri = (long long)b * S + s;
    const float* xr = x + ((long long)b * X + s) * E;
    float sum = 0.f;
    for (int c = lane * 4; c < E; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      sum += v.x + v.y + v.z + v.w;
    }
    const float mu = warp_sum(sum) / E;
    float var = 0.f;
    for (int c = lane * 4; c < E; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      const float a = v.x - mu, b2 = v.y - mu, c2 = v.z - mu, d = v.w - mu;
      var += a * a + b2 * b2 + c2 * c2 + d * d;
    }
    const float rs = rsqrtf(warp_sum(var) / E + eps);
    if (lane == 0) {
      mean[ri] = mu;
      rstd[ri] = rs;
    }
    for (int c = lane * 4; c < E; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      const float4 g = *reinterpret_cast<const float4*>(gamma + c);
      const float4 bb = *reinterpret_cast<const float4*>(beta + c);
      float4 y;
      y.x = (v.x - mu) * rs * g.x + bb.x;
      y.y = (v.y - mu) * rs * g.y + bb.y;
      y.z = (v.z - mu) * rs * g.z + bb.z;
      y.w = (v.w - mu) * rs * g.w + bb.w;
      if (o32) *reinterpret_cast<float4*>(o32 + ro * E + c) = y;
      if (o16) {
        uint2 u;
        *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(y.x, y.y);
        *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(y.z, y.w);
        *reinterpret_cast<uint2*>(o16 + ro * E + c) = u;
      }
    }
  }
}

// dy: [B, n_out, E] (f32) at row offset pad.  dx = [add +] rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*gamma.
// x, dx, add: [B, X, E]; rows S..X-1 of a slide did not take part in the forward and receive dx = add (or 0).
// dgamma/dbeta accumulate through shared-memory partials and one atomicAdd per block and column.
__global__ void __launch_bounds__(kWarps * 32)
ln_bwd_kernel(const float* __restrict__ dy, const bf16* __restrict__ dy16, const float* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ mean, const float* __restrict__ rstd, int B, int S, int X, int E, int n_out, int pad,
              float* dx, const float* add, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float sh[];  // [2][E]
  float* sg = sh;
  float* sb = sh + E;
  for (int c = threadIdx.x; c < 2 * E; c += blockDim.x) sh[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * X;
  for (long long ri = blockIdx.x * (long long)kWarps + (threadIdx.x >> 5); ri < rows; ri += (long long)gridDim.x * kWarps) {
    const int b = (int)(ri / X), s = (int)(ri % X);
    if (s >= S) {
      for (int c = lane * 4; c < E; c += 128)
        *reinterpret_cast<float4*>(dx + ri * E + c) = add ? *reinterpret_cast<const float4*>(add + ri * E + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const float* xr = x + ri * E;
    const long long goff = ((long long)b * n_out + pad + s) * E;
    const float* gr = dy16 ? nullptr : dy + goff;
    const bf16* gr16 = dy16 ? dy16 + goff : nullptr;
    const float mu = mean[(long long)b * S + s], rs = rstd[(long long)b * S + s];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane * 4; c < E; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      const float4 g = load_g4(gr, gr16, c);
      const float4 w = *reinterpret_cast<const float4*>(gamma + c);
      const float gx = g.x * w.x, gy = g.y * w.y, gz = g.z * w.z, gw = g.w * w.w;
      s1 += gx + gy + gz + gw;
      s2 += gx * (v.x - mu) + gy * (v.y - mu) + gz * (v.z - mu) + gw * (v.w - mu);
    }
    s1 = warp_sum(s1) / E;
    s2 = warp_sum(s2) * rs / E;  // mean(g * xhat)
    for (int c = lane * 4; c < E; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      const float4 g = load_g4(gr, gr16, c);
      const float4 w = *reinterpret_cast<const float4*>(gamma + c);
      const float hx = (v.x - mu) * rs, hy = (v.y - mu) * rs, hz = (v.z - mu) * rs, hw = (v.w - mu) * rs;
      float4 d;
      d.x = rs * (g.x * w.x - s1 - hx * s2);
      d.y = rs * (g.y * w.y - s1 - hy * s2);
      d.z = rs * (g.z * w.z - s1 - hz * s2);
      d.w = rs * (g.w * w.w - s1 - hw * s2);
      float4* dp = reinterpret_cast<float4*>(dx + ri * E + c);
      if (add) {  // dx = add + LN-gradient (add may alias dx)
        const float4 o = *reinterpret_cast<const float4*>(add + ri * E + c);
        d.x += o.x; d.y += o.y; d.z += o.z; d.w += o.w;
      }
      *dp = d;
      atomicAdd(sg + c, g.x * hx); atomicAdd(sg + c + 1, g.y * hy);
      atomicAdd(sg + c + 2, g.z * hz); atomicAdd(sg + c + 3, g.w * hw);
      atomicAdd(sb + c, g.x); atomicAdd(sb + c + 1, g.y);
      atomicAdd(sb + c + 2, g.z); atomicAdd(sb + c + 3, g.w);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    atomicAdd(dgamma + c, sg[c]);
    atomicAdd(dbeta + c, sb[c]);
  }
}

// Same contract, E == ITERS*128: the per-column partial sums of dgamma / dbeta stay in registers (lane owns columns
// lane*4 + 128*i), are combined across the 8 warps through shared memory once per CTA and leave as one global atomic
// per column and CTA -- no shared-memory atomics in the row loop.
// 4-warp CTAs, three per SM (166 registers per thread): 12 rows in flight per SM instead of the 8 of one 8-warp CTA.
constexpr int kRegWarps = 4;
template <int ITERS>
__global__ void __launch_bounds__(kRegWarps * 32, 3)
ln_bwd_reg_kernel(const float* __restrict__ dy, const bf16* __restrict__ dy16, const float* __restrict__ x, const float* __restrict__ gamma,
                  const float* __restrict__ mean, const float* __restrict__ rstd, int B, int S, int X, int n_out, int pad,
                  float* dx, const float* add, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  constexpr int E = ITERS * 128;
  __shared__ float sh[2 * E];
  for (int c = threadIdx.x; c < 2 * E; c += blockDim.x) sh[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float4 w[ITERS], ag[ITERS], ab[ITERS];
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    w[i] = *reinterpret_cast<const float4*>(gamma + lane * 4 + 128 * i);
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const long long rows = (long long)B * X;
  for (long long ri = blockIdx.x * (long long)kRegWarps + (threadIdx.x >> 5); ri < rows; ri += (long long)gridDim.x * kRegWarps) {
    const int b = (int)(ri / X), s = (int)(ri % X);
    if (s >= S) {
#pragma unroll
      for (int i = 0; i < ITERS; ++i) {
        const long long o = ri * E + lane * 4 + 128 * i;
        *reinterpret_cast<float4*>(dx + o) = add ? *reinterpret_cast<const float4*>(add + o) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      continue;
    }
    const float* xr = x + ri * E;
    const long long goff = ((long long)b * n_out + pad + s) * E;
    const float* gr = dy16 ? nullptr : dy + goff;
    const bf16* gr16 = dy16 ? dy16 + goff : nullptr;
    const float mu = mean[(long long)b * S + s], rs = rstd[(long long)b * S + s];
    float4 xv[ITERS], gv[ITERS];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      xv[i] = *reinterpret_cast<const float4*>(xr + lane * 4 + 128 * i);
      gv[i] = load_g4(gr, gr16, lane * 4 + 128 * i);
      xv[i].x = (xv[i].x - mu) * rs; xv[i].y = (xv[i].y - mu) * rs; xv[i].z = (xv[i].z - mu) * rs; xv[i].w = (xv[i].w - mu) * rs;
      const float gx = gv[i].x * w[i].x, gy = gv[i].y * w[i].y, gz = gv[i].z * w[i].z, gw = gv[i].w * w[i].w;
      s1 += gx + gy + gz + gw;
      s2 += gx * xv[i].x + gy * xv[i].y + gz * xv[i].z + gw * xv[i].w;
    }
    s1 = warp_sum(s1) / E;
    s2 = warp_sum(s2) / E;
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      float4 d;
      d.x = rs * (gv[i].x * w[i].x - s1 - xv[i].x * s2);
      d.y = rs * (gv[i].y * w[i].y - s1 - xv[i].y * s2);
      d.z = rs * (gv[i].z * w[i].z - s1 - xv[i].z * s2);
      d.w = rs * (gv[i].w * w[i].w - s1 - xv[i].w * s2);
      if (add) {
        const float4 o = *reinterpret_cast<const float4*>(add + ri * E + lane * 4 + 128 * i);
        d.x += o.x; d.y += o.y; d.z += o.z; d.w += o.w;
      }
      *reinterpret_cast<float4*>(dx + ri * E + lane * 4 + 128 * i) = d;
      ag[i].x += gv[i].x * xv[i].x; ag[i].y += gv[i].y * xv[i].y; ag[i].z += gv[i].z * xv[i].z; ag[i].w += gv[i].w * xv[i].w;
      ab[i].x += gv[i].x; ab[i].y += gv[i].y; ab[i].z += gv[i].z; ab[i].w += gv[i].w;
    }
  }
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int c = lane * 4 + 128 * i;
    atomicAdd(sh + c, ag[i].x); atomicAdd(sh + c + 1, ag[i].y); atomicAdd(sh + c + 2, ag[i].z); atomicAdd(sh + c + 3, ag[i].w);
    atomicAdd(sh + E + c, ab[i].x); atomicAdd(sh + E + c + 1, ab[i].y); atomicAdd(sh + E + c + 2, ab[i].z); atomicAdd(sh + E + c + 3, ab[i].w);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    atomicAdd(dgamma + c, sh[c]);
    atomicAdd(dbeta + c, sh[E + c]);
  }
}

// -------------------------------------------------------------------- softmax
// One warp per row, row staged in shared memory (cols floats per warp).
__global__ void softmax_fwd_kernel(const float* __restrict__ x, long long rows, int cols, bf16* __restrict__ y16,
                                   float* __restrict__ y32, int warps) {
  extern __shared__ float sh[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* row = sh + (long long)w * cols;
  for (long long r = blockIdx.x * (long long)warps + w; r < rows; r += (long long)gridDim.x * warps) {
    const float* xr = x + r * cols;
    float mx = -INFINITY;
    for (int c = lane; c < cols; c += 32) {
      const float v = xr[c];
      row[c] = v;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float e = __expf(row[c] - mx);
      row[c] = e;
      sum += e;
    }
    const float inv = 1.f / warp_sum(sum);
    for (int c = lane; c < cols; c += 32) {
      const float p = row[c] * inv;
      if (y16) y16[r * cols + c] = __float2bfloat16(p);
      if (y32) y32[r * cols + c] = p;
    }
    __syncwarp();
  }
}
// dx = scale * y * (dy - sum(dy*y))
__global__ void softmax_bwd_kernel(const bf16* __restrict__ y, const float* __restrict__ dy, long long rows, int cols,
                                   float scale, bf16* __restrict__ dx16, float* __restrict__ dx32, int warps) {
  extern __shared__ float sh[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* row = sh + (long long)w * cols;
  for (long long r = blockIdx.x * (long long)warps + w; r < rows; r += (long long)gridDim.x * warps) {
    float dot = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float g = dy[r * cols + c];
      row[c] = g;
      dot += g * __bfloat162float(y[r * cols + c]);
    }
    dot = warp_sum(dot);
    for (int c = lane; c < cols; c += 32) {
      const float v = scale * __bfloat162float(y[r * cols + c]) * (row[c] - dot);
      if (dx16) dx16[r * cols + c] = __float2bfloat16(v);
      if (dx32) dx32[r * cols + c] = v;
    }
    __syncwarp();
  }
}

// --------------------------------------------------------------- L2 normalise
// y = x / max(||x||, eps)  (F.normalize, models/mirror.py:540,683).  Rows may be strided (cls rows).
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, long long ldx, int rows, int cols, float eps,
                                  bf16* __restrict__ y16, float* __restrict__ y32, long long ldy, float* __restrict__ norm) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* xr = x + (long long)r * ldx;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += xr[c] * xr[c];
  const float nrm = fmaxf(sqrtf(warp_sum(s)), eps);
  if (lane == 0 && norm) norm[r] = nrm;
  const float inv = 1.f / nrm;
  for (int c = lane; c < cols; c += 32) {
    const float v = xr[c] * inv;
    if (y16) y16[(long long)r * ldy + c] = __float2bfloat16(v);
    if (y32) y32[(long long)r * ldy + c] = v;
  }
}
// dx (+)= (dy - y*(y.dy)) / norm, y recomputed from x
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ x, long long ldx,
                                  const float* __restrict__ norm, int rows, int cols, float* __restrict__ dx, long long lddx,
                                  int accumulate) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float inv = 1.f / norm[r];
  const float* xr = x + (long long)r * ldx;
  const float* gr = dy + (long long)r * lddy;
  float dot = 0.f;
  for (int c = lane; c < cols; c += 32) dot += gr[c] * xr[c] * inv;
  dot = warp_sum(dot);
  for (int c = lane; c < cols; c += 32) {
    const float v = (gr[c] - xr[c] * inv * dot) * inv;
    float* p = dx + (long long)r * lddx + c;
    *p = accumulate ? *p + v : v;
  }
}

}  // namespace
}  // namespace mb

using namespace mb;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int mirror_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int32_t B, int32_t S,
                                    int32_t x_rows, int32_t E, int32_t n_out, int32_t pad, void* out_bf16, float* out_f32,
                                    float* mean, float* rstd, mirror_stream_t stream) {
  MB_CHECK_ARG(x && gamma && beta && mean && rstd && (out_bf16 || out_f32), "layernorm_fwd: null pointer");
  MB_CHECK_ARG(B > 0 && S > 0 && x_rows >= S && E % 4 == 0 && pad >= 0 && n_out >= S + pad,
               "layernorm_fwd: bad shape (E must be a multiple of 4)");
  const long long rows = (long long)B * n_out;
  long long grid = (rows + kWarps - 1) / kWarps;
  if (grid > (long long)num_sms() * 8) grid = (long long)num_sms() * 8;
  ln_fwd_kernel<<<(int)grid, kWarps * 32, 0, STREAM>>>(x, gamma, beta, eps, B, S, x_rows, E, n_out, pad,
                                                       reinterpret_cast<bf16*>(out_bf16), out_f32, mean, rstd);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_layernorm_bwd(const void* dy_, int32_t dy_is_bf16, const float* x, const float* gamma, const float* mean, const float* rstd,
                                    int32_t B, int32_t S, int32_t x_rows, int32_t E, int32_t n_out, int32_t pad, float* dx,
                                    const float* add, float* dgamma, float* dbeta, mirror_stream_t stream) {
  MB_CHECK_ARG(dy_ && x && gamma && mean && rstd && dx && dgamma && dbeta, "layernorm_bwd: null pointer");
  const float* dy = dy_is_bf16 ? nullptr : reinterpret_cast<const float*>(dy_);
  const bf16* dy16 = dy_is_bf16 ? reinterpret_cast<const bf16*>(dy_) : nullptr;
  MB_CHECK_ARG(B > 0 && S > 0 && x_rows >= S && E % 4 == 0 && pad >= 0 && n_out >= S + pad, "layernorm_bwd: bad shape");
  const long long rows = (long long)B * x_rows;
  long long grid = (rows + kWarps * 4 - 1) / (kWarps * 4);  // >= 4 rows per warp to amortise the column atomics
  if (grid > (long long)num_sms() * 2) grid = (long long)num_sms() * 2;
  if (grid < 1) grid = 1;
  if (E == 768) {
    long long rgrid = (rows + kRegWarps * 4 - 1) / (kRegWarps * 4);
    if (rgrid > (long long)num_sms() * 3) rgrid = (long long)num_sms() * 3;
    if (rgrid < 1) rgrid = 1;
    ln_bwd_reg_kernel<6><<<(int)rgrid, kRegWarps * 32, 0, STREAM>>>(dy, dy16, x, gamma, mean, rstd, B, S, x_rows, n_out, pad, dx, add,
                                                                      dgamma, dbeta);
  } else {
    ln_bwd_kernel<<<(int)grid, kWarps * 32, 2 * E * sizeof(float), STREAM>>>(dy, dy16, x, gamma, mean, rstd, B, S, x_rows, E, n_out, pad, dx,
                                                                             add, dgamma, dbeta);
  }
  MB_LAUNCH_CHECK();
  return 0;
}

static int softmax_cfg(int cols, int* warps, size_t* smem) {
  int w = 8;
  while (w > 1 && (size_t)w * cols * sizeof(float) > 96 * 1024) w >>= 1;
  *warps = w;
  *smem = (size_t)w * cols * sizeof(float);
  return *smem <= 200 * 1024 ? 0 : -1;
}

extern "C" int mirror_softmax_fwd(const float* x, int64_t rows, int32_t cols, void* y_bf16, float* y_f32, mirror_stream_t stream) {
  MB_CHECK_ARG(x && rows > 0 && cols > 0 && (y_bf16 || y_f32), "softmax_fwd: bad args");
  int warps;
  size_t smem;
  MB_CHECK_ARG(softmax_cfg(cols, &warps, &smem) == 0, "softmax_fwd: row of %d columns does not fit shared memory", cols);
  MB_CUDA(cudaFuncSetAttribute(softmax_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long grid = (rows + warps - 1) / warps;
  if (grid > (long long)num_sms() * 8) grid = (long long)num_sms() * 8;
  softmax_fwd_kernel<<<(int)grid, warps * 32, smem, STREAM>>>(x, rows, cols, reinterpret_cast<bf16*>(y_bf16), y_f32, warps);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_softmax_bwd(const void* y_bf16, const float* dy, int64_t rows, int32_t cols, float scale, void* dx_bf16,
                                  float* dx_f32, mirror_stream_t stream) {
  MB_CHECK_ARG(y_bf16 && dy && rows > 0 && cols > 0 && (dx_bf16 || dx_f32), "softmax_bwd: bad args");
  int warps;
  size_t smem;
  MB_CHECK_ARG(softmax_cfg(cols, &warps, &smem) == 0, "softmax_bwd: row of %d columns does not fit shared memory", cols);
  MB_CUDA(cudaFuncSetAttribute(softmax_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long grid = (rows + warps - 1) / warps;
  if (grid > (long long)num_sms() * 8) grid = (long long)num_sms() * 8;
  softmax_bwd_kernel<<<(int)grid, warps * 32, smem, STREAM>>>(reinterpret_cast<const bf16*>(y_bf16), dy, rows, cols, scale,
                                                              reinterpret_cast<bf16*>(dx_bf16), dx_f32, warps);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_l2norm_fwd(const float* x, int64_t ldx, int32_t rows, int32_t cols, float eps, void* y_bf16, float* y_f32,
                                 int64_t ldy, float* norm, mirror_stream_t stream) {
  MB_CHECK_ARG(x && rows > 0 && cols > 0 && (y_bf16 || y_f32), "l2norm_fwd: bad args");
  l2norm_fwd_kernel<<<(rows + 3) / 4, 128, 0, STREAM>>>(x, ldx, rows, cols, eps, reinterpret_cast<bf16*>(y_bf16), y_f32, ldy, norm);
  MB_LAUNCH_CHECK();
  return 0;
}
extern "C" int mirror_l2norm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* norm, int32_t rows,
                                 int32_t cols, float* dx, int64_t lddx, int32_t accumulate, mirror_stream_t stream) {
  MB_CHECK_ARG(dy && x && norm && dx && rows > 0 && cols > 0, "l2norm_bwd: bad args");
  l2norm_bwd_kernel<<<(rows + 3) / 4, 128, 0, STREAM>>>(dy, lddy, x, ldx, norm, rows, cols, dx, lddx, accumulate);
  MB_LAUNCH_CHECK();
  return 0;
}
