// Loss reductions of the MIRROR step (losses/mirror_loss.py:37-135, losses/info_nce.py:144-164).
// Scalars (losses, the temperature scale, upstream gradients) live in DEVICE memory so that no entry
// point forces a host synchronisation.
#include "common.cuh"

namespace mb {
namespace {

// ------------------------------------------------------------------ masked MSE (retention losses)
// scratch[0] += sum mask[row] * (a-b)^2 / E ; scratch[1] += sum mask[row]     rows = B*T
__global__ void masked_mse_fwd_kernel(const float* __restrict__ a, long long a_bs, const float* __restrict__ b, long long b_bs,
                                      const float* __restrict__ mask, int B, int T, int E, float* __restrict__ scratch) {
  __shared__ float sh[32];
  const long long total = (long long)B * T * E;
  float num = 0.f, den = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % E);
    const long long row = i / E;
    const int t = (int)(row % T);
    const long long bb = row / T;
    const float m = mask[row];
    if (e == 0) den += m;
    if (m != 0.f) {
      const float d = a[bb * a_bs + (long long)t * E + e] - b[bb * b_bs + (long long)t * E + e];
      num += m * d * d;
    }
  }
  num = block_sum(num, sh);
  den = block_sum(den, sh);
  if (threadIdx.x == 0) {
    atomicAdd(scratch, num / E);
    atomicAdd(scratch + 1, den);
  }
}
// The same two kernels for E % 4 == 0 (the WSI retention term: 2048 x 768 tokens per slide): one warp per token row,
// float4 lanes, the (slide, token) split and the mask test once per row -- the scalar versions spend their time in
// 64-bit div/mod per element (2.8 TB/s); unmasked rows (25 %) are never read.
__global__ void __launch_bounds__(256)
masked_mse_fwd_rows_kernel(const float* __restrict__ a, long long a_bs, const float* __restrict__ b, long long b_bs,
                           const float* __restrict__ mask, int B, int T, int E, float* __restrict__ scratch) {
  __shared__ float sh[32];
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const long long rows = (long long)B * T;
  float num = 0.f, den = 0.f;
  for (long long row = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    const float m = mask[row];
    if (lane == 0) den += m;
    if (m == 0.f) continue;
    const long long bb = row / T;
    const int t = (int)(row - bb * T);
    const float* ar = a + bb * a_bs + (long long)t * E;
    const float* br = b + bb * b_bs + (long long)t * E;
    float acc = 0.f;
    for (int c = lane * 4; c < E; c += 128) {
      const float4 x = *reinterpret_cast<const float4*>(ar + c);
      const float4 y = *reinterpret_cast<const float4*>(br + c);
      const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
      acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    num += m * acc;
  }
  num = block_sum(num, sh);
  den = block_sum(den, sh);
  if (threadIdx.x == 0) {
    atomicAdd(scratch, num / E);
    atomicAdd(scratch + 1, den);
  }
}
__global__ void __launch_bounds__(256)
masked_mse_bwd_rows_kernel(const float* __restrict__ a, long long a_bs, const float* __restrict__ b, long long b_bs,
                           const float* __restrict__ mask, int B, int T, int E, const float* __restrict__ scratch,
                           const float* __restrict__ gout, float gw, float* __restrict__ da, long long da_bs, int acc_a,
                           float* __restrict__ db, long long db_bs, int acc_b) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const long long rows = (long long)B * T;
  const float k = *gout * gw * 2.f / (E * scratch[1]);
  for (long long row = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    const float m = mask[row];
    if (m == 0.f && (!da || acc_a) && (!db || acc_b)) continue;  // nothing to add
    const long long bb = row / T;
    const int t = (int)(row - bb * T);
    const long long ro = (long long)t * E;
    const float km = k * m;
    for (int c = lane * 4; c < E; c += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m != 0.f) {
        const float4 x = *reinterpret_cast<const float4*>(a + bb * a_bs + ro + c);
        const float4 y = *reinterpret_cast<const float4*>(b + bb * b_bs + ro + c);
        v = make_float4(km * (x.x - y.x), km * (x.y - y.y), km * (x.z - y.z), km * (x.w - y.w));
      }
      if (da) {
        float4* p = reinterpret_cast<float4*>(da + bb * da_bs + ro + c);
        if (acc_a) { const float4 o = *p; *p = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w); }
        else *p = v;
      }
      if (db) {
        float4* p = reinterpret_cast<float4*>(db + bb * db_bs + ro + c);
        if (acc_b) { const float4 o = *p; *p = make_float4(o.x - v.x, o.y - v.y, o.z - v.z, o.w - v.w); }
        else *p = make_float4(-v.x, -v.y, -v.z, -v.w);
      }
    }
  }
}
__global__ void ratio_kernel(const float* __restrict__ scratch, float* __restrict__ out) { *out = scratch[0] / scratch[1]; }
// da (+)= g * mask * 2 (a-b) / (E * den) ; db (+)= -that
__global__ void masked_mse_bwd_kernel(const float* __restrict__ a, long long a_bs, const float* __restrict__ b, long long b_bs,
                                      const float* __restrict__ mask, int B, int T, int E, const float* __restrict__ scratch,
                                      const float* __restrict__ gout, float gw, float* __restrict__ da, long long da_bs, int acc_a,
                                      float* __restrict__ db, long long db_bs, int acc_b) {
  const long long total = (long long)B * T * E;
  const float k = *gout * gw * 2.f / (E * scratch[1]);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % E);
    const long long row = i / E;
    const int t = (int)(row % T);
    const long long bb = row / T;
    const float m = mask[row];
    const float v = m != 0.f ? k * m * (a[bb * a_bs + (long long)t * E + e] - b[bb * b_bs + (long long)t * E + e]) : 0.f;
    if (da) {
      float* p = da + bb * da_bs + (long long)t * E + e;
      *p = acc_a ? *p + v : v;
    }
    if (db) {
      float* p = db + bb * db_bs + (long long)t * E + e;
      *p = acc_b ? *p - v : -v;
    }
  }
}

// ------------------------------------------------------------------ Gaussian KL ("style" term)
// out = 0.5/B * sum_{r,d} (exp(lv) + mu^2 - 1 - lv)   over R rows (both modalities stacked: R = 2B)
__global__ void gauss_kl_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, long long n, float inv_b,
                                    float* __restrict__ out) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += __expf(lv[i]) + mu[i] * mu[i] - 1.f - lv[i];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(out, 0.5f * inv_b * acc);
}
__global__ void gauss_kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, long long n, float inv_b,
                                    const float* __restrict__ gout, float gw, float* __restrict__ dmu, float* __restrict__ dlv) {
  const float k = *gout * gw * inv_b;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    dmu[i] += k * mu[i];
    dlv[i] += k * 0.5f * (__expf(lv[i]) - 1.f);
  }
}

// ------------------------------------------------------------------ symmetric KL over prototype scores ("cluster" term)
// scores: [2B, P] (rows 0..B-1 WSI, B..2B-1 RNA).  out += 0.5/B * sum_p (r-w)(log r - log w) per pair.  One CTA per pair.
__device__ __forceinline__ float block_lse(const float* __restrict__ x, int P, float* sh) {
  float mx = -INFINITY;
  for (int p = threadIdx.x; p < P; p += blockDim.x) mx = fmaxf(mx, x[p]);
  mx = block_max(mx, sh);
  float s = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) s += __expf(x[p] - mx);
  s = block_sum(s, sh);
  return mx + __logf(s);
}
__global__ void sym_kl_fwd_kernel(const float* __restrict__ scores, int B, int P, float* __restrict__ out) {
  __shared__ float sh[32];
  const float* w = scores + (long long)blockIdx.x * P;
  const float* r = scores + (long long)(B + blockIdx.x) * P;
  const float lw = block_lse(w, P, sh), lr = block_lse(r, P, sh);
  float acc = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const float a = w[p] - lw, b = r[p] - lr;
    acc += (__expf(b) - __expf(a)) * (b - a);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(out, 0.5f * acc / B);
}
// d/d(lw_p) f = -w_p (lr_p - lw_p) - (r_p - w_p);  through log-softmax: ds = g_l - softmax * sum(g_l)
__global__ void sym_kl_bwd_kernel(const float* __restrict__ scores, int B, int P, const float* __restrict__ gout, float gw,
                                  float* __restrict__ dscores32, bf16* __restrict__ dscores16) {
  __shared__ float sh[32];
  const float* w = scores + (long long)blockIdx.x * P;
  const float* r = scores + (long long)(B + blockIdx.x) * P;
  const float lw = block_lse(w, P, sh), lr = block_lse(r, P, sh);
  float sw = 0.f, sr = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const float a = w[p] - lw, b = r[p] - lr;
    const float pw = __expf(a), pr = __expf(b);
    sw += -pw * (b - a) - (pr - pw);
    sr += pr * (b - a) + (pr - pw);
  }
  sw = block_sum(sw, sh);
  sr = block_sum(sr, sh);
  const float k = *gout * gw * 0.5f / B;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const float a = w[p] - lw, b = r[p] - lr;
    const float pw = __expf(a), pr = __expf(b);
    const float gwv = k * ((-pw * (b - a) - (pr - pw)) - pw * sw);
    const float grv = k * ((pr * (b - a) + (pr - pw)) - pr * sr);
    const long long ow = (long long)blockIdx.x * P + p, orr = (long long)(B + blockIdx.x) * P + p;
    if (dscores32) { dscores32[ow] = gwv; dscores32[orr] = grv; }
    if (dscores16) { dscores16[ow] = __float2bfloat16(gwv); dscores16[orr] = __float2bfloat16(grv); }
  }
}

// total = sum_i w_i * term_i
__global__ void combine_kernel(const float* __restrict__ terms, float w0, float w1, float w2, float w3, float w4,
                               float* __restrict__ total) {
  *total = w0 * terms[0] + w1 * terms[1] + w2 * terms[2] + w3 * terms[3] + w4 * terms[4];
}

}  // namespace
}  // namespace mb

using namespace mb;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

static int ew_grid(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms() * 8;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

/* raw: [B,B] unscaled similarities; scale: device scalar; row_lse/col_lse: [B]; loss: device scalar */

/* scratch: 2 floats, zeroed here */
// float4 rows: E and every batch stride a multiple of 4 elements, bases 16-byte aligned, rows wide enough for a warp
static bool mse_rows_ok(int E, const float* a, long long a_bs, const float* b, long long b_bs, const float* da, long long da_bs,
                        const float* db, long long db_bs) {
  auto al = [](const float* p, long long bs) { return !p || ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && bs % 4 == 0); };
  return E % 4 == 0 && E >= 128 && al(a, a_bs) && al(b, b_bs) && al(da, da_bs) && al(db, db_bs);
}

extern "C" int mirror_masked_mse_fwd(const float* a, int64_t a_bs, const float* b, int64_t b_bs, const float* mask, int32_t B,
                                     int32_t T, int32_t E, float* scratch, float* out, mirror_stream_t stream) {
  MB_CHECK_ARG(a && b && mask && scratch && out && B > 0 && T > 0 && E > 0, "masked_mse_fwd: bad args");
  MB_CUDA(cudaMemsetAsync(scratch, 0, 8, STREAM));
  if (mse_rows_ok(E, a, a_bs, b, b_bs, nullptr, 0, nullptr, 0)) {
    masked_mse_fwd_rows_kernel<<<ew_grid((long long)B * T, 8), 256, 0, STREAM>>>(a, a_bs, b, b_bs, mask, B, T, E, scratch);
  } else {
    masked_mse_fwd_kernel<<<ew_grid((long long)B * T * E, 256 * 4), 256, 0, STREAM>>>(a, a_bs, b, b_bs, mask, B, T, E, scratch);
  }
  MB_LAUNCH_CHECK();
  ratio_kernel<<<1, 1, 0, STREAM>>>(scratch, out);
  MB_LAUNCH_CHECK();
  return 0;
}
extern "C" int mirror_masked_mse_bwd(const float* a, int64_t a_bs, const float* b, int64_t b_bs, const float* mask, int32_t B,
                                     int32_t T, int32_t E, const float* scratch, const float* gout, float gw, float* da,
                                     int64_t da_bs, int32_t acc_a, float* db, int64_t db_bs, int32_t acc_b, mirror_stream_t stream) {
  MB_CHECK_ARG(a && b && mask && scratch && gout && (da || db) && B > 0 && T > 0 && E > 0, "masked_mse_bwd: bad args");
  if (mse_rows_ok(E, a, a_bs, b, b_bs, da, da_bs, db, db_bs)) {
    masked_mse_bwd_rows_kernel<<<ew_grid((long long)B * T, 8), 256, 0, STREAM>>>(a, a_bs, b, b_bs, mask, B, T, E, scratch, gout, gw, da,
                                                                             da_bs, acc_a, db, db_bs, acc_b);
  } else {
    masked_mse_bwd_kernel<<<ew_grid((long long)B * T * E, 256 * 2), 256, 0, STREAM>>>(a, a_bs, b, b_bs, mask, B, T, E, scratch, gout,
                                                                                    gw, da, da_bs, acc_a, db, db_bs, acc_b);
  }
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_gauss_kl_fwd(const float* mu, const float* logvar, int64_t n, int32_t B, float* out, mirror_stream_t stream) {
  MB_CHECK_ARG(mu && logvar && out && n > 0 && B > 0, "gauss_kl_fwd: bad args");
  MB_CUDA(cudaMemsetAsync(out, 0, 4, STREAM));
  gauss_kl_fwd_kernel<<<ew_grid(n, 256), 256, 0, STREAM>>>(mu, logvar, n, 1.f / B, out);
  MB_LAUNCH_CHECK();
  return 0;
}
extern "C" int mirror_gauss_kl_bwd(const float* mu, const float* logvar, int64_t n, int32_t B, const float* gout, float gw,
                                   float* dmu, float* dlogvar, mirror_stream_t stream) {
  MB_CHECK_ARG(mu && logvar && gout && dmu && dlogvar && n > 0 && B > 0, "gauss_kl_bwd: bad args");
  gauss_kl_bwd_kernel<<<ew_grid(n, 256), 256, 0, STREAM>>>(mu, logvar, n, 1.f / B, gout, gw, dmu, dlogvar);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_sym_kl_fwd(const float* scores, int32_t B, int32_t P, float* out, mirror_stream_t stream) {
  MB_CHECK_ARG(scores && out && B > 0 && P > 0, "sym_kl_fwd: bad args");
  MB_CUDA(cudaMemsetAsync(out, 0, 4, STREAM));
  sym_kl_fwd_kernel<<<B, 256, 0, STREAM>>>(scores, B, P, out);
  MB_LAUNCH_CHECK();
  return 0;
}
extern "C" int mirror_sym_kl_bwd(const float* scores, int32_t B, int32_t P, const float* gout, float gw, float* dscores_f32,
                                 void* dscores_bf16, mirror_stream_t stream) {
  MB_CHECK_ARG(scores && gout && (dscores_f32 || dscores_bf16) && B > 0 && P > 0, "sym_kl_bwd: bad args");
  sym_kl_bwd_kernel<<<B, 256, 0, STREAM>>>(scores, B, P, gout, gw, dscores_f32, reinterpret_cast<bf16*>(dscores_bf16));
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_loss_combine(const float* terms5, const float* weights5_host, float* total, mirror_stream_t stream) {
  MB_CHECK_ARG(terms5 && weights5_host && total, "loss_combine: bad args");
  combine_kernel<<<1, 1, 0, STREAM>>>(terms5, weights5_host[0], weights5_host[1], weights5_host[2], weights5_host[3],
                                     weights5_host[4], total);
  MB_LAUNCH_CHECK();
  return 0;
}
