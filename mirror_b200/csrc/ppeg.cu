// PPEG positional layer (models/mirror.py:317-331): x + dwconv7(x) + dwconv5(x) + dwconv3(x) on the H x H token
// grid, cls row bypassed.  The three depthwise kernels and the identity are merged into ONE 7x7 stencil
// (tap-major [49][E] so that channel accesses are coalesced); tokens stay token-major (no transposes).
// Backward: data-grad is the same stencil with flipped taps; the 7x7 weight-grad is reduced once and the
// 5x5 / 3x3 / bias gradients are slices / sums of it.
#include "common.cuh"

namespace mb {
namespace {

// wm[tap][c] = w7[c][tap] + w5 (centre 5x5) + w3 (centre 3x3) + identity(centre);  bm[c] = b7+b5+b3
__global__ void ppeg_merge_kernel(const float* __restrict__ w7, const float* __restrict__ w5, const float* __restrict__ w3,
                                  const float* __restrict__ b7, const float* __restrict__ b5, const float* __restrict__ b3, int E,
                                  float* __restrict__ wm, float* __restrict__ bm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 49 * E) return;
  const int c = i % E, tap = i / E;
  const int dy = tap / 7, dx = tap % 7;
  float v = w7[c * 49 + tap];
  if (dy >= 1 && dy <= 5 && dx >= 1 && dx <= 5) v += w5[c * 25 + (dy - 1) * 5 + (dx - 1)];
  if (dy >= 2 && dy <= 4 && dx >= 2 && dx <= 4) v += w3[c * 9 + (dy - 2) * 3 + (dx - 2)];
  if (dy == 3 && dx == 3) v += 1.f;
  wm[i] = v;
  if (tap == 0) bm[c] = b7[c] + b5[c] + b3[c];
}

// y[b,1+t,c] = bm[c]*use_bias + sum_tap wm[tap or flipped][c] * x[b,1+nbr(t,tap),c];  y[b,0,:] = x[b,0,:]
template <bool FLIP>
__global__ void ppeg_stencil_kernel(const float* __restrict__ x, const float* __restrict__ wm, const float* __restrict__ bm,
                                    int B, int H, int E, float* __restrict__ y, int accumulate) {
  const int S = H * H + 1;
  const int E4 = E / 4;
  const long long total = (long long)B * S * E4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = 4 * (int)(i % E4);
    const long long bs = i / E4;
    const int s = (int)(bs % S);
    const long long b = bs / S;
    float4 acc;
    if (s == 0) {
      acc = *reinterpret_cast<const float4*>(x + bs * E + c);
    } else {
      const int t = s - 1, ty = t / H, tx = t % H;
      acc = (FLIP || !bm) ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(bm + c);
#pragma unroll
      for (int dy = 0; dy < 7; ++dy) {
        const int yy = ty + dy - 3;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int dx = 0; dx < 7; ++dx) {
          const int xx = tx + dx - 3;
          if (xx < 0 || xx >= H) continue;
          const int tap = FLIP ? (6 - dy) * 7 + (6 - dx) : dy * 7 + dx;
          const float4 w = *reinterpret_cast<const float4*>(wm + (long long)tap * E + c);
          const float4 v = *reinterpret_cast<const float4*>(x + ((b * S) + 1 + yy * H + xx) * E + c);
          acc.x += w.x * v.x; acc.y += w.y * v.y; acc.z += w.z * v.z; acc.w += w.w * v.w;
        }
      }
    }
    float4* p = reinterpret_cast<float4*>(y + bs * E + c);
    if (accumulate) {
      const float4 o = *p;
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    *p = acc;
  }
}

// dwm[tap][c] += sum_{b,t} dy[b,1+t,c] * x[b,1+nbr(t,tap),c]; dbm[c] += sum dy.  grid (E/32, chunks), block (32, 8)
__global__ void ppeg_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, int B, int H, int E,
                                  float* __restrict__ dwm, float* __restrict__ dbm, int toks_per_block) {
  __shared__ float sh[8][33];
  const int S = H * H + 1, T = H * H;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long toks = (long long)B * T;
  const long long t0 = (long long)blockIdx.y * toks_per_block;
  const long long t1 = t0 + toks_per_block < toks ? t0 + toks_per_block : toks;
  float acc[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) acc[k] = 0.f;
  float accb = 0.f;
  if (c < E) {
    for (long long q = t0 + threadIdx.y; q < t1; q += 8) {
      const long long b = q / T;
      const int t = (int)(q % T), ty = t / H, tx = t % H;
      const float g = dy[(b * S + 1 + t) * E + c];
      accb += g;
#pragma unroll
      for (int dyy = 0; dyy < 7; ++dyy) {
        const int yy = ty + dyy - 3;
#pragma unroll
        for (int dxx = 0; dxx < 7; ++dxx) {
          const int xx = tx + dxx - 3;
          if (yy >= 0 && yy < H && xx >= 0 && xx < H) acc[dyy * 7 + dxx] += g * x[(b * S + 1 + yy * H + xx) * E + c];
        }
      }
    }
  }
#pragma unroll 1
  for (int k = 0; k < 50; ++k) {
    float v = accb;
#pragma unroll
    for (int kk = 0; kk < 49; ++kk) v = (k == kk) ? acc[kk] : v;  // keeps acc[] in registers
    sh[threadIdx.y][threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.y == 0 && c < E) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) s += sh[r][threadIdx.x];
      atomicAdd(k < 49 ? dwm + (long long)k * E + c : dbm + c, s);
    }
    __syncthreads();
  }
}

// scatter the merged gradient back: dw7 += dwm, dw5 += centre 5x5, dw3 += centre 3x3, db7/5/3 += dbm
__global__ void ppeg_split_kernel(const float* __restrict__ dwm, const float* __restrict__ dbm, int E, float* __restrict__ dw7,
                                  float* __restrict__ dw5, float* __restrict__ dw3, float* __restrict__ db7,
                                  float* __restrict__ db5, float* __restrict__ db3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 49 * E) return;
  const int c = i % E, tap = i / E;
  const int dy = tap / 7, dx = tap % 7;
  const float v = dwm[i];
  dw7[c * 49 + tap] += v;
  if (dy >= 1 && dy <= 5 && dx >= 1 && dx <= 5) dw5[c * 25 + (dy - 1) * 5 + (dx - 1)] += v;
  if (dy >= 2 && dy <= 4 && dx >= 2 && dx <= 4) dw3[c * 9 + (dy - 2) * 3 + (dx - 2)] += v;
  if (tap == 0) {
    const float b = dbm[c];
    db7[c] += b;
    db5[c] += b;
    db3[c] += b;
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

/* wm: [49*E] f32 scratch, bm: [E] f32 scratch (both caller-owned, reused by the backward) */
extern "C" int mirror_ppeg_fwd(const float* x, const float* w7, const float* w5, const float* w3, const float* b7,
                               const float* b5, const float* b3, int32_t B, int32_t H, int32_t E, float* wm, float* bm, float* y,
                               mirror_stream_t stream) {
  MB_CHECK_ARG(x && w7 && w5 && w3 && b7 && b5 && b3 && wm && bm && y && B > 0 && H > 0 && E % 4 == 0, "ppeg_fwd: bad args");
  ppeg_merge_kernel<<<(49 * E + 255) / 256, 256, 0, STREAM>>>(w7, w5, w3, b7, b5, b3, E, wm, bm);
  MB_LAUNCH_CHECK();
  ppeg_stencil_kernel<false><<<ew_grid((long long)B * (H * H + 1) * E / 4, 256), 256, 0, STREAM>>>(x, wm, bm, B, H, E, y, 0);
  MB_LAUNCH_CHECK();
  return 0;
}

/* dx (+)= stencil^T(dy);  dw*, db* += weight gradients.  dwm: [49*E] scratch, dbm: [E] scratch (zeroed here). */
extern "C" int mirror_ppeg_bwd(const float* dy, const float* x, const float* wm, int32_t B, int32_t H, int32_t E, float* dx,
                               int32_t accumulate, float* dwm, float* dbm, float* dw7, float* dw5, float* dw3, float* db7,
                               float* db5, float* db3, mirror_stream_t stream) {
  MB_CHECK_ARG(dy && x && wm && dx && dwm && dbm && dw7 && dw5 && dw3 && db7 && db5 && db3 && B > 0 && H > 0 && E % 4 == 0,
               "ppeg_bwd: bad args");
  ppeg_stencil_kernel<true><<<ew_grid((long long)B * (H * H + 1) * E / 4, 256), 256, 0, STREAM>>>(dy, wm, nullptr, B, H, E, dx,
                                                                                                accumulate);
  MB_LAUNCH_CHECK();
  MB_CUDA(cudaMemsetAsync(dwm, 0, sizeof(float) * 49 * E, STREAM));
  MB_CUDA(cudaMemsetAsync(dbm, 0, sizeof(float) * E, STREAM));
  const long long toks = (long long)B * H * H;
  const int gx = (E + 31) / 32;
  long long gy = (long long)num_sms() * 4 / gx;
  if (gy < 1) gy = 1;
  long long tpb = (toks + gy - 1) / gy;
  if (tpb < 64) tpb = 64;
  gy = (toks + tpb - 1) / tpb;
  ppeg_wgrad_kernel<<<dim3(gx, (unsigned)gy), dim3(32, 8), 0, STREAM>>>(dy, x, B, H, E, dwm, dbm, (int)tpb);
  MB_LAUNCH_CHECK();
  ppeg_split_kernel<<<(49 * E + 255) / 256, 256, 0, STREAM>>>(dwm, dbm, E, dw7, dw5, dw3, db7, db5, db3);
  MB_LAUNCH_CHECK();
  return 0;
}
