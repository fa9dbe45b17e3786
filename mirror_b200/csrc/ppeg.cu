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

// y[b,1+t,c] = bm[c] + sum_tap wm[tap][c] * x[b,1+nbr(t,tap),c];  y[b,0,:] = x[b,0,:].  FLIP: transposed stencil (data grad).
// One thread = one channel (a warp reads 128 contiguous bytes per token) and a strip of TX outputs of one grid row:
// per input row the TX+6 inputs and 7 weights are loaded once and reused by 7*TX FMAs.
constexpr int TX = 16;

template <bool FLIP>
__global__ void __launch_bounds__(128)
ppeg_stencil_kernel(const float* __restrict__ x, const float* __restrict__ wm, const float* __restrict__ bm, int H, int E,
                    float* __restrict__ y, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= E) return;
  const int S = H * H + 1;
  const long long b = blockIdx.z;
  const int strips = (H + TX - 1) / TX;
  if ((int)blockIdx.y == H * strips) {  // cls row bypasses the stencil
    const long long o = b * S * E + c;
    y[o] = accumulate ? y[o] + x[o] : x[o];
    return;
  }
  const int ty = blockIdx.y / strips, tx0 = (blockIdx.y % strips) * TX;
  float acc[TX];
  const float bias = (FLIP || !bm) ? 0.f : bm[c];
#pragma unroll
  for (int i = 0; i < TX; ++i) acc[i] = bias;
#pragma unroll
  for (int dy = 0; dy < 7; ++dy) {
    const int yy = ty + dy - 3;
    if (yy < 0 || yy >= H) continue;
    float in[TX + 6], w[7];
#pragma unroll
    for (int i = 0; i < TX + 6; ++i) {
      const int xx = tx0 + i - 3;
      in[i] = (xx >= 0 && xx < H) ? x[(b * S + 1 + yy * H + xx) * E + c] : 0.f;
    }
#pragma unroll
    for (int dx = 0; dx < 7; ++dx) w[dx] = wm[(long long)(FLIP ? (6 - dy) * 7 + (6 - dx) : dy * 7 + dx) * E + c];
#pragma unroll
    for (int dx = 0; dx < 7; ++dx)
#pragma unroll
      for (int i = 0; i < TX; ++i) acc[i] += w[dx] * in[i + dx];
  }
#pragma unroll
  for (int i = 0; i < TX; ++i) {
    const int tx = tx0 + i;
    if (tx >= H) break;
    const long long o = (b * S + 1 + ty * H + tx) * E + c;
    y[o] = accumulate ? y[o] + acc[i] : acc[i];
  }
}

// dwm[tap][c] += sum_{b,t} dy[b,1+t,c] * x[b,1+nbr(t,tap),c]; dbm[c] += sum dy.
// One thread = one channel; a CTA walks RY grid rows of one slide.  Along a row the 7x7 input window lives in registers:
// column j sits in slot j mod 7, the row loop is unrolled by 7 so every slot index is a compile-time constant
// (7 loads + 49 FMAs per token, no register shuffling).  49+1 coalesced global atomics per thread at the end.
constexpr int RY = 8;

__global__ void __launch_bounds__(128)
ppeg_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, int H, int E, float* __restrict__ dwm,
                  float* __restrict__ dbm) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= E) return;
  const int S = H * H + 1;
  const long long b = blockIdx.z;
  const float* xb = x + (b * S + 1) * E + c;
  const float* gb = dy + (b * S + 1) * E + c;
  float acc[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) acc[k] = 0.f;
  float accb = 0.f;
  const int y0 = blockIdx.y * RY, y1 = min(H, y0 + RY);
  for (int ty = y0; ty < y1; ++ty) {
    float win[7][7];  // win[dy][slot]
#pragma unroll
    for (int dyy = 0; dyy < 7; ++dyy) {
      const int yy = ty + dyy - 3;
      const bool rv = yy >= 0 && yy < H;
#pragma unroll
      for (int sl = 0; sl < 7; ++sl) win[dyy][sl] = 0.f;
#pragma unroll
      for (int j = 0; j < 3; ++j) win[dyy][j] = (rv && j < H) ? xb[(long long)(yy * H + j) * E] : 0.f;
    }
    for (int tx0 = 0; tx0 < H; tx0 += 7) {
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const int tx = tx0 + k;
        const int xn = tx + 3;  // new column enters slot (k+3)%7
#pragma unroll
        for (int dyy = 0; dyy < 7; ++dyy) {
          const int yy = ty + dyy - 3;
          win[dyy][(k + 3) % 7] = (yy >= 0 && yy < H && xn < H) ? xb[(long long)(yy * H + xn) * E] : 0.f;
        }
        const float g = tx < H ? gb[(long long)(ty * H + tx) * E] : 0.f;
        accb += g;
#pragma unroll
        for (int dyy = 0; dyy < 7; ++dyy)
#pragma unroll
          for (int dxx = 0; dxx < 7; ++dxx) acc[dyy * 7 + dxx] += g * win[dyy][(k + dxx + 4) % 7];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 49; ++k) atomicAdd(dwm + (long long)k * E + c, acc[k]);
  atomicAdd(dbm + c, accb);
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

/* wm: [49*E] f32 scratch, bm: [E] f32 scratch (both caller-owned, reused by the backward) */
extern "C" int mirror_ppeg_fwd(const float* x, const float* w7, const float* w5, const float* w3, const float* b7,
                               const float* b5, const float* b3, int32_t B, int32_t H, int32_t E, float* wm, float* bm, float* y,
                               mirror_stream_t stream) {
  MB_CHECK_ARG(x && w7 && w5 && w3 && b7 && b5 && b3 && wm && bm && y && B > 0 && H > 0 && E % 4 == 0, "ppeg_fwd: bad args");
  ppeg_merge_kernel<<<(49 * E + 255) / 256, 256, 0, STREAM>>>(w7, w5, w3, b7, b5, b3, E, wm, bm);
  MB_LAUNCH_CHECK();
  const dim3 grid((E + 127) / 128, H * ((H + TX - 1) / TX) + 1, B);
  ppeg_stencil_kernel<false><<<grid, 128, 0, STREAM>>>(x, wm, bm, H, E, y, 0);
  MB_LAUNCH_CHECK();
  return 0;
}

/* dx (+)= stencil^T(dy);  dw*, db* += weight gradients.  dwm: [49*E] scratch, dbm: [E] scratch (zeroed here). */
extern "C" int mirror_ppeg_bwd(const float* dy, const float* x, const float* wm, int32_t B, int32_t H, int32_t E, float* dx,
                               int32_t accumulate, float* dwm, float* dbm, float* dw7, float* dw5, float* dw3, float* db7,
                               float* db5, float* db3, mirror_stream_t stream) {
  MB_CHECK_ARG(dy && x && wm && dx && dwm && dbm && dw7 && dw5 && dw3 && db7 && db5 && db3 && B > 0 && H > 0 && E % 4 == 0,
               "ppeg_bwd: bad args");
  const dim3 grid((E + 127) / 128, H * ((H + TX - 1) / TX) + 1, B);
  ppeg_stencil_kernel<true><<<grid, 128, 0, STREAM>>>(dy, wm, nullptr, H, E, dx, accumulate);
  MB_LAUNCH_CHECK();
  MB_CUDA(cudaMemsetAsync(dwm, 0, sizeof(float) * 49 * E, STREAM));
  MB_CUDA(cudaMemsetAsync(dbm, 0, sizeof(float) * E, STREAM));
  ppeg_wgrad_kernel<<<dim3((E + 127) / 128, (H + RY - 1) / RY, B), 128, 0, STREAM>>>(dy, x, H, E, dwm, dbm);
  MB_LAUNCH_CHECK();
  ppeg_split_kernel<<<(49 * E + 255) / 256, 256, 0, STREAM>>>(dwm, dbm, E, dw7, dw5, dw3, db7, db5, db3);
  MB_LAUNCH_CHECK();
  return 0;
}
