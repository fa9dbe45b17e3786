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
// One thread = one channel (a warp reads 128 contiguous bytes per token) and a strip of TX output columns; it walks DOWN
// the grid rows of its chunk.  Every input row is loaded once (TX+6 values, fetched one row ahead of its use) and
// scattered into the 7 output rows it touches, whose partial sums wait in registers (7 x TX accumulators, shifted by one
// row per step); the 49 weights of the channel stay in registers.  49 x TX FMAs per TX+6 loads, no re-reads along y.
// (The first version re-loaded 7 input rows per output row and unrolled everything: I-cache misses, 1.2 ms per call.)
constexpr int TX = 16;

template <bool FLIP>
__global__ void __launch_bounds__(128)
ppeg_stencil_kernel(const float* __restrict__ x, const float* __restrict__ wm, const float* __restrict__ bm, int H, int E,
                    float* __restrict__ y, int accumulate, int rows_per_chunk) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= E) return;
  const int S = H * H + 1;
  const long long b = blockIdx.z;
  const int strips = (H + TX - 1) / TX;
  const int chunks = (H + rows_per_chunk - 1) / rows_per_chunk;
  if ((int)blockIdx.y == chunks * strips) {  // cls row bypasses the stencil
    const long long o = b * S * E + c;
    y[o] = accumulate ? y[o] + x[o] : x[o];
    return;
  }
  const int y0 = (blockIdx.y / strips) * rows_per_chunk, y1 = min(H, y0 + rows_per_chunk);
  const int tx0 = (blockIdx.y % strips) * TX;
  const float* xb = x + (b * S + 1) * E + c;
  float* yb = y + (b * S + 1) * E + c;
  float w[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) w[k] = wm[(long long)(FLIP ? 48 - k : k) * E + c];
  const float bias = (FLIP || !bm) ? 0.f : bm[c];
  float acc[7][TX];  // acc[r] = output row (yy - 3 + r) while input row yy is being consumed ... see the shift below
#pragma unroll
  for (int r = 0; r < 7; ++r)
#pragma unroll
    for (int i = 0; i < TX; ++i) acc[r][i] = bias;
  auto load_row = [&](int yy, float (&dst)[TX + 6]) {
    const bool rv = yy >= 0 && yy < H;
#pragma unroll
    for (int i = 0; i < TX + 6; ++i) {
      const int xx = tx0 + i - 3;
      dst[i] = (rv && xx >= 0 && xx < H) ? xb[(long long)(yy * H + xx) * E] : 0.f;
    }
  };
  float nxt[TX + 6];
  load_row(y0 - 3, nxt);
  // input rows y0-3 .. y1+2; after consuming input row yy, output row yy-3 is complete (it has seen rows yy-6 .. yy)
  for (int yy = y0 - 3; yy < y1 + 3; ++yy) {
    float in[TX + 6];
#pragma unroll
    for (int i = 0; i < TX + 6; ++i) in[i] = nxt[i];
    if (yy + 1 < y1 + 3) load_row(yy + 1, nxt);
    // input row yy feeds output row ty = yy + 3 - dy with tap row dy; acc[r] holds output row yy - 3 + r  ->  r = 6 - dy
#pragma unroll
    for (int dy = 0; dy < 7; ++dy)
#pragma unroll
      for (int dx = 0; dx < 7; ++dx)
#pragma unroll
        for (int i = 0; i < TX; ++i) acc[6 - dy][i] += w[dy * 7 + dx] * in[i + dx];
    const int ty = yy - 3;
    if (ty >= y0) {
#pragma unroll
      for (int i = 0; i < TX; ++i) {
        const int tx = tx0 + i;
        if (tx < H) {
          const long long o = (long long)(ty * H + tx) * E;
          yb[o] = accumulate ? yb[o] + acc[0][i] : acc[0][i];
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int i = 0; i < TX; ++i) acc[r][i] = acc[r + 1][i];
#pragma unroll
    for (int i = 0; i < TX; ++i) acc[6][i] = bias;
  }
}

// dwm[tap][c] += sum_{b,t} dy[b,1+t,c] * x[b,1+nbr(t,tap),c]; dbm[c] += sum dy.
// The mirror image of the stencil walk: the thread keeps the gradients of the 7 output rows an input row touches in
// registers (7 x TX, shifted per step) and its 49 partial sums; input row yy meets gradient row yy + 3 - dy with tap row dy.
// 49+1 coalesced global atomics per thread at the end.
__global__ void __launch_bounds__(128)
ppeg_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, int H, int E, float* __restrict__ dwm,
                  float* __restrict__ dbm, int rows_per_chunk) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= E) return;
  const int S = H * H + 1;
  const long long b = blockIdx.z;
  const int strips = (H + TX - 1) / TX;
  const int y0 = (blockIdx.y / strips) * rows_per_chunk, y1 = min(H, y0 + rows_per_chunk);
  const int tx0 = (blockIdx.y % strips) * TX;
  const float* xb = x + (b * S + 1) * E + c;
  const float* gb = dy + (b * S + 1) * E + c;
  float acc[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) acc[k] = 0.f;
  float accb = 0.f;
  float g[7][TX];  // g[r] = gradient row (yy - 3 + r) while input row yy is consumed; rows outside the chunk are zero
#pragma unroll
  for (int r = 0; r < 7; ++r)
#pragma unroll
    for (int i = 0; i < TX; ++i) g[r][i] = 0.f;
  auto load_x = [&](int yy, float (&dst)[TX + 6]) {
    const bool rv = yy >= 0 && yy < H;
#pragma unroll
    for (int i = 0; i < TX + 6; ++i) {
      const int xx = tx0 + i - 3;
      dst[i] = (rv && xx >= 0 && xx < H) ? xb[(long long)(yy * H + xx) * E] : 0.f;
    }
  };
  auto load_g = [&](int ty, float (&dst)[TX]) {
    const bool rv = ty >= y0 && ty < y1;
#pragma unroll
    for (int i = 0; i < TX; ++i) dst[i] = (rv && tx0 + i < H) ? gb[(long long)(ty * H + tx0 + i) * E] : 0.f;
  };
  float nxt[TX + 6], gn[TX];
  load_x(y0 - 3, nxt);
  load_g(y0, gn);  // gradient row yy + 3 enters slot 6 when input row yy is consumed
  for (int yy = y0 - 3; yy < y1 + 3; ++yy) {
    float in[TX + 6];
#pragma unroll
    for (int i = 0; i < TX + 6; ++i) in[i] = nxt[i];
#pragma unroll
    for (int i = 0; i < TX; ++i) {
      g[6][i] = gn[i];
      accb += gn[i];
    }
    if (yy + 1 < y1 + 3) {
      load_x(yy + 1, nxt);
      load_g(yy + 4, gn);
    }
#pragma unroll
    for (int dyy = 0; dyy < 7; ++dyy)
#pragma unroll
      for (int dxx = 0; dxx < 7; ++dxx)
#pragma unroll
        for (int i = 0; i < TX; ++i) acc[dyy * 7 + dxx] += g[6 - dyy][i] * in[i + dxx];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int i = 0; i < TX; ++i) g[r][i] = g[r + 1][i];
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

// grid rows one thread walks: each chunk re-reads 6 halo rows, so whole columns when the batch alone fills the GPU
// (measured at the benchmark shape: 46 rows 0.41 ms, 16 rows 0.49 ms, 8 rows 0.62 ms).  MIRROR_B200_PPEG_RPC overrides.
static int ppeg_rows_per_chunk(int B, int H, int E) {
  static const int env = [] { const char* v = getenv("MIRROR_B200_PPEG_RPC"); return v && *v ? atoi(v) : 0; }();
  const long long per_chunk = (long long)((E + 127) / 128) * ((H + TX - 1) / TX) * B;
  const long long want_chunks = (4LL * num_sms() + per_chunk - 1) / per_chunk;
  int rpc = env > 0 ? env : (int)(H / want_chunks);
  if (env <= 0 && rpc < 8) rpc = 8;
  return rpc > H ? H : rpc;
}

/* wm: [49*E] f32 scratch, bm: [E] f32 scratch (both caller-owned, reused by the backward) */
extern "C" int mirror_ppeg_fwd(const float* x, const float* w7, const float* w5, const float* w3, const float* b7,
                               const float* b5, const float* b3, int32_t B, int32_t H, int32_t E, float* wm, float* bm, float* y,
                               mirror_stream_t stream) {
  MB_CHECK_ARG(x && w7 && w5 && w3 && b7 && b5 && b3 && wm && bm && y && B > 0 && H > 0 && E % 4 == 0, "ppeg_fwd: bad args");
  ppeg_merge_kernel<<<(49 * E + 255) / 256, 256, 0, STREAM>>>(w7, w5, w3, b7, b5, b3, E, wm, bm);
  MB_LAUNCH_CHECK();
  const int rpc = ppeg_rows_per_chunk(B, H, E);
  const dim3 grid((E + 127) / 128, ((H + rpc - 1) / rpc) * ((H + TX - 1) / TX) + 1, B);
  ppeg_stencil_kernel<false><<<grid, 128, 0, STREAM>>>(x, wm, bm, H, E, y, 0, rpc);
  MB_LAUNCH_CHECK();
  return 0;
}

/* dx (+)= stencil^T(dy);  dw*, db* += weight gradients.  dwm: [49*E] scratch, dbm: [E] scratch (zeroed here). */
extern "C" int mirror_ppeg_bwd(const float* dy, const float* x, const float* wm, int32_t B, int32_t H, int32_t E, float* dx,
                               int32_t accumulate, float* dwm, float* dbm, float* dw7, float* dw5, float* dw3, float* db7,
                               float* db5, float* db3, mirror_stream_t stream) {
  MB_CHECK_ARG(dy && x && wm && dx && dwm && dbm && dw7 && dw5 && dw3 && db7 && db5 && db3 && B > 0 && H > 0 && E % 4 == 0,
               "ppeg_bwd: bad args");
  const int rpc = ppeg_rows_per_chunk(B, H, E);
  const int chunks = (H + rpc - 1) / rpc, strips = (H + TX - 1) / TX;
  const dim3 grid((E + 127) / 128, chunks * strips + 1, B);
  ppeg_stencil_kernel<true><<<grid, 128, 0, STREAM>>>(dy, wm, nullptr, H, E, dx, accumulate, rpc);
  MB_LAUNCH_CHECK();
  MB_CUDA(cudaMemsetAsync(dwm, 0, sizeof(float) * 49 * E, STREAM));
  MB_CUDA(cudaMemsetAsync(dbm, 0, sizeof(float) * E, STREAM));
  ppeg_wgrad_kernel<<<dim3((E + 127) / 128, chunks * strips, B), 128, 0, STREAM>>>(dy, x, H, E, dwm, dbm, rpc);
  MB_LAUNCH_CHECK();
  ppeg_split_kernel<<<(49 * E + 255) / 256, 256, 0, STREAM>>>(dwm, dbm, E, dw7, dw5, dw3, db7, db5, db3);
  MB_LAUNCH_CHECK();
  return 0;
}
