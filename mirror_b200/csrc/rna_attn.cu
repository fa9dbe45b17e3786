// RNA encoder attention (models/mirror.py:77-102): the embedding of ONE sample is split into 12 chunks of
// hd = E/12; scaled-dot-product attention runs over those 12 chunks (seq = 12, dim = hd) and the result is
// interleaved dim-major:  out[b, c*12 + i] = o_i[c].  One CTA per sample, everything in shared memory.
#include "common.cuh"

namespace mb {
namespace {

constexpr int L = 12;  // RNA_HEADS: the "sequence" length

__global__ void rna_attn_fwd_kernel(const float* __restrict__ qkv, int E, bf16* __restrict__ out16, float* __restrict__ out32) {
  extern __shared__ float sh[];
  const int hd = E / L;
  float* s = sh;               // 3E
  float* att = sh + 3 * E;     // L*L
  const float* src = qkv + (long long)blockIdx.x * 3 * E;
  for (int i = threadIdx.x; i < 3 * E; i += blockDim.x) s[i] = src[i];
  __syncthreads();
  const float scale = rsqrtf((float)hd);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int p = w; p < L * L; p += nw) {
    const int i = p / L, j = p % L;
    float d = 0.f;
    for (int c = lane; c < hd; c += 32) d += s[i * hd + c] * s[E + j * hd + c];
    d = warp_sum(d);
    if (lane == 0) att[p] = d * scale;
  }
  __syncthreads();
  if (threadIdx.x < L) {
    float* r = att + threadIdx.x * L;
    float mx = r[0];
    for (int j = 1; j < L; ++j) mx = fmaxf(mx, r[j]);
    float sum = 0.f;
    for (int j = 0; j < L; ++j) { r[j] = __expf(r[j] - mx); sum += r[j]; }
    const float inv = 1.f / sum;
    for (int j = 0; j < L; ++j) r[j] *= inv;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < E; o += blockDim.x) {
    const int c = o / L, i = o % L;  // out[c*12 + i]
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < L; ++j) acc += att[i * L + j] * s[2 * E + j * hd + c];
    if (out16) out16[(long long)blockIdx.x * E + o] = __float2bfloat16(acc);
    if (out32) out32[(long long)blockIdx.x * E + o] = acc;
  }
}

__global__ void rna_attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dout, int E,
                                    bf16* __restrict__ dqkv16, float* __restrict__ dqkv32) {
  extern __shared__ float sh[];
  const int hd = E / L;
  float* s = sh;                 // 3E  q,k,v
  float* g = sh + 3 * E;         // E   do_i[c] at [i*hd + c]
  float* att = g + E;            // L*L
  float* dat = att + L * L;      // L*L
  const float* src = qkv + (long long)blockIdx.x * 3 * E;
  for (int i = threadIdx.x; i < 3 * E; i += blockDim.x) s[i] = src[i];
  for (int o = threadIdx.x; o < E; o += blockDim.x) {
    const int c = o / L, i = o % L;
    g[i * hd + c] = dout[(long long)blockIdx.x * E + o];
  }
  __syncthreads();
  const float scale = rsqrtf((float)hd);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int p = w; p < L * L; p += nw) {
    const int i = p / L, j = p % L;
    float d = 0.f, e = 0.f;
    for (int c = lane; c < hd; c += 32) {
      d += s[i * hd + c] * s[E + j * hd + c];
      e += g[i * hd + c] * s[2 * E + j * hd + c];  // dattn[i,j] = do_i . v_j
    }
    d = warp_sum(d);
    e = warp_sum(e);
    if (lane == 0) { att[p] = d * scale; dat[p] = e; }
  }
  __syncthreads();
  if (threadIdx.x < L) {
    float* r = att + threadIdx.x * L;
    float* dr = dat + threadIdx.x * L;
    float mx = r[0];
    for (int j = 1; j < L; ++j) mx = fmaxf(mx, r[j]);
    float sum = 0.f;
    for (int j = 0; j < L; ++j) { r[j] = __expf(r[j] - mx); sum += r[j]; }
    const float inv = 1.f / sum;
    float dot = 0.f;
    for (int j = 0; j < L; ++j) { r[j] *= inv; dot += r[j] * dr[j]; }
    for (int j = 0; j < L; ++j) dr[j] = r[j] * (dr[j] - dot) * scale;  // d(sim)*scale
  }
  __syncthreads();
  const long long base = (long long)blockIdx.x * 3 * E;
  for (int o = threadIdx.x; o < E; o += blockDim.x) {
    const int i = o / hd, c = o % hd;  // row i of q / k / v, component c
    float dq = 0.f, dk = 0.f, dv = 0.f;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      dq += dat[i * L + j] * s[E + j * hd + c];
      dk += dat[j * L + i] * s[j * hd + c];
      dv += att[j * L + i] * g[j * hd + c];
    }
    if (dqkv16) {
      dqkv16[base + o] = __float2bfloat16(dq);
      dqkv16[base + E + o] = __float2bfloat16(dk);
      dqkv16[base + 2 * E + o] = __float2bfloat16(dv);
    }
    if (dqkv32) {
      dqkv32[base + o] = dq;
      dqkv32[base + E + o] = dk;
      dqkv32[base + 2 * E + o] = dv;
    }
  }
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" int mirror_rna_attn_fwd(const float* qkv, int32_t B, int32_t E, void* out_bf16, float* out_f32, mirror_stream_t stream) {
  MB_CHECK_ARG(qkv && (out_bf16 || out_f32) && B > 0 && E > 0 && E % L == 0, "rna_attn_fwd: E must be a multiple of 12");
  const size_t smem = (3 * E + L * L) * sizeof(float);
  rna_attn_fwd_kernel<<<B, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(qkv, E, reinterpret_cast<bf16*>(out_bf16), out_f32);
  MB_LAUNCH_CHECK();
  return 0;
}
extern "C" int mirror_rna_attn_bwd(const float* qkv, const float* dout, int32_t B, int32_t E, void* dqkv_bf16, float* dqkv_f32,
                                   mirror_stream_t stream) {
  MB_CHECK_ARG(qkv && dout && (dqkv_bf16 || dqkv_f32) && B > 0 && E > 0 && E % L == 0, "rna_attn_bwd: E must be a multiple of 12");
  const size_t smem = (4 * E + 2 * L * L) * sizeof(float);
  rna_attn_bwd_kernel<<<B, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(qkv, dout, E, reinterpret_cast<bf16*>(dqkv_bf16), dqkv_f32);
  MB_LAUNCH_CHECK();
  return 0;
}
