// Step tail of the trainer (train_mirror.py:1221-1230 optimizer.step() with opt: adam, :1222-1229 clip_grad mode "norm",
// :1254-1256 logit_scale clamp) over FLAT fp32 parameter / gradient / moment buffers: one HBM-bound pass instead of
// ~230 per-tensor optimizer launches.  Every scalar that changes from step to step (learning rate, step count, clip
// coefficient) is read from device memory, so the launches can live inside a replayed CUDA graph.
#include "common.cuh"

namespace mb {
namespace {

// torch.optim.Adam (amsgrad off, maximize off): g += wd * p (L2) | p *= 1 - lr * wd (decoupled, AdamW);
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
            const float* __restrict__ lr_p, float b1, float b2, float eps, float wd, int decoupled, const float* __restrict__ step_p,
            const float* __restrict__ gscale_p) {
  const float lr = *lr_p, t = *step_p;
  const float gs = gscale_p ? *gscale_p : 1.f;
  const float c1 = 1.f - powf(b1, t), c2s = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / c1;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto upd = [&](float& pv, float gv, float& mv, float& vv) {
    gv *= gs;
    if (wd != 0.f) {
      if (decoupled) pv *= 1.f - lr * wd;
      else gv = fmaf(wd, pv, gv);
    }
    mv = fmaf(b1, mv, (1.f - b1) * gv);
    vv = fmaf(b2, vv, (1.f - b2) * gv * gv);
    pv -= step_size * mv / (sqrtf(vv) / c2s + eps);
  };
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    upd(pv.x, gv.x, mv.x, vv.x);
    upd(pv.y, gv.y, mv.y, vv.y);
    upd(pv.z, gv.z, mv.z, vv.z);
    upd(pv.w, gv.w, mv.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) upd(p[i], g[i], m[i], v[i]);
}

// sum of squares of a flat buffer -> *out (pre-zeroed), one atomic per block
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  __shared__ float sh[8];
  float acc = 0.f;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 x = reinterpret_cast<const float4*>(g)[i];
    acc += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) acc += g[i] * g[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    acc = warp_sum(acc);
    if (threadIdx.x == 0) atomicAdd(out, acc);
  }
}

// tail scalars in one launch: clip coefficient from the squared gradient norm (torch.nn.utils.clip_grad_norm_:
// coef = min(1, max_norm / (norm + 1e-6))), step counter += 1, optional clamp of one scalar parameter
__global__ void tail_scalars_kernel(const float* sumsq, float max_norm, float* coef, float* step, float* clamp_p, float lo, float hi) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (coef) *coef = sumsq && max_norm > 0.f ? fminf(1.f, max_norm / (sqrtf(*sumsq) + 1e-6f)) : 1.f;
    if (step) *step += 1.f;
    if (clamp_p) *clamp_p = fminf(fmaxf(*clamp_p, lo), hi);
  }
}

}  // namespace
}  // namespace mb

using namespace mb;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int mirror_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* lr, float beta1, float beta2,
                                float eps, float weight_decay, int32_t decoupled, const float* step, const float* grad_scale,
                                mirror_stream_t stream) {
  MB_CHECK_ARG(p && g && m && v && n > 0 && lr && step, "adam_step: null operand");
  MB_CHECK_ARG(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v)) & 15) == 0, "adam_step: buffers must be 16-byte aligned");
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<(int)blocks, 256, 0, STREAM>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, decoupled, step, grad_scale);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_grad_sumsq(const float* g, int64_t n, float* out, mirror_stream_t stream) {
  MB_CHECK_ARG(g && out && n > 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0, "grad_sumsq: bad args");
  MB_CUDA(cudaMemsetAsync(out, 0, sizeof(float), STREAM));
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<(int)blocks, 256, 0, STREAM>>>(g, n, out);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_tail_scalars(const float* sumsq, float max_norm, float* coef, float* step, float* clamp_param, float lo, float hi,
                                   mirror_stream_t stream) {
  tail_scalars_kernel<<<1, 32, 0, STREAM>>>(sumsq, max_norm, coef, step, clamp_param, lo, hi);
  MB_LAUNCH_CHECK();
  return 0;
}
