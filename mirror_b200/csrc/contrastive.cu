// Fused contrastive loss for sm_100a (ClipLoss, losses/mirror_loss.py:37-52; InfoNCE implicit negatives,
// losses/info_nce.py:144-164).  The B x B logits never exist in HBM:
//
//   stats pass   S = X Y^T tile by tile in TMEM -> online row log-sum-exp + the positive's logit    (MODE 0)
//   grad  pass   S recomputed -> G = dLoss/dS in registers -> bf16 tile in shared memory -> second tcgen05.mma
//                dX += G Y accumulating in TMEM across all column blocks                             (MODE 1)
//
// Both directions of the symmetric loss are the same kernel with the operands swapped (S^T = Y X^T), and so is the
// global-negative form (rows = local samples, columns = the all-gathered batch, `diag0` = offset of the positives).
//   loss_i = w_r (lse_r[i] - L_ii) + w_c (lse_c[i] - L_ii),   L = s X Y^T
//   G_ij   = s ( a_r[i] e^{L_ij - lse_r[i]} + a_c[j] e^{L_ij - lse_c[j]} - [j == i + diag0] (a_r[i] + a_c[j]) )
//   dX     = G Y ;   d s += sum_ij (a_r[i] e^{L_ij - lse_r[i]} - [diag] a_r[i]) (X Y^T)_ij     (the row part only: the column
//   part is the row part of the swapped launch)
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-5 softmax / epilogue
// (thread = row of the 128-row tile).  "PRECISE": operands are bf16 split-3 ([hi|lo|hi] x [hi|hi|lo], K = 3 D), and G is
// split into hi + lo as well, so the small-batch (B <= 1024) loss and gradients are fp32-grade.
#include "tile.cuh"

namespace mb {
namespace {

constexpr int kThreadsC = 192;
constexpr int kTileBytes = 128 * 128;  // one [128 x 64] bf16 tile

struct CParams {
  int Br, Bc, K, D, E;  // rows, columns, contraction length of S, value width (multiple of 64), stored width of dX (<= D)
  int lo_off;           // PRECISE: first column of the "lo" block inside Y
  int diag0;            // positive of row i is column i + diag0
  int nb, nsplit;       // column blocks of 128; MODE 0: blocks are divided over nsplit CTAs per row tile
  const float* scale;
  const float* lse_r;   // [Br] natural log
  const float* lse_c;   // [Bc]
  const float* a_r;     // [Br]
  const float* a_c;     // [Bc] or NULL (one-sided loss)
  float* dx;            // [Br, E] f32
  long long lddx;
  float* dscale;        // accumulated, may be NULL
  float2* part;         // MODE 0: [nsplit, Br] (max2, sum)
  float* diag;          // MODE 0: [Br] scaled logit of the positive
};

template <int MODE, int DC, bool PRECISE>
struct CCfg {
  static constexpr int NS = MODE == 0 ? 4 : (PRECISE ? 2 : 3);
  static constexpr int RING = NS * 2 * kTileBytes;
  static constexpr int YV = MODE == 0 ? 0 : (DC / 64) * kTileBytes * (PRECISE ? 2 : 1);
  static constexpr int G = MODE == 0 ? 0 : 2 * kTileBytes * (PRECISE ? 2 : 1);
  static constexpr int CS = 2 * 128 * 16;
  static constexpr int BARS = (2 * NS + 16) * 8 + 16;
  static constexpr int SMEM = RING + YV + G + CS + BARS + 1024;
  static constexpr int TMEM_COLS = MODE == 0 ? 256 : 512;
};

template <int MODE, int DC, bool PRECISE>
__global__ void __launch_bounds__(kThreadsC, 1)
contrastive_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CParams p) {
  using C = CCfg<MODE, DC, PRECISE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* yv = ring + C::RING;
  uint8_t* gt = yv + C::YV;
  float4* colstat = reinterpret_cast<float4*>(gt + C::G);
  uint64_t* full = reinterpret_cast<uint64_t*>(gt + C::G + C::CS);
  uint64_t* empty = full + C::NS;
  uint64_t* s_full = empty + C::NS;   // [2]
  uint64_t* s_empty = s_full + 2;     // [2]
  uint64_t* g_full = s_empty + 2;
  uint64_t* g_empty = g_full + 1;
  uint64_t* yv_full = g_empty + 1;
  uint64_t* yv_empty = yv_full + 1;
  uint64_t* acc_full = yv_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work decomposition: MODE 0: blockIdx = (row tile, column split); MODE 1: blockIdx = (row tile, D chunk)
  const int nsec = MODE == 0 ? p.nsplit : (p.D + DC - 1) / DC;
  const int mt = blockIdx.x / nsec, sec = blockIdx.x % nsec;
  const int m0 = mt * 128;
  int jb0 = 0, jb1 = p.nb;
  if (MODE == 0) {
    const int per = (p.nb + p.nsplit - 1) / p.nsplit;
    jb0 = sec * per;
    jb1 = min(p.nb, jb0 + per);
  }
  const int d0 = MODE == 1 ? sec * DC : 0;
  const int n_valid = MODE == 1 ? min(DC, p.D - d0) : 0;  // multiple of 64
  const int nkb = (p.K + 63) / 64;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < C::NS; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&s_full[s], 1);
        mbar_init(&s_empty[s], 4);
      }
      mbar_init(g_full, 4);
      mbar_init(g_empty, 1);
      mbar_init(yv_full, 1);
      mbar_init(yv_empty, 1);
      mbar_init(acc_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, C::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------- TMA producer
    if (elect_one() && jb0 < jb1) {
      int stage = 0;
      uint32_t phase = 0;
      auto ring_load = [&](int jb) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], 2 * kTileBytes);
          tma_load_2d(&tmX, &full[stage], ring + stage * 2 * kTileBytes, kb * 64, m0);
          tma_load_2d(&tmY, &full[stage], ring + stage * 2 * kTileBytes + kTileBytes, kb * 64, jb * 128);
          if (++stage == C::NS) { stage = 0; phase ^= 1; }
        }
      };
      ring_load(jb0);
      for (int jb = jb0; jb < jb1; ++jb) {
        if (jb + 1 < jb1) ring_load(jb + 1);
        if (MODE == 1) {
          const int it = jb - jb0;
          mbar_wait(yv_empty, (it & 1) ^ 1);
          const int nch = n_valid / 64;
          mbar_expect_tx(yv_full, nch * kTileBytes * (PRECISE ? 2 : 1));
          for (int c = 0; c < nch; ++c) {
            tma_load_2d(&tmY, yv_full, yv + c * kTileBytes, d0 + c * 64, jb * 128);
            if (PRECISE) tma_load_2d(&tmY, yv_full, yv + (DC / 64 + c) * kTileBytes, p.lo_off + d0 + c * 64, jb * 128);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------- MMA issuer
    if (elect_one() && jb0 < jb1) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_dx = make_idesc_bf16(128, MODE == 1 ? n_valid : 64, 0, 1);
      int stage = 0;
      uint32_t phase = 0;
      auto issue_s = [&](int jb) {
        const int it = jb - jb0, sb = it & 1;
        mbar_wait(&s_empty[sb], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + sb * 128;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a = smem_u32(ring + stage * 2 * kTileBytes), b = a + kTileBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, desc_kmajor(a, k), desc_kmajor(b, k), idesc_s, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (++stage == C::NS) { stage = 0; phase ^= 1; }
        }
        umma_commit(&s_full[sb]);
      };
      issue_s(jb0);
      for (int jb = jb0; jb < jb1; ++jb) {
        if (jb + 1 < jb1) issue_s(jb + 1);
        if (MODE == 1) {
          const int it = jb - jb0;
          mbar_wait(g_full, it & 1);
          mbar_wait(yv_full, it & 1);
          tc_fence_after();
          const uint32_t d = tmem_base + 256;
          const uint32_t g_hi = smem_u32(gt), g_lo = g_hi + 2 * kTileBytes;
          const uint32_t y_hi = smem_u32(yv), y_lo = y_hi + (DC / 64) * kTileBytes;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {  // contraction over the 128 columns j of this block
            const uint32_t ga = g_hi + (kk >> 2) * kTileBytes, gl = g_lo + (kk >> 2) * kTileBytes;
            umma_f16(d, desc_kmajor(ga, kk & 3), desc_mnmajor(y_hi, kTileBytes, kk), idesc_dx, (it > 0 || kk > 0) ? 1u : 0u);
            if (PRECISE) {
              umma_f16(d, desc_kmajor(gl, kk & 3), desc_mnmajor(y_hi, kTileBytes, kk), idesc_dx, 1u);
              umma_f16(d, desc_kmajor(ga, kk & 3), desc_mnmajor(y_lo, kTileBytes, kk), idesc_dx, 1u);
            }
          }
          umma_commit(g_empty);
          umma_commit(yv_empty);
        }
      }
      umma_commit(acc_full);
    }
  } else {
    // ------------------------------------------------------------------------------------------- softmax / epilogue warps
    const int q = warp & 3, t = threadIdx.x - 64;
    const int rl = q * 32 + lane, row = m0 + rl;
    const bool row_ok = row < p.Br;
    const float sc = *p.scale, s2 = sc * 1.4426950408889634f;
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
    const int jdiag = row + p.diag0;  // global column of this row's positive
    float run_m = -INFINITY, run_s = 0.f, dval = 0.f;
    float lr2 = INFINITY, ar = 0.f, ds_acc = 0.f;
    if (MODE == 1 && row_ok) {
      lr2 = p.lse_r[row] * 1.4426950408889634f;
      ar = p.a_r[row] * sc;
    }
    const bool has_c = MODE == 1 && p.a_c != nullptr;
    for (int jb = jb0; jb < jb1; ++jb) {
      const int it = jb - jb0, sb = it & 1;
      {  // this block's column statistics: (bias, lse_c in base 2, a_c * s)
        const int j = jb * 128 + t;
        const bool ok = j < p.Bc;
        float4 cs = make_float4(ok ? 0.f : -INFINITY, INFINITY, 0.f, 0.f);
        if (has_c && ok) {
          cs.y = p.lse_c[j] * 1.4426950408889634f;
          cs.z = p.a_c[j] * sc;
        }
        colstat[sb * 128 + t] = cs;
      }
      named_bar_sync(1, 128);
      mbar_wait(&s_full[sb], (it >> 1) & 1);
      tc_fence_after();
      const float4* cs = colstat + sb * 128;
      const int jd = jdiag - jb * 128;  // diagonal column inside this block, if in [0,128)
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t acc[32];
        tmem_ld_32x32(lane_base + sb * 128 + ch * 32, acc);
        tmem_ld_wait();
        if (MODE == 0) {
          float v[32];
          float cm = -INFINITY;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            v[e] = fmaf(s2, __uint_as_float(acc[e]), cs[ch * 32 + e].x);
            cm = fmaxf(cm, v[e]);
          }
          if (cm > run_m) {
            run_s *= fast_exp2(run_m - cm);
            run_m = cm;
          }
          if (run_m > -INFINITY) {
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 32; ++e) s4[e & 3] += fast_exp2(v[e] - run_m);
            run_s += (s4[0] + s4[1]) + (s4[2] + s4[3]);
          }
          if (jd >= ch * 32 && jd < ch * 32 + 32) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (e == jd - ch * 32) dval = sc * __uint_as_float(acc[e]);
          }
        } else {
          float g[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float4 c4 = cs[ch * 32 + e];
            const float raw = __uint_as_float(acc[e]);
            const float l2 = fmaf(s2, raw, c4.x);
            const float pr = ar * fast_exp2(l2 - lr2);
            ds_acc = fmaf(pr, raw, ds_acc);
            g[e] = has_c ? fmaf(c4.z, fast_exp2(l2 - c4.y), pr) : pr;
          }
          if (jd >= ch * 32 && jd < ch * 32 + 32) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (e == jd - ch * 32) {
                g[e] -= ar + cs[ch * 32 + e].z;
                ds_acc = fmaf(-ar, __uint_as_float(acc[e]), ds_acc);
              }
          }
          if (ch == 0) mbar_wait(g_empty, (it & 1) ^ 1);  // dX of the previous block has consumed the G tile
          uint4 hi[4], lo[4];
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            uint32_t w[4], wl[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float a = g[pc * 8 + 2 * u], b = g[pc * 8 + 2 * u + 1];
              w[u] = pack_bf16(a, b);
              if (PRECISE) wl[u] = pack_bf16(a - __uint_as_float(w[u] << 16), b - __uint_as_float(w[u] & 0xffff0000u));
            }
            hi[pc] = make_uint4(w[0], w[1], w[2], w[3]);
            if (PRECISE) lo[pc] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
          }
          const uint32_t gbase = smem_u32(gt) + (ch >> 1) * kTileBytes;
          tile_store_32cols(gbase, rl, ch & 1, hi);
          if (PRECISE) tile_store_32cols(gbase + 2 * kTileBytes, rl, ch & 1, lo);
        }
      }
      tc_fence_before();
      if (MODE == 1) fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_empty[sb]);
        if (MODE == 1) mbar_arrive(g_full);
      }
    }
    if (MODE == 0) {
      if (row_ok && jb0 < jb1) {
        p.part[(long long)sec * p.Br + row] = make_float2(run_m, run_s);
        if (jdiag >= jb0 * 128 && jdiag < jb1 * 128) p.diag[row] = dval;
      }
    } else if (jb0 < jb1) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < n_valid / 32; ++ch) {
        uint32_t acc[32];
        tmem_ld_32x32(lane_base + 256 + ch * 32, acc);
        tmem_ld_wait();
        const int c0 = d0 + ch * 32;
        if (row_ok && c0 < p.E) {
          float* o = p.dx + (long long)row * p.lddx + c0;
          if (c0 + 32 <= p.E && (p.lddx & 3) == 0) {
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              *reinterpret_cast<float4*>(o + e) = make_float4(__uint_as_float(acc[e]), __uint_as_float(acc[e + 1]), __uint_as_float(acc[e + 2]),
                                                             __uint_as_float(acc[e + 3]));
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (c0 + e < p.E) o[e] = __uint_as_float(acc[e]);
          }
        }
      }
      if (p.dscale && sec == 0) {
        const float v = warp_sum(ds_acc);
        if (lane == 0) atomicAdd(p.dscale, v / sc);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// lse[i] = ln sum_j e^{L_ij} from the per-split (max, sum) partials (base-2 domain)
__global__ void contrastive_merge_kernel(const float2* __restrict__ part, int nsplit, int Br, float* __restrict__ lse) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Br) return;
  float m = -INFINITY;
  for (int s = 0; s < nsplit; ++s) m = fmaxf(m, part[(long long)s * Br + i].x);
  float sum = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float2 v = part[(long long)s * Br + i];
    if (v.x > -INFINITY) sum += v.y * exp2f(v.x - m);
  }
  lse[i] = (m + log2f(sum)) * 0.6931471805599453f;
}

// per-sample losses and their reduction: loss_i = w_r (lse_r[i] - d_i) + w_c (lse_c[i] - d_i); out = mult * sum_i loss_i
__global__ void contrastive_loss_kernel(const float* __restrict__ lse_r, const float* __restrict__ lse_c, const float* __restrict__ diag, int B,
                                        float w_r, float w_c, float mult, float* __restrict__ per_sample, float* __restrict__ out) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float d = diag[i];
    float l = w_r * (lse_r[i] - d);
    if (w_c != 0.f) l += w_c * (lse_c[i] - d);
    if (per_sample) per_sample[i] = l;
    acc += l;
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0 && out) *out = acc * mult;
}

// a_r[i] = g_i * w_r * mult, a_c[i] = g_i * w_c * mult with g a device scalar (g_stride 0) or a per-sample vector (1)
__global__ void contrastive_coef_kernel(const float* __restrict__ g, int g_stride, int B, float w_r, float w_c, float mult,
                                        float* __restrict__ a_r, float* __restrict__ a_c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float v = g[(long long)i * g_stride] * mult;
  a_r[i] = v * w_r;
  if (a_c) a_c[i] = v * w_c;
}

template <int MODE, int DC, bool PRECISE>
int launch_contrastive(const CUtensorMap& tmX, const CUtensorMap& tmY, const CParams& p, int grid, cudaStream_t stream) {
  using C = CCfg<MODE, DC, PRECISE>;
  auto kern = contrastive_kernel<MODE, DC, PRECISE>;
  static DeviceOnce once;
  if (once.first()) MB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  kern<<<grid, kThreadsC, C::SMEM, stream>>>(tmX, tmY, p);
  MB_LAUNCH_CHECK();
  return 0;
}

}  // namespace
}  // namespace mb

using namespace mb;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int mirror_contrastive_nsplit(int32_t Br, int32_t Bc) {
  const int tiles = (Br + 127) / 128, nb = (Bc + 127) / 128;
  int ns = num_sms() / (tiles > 0 ? tiles : 1);
  if (ns < 1) ns = 1;
  if (ns > nb) ns = nb;
  if (ns > 16) ns = 16;
  return ns;
}

extern "C" int mirror_contrastive_stats(const void* x, int64_t ldx, const void* y, int64_t ldy, int32_t Br, int32_t Bc, int32_t K,
                                        const float* scale, int32_t diag0, float* part, int32_t nsplit, float* lse, float* diag,
                                        mirror_stream_t stream) {
  MB_CHECK_ARG(x && y && scale && part && lse && diag && Br > 0 && Bc > 0 && K > 0 && nsplit >= 1, "contrastive_stats: bad arguments");
  CUtensorMap tmX, tmY;
  int rc = make_map_2d(&tmX, x, Br, K, ldx, 128);
  if (rc) return rc;
  rc = make_map_2d(&tmY, y, Bc, K, ldy, 128);
  if (rc) return rc;
  CParams p = {};
  p.Br = Br; p.Bc = Bc; p.K = K; p.D = 0; p.E = 0; p.diag0 = diag0;
  p.nb = (Bc + 127) / 128;
  p.nsplit = nsplit < p.nb ? nsplit : p.nb;
  {  // every split must own at least one column block (the merge reads all of them)
    const int per = (p.nb + p.nsplit - 1) / p.nsplit;
    p.nsplit = (p.nb + per - 1) / per;
  }
  p.scale = scale;
  p.part = reinterpret_cast<float2*>(part);
  p.diag = diag;
  const int tiles = (Br + 127) / 128;
  rc = launch_contrastive<0, 256, false>(tmX, tmY, p, tiles * p.nsplit, STREAM);
  if (rc) return rc;
  contrastive_merge_kernel<<<(Br + 127) / 128, 128, 0, STREAM>>>(p.part, p.nsplit, Br, lse);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_contrastive_grad(const void* x, int64_t ldx, const void* y, int64_t ldy, int32_t Br, int32_t Bc, int32_t K,
                                       int32_t D, int32_t E, int32_t precise, int32_t lo_off, const float* scale, int32_t diag0,
                                       const float* lse_r, const float* lse_c, const float* a_r, const float* a_c, float* dx,
                                       int64_t lddx, float* dscale, mirror_stream_t stream) {
  MB_CHECK_ARG(x && y && scale && lse_r && lse_c && a_r && dx && Br > 0 && Bc > 0 && K > 0, "contrastive_grad: bad arguments");
  MB_CHECK_ARG(D > 0 && D % 64 == 0 && E > 0 && E <= D, "contrastive_grad: D must be a multiple of 64 and E <= D (D=%d E=%d)", D, E);
  MB_CHECK_ARG(precise ? (K == 3 * D && lo_off >= D && lo_off + D <= K) : (K == D), "contrastive_grad: K does not match D (K=%d D=%d)", K, D);
  CUtensorMap tmX, tmY;
  int rc = make_map_2d(&tmX, x, Br, K, ldx, 128);
  if (rc) return rc;
  rc = make_map_2d(&tmY, y, Bc, K, ldy, 128);
  if (rc) return rc;
  CParams p = {};
  p.Br = Br; p.Bc = Bc; p.K = K; p.D = D; p.E = E; p.lo_off = lo_off; p.diag0 = diag0;
  p.nb = (Bc + 127) / 128;
  p.nsplit = 1;
  p.scale = scale; p.lse_r = lse_r; p.lse_c = lse_c; p.a_r = a_r; p.a_c = a_c;
  p.dx = dx; p.lddx = lddx; p.dscale = dscale;
  const int tiles = (Br + 127) / 128;
  if (precise) return launch_contrastive<1, 128, true>(tmX, tmY, p, tiles * ((D + 127) / 128), STREAM);
  return launch_contrastive<1, 256, false>(tmX, tmY, p, tiles * ((D + 255) / 256), STREAM);
}

extern "C" int mirror_contrastive_loss(const float* lse_r, const float* lse_c, const float* diag, int32_t B, float w_r, float w_c,
                                       float mult, float* per_sample, float* out, mirror_stream_t stream) {
  MB_CHECK_ARG(lse_r && diag && B > 0 && (w_c == 0.f || lse_c) && (per_sample || out), "contrastive_loss: bad arguments");
  contrastive_loss_kernel<<<1, 256, 0, STREAM>>>(lse_r, lse_c, diag, B, w_r, w_c, mult, per_sample, out);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_contrastive_coef(const float* g, int32_t g_stride, int32_t B, float w_r, float w_c, float mult, float* a_r,
                                       float* a_c, mirror_stream_t stream) {
  MB_CHECK_ARG(g && a_r && B > 0 && (g_stride == 0 || g_stride == 1), "contrastive_coef: bad arguments");
  contrastive_coef_kernel<<<(B + 255) / 256, 256, 0, STREAM>>>(g, g_stride, B, w_r, w_c, mult, a_r, a_c);
  MB_LAUNCH_CHECK();
  return 0;
}
