// Batched bf16 GEMM for sm_100a: TMA -> 128B-swizzled shared memory -> tcgen05.mma (TMEM
// accumulators, double buffered) -> fused epilogue.  Persistent, warp specialised:
//   warp 0      TMA producer (one elected lane)
//   warp 1      TMEM allocator + MMA issuer (one elected lane)
//   warps 2..9  epilogue (two warps per TMEM lane quarter, each owning half of the tile's columns):
//               residual prefetch -> tcgen05.ld -> alpha/diag/bias/act/dropout/residual/beta -> 128-bit global stores
// See include/mirror_b200.h (mirror_gemm_bf16) for the contract and the reference call sites.
#include <cudaTypedefs.h>

#include "common.cuh"
#include "ptx.cuh"

namespace mb {
namespace {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 64;           // one 128-byte swizzle row of bf16
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr int kAccStages = 2;

struct KParams {
  Epi e;
  int K, batch1, batch2, tiles_m, tiles_n, split_k;
  int a_b1, a_b2, b_b1, b_b2;  // 0 where an operand is broadcast along that batch dim (batch stride 0), else 1
};

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = kAccStages * BN;  // 256 or 512, power of two
  static constexpr int BAR_BYTES = (2 * STAGES + 2 * kAccStages) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
};

// 32 columns of one output row of the residual (or of the accumulate target), fetched BEFORE the accumulator is ready so
// that the global-load latency hides behind the tile's MMAs.  fp32: 8 x 16 B, bf16: 4 x 16 B.
struct ResBuf {
  uint4 v[8];
};

__device__ __forceinline__ void res_prefetch(const Epi& e, ResBuf& rb, int b1, int b2, int row, int col0) {
  if (e.res) {
    const long long off = b2 * e.r_bs2 + b1 * e.r_bs1 + (long long)row * e.ldr + col0;
    if (e.res_is_bf16) {
      const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(e.res) + off);
#pragma unroll
      for (int j = 0; j < 4; ++j) rb.v[j] = rp[j];
    } else {
      const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(e.res) + off);
#pragma unroll
      for (int j = 0; j < 8; ++j) rb.v[j] = rp[j];
    }
  } else if (e.beta != 0.f) {
    const uint4* rp = reinterpret_cast<const uint4*>(e.o32 + b2 * e.c32_bs2 + b1 * e.c32_bs1 + (long long)row * e.ldc32 + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) rb.v[j] = rp[j];
  }
}

__device__ __forceinline__ void epilogue_chunk_scalar(const Epi& e, const uint32_t (&acc)[32], int b1, int b2, int row, int col0) {
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (col0 + j < e.N) epi_store_scalar(e, __uint_as_float(acc[j]), b1, b2, row, col0 + j);
}

// full, 16-byte aligned 32-column chunk
__device__ __forceinline__ void epilogue_chunk_vec(const Epi& e, const uint32_t (&acc)[32], const ResBuf& rb, int b1, int b2,
                                                   int row, int col0) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = e.alpha * __uint_as_float(acc[j]);
  if (e.diag != 0.f && row >= col0 && row < col0 + 32) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += (col0 + j == row) ? e.diag : 0.f;
  }
  if (e.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(e.bias + col0 + j);
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (e.act == MIRROR_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (e.act == MIRROR_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  }
  if (e.drop_p > 0.f) {
    const uint64_t base = ((uint64_t)(b2 * e.batch1 + b1) * e.M + row) * e.N + col0;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = hash_u01(e.drop_seed, base + j) >= e.drop_p ? v[j] * e.drop_scale : 0.f;
  }
  if (e.res) {
    if (e.res_is_bf16) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rb.v[j]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __bfloat1622float2(h[t]);
          v[j * 8 + 2 * t] += e.gamma * f.x;
          v[j * 8 + 2 * t + 1] += e.gamma * f.y;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 f = *reinterpret_cast<const float4*>(&rb.v[j]);
        v[4 * j] += e.gamma * f.x; v[4 * j + 1] += e.gamma * f.y;
        v[4 * j + 2] += e.gamma * f.z; v[4 * j + 3] += e.gamma * f.w;
      }
    }
  }
  if (e.o32) {
    float4* op = reinterpret_cast<float4*>(e.o32 + b2 * e.c32_bs2 + b1 * e.c32_bs1 + (long long)row * e.ldc32 + col0);
    if (e.beta != 0.f) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 f = e.res ? op[j] : *reinterpret_cast<const float4*>(&rb.v[j]);  // prefetched unless res took the buffer
        v[4 * j] += e.beta * f.x; v[4 * j + 1] += e.beta * f.y;
        v[4 * j + 2] += e.beta * f.z; v[4 * j + 3] += e.beta * f.w;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  if (e.o16) {
    uint4* op = reinterpret_cast<uint4*>(e.o16 + b2 * e.c16_bs2 + b1 * e.c16_bs1 + (long long)row * e.ldc16 + col0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[j * 8 + 2 * t], v[j * 8 + 2 * t + 1]);
      op[j] = u;
    }
  }
}

template <int BN, int A_MN, int B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ KParams p, const int vec_ok) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * C::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kAccStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < C::STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int s = 0; s < kAccStages; ++s) {
        mbar_init(&tfull[s], 1);
        mbar_init(&tempty[s], kEpiWarps);  // one arrive per epilogue warp
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, C::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles = p.tiles_m * p.tiles_n * p.batch1 * p.batch2;
  const int total = tiles * p.split_k;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_per = (kb_total + p.split_k - 1) / p.split_k;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int tile = w / p.split_k, ks = w - tile * p.split_k;
        const int kb0 = ks * kb_per, kb1 = min(kb_total, kb0 + kb_per);
        if (kb0 >= kb1) continue;
        const int nb = tile % p.tiles_n;
        int t = tile / p.tiles_n;
        const int mb_ = t % p.tiles_m;
        t /= p.tiles_m;
        const int b1 = t % p.batch1, b2 = t / p.batch1;
        const int m0 = mb_ * BM, n0 = nb * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], C::STAGE_BYTES);
          uint8_t* a = sA + stage * C::A_BYTES;
          uint8_t* b = sB + stage * C::B_BYTES;
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_4d(&tmA, &full[stage], a + i * (BK * 128), m0 + i * 64, kb * BK, b1 * p.a_b1, b2 * p.a_b2);
          } else {
            tma_load_4d(&tmA, &full[stage], a, kb * BK, m0, b1 * p.a_b1, b2 * p.a_b2);
          }
          if (B_MN) {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_4d(&tmB, &full[stage], b + i * (BK * 128), n0 + i * 64, kb * BK, b1 * p.b_b1, b2 * p.b_b2);
          } else {
            tma_load_4d(&tmB, &full[stage], b, kb * BK, n0, b1 * p.b_b1, b2 * p.b_b2);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      // K-major: 8-row groups 1024 B apart (SBO), K advance 32 B inside the swizzle row.
      // MN-major: 64-element chunks BK*128 B apart (LBO), 8-k groups 1024 B apart (SBO), K advance 2048 B.
      constexpr uint32_t a_lbo = A_MN ? BK * 128 : 0, b_lbo = B_MN ? BK * 128 : 0;
      constexpr uint32_t a_kstep = A_MN ? (UMMA_K / 8) * 1024 : UMMA_K * 2;
      constexpr uint32_t b_kstep = B_MN ? (UMMA_K / 8) * 1024 : UMMA_K * 2;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int tile = w / p.split_k, ks = w - tile * p.split_k;
        const int kb0 = ks * kb_per, kb1 = min(kb_total, kb0 + kb_per);
        if (kb0 >= kb1) continue;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * C::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = make_smem_desc_sw128(a_addr + k * a_kstep, a_lbo, 1024);
            const uint64_t bd = make_smem_desc_sw128(b_addr + k * b_kstep, b_lbo, 1024);
            umma_f16(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[as]);  // accumulator complete
        ++it;
      }
    }
  } else {
    const int q = warp & 3;            // TMEM lane quarter this warp may read (hardware rule: warp id % 4)
    const int half = (warp - 2) >> 2;  // which half of the tile's columns this warp drains
    constexpr int NCH = BN / 64;       // 32-column chunks per thread
    const bool fast = vec_ok != 0 && !p.e.atomic;
    int it = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int tile = w / p.split_k, ks = w - tile * p.split_k;
      const int kb0 = ks * kb_per, kb1 = min(kb_total, kb0 + kb_per);
      if (kb0 >= kb1) continue;
      const int nb = tile % p.tiles_n;
      int t = tile / p.tiles_n;
      const int mb_ = t % p.tiles_m;
      t /= p.tiles_m;
      const int b1 = t % p.batch1, b2 = t / p.batch1;
      const int cbase = nb * BN + half * (BN / 2);
      const int row = mb_ * BM + q * 32 + lane;
      const bool row_ok = row < p.e.M;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ResBuf rb[2];
      if (fast && row_ok) {  // start the residual loads of the first two chunks while the MMAs of this tile still run
        if (cbase + 32 <= p.e.N) res_prefetch(p.e, rb[0], b1, b2, row, cbase);
        if (NCH > 1 && cbase + 64 <= p.e.N) res_prefetch(p.e, rb[1], b1, b2, row, cbase + 32);
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * BN + half * (BN / 2);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col0 = cbase + c * 32;
        if (col0 < p.e.N) {  // warp-uniform
          uint32_t acc[32];
          tmem_ld_32x32(taddr + c * 32, acc);
          tmem_ld_wait();
          if (row_ok) {
            if (fast && col0 + 32 <= p.e.N) epilogue_chunk_vec(p.e, acc, rb[c & 1], b1, b2, row, col0);
            else epilogue_chunk_scalar(p.e, acc, b1, b2, row, col0);
          }
          if (c + 2 < NCH && fast && row_ok && col0 + 96 <= p.e.N) res_prefetch(p.e, rb[c & 1], b1, b2, row, col0 + 64);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      ++it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  }
  return fn;
}

// Operand is logically [MN, K]; K-major: K contiguous.  MN-major: MN contiguous.
int make_operand_map(CUtensorMap* map, const void* ptr, int mn_major, long long mn, long long k, long long ld,
                     long long bs1, int batch1, long long bs2, int batch2, int box_mn) {
  auto enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return MB_ERR_DRIVER;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 7) || (batch1 > 1 && (bs1 & 7)) || (batch2 > 1 && (bs2 & 7))) {
    set_error("gemm operand misaligned: ptr=%p ld=%lld bs1=%lld bs2=%lld (need 16 B / multiples of 8 elements)", ptr, ld,
              bs1, bs2);
    return MB_ERR_ALIGN;
  }
  const cuuint64_t inner = mn_major ? mn : k, outer = mn_major ? k : mn;
  cuuint64_t dims[4] = {inner, outer, (cuuint64_t)batch1, (cuuint64_t)batch2};
  const cuuint64_t fallback = (cuuint64_t)ld * 2;
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, batch1 > 1 ? (cuuint64_t)bs1 * 2 : fallback,
                           batch2 > 1 ? (cuuint64_t)bs2 * 2 : fallback};
  cuuint32_t box[4] = {64, (cuuint32_t)(mn_major ? BK : box_mn), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): dims=[%llu,%llu,%d,%d] ld=%lld", (int)r, (unsigned long long)inner,
              (unsigned long long)outer, batch1, batch2, ld);
    return MB_ERR_DRIVER;
  }
  return 0;
}

template <int BN, int A_MN, int B_MN>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const KParams& p, int vec_ok, cudaStream_t stream) {
  using C = Cfg<BN>;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN>;
  static bool configured = false;  // benign race: attribute set is idempotent
  if (!configured) {
    MB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  const long long total = (long long)p.tiles_m * p.tiles_n * p.batch1 * p.batch2 * p.split_k;
  const int grid = (int)(total < num_sms() ? total : num_sms());
  kern<<<grid, kThreads, C::SMEM_BYTES, stream>>>(tmA, tmB, p, vec_ok);
  MB_LAUNCH_CHECK();
  return 0;
}

int fill_epi(const mirror_gemm_args* g, Epi* e) {
  MB_CHECK_ARG(g && g->a && g->b, "gemm: null operand");
  MB_CHECK_ARG(g->M > 0 && g->N > 0 && g->K > 0 && g->batch1 > 0 && g->batch2 > 0, "gemm: bad shape M=%d N=%d K=%d", g->M,
               g->N, g->K);
  MB_CHECK_ARG(g->out_f32 || g->out_bf16, "gemm: no output");
  MB_CHECK_ARG(g->beta == 0.f || g->out_f32, "gemm: beta needs out_f32");
  MB_CHECK_ARG(g->drop_p >= 0.f && g->drop_p < 1.f, "gemm: drop_p out of range");
  MB_CHECK_ARG(g->split_k <= 1 || (g->out_f32 && !g->out_bf16 && !g->bias && !g->res && g->act == 0 && g->drop_p == 0.f && g->diag == 0.f),
               "gemm: split_k supports only alpha and an fp32 accumulate target");
  e->M = g->M; e->N = g->N; e->batch1 = g->batch1;
  e->alpha = g->alpha; e->diag = g->diag; e->bias = g->bias; e->act = g->act;
  e->drop_p = g->drop_p; e->drop_scale = 1.f / (1.f - g->drop_p); e->drop_seed = g->drop_seed;
  e->res = g->res; e->res_is_bf16 = g->res_is_bf16; e->gamma = g->gamma;
  e->ldr = g->ldr; e->r_bs1 = g->r_bs1; e->r_bs2 = g->r_bs2;
  e->beta = g->beta;
  e->o32 = g->out_f32; e->ldc32 = g->ldc32; e->c32_bs1 = g->c32_bs1; e->c32_bs2 = g->c32_bs2;
  e->o16 = reinterpret_cast<bf16*>(g->out_bf16); e->ldc16 = g->ldc16; e->c16_bs1 = g->c16_bs1; e->c16_bs2 = g->c16_bs2;
  e->atomic = g->split_k > 1;
  return 0;
}

bool epi_vec_ok(const mirror_gemm_args* g) {
  auto al = [](const void* p, int bytes) { return (reinterpret_cast<uintptr_t>(p) % bytes) == 0; };
  bool ok = true;
  if (g->out_f32) ok = ok && al(g->out_f32, 16) && g->ldc32 % 4 == 0 && g->c32_bs1 % 4 == 0 && g->c32_bs2 % 4 == 0;
  if (g->out_bf16) ok = ok && al(g->out_bf16, 16) && g->ldc16 % 8 == 0 && g->c16_bs1 % 8 == 0 && g->c16_bs2 % 8 == 0;
  if (g->bias) ok = ok && al(g->bias, 16);
  if (g->res) {
    const int q = g->res_is_bf16 ? 8 : 4;
    ok = ok && al(g->res, 16) && g->ldr % q == 0 && g->r_bs1 % q == 0 && g->r_bs2 % q == 0;
  }
  return ok;
}

// ---------------------------------------------------------------- SIMT cross-check
__global__ void gemm_simt_kernel(const bf16* __restrict__ A, const bf16* __restrict__ B, int a_mn, int b_mn,
                                 long long lda, long long ldb, long long a_bs1, long long a_bs2, long long b_bs1,
                                 long long b_bs2, int K, int batch2, const __grid_constant__ Epi e) {
  const long long per = (long long)e.M * e.N;
  const long long total = per * e.batch1 * batch2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % e.N);
    const int row = (int)((i / e.N) % e.M);
    const int b = (int)(i / per);
    const int b1 = b % e.batch1, b2 = b / e.batch1;
    const bf16* a = A + b2 * a_bs2 + b1 * a_bs1;
    const bf16* bb = B + b2 * b_bs2 + b1 * b_bs1;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      const float x = __bfloat162float(a_mn ? a[(long long)k * lda + row] : a[(long long)row * lda + k]);
      const float y = __bfloat162float(b_mn ? bb[(long long)k * ldb + col] : bb[(long long)col * ldb + k]);
      acc = fmaf(x, y, acc);
    }
    epi_store_scalar(e, acc, b1, b2, row, col);
  }
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" int mirror_gemm_bf16(const mirror_gemm_args* g, mirror_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KParams p;
  int rc = fill_epi(g, &p.e);
  if (rc) return rc;
  // N tile: 256 when it divides N (or N is large), otherwise 128 (less padding waste for N = 96, 384, ...).
  const int BN = (g->N % 256 == 0 || g->N >= 1024) ? 256 : 128;
  p.K = g->K;
  p.batch1 = g->batch1;
  p.batch2 = g->batch2;
  p.tiles_m = (g->M + BM - 1) / BM;
  p.tiles_n = (g->N + BN - 1) / BN;
  p.split_k = g->split_k > 1 ? g->split_k : 1;
  // a batch stride of 0 broadcasts that operand (e.g. one weight matrix for every slide)
  p.a_b1 = (g->batch1 > 1 && g->a_bs1 == 0) ? 0 : 1;
  p.a_b2 = (g->batch2 > 1 && g->a_bs2 == 0) ? 0 : 1;
  p.b_b1 = (g->batch1 > 1 && g->b_bs1 == 0) ? 0 : 1;
  p.b_b2 = (g->batch2 > 1 && g->b_bs2 == 0) ? 0 : 1;
  CUtensorMap tmA, tmB;
  rc = make_operand_map(&tmA, g->a, g->a_mn_major, g->M, g->K, g->lda, g->a_bs1, p.a_b1 ? g->batch1 : 1, g->a_bs2,
                        p.a_b2 ? g->batch2 : 1, BM);
  if (rc) return rc;
  rc = make_operand_map(&tmB, g->b, g->b_mn_major, g->N, g->K, g->ldb, g->b_bs1, p.b_b1 ? g->batch1 : 1, g->b_bs2,
                        p.b_b2 ? g->batch2 : 1, BN);
  if (rc) return rc;
  const int vec = epi_vec_ok(g) ? 1 : 0;
  const int key = (BN == 256 ? 4 : 0) | (g->a_mn_major ? 2 : 0) | (g->b_mn_major ? 1 : 0);
  switch (key) {
    case 0: return launch<128, 0, 0>(tmA, tmB, p, vec, stream);
    case 1: return launch<128, 0, 1>(tmA, tmB, p, vec, stream);
    case 2: return launch<128, 1, 0>(tmA, tmB, p, vec, stream);
    case 3: return launch<128, 1, 1>(tmA, tmB, p, vec, stream);
    case 4: return launch<256, 0, 0>(tmA, tmB, p, vec, stream);
    case 5: return launch<256, 0, 1>(tmA, tmB, p, vec, stream);
    case 6: return launch<256, 1, 0>(tmA, tmB, p, vec, stream);
    default: return launch<256, 1, 1>(tmA, tmB, p, vec, stream);
  }
}

extern "C" int mirror_gemm_bf16_simt(const mirror_gemm_args* g, mirror_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  Epi e;
  int rc = fill_epi(g, &e);
  if (rc) return rc;
  const long long total = (long long)g->M * g->N * g->batch1 * g->batch2;
  const int block = 256;
  const int grid = (int)((total + block - 1) / block < 65535 * 4 ? (total + block - 1) / block : 65535 * 4);
  gemm_simt_kernel<<<grid, block, 0, stream>>>(reinterpret_cast<const bf16*>(g->a), reinterpret_cast<const bf16*>(g->b),
                                               g->a_mn_major, g->b_mn_major, g->lda, g->ldb, g->a_bs1, g->a_bs2, g->b_bs1,
                                               g->b_bs2, g->K, g->batch2, e);
  MB_LAUNCH_CHECK();
  return 0;
}
