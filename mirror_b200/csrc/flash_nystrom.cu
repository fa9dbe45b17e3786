// Flash-style fused softmax products of the Nystrom attention (nystrom_attention forward, call site models/mirror.py:299-312;
// SURVEY.md §2.1 kernels K-A / K-C): the [rows x keys] probability matrices
//     attn1 = softmax(q k_l^T)   [n x m]   ->  out = attn1 W + res_conv(v)
//     attn3 = softmax(q_l k^T)   [m x n]   ->  kv  = attn3 v
// never exist in HBM.  Forward: O = softmax(alpha X Y^T) V with the logits tile in TMEM, probabilities written as a bf16
// shared-memory tile and consumed by a second tcgen05.mma; two passes over the (cheap, K = head_dim) logits -- row maxima
// first, then exp / row sums / P V -- so the accumulator never has to be rescaled.  Backward (flash_bwd below): logits and
// dP = dO V^T are recomputed per 128 x 64 block, dS = alpha P (dP - D) goes through shared memory into
// dX += dS Y (row-stationary) or dY += dS^T X, dV += P^T dO (column-stationary).
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2..5 softmax / epilogue
// (thread = row of the 128-row tile).  Persistent: a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
#include "tile.cuh"

namespace mb {
namespace {

constexpr int kThreadsF = 192;
constexpr int kTB = 128 * 128;  // bytes of one [128 rows x 64 cols] bf16 tile
constexpr float kLog2e = 1.4426950408889634f;

struct FwdParams {
  int R, C;        // softmax rows / keys per (batch, head)
  int d, dpad;     // head dim, head dim rounded up to 16
  int heads, batch;
  float alpha;
  bf16* out;       // [batch, heads, R, d] through (o_bs, o_hs, o_ld)
  long long o_ld, o_hs, o_bs;
  const bf16* res;  // optional residual added to out, same indexing through (r_bs, r_hs, r_ld)
  long long r_ld, r_hs, r_bs;
  float* lse2;     // [batch, heads, R]: log2 sum_j 2^(alpha log2e S_ij)  (base-2 log-sum-exp of the scaled logits)
  int tiles_r, nb;  // row tiles of 128, key blocks of 128
};

struct FwdSmem {
  static constexpr int X = 2 * 2 * kTB;   // two row tiles in flight x two K blocks
  static constexpr int Y = 2 * 2 * kTB;   // ring of two key blocks x two K blocks
  static constexpr int V = 2 * kTB;       // one key block of values: two 64-column chunks [128 keys x 64]
  static constexpr int P = 2 * kTB;       // probabilities [128 rows x 128 keys] as two K blocks
  static constexpr int BARS = 32 * 8 + 16;
  static constexpr int TOTAL = X + Y + V + P + BARS + 1024;
};

__global__ void __launch_bounds__(kThreadsF, 1)
flash_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sY = sX + FwdSmem::X;
  uint8_t* sV = sY + FwdSmem::Y;
  uint8_t* sP = sV + FwdSmem::V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + FwdSmem::P);
  uint64_t* x_full = bars;          // [2]
  uint64_t* x_empty = bars + 2;     // [2]
  uint64_t* y_full = bars + 4;      // [2]
  uint64_t* y_empty = bars + 6;     // [2]
  uint64_t* s_full = bars + 8;      // [2]
  uint64_t* s_empty = bars + 10;    // [2]
  uint64_t* v_full = bars + 12;
  uint64_t* v_empty = bars + 13;
  uint64_t* p_full = bars + 14;
  uint64_t* p_empty = bars + 15;
  uint64_t* o_full = bars + 16;
  uint64_t* o_empty = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.d + 63) / 64;      // 64-column K blocks of the head dim
  const int nks = (p.d + 15) / 16;      // UMMA k-steps of the logits products
  const int nvc = (p.dpad + 63) / 64;   // 64-column chunks of the value tile
  const int total = p.batch * p.heads * p.tiles_r;
  const int ns = 2 * p.nb;              // S tiles per row tile: pass 0 (row maxima) then pass 1 (exp, P V)

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&x_full[i], 1);
        mbar_init(&x_empty[i], 1);
        mbar_init(&y_full[i], 1);
        mbar_init(&y_empty[i], 1);
        mbar_init(&s_full[i], 1);
        mbar_init(&s_empty[i], 4);
      }
      mbar_init(v_full, 1);
      mbar_init(v_empty, 1);
      mbar_init(p_full, 4);
      mbar_init(p_empty, 1);
      mbar_init(o_full, 1);
      mbar_init(o_empty, 4);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColO = 256;  // TMEM: logits buffers at columns [0,128) and [128,256), output accumulator at [256, 256+dpad)

  if (warp == 0) {
    // ----------------------------------------------------------------------------------------------- TMA producer
    if (elect_one()) {
      uint32_t ycount = 0, vcount = 0, ti = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
        const int rt = w % p.tiles_r;
        const int bh = w / p.tiles_r;
        const int h = bh % p.heads, b = bh / p.heads;
        const int xb = ti & 1;
        mbar_wait(&x_empty[xb], ((ti >> 1) & 1) ^ 1);
        mbar_expect_tx(&x_full[xb], nkb * kTB);
        for (int kb = 0; kb < nkb; ++kb) tma_load_4d(&tmX, &x_full[xb], sX + (xb * 2 + kb) * kTB, kb * 64, rt * 128, h, b);
        auto y_load = [&](int s) {
          const int jb = s % p.nb, st = ycount & 1;
          mbar_wait(&y_empty[st], ((ycount >> 1) & 1) ^ 1);
          mbar_expect_tx(&y_full[st], nkb * kTB);
          for (int kb = 0; kb < nkb; ++kb) tma_load_4d(&tmY, &y_full[st], sY + (st * 2 + kb) * kTB, kb * 64, jb * 128, h, b);
          ++ycount;
        };
        y_load(0);
        for (int s = 0; s < ns; ++s) {
          if (s + 1 < ns) y_load(s + 1);
          if (s >= p.nb) {
            const int jb = s - p.nb;
            mbar_wait(v_empty, (vcount & 1) ^ 1);
            mbar_expect_tx(v_full, nvc * kTB);
            for (int c = 0; c < nvc; ++c) tma_load_4d(&tmV, v_full, sV + c * kTB, c * 64, jb * 128, h, b);
            ++vcount;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ----------------------------------------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, p.dpad, 0, 1);
      uint32_t ycount = 0, scount = 0, pvcount = 0, ti = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
        const int xb = ti & 1;
        mbar_wait(&x_full[xb], (ti >> 1) & 1);
        tc_fence_after();
        const uint32_t xa = smem_u32(sX + xb * 2 * kTB);
        auto issue_s = [&]() {
          const int sb = scount & 1, st = ycount & 1;
          mbar_wait(&s_empty[sb], ((scount >> 1) & 1) ^ 1);
          mbar_wait(&y_full[st], (ycount >> 1) & 1);
          tc_fence_after();
          const uint32_t ya = smem_u32(sY + st * 2 * kTB);
          for (int ks = 0; ks < nks; ++ks)
            umma_f16(tmem_base + sb * 128, desc_kmajor(xa + (ks >> 2) * kTB, ks & 3), desc_kmajor(ya + (ks >> 2) * kTB, ks & 3), idesc_s,
                     ks > 0 ? 1u : 0u);
          umma_commit(&y_empty[st]);
          umma_commit(&s_full[sb]);
          ++scount;
          ++ycount;
        };
        issue_s();
        for (int s = 0; s < ns; ++s) {
          if (s + 1 < ns) issue_s();
          if (s + 1 == ns) umma_commit(&x_empty[xb]);  // every logits product of this row tile has been issued
          if (s >= p.nb) {
            if (s == p.nb) {  // first P V of the tile overwrites the accumulator: the previous tile's epilogue must have read it
              mbar_wait(o_empty, (ti & 1) ^ 1);
            }
            mbar_wait(p_full, pvcount & 1);
            mbar_wait(v_full, pvcount & 1);
            tc_fence_after();
            const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_f16(tmem_base + kColO, desc_kmajor(pa + (kk >> 2) * kTB, kk & 3), desc_mnmajor(va, kTB, kk), idesc_pv,
                       (s > p.nb || kk > 0) ? 1u : 0u);
            umma_commit(p_empty);
            umma_commit(v_empty);
            ++pvcount;
          }
        }
        umma_commit(o_full);
      }
    }
  } else {
    // ----------------------------------------------------------------------------------------------- softmax / epilogue warps
    const int q = warp & 3;
    const int rl = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
    const float a2 = p.alpha * kLog2e;
    uint32_t scount = 0, pvcount = 0, ti = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
      const int rt = w % p.tiles_r;
      const int bh = w / p.tiles_r;
      const int h = bh % p.heads, b = bh / p.heads;
      const int row = rt * 128 + rl;
      const bool row_ok = row < p.R;
      // ---- pass 0: row maxima of the (unscaled) logits; alpha > 0 so max commutes with the scaling
      float mx = -INFINITY;
      for (int jb = 0; jb < p.nb; ++jb, ++scount) {
        const int sb = scount & 1;
        mbar_wait(&s_full[sb], (scount >> 1) & 1);
        tc_fence_after();
        const int cvalid = p.C - jb * 128;
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          if (ch * 32 >= cvalid) break;
          uint32_t acc[32];
          tmem_ld_32x32(lane_base + sb * 128 + ch * 32, acc);
          tmem_ld_wait();
          if (cvalid - ch * 32 >= 32) {
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(acc[e]));
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (e < cvalid - ch * 32) mx = fmaxf(mx, __uint_as_float(acc[e]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[sb]);
      }
      const float m2 = mx * a2;
      // ---- pass 1: P = 2^(a2 S - m2) -> bf16 tile; row sums
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int jb = 0; jb < p.nb; ++jb, ++scount, ++pvcount) {
        const int sb = scount & 1;
        mbar_wait(&s_full[sb], (scount >> 1) & 1);
        tc_fence_after();
        const int cvalid = p.C - jb * 128;
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          float v[32];
          if (ch * 32 < cvalid) {
            uint32_t acc[32];
            tmem_ld_32x32(lane_base + sb * 128 + ch * 32, acc);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = fast_exp2(fmaf(a2, __uint_as_float(acc[e]), -m2));
            if (cvalid - ch * 32 < 32) {
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (e >= cvalid - ch * 32) v[e] = 0.f;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = 0.f;
          }
          uint4 pc[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t wd[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              wd[u] = pack_bf16(v[g * 8 + 2 * u], v[g * 8 + 2 * u + 1]);
              // the row sum is taken over the ROUNDED probabilities: O / l is then an exact convex combination of the values
              l4[u] += __uint_as_float(wd[u] << 16) + __uint_as_float(wd[u] & 0xffff0000u);
            }
            pc[g] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
          }
          if (ch == 0) mbar_wait(p_empty, (pvcount & 1) ^ 1);  // the previous block's P V has consumed the tile
          tile_store_32cols(smem_u32(sP) + (ch >> 1) * kTB, rl, ch & 1, pc);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&s_empty[sb]);
          mbar_arrive(p_full);
        }
      }
      const float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      const float inv = 1.f / l;
      // ---- epilogue: out = O / l (+ residual), bf16
      mbar_wait(o_full, ti & 1);
      tc_fence_after();
      bf16* orow = p.out + b * p.o_bs + h * p.o_hs + (long long)row * p.o_ld;
      const bf16* rrow = p.res ? p.res + b * p.r_bs + h * p.r_hs + (long long)row * p.r_ld : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < p.dpad; c0 += 32) {
        uint32_t acc[32];
        tmem_ld_32x32(lane_base + kColO + c0, acc);
        tmem_ld_wait();
        if (!row_ok) continue;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int c = c0 + g * 8;
          if (c >= p.d) break;
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(acc[g * 8 + e]) * inv;
          if (c + 8 <= p.d) {
            if (rrow) {
              const uint4 rv = *reinterpret_cast<const uint4*>(rrow + c);
              const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                f[2 * u] += __uint_as_float(rw[u] << 16);
                f[2 * u + 1] += __uint_as_float(rw[u] & 0xffff0000u);
              }
            }
            *reinterpret_cast<uint4*>(orow + c) = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
          } else {
            for (int e = 0; e < 8 && c + e < p.d; ++e) orow[c + e] = __float2bfloat16(f[e] + (rrow ? __bfloat162float(rrow[c + e]) : 0.f));
          }
        }
      }
      if (row_ok && p.lse2) p.lse2[((long long)b * p.heads + h) * p.R + row] = m2 + fast_log2(l);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}


// =================================================================================================================
// Backward.  One kernel, two orientations of the same block computation (P never read from / written to HBM):
//   ROWS (COLS = false): tile = 128 softmax rows i, blocks = 64 keys j
//        S = A_T B_blk^T (x y^T), dP = C_T D_blk^T (dO v^T), statistics (lse2, dot) per TILE row
//        out1[i,:] += sum_j dS_ij B_blk[j,:]                                    -> dX
//   COLS (COLS = true): tile = 128 keys j, blocks = 64 softmax rows i
//        S' = A_T B_blk^T (y x^T), dP' = C_T D_blk^T (v dO^T), statistics per BLOCK column
//        out1[j,:] += sum_i dS'_ji B_blk[i,:]  -> dY          out2[j,:] += sum_i P'_ji D_blk[i,:]  -> dV
// with P = 2^(a2 S - lse2), dS = alpha P (dP - dot), dot_i = dO_i . O_i (row dots of the forward product, precomputed).
// The B_blk / D_blk tiles are loaded once and used twice: as the K-major operand of the logits products and as the
// MN-major operand of the output products (the [64 x 64] swizzled tile is the same bytes under both readings).
constexpr int kThreadsB = 64 + 8 * 32;
constexpr int kHB = 64 * 128;  // bytes of one [64 rows x 64 cols] bf16 tile

struct BwdOut {
  void* ptr;        // bf16 or f32 [batch, heads, rows, d] through (bs, hs, ld)
  int is_f32;
  long long ld, hs, bs;
  const bf16* res;  // optional bf16 residual: out = acc + rscale * res[row / row_div]
  long long r_ld, r_hs, r_bs;
  int row_div;
  float rscale;
};

struct BwdParams {
  int T, L;         // extent of the tile dimension / of the block dimension per (batch, head)
  int d, dpad, heads, batch;
  float alpha;
  const float* lse2;  // [batch, heads, softmax rows]
  const float* dot;   // [batch, heads, softmax rows]
  int n_rows;         // softmax rows per (batch, head) (= T for ROWS, = L for COLS)
  BwdOut o1, o2;
  int tiles, nblk;
};

struct BwdSmem {
  static constexpr int AT = 2 * kTB;      // A_T: two K blocks of [128 x 64]
  static constexpr int CT = 2 * kTB;
  static constexpr int STAGE = 4 * kHB;   // B_blk (two K blocks of [64 x 64]) + D_blk
  static constexpr int NST = 3;
  static constexpr int DS = kTB;          // dS tile [128 x 64]
  static constexpr int PT = kTB;          // P tile (COLS only)
  static constexpr int CS = 2 * 64 * 8;   // per-block column statistics (COLS only), double buffered
  static constexpr int BARS = 32 * 8 + 16;
  static constexpr int TOTAL = AT + CT + NST * STAGE + DS + PT + CS + BARS + 1024;
};

template <bool COLS>
__device__ __forceinline__ void bwd_store_out(const BwdOut& o, uint32_t taddr, int dpad, int d, int b, int h, int row, bool row_ok) {
  const long long base = b * o.bs + h * o.hs + (long long)row * o.ld;
  const bf16* rrow = o.res ? o.res + b * o.r_bs + h * o.r_hs + (long long)(row / o.row_div) * o.r_ld : nullptr;
#pragma unroll 1
  for (int c0 = 0; c0 < dpad; c0 += 32) {
    uint32_t acc[32];
    tmem_ld_32x32(taddr + c0, acc);
    tmem_ld_wait();
    if (!row_ok) continue;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int c = c0 + g * 8;
      if (c >= d) break;
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(acc[g * 8 + e]);
      if (rrow) {
        const uint4 rv = *reinterpret_cast<const uint4*>(rrow + c);
        const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          f[2 * u] = fmaf(o.rscale, __uint_as_float(rw[u] << 16), f[2 * u]);
          f[2 * u + 1] = fmaf(o.rscale, __uint_as_float(rw[u] & 0xffff0000u), f[2 * u + 1]);
        }
      }
      if (o.is_f32) {
        float* op = reinterpret_cast<float*>(o.ptr) + base + c;
        *reinterpret_cast<float4*>(op) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(op + 4) = make_float4(f[4], f[5], f[6], f[7]);
      } else {
        *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(o.ptr) + base + c) =
            make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
      }
    }
  }
}

template <bool COLS>
__global__ void __launch_bounds__(kThreadsB, 1)
flash_bwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                 const __grid_constant__ CUtensorMap tmD, const __grid_constant__ BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sC = sA + BwdSmem::AT;
  uint8_t* sSt = sC + BwdSmem::CT;
  uint8_t* sDS = sSt + BwdSmem::NST * BwdSmem::STAGE;
  uint8_t* sPT = sDS + BwdSmem::DS;
  float2* colstat = reinterpret_cast<float2*>(sPT + BwdSmem::PT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPT + BwdSmem::PT + BwdSmem::CS);
  uint64_t* at_full = bars;
  uint64_t* at_empty = bars + 1;
  uint64_t* st_full = bars + 2;    // [3]
  uint64_t* st_empty = bars + 5;   // [3]
  uint64_t* s_full = bars + 8;     // [2]
  uint64_t* s_empty = bars + 10;   // [2]
  uint64_t* ds_full = bars + 12;
  uint64_t* ds_empty = bars + 13;
  uint64_t* o_full = bars + 14;
  uint64_t* o_empty = bars + 15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.d + 63) / 64;
  const int nks = (p.d + 15) / 16;
  const int total = p.batch * p.heads * p.tiles;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    tma_prefetch_desc(&tmD);
  }
  if (warp == 1) {
    if (elect_one()) {
      mbar_init(at_full, 1);
      mbar_init(at_empty, 1);
      for (int i = 0; i < BwdSmem::NST; ++i) {
        mbar_init(&st_full[i], 1);
        mbar_init(&st_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&s_empty[i], 8);
      }
      mbar_init(ds_full, 8);
      mbar_init(ds_empty, 1);
      mbar_init(o_full, 1);
      mbar_init(o_empty, 8);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kCol1 = 256, kCol2 = 384;  // TMEM: (S | dP) buffers at [0,128) and [128,256); out1 at 256, out2 at 384

  if (warp == 0) {
    // ----------------------------------------------------------------------------------------------- TMA producer
    if (elect_one()) {
      uint32_t bc = 0, ti = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
        const int tt = w % p.tiles;
        const int bh = w / p.tiles;
        const int h = bh % p.heads, b = bh / p.heads;
        mbar_wait(at_empty, (ti & 1) ^ 1);
        mbar_expect_tx(at_full, 2 * nkb * kTB);
        for (int kb = 0; kb < nkb; ++kb) {
          tma_load_4d(&tmA, at_full, sA + kb * kTB, kb * 64, tt * 128, h, b);
          tma_load_4d(&tmC, at_full, sC + kb * kTB, kb * 64, tt * 128, h, b);
        }
        for (int blk = 0; blk < p.nblk; ++blk, ++bc) {
          const int st = bc % BwdSmem::NST;
          mbar_wait(&st_empty[st], ((bc / BwdSmem::NST) & 1) ^ 1);
          mbar_expect_tx(&st_full[st], 2 * nkb * kHB);
          uint8_t* base = sSt + st * BwdSmem::STAGE;
          for (int kb = 0; kb < nkb; ++kb) {
            tma_load_4d(&tmB, &st_full[st], base + kb * kHB, kb * 64, blk * 64, h, b);
            tma_load_4d(&tmD, &st_full[st], base + (2 + kb) * kHB, kb * 64, blk * 64, h, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ----------------------------------------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, p.dpad, 0, 1);
      uint32_t sdc = 0, oc = 0, ti = 0;  // blocks whose logits products / output products have been issued
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
        mbar_wait(at_full, ti & 1);
        tc_fence_after();
        const uint32_t aa = smem_u32(sA), ca = smem_u32(sC);
        auto issue_sd = [&]() {
          const int buf = sdc & 1, st = sdc % BwdSmem::NST;
          mbar_wait(&s_empty[buf], ((sdc >> 1) & 1) ^ 1);
          mbar_wait(&st_full[st], (sdc / BwdSmem::NST) & 1);
          tc_fence_after();
          const uint32_t ba = smem_u32(sSt + st * BwdSmem::STAGE), da = ba + 2 * kHB;
          for (int ks = 0; ks < nks; ++ks)
            umma_f16(tmem_base + buf * 128, desc_kmajor(aa + (ks >> 2) * kTB, ks & 3), desc_kmajor(ba + (ks >> 2) * kHB, ks & 3), idesc_s,
                     ks > 0 ? 1u : 0u);
          for (int ks = 0; ks < nks; ++ks)
            umma_f16(tmem_base + buf * 128 + 64, desc_kmajor(ca + (ks >> 2) * kTB, ks & 3), desc_kmajor(da + (ks >> 2) * kHB, ks & 3), idesc_s,
                     ks > 0 ? 1u : 0u);
          umma_commit(&s_full[buf]);
          ++sdc;
        };
        issue_sd();
        for (int blk = 0; blk < p.nblk; ++blk) {
          if (blk + 1 < p.nblk) issue_sd();
          if (blk + 1 == p.nblk) umma_commit(at_empty);  // all logits products of this tile are issued
          if (blk == 0) mbar_wait(o_empty, (ti & 1) ^ 1);
          mbar_wait(ds_full, oc & 1);
          tc_fence_after();
          const int st = oc % BwdSmem::NST;
          const uint32_t ba = smem_u32(sSt + st * BwdSmem::STAGE), da = ba + 2 * kHB;
          const uint32_t dsa = smem_u32(sDS), pta = smem_u32(sPT);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base + kCol1, desc_kmajor(dsa, kk), desc_mnmajor(ba, kHB, kk), idesc_o, (blk > 0 || kk > 0) ? 1u : 0u);
          if (COLS) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base + kCol2, desc_kmajor(pta, kk), desc_mnmajor(da, kHB, kk), idesc_o, (blk > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&st_empty[st]);
          umma_commit(ds_empty);
          ++oc;
        }
        umma_commit(o_full);
      }
    }
  } else {
    // ----------------------------------------------------------------------------------------------- softmax / epilogue warps
    const int q = warp & 3, hf = (warp - 2) >> 2;  // TMEM lane quarter; which 32 of the block's 64 columns
    const int rl = q * 32 + lane;
    const int t = threadIdx.x - 64;  // 0..255
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
    const float a2 = p.alpha * kLog2e;
    uint32_t bc = 0, ti = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
      const int tt = w % p.tiles;
      const int bh = w / p.tiles;
      const int h = bh % p.heads, b = bh / p.heads;
      const int row = tt * 128 + rl;
      const bool row_ok = row < p.T;
      const float* lse_bh = p.lse2 + (long long)bh * p.n_rows;
      const float* dot_bh = p.dot + (long long)bh * p.n_rows;
      float r_lse = INFINITY, r_dot = 0.f;
      if (!COLS && row_ok) {
        r_lse = lse_bh[row];
        r_dot = dot_bh[row];
      }
      for (int blk = 0; blk < p.nblk; ++blk, ++bc) {
        const int buf = bc & 1;
        if (COLS) {  // statistics of this block's 64 softmax rows
          if (t < 64) {
            const int i = blk * 64 + t;
            colstat[buf * 64 + t] = i < p.L ? make_float2(lse_bh[i], dot_bh[i]) : make_float2(INFINITY, 0.f);
          }
          named_bar_sync(1, 256);
        }
        mbar_wait(&s_full[buf], (bc >> 1) & 1);
        tc_fence_after();
        uint32_t sa[32], da[32];
        tmem_ld_32x32(lane_base + buf * 128 + hf * 32, sa);
        tmem_ld_32x32(lane_base + buf * 128 + 64 + hf * 32, da);
        tmem_ld_wait();
        float pv[32], dv[32];
        if (COLS) {
          const float2* cs = colstat + buf * 64 + hf * 32;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float2 c = cs[e];
            pv[e] = fast_exp2(fmaf(a2, __uint_as_float(sa[e]), -c.x));
            dv[e] = p.alpha * pv[e] * (__uint_as_float(da[e]) - c.y);
          }
        } else {
          const int cvalid = p.L - blk * 64 - hf * 32;  // keys of this chunk that exist
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            float pe = fast_exp2(fmaf(a2, __uint_as_float(sa[e]), -r_lse));
            if (e >= cvalid) pe = 0.f;
            dv[e] = p.alpha * pe * (__uint_as_float(da[e]) - r_dot);
          }
        }
        uint4 pd[4], pp[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          pd[g] = make_uint4(pack_bf16(dv[g * 8], dv[g * 8 + 1]), pack_bf16(dv[g * 8 + 2], dv[g * 8 + 3]), pack_bf16(dv[g * 8 + 4], dv[g * 8 + 5]),
                             pack_bf16(dv[g * 8 + 6], dv[g * 8 + 7]));
          if (COLS)
            pp[g] = make_uint4(pack_bf16(pv[g * 8], pv[g * 8 + 1]), pack_bf16(pv[g * 8 + 2], pv[g * 8 + 3]), pack_bf16(pv[g * 8 + 4], pv[g * 8 + 5]),
                               pack_bf16(pv[g * 8 + 6], pv[g * 8 + 7]));
        }
        mbar_wait(ds_empty, (bc & 1) ^ 1);  // the previous block's output products have consumed the tiles
        tile_store_32cols(smem_u32(sDS), rl, hf, pd);
        if (COLS) tile_store_32cols(smem_u32(sPT), rl, hf, pp);
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&s_empty[buf]);
          mbar_arrive(ds_full);
        }
      }
      // ---- epilogue: each of the two warps sharing a lane quarter stores one of the outputs (ROWS: the two halves of out1's columns)
      mbar_wait(o_full, ti & 1);
      tc_fence_after();
      if (COLS) {
        if (hf == 0) bwd_store_out<COLS>(p.o1, lane_base + kCol1, p.dpad, p.d, b, h, row, row_ok);
        else bwd_store_out<COLS>(p.o2, lane_base + kCol2, p.dpad, p.d, b, h, row, row_ok);
      } else if (hf == 0) {
        bwd_store_out<COLS>(p.o1, lane_base + kCol1, p.dpad, p.d, b, h, row, row_ok);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

int fill_bwd_out(BwdOut* o, const mirror_flash_out* g, const char* what) {
  o->ptr = g->ptr;
  o->is_f32 = g->is_f32;
  o->ld = g->ld; o->hs = g->hs; o->bs = g->bs;
  o->res = reinterpret_cast<const bf16*>(g->res);
  o->r_ld = g->r_ld; o->r_hs = g->r_hs; o->r_bs = g->r_bs;
  o->row_div = g->row_div > 1 ? g->row_div : 1;
  o->rscale = g->rscale;
  const int q = g->is_f32 ? 4 : 8;
  MB_CHECK_ARG(g->ptr && (reinterpret_cast<uintptr_t>(g->ptr) & 15) == 0 && g->ld % q == 0 && g->hs % q == 0 && g->bs % q == 0,
               "flash_bwd: %s must be 16-byte aligned with 16-byte strides", what);
  MB_CHECK_ARG(!g->res || ((reinterpret_cast<uintptr_t>(g->res) & 15) == 0 && g->r_ld % 8 == 0 && g->r_hs % 8 == 0 && g->r_bs % 8 == 0),
               "flash_bwd: residual of %s must be 16-byte aligned with 16-byte strides", what);
  return 0;
}

}  // namespace
}  // namespace mb

using namespace mb;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int mirror_flash_softmax_pv(const mirror_flash_args* a, mirror_stream_t stream) {
  MB_CHECK_ARG(a && a->x && a->y && a->v && a->out, "flash_softmax_pv: null operand");
  MB_CHECK_ARG(a->R > 0 && a->C > 0 && a->d >= 8 && a->d <= 128 && a->d % 8 == 0 && a->heads > 0 && a->batch > 0 && a->alpha > 0.f,
               "flash_softmax_pv: bad shape R=%d C=%d d=%d", a->R, a->C, a->d);
  MB_CHECK_ARG((reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && a->o_ld % 8 == 0 && a->o_hs % 8 == 0 && a->o_bs % 8 == 0,
               "flash_softmax_pv: output must be 16-byte aligned with strides that are multiples of 8 elements");
  MB_CHECK_ARG(!a->res || ((reinterpret_cast<uintptr_t>(a->res) & 15) == 0 && a->r_ld % 8 == 0 && a->r_hs % 8 == 0 && a->r_bs % 8 == 0),
               "flash_softmax_pv: residual must be 16-byte aligned with strides that are multiples of 8 elements");
  CUtensorMap tmX, tmY, tmV;
  int rc = make_map_4d(&tmX, a->x, a->d, a->R, a->x_ld, a->heads, a->x_hs, a->batch, a->x_bs, 128);
  if (rc) return rc;
  rc = make_map_4d(&tmY, a->y, a->d, a->C, a->y_ld, a->heads, a->y_hs, a->batch, a->y_bs, 128);
  if (rc) return rc;
  rc = make_map_4d(&tmV, a->v, a->d, a->C, a->v_ld, a->heads, a->v_hs, a->batch, a->v_bs, 128);
  if (rc) return rc;
  FwdParams p = {};
  p.R = a->R; p.C = a->C; p.d = a->d; p.dpad = (a->d + 15) / 16 * 16;
  p.heads = a->heads; p.batch = a->batch; p.alpha = a->alpha;
  p.out = reinterpret_cast<bf16*>(a->out); p.o_ld = a->o_ld; p.o_hs = a->o_hs; p.o_bs = a->o_bs;
  p.res = reinterpret_cast<const bf16*>(a->res); p.r_ld = a->r_ld; p.r_hs = a->r_hs; p.r_bs = a->r_bs;
  p.lse2 = a->lse2;
  p.tiles_r = (a->R + 127) / 128;
  p.nb = (a->C + 127) / 128;
  static DeviceOnce once;
  if (once.first()) MB_CUDA(cudaFuncSetAttribute(flash_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL));
  const long long total = (long long)p.batch * p.heads * p.tiles_r;
  const int grid = (int)(total < num_sms() ? total : num_sms());
  flash_fwd_kernel<<<grid, kThreadsF, FwdSmem::TOTAL, STREAM>>>(tmX, tmY, tmV, p);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mirror_flash_bwd(const mirror_flash_bwd_args* a, mirror_stream_t stream) {
  MB_CHECK_ARG(a && a->a && a->b && a->c && a->dd && a->lse2 && a->dot, "flash_bwd: null operand");
  MB_CHECK_ARG(a->T > 0 && a->L > 0 && a->d >= 8 && a->d <= 128 && a->d % 8 == 0 && a->heads > 0 && a->batch > 0 && a->alpha > 0.f,
               "flash_bwd: bad shape T=%d L=%d d=%d", a->T, a->L, a->d);
  CUtensorMap tmA, tmB, tmC, tmD;
  int rc = make_map_4d(&tmA, a->a, a->d, a->T, a->a_ld, a->heads, a->a_hs, a->batch, a->a_bs, 128);
  if (rc) return rc;
  rc = make_map_4d(&tmC, a->c, a->d, a->T, a->c_ld, a->heads, a->c_hs, a->batch, a->c_bs, 128);
  if (rc) return rc;
  rc = make_map_4d(&tmB, a->b, a->d, a->L, a->b_ld, a->heads, a->b_hs, a->batch, a->b_bs, 64);
  if (rc) return rc;
  rc = make_map_4d(&tmD, a->dd, a->d, a->L, a->d_ld, a->heads, a->d_hs, a->batch, a->d_bs, 64);
  if (rc) return rc;
  BwdParams p = {};
  p.T = a->T; p.L = a->L; p.d = a->d; p.dpad = (a->d + 15) / 16 * 16;
  p.heads = a->heads; p.batch = a->batch; p.alpha = a->alpha;
  p.lse2 = a->lse2; p.dot = a->dot;
  p.n_rows = a->cols ? a->L : a->T;
  rc = fill_bwd_out(&p.o1, &a->out1, "out1");
  if (rc) return rc;
  if (a->cols) {
    rc = fill_bwd_out(&p.o2, &a->out2, "out2");
    if (rc) return rc;
  }
  p.tiles = (a->T + 127) / 128;
  p.nblk = (a->L + 63) / 64;
  const long long total = (long long)p.batch * p.heads * p.tiles;
  const int grid = (int)(total < num_sms() ? total : num_sms());
  if (a->cols) {
    static DeviceOnce once;
    if (once.first()) MB_CUDA(cudaFuncSetAttribute(flash_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::TOTAL));
    flash_bwd_kernel<true><<<grid, kThreadsB, BwdSmem::TOTAL, STREAM>>>(tmA, tmB, tmC, tmD, p);
  } else {
    static DeviceOnce once;
    if (once.first()) MB_CUDA(cudaFuncSetAttribute(flash_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::TOTAL));
    flash_bwd_kernel<false><<<grid, kThreadsB, BwdSmem::TOTAL, STREAM>>>(tmA, tmB, tmC, tmD, p);
  }
  MB_LAUNCH_CHECK();
  return 0;
}
