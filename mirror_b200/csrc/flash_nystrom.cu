// Flash-style fused softmax products of the Nystrom attention (nystrom_attention forward, call site models/mirror.py:299-312;
// SURVEY.md §2.1 kernels K-A / K-C): the [rows x keys] probability matrices
//     attn1 = softmax(q k_l^T)   [n x m]   ->  out = attn1 W + res_conv(v)
//     attn3 = softmax(q_l k^T)   [m x n]   ->  kv  = attn3 v
// never exist in HBM -- nor in shared memory: the probabilities are written back into TENSOR MEMORY as bf16 and consumed as
// the A operand of the second tcgen05.mma (the "TS" form: A from TMEM, B from shared memory).
// Forward: O = softmax(alpha X Y^T) V, logits tile in TMEM, ONE pass over the key blocks: tensor-memory reads (~57 B/clk per SM,
// tools/tmem_bw.cu) are what bounds these kernels, so every logit is read exactly once.  The exponentials are taken relative to
// the running row maximum of the FIRST block and the reference is only moved (accumulator and row sums rescaled, a rare path)
// when a later block exceeds it by more than 2^20: P = 2^(a2 (S - m_ref)) <= 2^20 is exact enough in bf16 / fp32 at any offset,
// and O / l does not depend on the reference.  Backward (flash_bwd below): logits and
// dP = dO V^T are recomputed per 128 x 64 block, dS = alpha P (dP - D) overwrites the logits in TMEM and feeds
// dX += dS Y (row-stationary) or dY += dS^T X, dV += P^T dO (column-stationary).
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2..9 softmax / epilogue
// (thread = row of the 128-row tile x half of a block's columns).  Persistent: a CTA walks tiles blockIdx.x,
// blockIdx.x + gridDim.x, ...  Every hand-off is at least double buffered (logits and probabilities in TMEM, operand tiles
// in shared memory), so TMA, tensor core and the softmax warps each run at their own pace.
#include "tile.cuh"

namespace mb {
namespace {

constexpr int kThreadsF = 64 + 8 * 32;
constexpr int kTB = 128 * 128;  // bytes of one [128 rows x 64 cols] bf16 tile
constexpr int kHB = 64 * 128;   // bytes of one [64 rows x 64 cols] bf16 tile
constexpr float kLog2e = 1.4426950408889634f;

// Optional event trace of CTA 0 (debug / measurement only; mirror_debug_flash_trace installs a buffer, NULL = off).  Each traced
// role (region 0 producer, 1 MMA issuer, 2 softmax warp 2) appends plain stores to its own quarter of the buffer -- no atomics, so
// an event costs a few cycles: entry = (event id << 48) | (clock64 & 2^48-1); the first word of a region counts its entries.
__device__ unsigned long long* g_trace = nullptr;
__device__ unsigned long long g_trace_cap = 0;
#ifdef MIRROR_FLASH_TRACE
struct Tracer {
  unsigned long long* base;
  unsigned int n, cap;
  __device__ __forceinline__ void init(int region) {
    base = (blockIdx.x == 0 && g_trace) ? g_trace + region * (g_trace_cap / 4) : nullptr;
    cap = (unsigned int)(g_trace_cap / 4);
    n = 0;
  }
  __device__ __forceinline__ void ev(int id) {
    if (base && n + 1 < cap) {
      base[++n] = ((unsigned long long)id << 48) | ((unsigned long long)clock64() & 0xffffffffffffull);
      base[0] = n;
    }
  }
};
#else  // product build: the trace hooks compile to nothing (code size matters: the kernels run out of the instruction cache)
struct Tracer {
  __device__ __forceinline__ void init(int) {}
  __device__ __forceinline__ void ev(int) {}
};
#endif

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
  static constexpr int V = 2 * 2 * kTB;   // ring of two value blocks: two 64-column chunks [128 keys x 64] each
  static constexpr int ROWS = (2 * 2 * 128 + 2 * 128) * 4;  // block maxima of the two column halves (two parities), row sums of the halves
  static constexpr int BARS = 32 * 8 + 16;
  static constexpr int TOTAL = X + Y + V + ROWS + BARS + 1024;
};

// Two-pass variant (row maxima first, then exp / row sums / P V; the logits products are issued twice): faster when a row tile
// has only a few key blocks (K-C: 3 blocks of landmarks per 128 tokens), where the single-pass kernel's per-block maximum exchange
// and shorter software pipeline cost more than the second tensor-memory read (measured 0.44 vs 0.53 ms at the benchmark shape).
__global__ void __launch_bounds__(kThreadsF, 1)
flash_fwd_twopass_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sY = sX + FwdSmem::X;
  uint8_t* sV = sY + FwdSmem::Y;
  float* rowmax = reinterpret_cast<float*>(sV + FwdSmem::V);
  float* rowsum = rowmax + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + FwdSmem::V + FwdSmem::ROWS);
  uint64_t* x_full = bars;          // [2]
  uint64_t* x_empty = bars + 2;     // [2]
  uint64_t* y_full = bars + 4;      // [2]
  uint64_t* y_empty = bars + 6;     // [2]
  uint64_t* s_full = bars + 8;      // [2]
  uint64_t* s_empty = bars + 10;    // [2]
  uint64_t* v_full = bars + 12;     // [2]
  uint64_t* v_empty = bars + 14;    // [2]
  uint64_t* p_full = bars + 16;     // [2]
  uint64_t* p_empty = bars + 18;    // [2]
  uint64_t* o_full = bars + 20;
  uint64_t* o_empty = bars + 21;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

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
        mbar_init(&s_empty[i], 8);
        mbar_init(&v_full[i], 1);
        mbar_init(&v_empty[i], 1);
        mbar_init(&p_full[i], 8);
        mbar_init(&p_empty[i], 1);
      }
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
  pdl_wait();  // programmatic dependent launch (common.cuh): the setup above overlapped the previous grid's tail
  pdl_launch_dependents();
  // TMEM columns: logits S[2] at 0 / 128, output accumulator at 256 (dpad <= 128), probabilities P[2] (bf16 pairs) at 384 / 448
  constexpr uint32_t kColO = 256, kColP = 384;

  if (warp == 0) {
    // ----------------------------------------------------------------------------------------------- TMA producer
    if (elect_one()) {
      Tracer tr;
      tr.init(0);
      uint32_t ycount = 0, vcount = 0, ti = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
        const int rt = w % p.tiles_r;
        const int bh = w / p.tiles_r;
        const int h = bh % p.heads, b = bh / p.heads;
        const int xb = ti & 1;
        mbar_wait(&x_empty[xb], ((ti >> 1) & 1) ^ 1);
        mbar_expect_tx(&x_full[xb], nkb * kTB);
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb) tma_load_4d(&tmX, &x_full[xb], sX + (xb * 2 + kb) * kTB, kb * 64, rt * 128, h, b);
        tr.ev(1);
        // loads are issued in the order the MMA warp consumes them: S(0), S(1), then per step [P V(s)], S(s+2).  ONE load site per
        // operand and no unrolling: this warp's code shares the 32 KB instruction cache with the softmax warps' hot loops.
#pragma unroll 1
        for (int s = -2; s < ns; ++s) {
          if (s >= p.nb) {
            const int jb = s - p.nb, st = vcount & 1;
            mbar_wait(&v_empty[st], ((vcount >> 1) & 1) ^ 1);
            mbar_expect_tx(&v_full[st], nvc * kTB);
#pragma unroll 1
            for (int c = 0; c < nvc; ++c) tma_load_4d(&tmV, &v_full[st], sV + (st * 2 + c) * kTB, c * 64, jb * 128, h, b);
            tr.ev(3);
            ++vcount;
          }
          if (s + 2 < ns) {
            const int jb = (s + 2) % p.nb, st = ycount & 1;
            mbar_wait(&y_empty[st], ((ycount >> 1) & 1) ^ 1);
            mbar_expect_tx(&y_full[st], nkb * kTB);
#pragma unroll 1
            for (int kb = 0; kb < nkb; ++kb) tma_load_4d(&tmY, &y_full[st], sY + (st * 2 + kb) * kTB, kb * 64, jb * 128, h, b);
            tr.ev(2);
            ++ycount;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ----------------------------------------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, p.dpad, 0, 1);
      Tracer tr;
      tr.init(1);
      uint32_t scount = 0, pvcount = 0, ti = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
        const int xb = ti & 1;
        mbar_wait(&x_full[xb], (ti >> 1) & 1);
        tc_fence_after();
        const uint32_t xa = smem_u32(sX + xb * 2 * kTB);
        tr.ev(10);
#pragma unroll 1
        for (int s = -2; s < ns; ++s) {  // S(0), S(1), then per step [P V(s)], S(s+2): one issue site each (code size)
          if (s >= p.nb) {
            const int pb = pvcount & 1;
            if (s == p.nb) mbar_wait(o_empty, (ti & 1) ^ 1);  // the previous tile's epilogue has read the accumulator
            mbar_wait(&p_full[pb], (pvcount >> 1) & 1);
            tr.ev(15);
            mbar_wait(&v_full[pb], (pvcount >> 1) & 1);
            tc_fence_after();
            tr.ev(12);
            const uint32_t va = smem_u32(sV + pb * 2 * kTB);
#pragma unroll 2
            for (int kk = 0; kk < 8; ++kk)  // contraction over the block's 128 keys: 8 TMEM columns (16 bf16) of P per step
              umma_f16_ts(tmem_base + kColO, tmem_base + kColP + pb * 64 + kk * 8, desc_mnmajor(va, kTB, kk), idesc_pv,
                          (s > p.nb || kk > 0) ? 1u : 0u);
            umma_commit(&p_empty[pb]);
            umma_commit(&v_empty[pb]);
            ++pvcount;
          }
          if (s + 2 < ns) {
            const int sb = scount & 1;  // logits buffer and Y ring stage advance together
            mbar_wait(&s_empty[sb], ((scount >> 1) & 1) ^ 1);
            tr.ev(14);
            mbar_wait(&y_full[sb], (scount >> 1) & 1);
            tc_fence_after();
            tr.ev(11);
            const uint32_t ya = smem_u32(sY + sb * 2 * kTB);
#pragma unroll 2
            for (int ks = 0; ks < nks; ++ks)
              umma_f16(tmem_base + sb * 128, desc_kmajor(xa + (ks >> 2) * kTB, ks & 3), desc_kmajor(ya + (ks >> 2) * kTB, ks & 3), idesc_s,
                       ks > 0 ? 1u : 0u);
            umma_commit(&y_empty[sb]);
            umma_commit(&s_full[sb]);
            ++scount;
          }
          if (s + 3 == ns) umma_commit(&x_empty[xb]);  // every logits product of this row tile has been issued
        }
        umma_commit(o_full);
        tr.ev(13);
      }
    }
  } else {
    // ----------------------------------------------------------------------------------------------- softmax / epilogue warps
    // eight warps: two per TMEM lane quarter, each owning 64 of a key block's 128 columns (thread = row x column half)
    const int q = warp & 3, hf = (warp - 2) >> 2;
    const int rl = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
    const float a2 = p.alpha * kLog2e;
    Tracer tr;
    tr.init(2);
    uint32_t scount = 0, pvcount = 0, ti = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
      const int rt = w % p.tiles_r;
      const int bh = w / p.tiles_r;
      const int h = bh % p.heads, b = bh / p.heads;
      const int row = rt * 128 + rl;
      const bool row_ok = row < p.R;
      const bf16* rrow = p.res ? p.res + b * p.r_bs + h * p.r_hs + (long long)row * p.r_ld : nullptr;
      if (rrow && row_ok) prefetch_l2(rrow + (hf * 64 < p.d ? hf * 64 : 0));
      if (warp == 2 && lane == 0) tr.ev(33);
      // ---- pass 0: row maxima of the (unscaled) logits; alpha > 0 so max commutes with the scaling.
      // (All per-chunk loops below are deliberately NOT unrolled across chunks: the kernel's code must stay well inside the
      // 32 KB instruction cache -- with everything unrolled the once-per-tile epilogue ran at ~17 cycles per instruction.)
      float mx = -INFINITY;
      for (int jb = 0; jb < p.nb; ++jb, ++scount) {
        const int sb = scount & 1;
        mbar_wait(&s_full[sb], (scount >> 1) & 1);
        tc_fence_after();
        if (warp == 2 && lane == 0) tr.ev(20);
        const int cvalid = p.C - jb * 128 - hf * 64;  // keys of this warp's half that exist
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int cv = cvalid - cc * 32;
          if (cv <= 0) break;  // warp-uniform
          uint32_t acc[32];
          tmem_ld_32x32(lane_base + sb * 128 + hf * 64 + cc * 32, acc);
          tmem_ld_wait();
          if (cv >= 32) {  // full chunk (every chunk of the model's shapes: keys are multiples of 128): no per-element masks
#pragma unroll
            for (int e = 0; e < 32; e += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(acc[e]), __uint_as_float(acc[e + 1])));
          } else {  // ragged tail of the last key block (cold: the model's key counts are multiples of 128)
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, e < cv ? __uint_as_float(acc[e]) : -INFINITY);
          }
        }
        if (warp == 2 && lane == 0) tr.ev(21);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[sb]);
      }
      rowmax[hf * 128 + rl] = mx;
      named_bar_sync(1, 256);
      if (warp == 2 && lane == 0) tr.ev(22);
      mx = fmaxf(rowmax[rl], rowmax[128 + rl]);
      const float m2 = mx * a2;
      // ---- pass 1: P = 2^(a2 S - m2) -> bf16 -> tensor memory; row sums
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int jb = 0; jb < p.nb; ++jb, ++scount, ++pvcount) {
        const int sb = scount & 1, pb = pvcount & 1;
        mbar_wait(&s_full[sb], (scount >> 1) & 1);
        mbar_wait(&p_empty[pb], ((pvcount >> 1) & 1) ^ 1);  // P V of two blocks ago has consumed this probability buffer
        tc_fence_after();
        if (warp == 2 && lane == 0) tr.ev(23);
        const int cvalid = p.C - jb * 128 - hf * 64;
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {  // one 32-column chunk at a time: logits -> exp -> row sum -> bf16 pairs -> 16 TMEM columns
          const int cv = cvalid - cc * 32;
          uint32_t acc[32], wd[16];
          if (cv > 0) {  // warp-uniform
            tmem_ld_32x32(lane_base + sb * 128 + hf * 64 + cc * 32, acc);
            tmem_ld_wait();
          }
          if (cv >= 32) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const float x0 = fast_exp2(fmaf(a2, __uint_as_float(acc[2 * u]), -m2)), x1 = fast_exp2(fmaf(a2, __uint_as_float(acc[2 * u + 1]), -m2));
              l4[u & 3] += x0 + x1;
              wd[u] = pack_bf16(x0, x1);
            }
          } else {  // ragged tail of the last key block (cold)
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const float x0 = 2 * u < cv ? fast_exp2(fmaf(a2, __uint_as_float(acc[2 * u]), -m2)) : 0.f;
              const float x1 = 2 * u + 1 < cv ? fast_exp2(fmaf(a2, __uint_as_float(acc[2 * u + 1]), -m2)) : 0.f;
              l4[u & 3] += x0 + x1;
              wd[u] = pack_bf16(x0, x1);
            }
          }
          tmem_st_32x32_x16(lane_base + kColP + pb * 64 + hf * 32 + cc * 16, wd);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&s_empty[sb]);
          mbar_arrive(&p_full[pb]);
        }
        if (warp == 2 && lane == 0) tr.ev(26);
      }
      rowsum[hf * 128 + rl] = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      named_bar_sync(1, 256);
      if (warp == 2 && lane == 0) tr.ev(27);
      const float l = rowsum[rl] + rowsum[128 + rl];
      const float inv = 1.f / l;
      // ---- epilogue: out = O / l (+ residual), bf16.  The accumulator is drained in 16-column pieces; the two warps of a lane
      // quarter take alternate pieces, so both finish together.  Everything with a latency is issued up front: the residual
      // pieces travel while the last P V products retire, the accumulator pieces are fetched with ONE wait.
      bf16* orow = p.out + b * p.o_bs + h * p.o_hs + (long long)row * p.o_ld;
      constexpr int kPieces = 4;  // d <= 128: at most four 16-column pieces per warp
      uint4 rpre[kPieces][2];
#pragma unroll
      for (int i = 0; i < kPieces; ++i) {
        const int c0 = hf * 16 + i * 32;
        rpre[i][0] = rpre[i][1] = make_uint4(0u, 0u, 0u, 0u);
        if (rrow && row_ok && c0 < p.d) {
          rpre[i][0] = ldg_v4(rrow + c0);
          if (c0 + 8 < p.d) rpre[i][1] = ldg_v4(rrow + c0 + 8);
        }
      }
      mbar_wait(o_full, ti & 1);
      tc_fence_after();
      if (warp == 2 && lane == 0) tr.ev(28);
      uint32_t oacc[kPieces][16];
#pragma unroll
      for (int i = 0; i < kPieces; ++i)
        if (hf * 16 + i * 32 < p.d) tmem_ld_32x32_x16(lane_base + kColO + hf * 16 + i * 32, oacc[i]);
      tmem_ld_wait();
      if (warp == 2 && lane == 0) tr.ev(31);
      tc_fence_before();  // the accumulator is in registers: the next tile's P V products may overwrite it
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < kPieces; ++i) {
          const int c0 = hf * 16 + i * 32;
          if (c0 < p.d) {
            const uint32_t rw[8] = {rpre[i][0].x, rpre[i][0].y, rpre[i][0].z, rpre[i][0].w, rpre[i][1].x, rpre[i][1].y, rpre[i][1].z, rpre[i][1].w};
            uint32_t o[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
              o[u] = pack_bf16(fmaf(__uint_as_float(oacc[i][2 * u]), inv, __uint_as_float(rw[u] << 16)),
                               fmaf(__uint_as_float(oacc[i][2 * u + 1]), inv, __uint_as_float(rw[u] & 0xffff0000u)));
            stg_v4(orow + c0, make_uint4(o[0], o[1], o[2], o[3]));
            if (c0 + 8 < p.d) stg_v4(orow + c0 + 8, make_uint4(o[4], o[5], o[6], o[7]));
          }
        }
      }
      if (hf == 0 && row_ok && p.lse2) p.lse2[((long long)b * p.heads + h) * p.R + row] = m2 + fast_log2(l);
      if (warp == 2 && lane == 0) tr.ev(29);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}


// Single-pass variant (see the header): long rows of key blocks (K-A: 18+ blocks of tokens per 128 landmarks).
__global__ void __launch_bounds__(kThreadsF, 1)
flash_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sY = sX + FwdSmem::X;
  uint8_t* sV = sY + FwdSmem::Y;
  float* pairmax = reinterpret_cast<float*>(sV + FwdSmem::V);  // [parity][half][128]
  float* rowsum = pairmax + 512;                                // [half][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + FwdSmem::V + FwdSmem::ROWS);
  uint64_t* x_full = bars;          // [2]
  uint64_t* x_empty = bars + 2;     // [2]
  uint64_t* y_full = bars + 4;      // [2]
  uint64_t* y_empty = bars + 6;     // [2]
  uint64_t* s_full = bars + 8;      // [2]
  uint64_t* s_empty = bars + 10;    // [2]
  uint64_t* v_full = bars + 12;     // [2]
  uint64_t* v_empty = bars + 14;    // [2]
  uint64_t* p_full = bars + 16;     // [2]
  uint64_t* p_empty = bars + 18;    // [2]
  uint64_t* o_full = bars + 20;
  uint64_t* o_empty = bars + 21;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.d + 63) / 64;      // 64-column K blocks of the head dim
  const int nks = (p.d + 15) / 16;      // UMMA k-steps of the logits products
  const int nvc = (p.dpad + 63) / 64;   // 64-column chunks of the value tile
  const int total = p.batch * p.heads * p.tiles_r;
  const int ns = p.nb;                  // S tiles per row tile: one per key block

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
        mbar_init(&s_empty[i], 8);
        mbar_init(&v_full[i], 1);
        mbar_init(&v_empty[i], 1);
        mbar_init(&p_full[i], 8);
        mbar_init(&p_empty[i], 1);
      }
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
  pdl_wait();  // programmatic dependent launch (common.cuh): the setup above overlapped the previous grid's tail
  pdl_launch_dependents();
  // TMEM columns: logits S[2] at 0 / 128, output accumulator at 256 (dpad <= 128), probabilities P[2] (bf16 pairs) at 384 / 448
  constexpr uint32_t kColO = 256, kColP = 384;

  if (warp == 0) {
    // ----------------------------------------------------------------------------------------------- TMA producer
    if (elect_one()) {
      Tracer tr;
      tr.init(0);
      uint32_t ycount = 0, vcount = 0, ti = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
        const int rt = w % p.tiles_r;
        const int bh = w / p.tiles_r;
        const int h = bh % p.heads, b = bh / p.heads;
        const int xb = ti & 1;
        mbar_wait(&x_empty[xb], ((ti >> 1) & 1) ^ 1);
        mbar_expect_tx(&x_full[xb], nkb * kTB);
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb) tma_load_4d(&tmX, &x_full[xb], sX + (xb * 2 + kb) * kTB, kb * 64, rt * 128, h, b);
        tr.ev(1);
        // loads are issued in the order the MMA warp consumes them: S(0), S(1), then per step [P V(s)], S(s+2).  ONE load site per
        // operand and no unrolling: this warp's code shares the 32 KB instruction cache with the softmax warps' hot loops.
#pragma unroll 1
        for (int s = -2; s < ns; ++s) {
          if (s >= 0) {
            const int jb = s, st = vcount & 1;
            mbar_wait(&v_empty[st], ((vcount >> 1) & 1) ^ 1);
            mbar_expect_tx(&v_full[st], nvc * kTB);
#pragma unroll 1
            for (int c = 0; c < nvc; ++c) tma_load_4d(&tmV, &v_full[st], sV + (st * 2 + c) * kTB, c * 64, jb * 128, h, b);
            tr.ev(3);
            ++vcount;
          }
          if (s + 2 < ns) {
            const int jb = s + 2, st = ycount & 1;
            mbar_wait(&y_empty[st], ((ycount >> 1) & 1) ^ 1);
            mbar_expect_tx(&y_full[st], nkb * kTB);
#pragma unroll 1
            for (int kb = 0; kb < nkb; ++kb) tma_load_4d(&tmY, &y_full[st], sY + (st * 2 + kb) * kTB, kb * 64, jb * 128, h, b);
            tr.ev(2);
            ++ycount;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ----------------------------------------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, p.dpad, 0, 1);
      Tracer tr;
      tr.init(1);
      uint32_t scount = 0, pvcount = 0, ti = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
        const int xb = ti & 1;
        mbar_wait(&x_full[xb], (ti >> 1) & 1);
        tc_fence_after();
        const uint32_t xa = smem_u32(sX + xb * 2 * kTB);
        tr.ev(10);
#pragma unroll 1
        for (int s = -2; s < ns; ++s) {  // S(0), S(1), then per step [P V(s)], S(s+2): one issue site each (code size)
          if (s >= 0) {
            const int pb = pvcount & 1;
            if (s == 0) mbar_wait(o_empty, (ti & 1) ^ 1);  // the previous tile's epilogue has read the accumulator
            mbar_wait(&p_full[pb], (pvcount >> 1) & 1);
            tr.ev(15);
            mbar_wait(&v_full[pb], (pvcount >> 1) & 1);
            tc_fence_after();
            tr.ev(12);
            const uint32_t va = smem_u32(sV + pb * 2 * kTB);
#pragma unroll 2
            for (int kk = 0; kk < 8; ++kk)  // contraction over the block's 128 keys: 8 TMEM columns (16 bf16) of P per step
              umma_f16_ts(tmem_base + kColO, tmem_base + kColP + pb * 64 + kk * 8, desc_mnmajor(va, kTB, kk), idesc_pv,
                          (s > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&p_empty[pb]);
            umma_commit(&v_empty[pb]);
            ++pvcount;
          }
          if (s + 2 < ns) {
            const int sb = scount & 1;  // logits buffer and Y ring stage advance together
            mbar_wait(&s_empty[sb], ((scount >> 1) & 1) ^ 1);
            tr.ev(14);
            mbar_wait(&y_full[sb], (scount >> 1) & 1);
            tc_fence_after();
            tr.ev(11);
            const uint32_t ya = smem_u32(sY + sb * 2 * kTB);
#pragma unroll 2
            for (int ks = 0; ks < nks; ++ks)
              umma_f16(tmem_base + sb * 128, desc_kmajor(xa + (ks >> 2) * kTB, ks & 3), desc_kmajor(ya + (ks >> 2) * kTB, ks & 3), idesc_s,
                       ks > 0 ? 1u : 0u);
            umma_commit(&y_empty[sb]);
            umma_commit(&s_full[sb]);
            ++scount;
          }
          if (s + 3 == ns) umma_commit(&x_empty[xb]);  // every logits product of this row tile has been issued
        }
        umma_commit(o_full);
        tr.ev(13);
      }
    }
  } else {
    // ----------------------------------------------------------------------------------------------- softmax / epilogue warps
    // eight warps: two per TMEM lane quarter, each owning 64 of a key block's 128 columns (thread = row x column half)
    const int q = warp & 3, hf = (warp - 2) >> 2;
    const int rl = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
    const float a2 = p.alpha * kLog2e;
    Tracer tr;
    tr.init(2);
    uint32_t scount = 0, pvcount = 0, ti = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
      const int rt = w % p.tiles_r;
      const int bh = w / p.tiles_r;
      const int h = bh % p.heads, b = bh / p.heads;
      const int row = rt * 128 + rl;
      const bool row_ok = row < p.R;
      const bf16* rrow = p.res ? p.res + b * p.r_bs + h * p.r_hs + (long long)row * p.r_ld : nullptr;
      if (rrow && row_ok) prefetch_l2(rrow + (hf * 64 < p.d ? hf * 64 : 0));
      if (warp == 2 && lane == 0) tr.ev(33);
      // ---- one pass over the key blocks: logits -> registers (the ONLY tensor-memory read of a logit), block maximum exchanged
      // with the partner warp of the lane quarter, P = 2^(a2 (S - m_ref)) -> bf16 -> tensor memory, row sums.
      float m_ref = -INFINITY;  // reference maximum (raw logits; alpha > 0 commutes with max)
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
      const int pair_bar = 2 + q;  // named barrier of the two warps that share this lane quarter
      for (int jb = 0; jb < p.nb; ++jb, ++scount, ++pvcount) {
        const int sb = scount & 1, pb = pvcount & 1;
        mbar_wait(&s_full[sb], (scount >> 1) & 1);
        tc_fence_after();
        if (warp == 2 && lane == 0) tr.ev(23);
        const int cvalid = p.C - jb * 128 - hf * 64;  // keys of this warp's half that exist
        uint32_t a0[32], a1[32];
        if (cvalid > 0) tmem_ld_32x32(lane_base + sb * 128 + hf * 64, a0);
        if (cvalid > 32) tmem_ld_32x32(lane_base + sb * 128 + hf * 64 + 32, a1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[sb]);  // the logits are in registers: the buffer may take S(jb + 2)
        float mloc = -INFINITY;
        if (cvalid >= 64) {  // full half block (every block of the model's shapes): no per-element masks
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            mloc = fmaxf(mloc, fmaxf(__uint_as_float(a0[e]), __uint_as_float(a0[e + 1])));
            mloc = fmaxf(mloc, fmaxf(__uint_as_float(a1[e]), __uint_as_float(a1[e + 1])));
          }
        } else {  // ragged tail of the last key block (cold)
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (e < cvalid) mloc = fmaxf(mloc, __uint_as_float(a0[e]));
            if (e + 32 < cvalid) mloc = fmaxf(mloc, __uint_as_float(a1[e]));
          }
        }
        float* pm = pairmax + (jb & 1) * 256;  // parity-double-buffered: the slot is rewritten two blocks (= one pair barrier) later
        pm[hf * 128 + rl] = mloc;
        named_bar_sync(pair_bar, 64);
        const float mblk = fmaxf(pm[rl], pm[128 + rl]);
        // the reference only moves when a block exceeds it by more than 2^20 (or at the first block)
        const bool move = jb == 0 || (mblk - m_ref) * a2 > 20.f;
        if (jb > 0 && __any_sync(0xffffffffu, move)) {
          // rare path: rescale this warp's half of the accumulator columns and the row sums.  All P V products issued so far
          // (blocks < jb) must have retired: the last one releases p_empty of its buffer.
          const uint32_t prev = pvcount - 1;
          mbar_wait(&p_empty[prev & 1], (prev >> 1) & 1);
          tc_fence_after();
          const float f = move ? fast_exp2((m_ref - mblk) * a2) : 1.f;
#pragma unroll 1
          for (int c0 = hf * 16; c0 < p.dpad; c0 += 32) {
            uint32_t r[16];
            tmem_ld_32x32_x16(lane_base + kColO + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * f);
            tmem_st_32x32_x16(lane_base + kColO + c0, r);
          }
          tmem_st_wait();
#pragma unroll
          for (int e = 0; e < 4; ++e) l4[e] *= f;
        }
        if (move) m_ref = mblk;
        const float m2 = m_ref * a2;
        mbar_wait(&p_empty[pb], ((pvcount >> 1) & 1) ^ 1);  // P V of two blocks ago has consumed this probability buffer
        tc_fence_after();
        uint32_t wd[16];
        if (cvalid >= 64) {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float x0 = fast_exp2(fmaf(a2, __uint_as_float(a0[2 * u]), -m2)), x1 = fast_exp2(fmaf(a2, __uint_as_float(a0[2 * u + 1]), -m2));
            l4[u & 3] += x0 + x1;
            wd[u] = pack_bf16(x0, x1);
          }
          tmem_st_32x32_x16(lane_base + kColP + pb * 64 + hf * 32, wd);
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float x0 = fast_exp2(fmaf(a2, __uint_as_float(a1[2 * u]), -m2)), x1 = fast_exp2(fmaf(a2, __uint_as_float(a1[2 * u + 1]), -m2));
            l4[u & 3] += x0 + x1;
            wd[u] = pack_bf16(x0, x1);
          }
          tmem_st_32x32_x16(lane_base + kColP + pb * 64 + hf * 32 + 16, wd);
        } else {  // ragged tail of the last key block (cold)
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float x0 = 2 * u < cvalid ? fast_exp2(fmaf(a2, __uint_as_float(a0[2 * u]), -m2)) : 0.f;
            const float x1 = 2 * u + 1 < cvalid ? fast_exp2(fmaf(a2, __uint_as_float(a0[2 * u + 1]), -m2)) : 0.f;
            l4[u & 3] += x0 + x1;
            wd[u] = pack_bf16(x0, x1);
          }
          tmem_st_32x32_x16(lane_base + kColP + pb * 64 + hf * 32, wd);
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float x0 = 2 * u + 32 < cvalid ? fast_exp2(fmaf(a2, __uint_as_float(a1[2 * u]), -m2)) : 0.f;
            const float x1 = 2 * u + 33 < cvalid ? fast_exp2(fmaf(a2, __uint_as_float(a1[2 * u + 1]), -m2)) : 0.f;
            l4[u & 3] += x0 + x1;
            wd[u] = pack_bf16(x0, x1);
          }
          tmem_st_32x32_x16(lane_base + kColP + pb * 64 + hf * 32 + 16, wd);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[pb]);
        if (warp == 2 && lane == 0) tr.ev(26);
      }
      const float m2 = m_ref * a2;
      rowsum[hf * 128 + rl] = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      named_bar_sync(pair_bar, 64);
      if (warp == 2 && lane == 0) tr.ev(27);
      const float l = rowsum[rl] + rowsum[128 + rl];
      const float inv = 1.f / l;
      // ---- epilogue: out = O / l (+ residual), bf16.  The accumulator is drained in 16-column pieces; the two warps of a lane
      // quarter take alternate pieces, so both finish together.  Everything with a latency is issued up front: the residual
      // pieces travel while the last P V products retire, the accumulator pieces are fetched with ONE wait.
      bf16* orow = p.out + b * p.o_bs + h * p.o_hs + (long long)row * p.o_ld;
      constexpr int kPieces = 4;  // d <= 128: at most four 16-column pieces per warp
      uint4 rpre[kPieces][2];
#pragma unroll
      for (int i = 0; i < kPieces; ++i) {
        const int c0 = hf * 16 + i * 32;
        rpre[i][0] = rpre[i][1] = make_uint4(0u, 0u, 0u, 0u);
        if (rrow && row_ok && c0 < p.d) {
          rpre[i][0] = ldg_v4(rrow + c0);
          if (c0 + 8 < p.d) rpre[i][1] = ldg_v4(rrow + c0 + 8);
        }
      }
      mbar_wait(o_full, ti & 1);
      tc_fence_after();
      if (warp == 2 && lane == 0) tr.ev(28);
      uint32_t oacc[kPieces][16];
#pragma unroll
      for (int i = 0; i < kPieces; ++i)
        if (hf * 16 + i * 32 < p.d) tmem_ld_32x32_x16(lane_base + kColO + hf * 16 + i * 32, oacc[i]);
      tmem_ld_wait();
      if (warp == 2 && lane == 0) tr.ev(31);
      tc_fence_before();  // the accumulator is in registers: the next tile's P V products may overwrite it
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < kPieces; ++i) {
          const int c0 = hf * 16 + i * 32;
          if (c0 < p.d) {
            const uint32_t rw[8] = {rpre[i][0].x, rpre[i][0].y, rpre[i][0].z, rpre[i][0].w, rpre[i][1].x, rpre[i][1].y, rpre[i][1].z, rpre[i][1].w};
            uint32_t o[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
              o[u] = pack_bf16(fmaf(__uint_as_float(oacc[i][2 * u]), inv, __uint_as_float(rw[u] << 16)),
                               fmaf(__uint_as_float(oacc[i][2 * u + 1]), inv, __uint_as_float(rw[u] & 0xffff0000u)));
            stg_v4(orow + c0, make_uint4(o[0], o[1], o[2], o[3]));
            if (c0 + 8 < p.d) stg_v4(orow + c0 + 8, make_uint4(o[4], o[5], o[6], o[7]));
          }
        }
      }
      if (hf == 0 && row_ok && p.lse2) p.lse2[((long long)b * p.heads + h) * p.R + row] = m2 + fast_log2(l);
      if (warp == 2 && lane == 0) tr.ev(29);
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
// dS (and P') overwrite the logits (dP) they were computed from in tensor memory and are the TMEM A operand of the output
// products; the B_blk / D_blk tiles are loaded once and used twice: as the K-major operand of the logits products and as
// the MN-major operand of the output products (the [64 x 64] swizzled tile is the same bytes under both readings).
// The tensor-core instructions execute in issue order, so a logits buffer needs no "empty" barrier: the products that
// overwrite it are issued after the output products that read it.
constexpr int kThreadsB = 64 + 8 * 32;

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
  int ntb, nst;       // tile-operand buffers (1 or 2) and block stages (4 or 2): six 32 KB units in total
};

struct BwdSmem {
  static constexpr int UNIT = 4 * kHB;     // 32 KB: (A_T or C_T: two K blocks of [128 x 64]) or (one stage: B_blk + D_blk)
  static constexpr int UNITS = 6;
  static constexpr int CS = 8 * 32 * 8;    // per-warp column statistics (COLS only)
  static constexpr int BARS = 32 * 8 + 16;
  static constexpr int TOTAL = UNITS * UNIT + CS + BARS + 1024;
};

__device__ __forceinline__ void bwd_store_out(const BwdOut& o, uint32_t taddr, int c_lo, int c_hi, int d, int b, int h, int row, bool row_ok) {
  const long long base = b * o.bs + h * o.hs + (long long)row * o.ld;
  const bf16* rrow = o.res ? o.res + b * o.r_bs + h * o.r_hs + (long long)(row / o.row_div) * o.r_ld : nullptr;
#pragma unroll 1
  for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
    uint4 rres[4];
    if (rrow && row_ok) {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (c0 + g * 8 < d) rres[g] = ldg_v4(rrow + c0 + g * 8);
    }
    uint32_t acc[32];
    tmem_ld_32x32(taddr + c0, acc);
    tmem_ld_wait();
    if (!row_ok) continue;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int c = c0 + g * 8;
      if (c < d) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(acc[g * 8 + e]);
        if (rrow) {
          const uint32_t rw[4] = {rres[g].x, rres[g].y, rres[g].z, rres[g].w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            f[2 * u] = fmaf(o.rscale, __uint_as_float(rw[u] << 16), f[2 * u]);
            f[2 * u + 1] = fmaf(o.rscale, __uint_as_float(rw[u] & 0xffff0000u), f[2 * u + 1]);
          }
        }
        if (o.is_f32) {
          float* op = reinterpret_cast<float*>(o.ptr) + base + c;
          stg_v4(op, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
          stg_v4(op + 4, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
        } else {
          stg_v4(reinterpret_cast<bf16*>(o.ptr) + base + c,
                 make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7])));
        }
      }
    }
  }
}

// NKS: k-steps (of 16) of the head dim when known at compile time (6: the model's d = 96), 0 = run-time count.  The issue loops of
// the MMA warp are on the critical path of every block AND the kernel has to stay inside the 32 KB instruction cache: an exact
// unroll is shorter than a predicated unroll by eight.
template <bool COLS, int NKS>
__global__ void __launch_bounds__(kThreadsB, 1)
flash_bwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                 const __grid_constant__ CUtensorMap tmD, const __grid_constant__ BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sTile = smem;                               // ntb x (A_T | C_T)
  uint8_t* sSt = smem + p.ntb * 2 * BwdSmem::UNIT;     // nst stages
  float2* colstat = reinterpret_cast<float2*>(smem + BwdSmem::UNITS * BwdSmem::UNIT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::UNITS * BwdSmem::UNIT + BwdSmem::CS);
  uint64_t* at_full = bars;        // [2]
  uint64_t* at_empty = bars + 2;   // [2]
  uint64_t* st_full = bars + 4;    // [4]
  uint64_t* st_empty = bars + 8;   // [4]
  uint64_t* s_full = bars + 12;    // [2]
  uint64_t* ds_full = bars + 14;   // [2]
  uint64_t* o_full = bars + 16;
  uint64_t* o_empty = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.d + 63) / 64;
  const int nks = NKS ? NKS : (p.d + 15) / 16;
  const int total = p.batch * p.heads * p.tiles;
  const int ntb = p.ntb, nst = p.nst;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    tma_prefetch_desc(&tmD);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&at_full[i], 1);
        mbar_init(&at_empty[i], 1);
        mbar_init(&s_full[i], 1);
        mbar_init(&ds_full[i], 8);
      }
      for (int i = 0; i < 4; ++i) {
        mbar_init(&st_full[i], 1);
        mbar_init(&st_empty[i], 1);
      }
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
  pdl_wait();
  pdl_launch_dependents();
  constexpr uint32_t kCol1 = 256, kCol2 = 384;  // TMEM: (S | dP) buffers at [0,128) and [128,256); out1 at 256, out2 at 384

  if (warp == 0) {
    // ----------------------------------------------------------------------------------------------- TMA producer
    if (elect_one()) {
      Tracer tr;
      tr.init(0);
      uint32_t bc = 0, ti = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
        const int tt = w % p.tiles;
        const int bh = w / p.tiles;
        const int h = bh % p.heads, b = bh / p.heads;
        const int tb = ti % ntb;
        mbar_wait(&at_empty[tb], ((ti / ntb) & 1) ^ 1);
        mbar_expect_tx(&at_full[tb], 2 * nkb * kTB);
        uint8_t* ta = sTile + tb * 2 * BwdSmem::UNIT;
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb) {
          tma_load_4d(&tmA, &at_full[tb], ta + kb * kTB, kb * 64, tt * 128, h, b);
          tma_load_4d(&tmC, &at_full[tb], ta + BwdSmem::UNIT + kb * kTB, kb * 64, tt * 128, h, b);
        }
        tr.ev(60);
#pragma unroll 1
        for (int blk = 0; blk < p.nblk; ++blk, ++bc) {
          const int st = bc % nst;
          mbar_wait(&st_empty[st], ((bc / nst) & 1) ^ 1);
          mbar_expect_tx(&st_full[st], 2 * nkb * kHB);
          uint8_t* base = sSt + st * BwdSmem::UNIT;
#pragma unroll 1
          for (int kb = 0; kb < nkb; ++kb) {
            tma_load_4d(&tmB, &st_full[st], base + kb * kHB, kb * 64, blk * 64, h, b);
            tma_load_4d(&tmD, &st_full[st], base + (2 + kb) * kHB, kb * 64, blk * 64, h, b);
          }
          tr.ev(61);
        }
      }
    }
  } else if (warp == 1) {
    // ----------------------------------------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, p.dpad, 0, 1);
      uint32_t sdc = 0, oc = 0, ti = 0;  // blocks whose logits products / output products have been issued
      Tracer tr;
      tr.init(1);
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
        const int tb = ti % ntb;
        mbar_wait(&at_full[tb], (ti / ntb) & 1);
        tc_fence_after();
        tr.ev(40);
        const uint32_t aa = smem_u32(sTile + tb * 2 * BwdSmem::UNIT), ca = aa + BwdSmem::UNIT;
        int issued = 0;
        // Descriptors are formed ONCE per tile / stage and advanced by compile-time constants per k-step (one uniform add each):
        // this single thread's issue rate is on the critical path of every block (16-20 tcgen05.mma per 64-key block).
        const uint64_t dA0 = desc_kmajor(aa, 0), dC0 = desc_kmajor(ca, 0);
        auto issue_sd = [&]() {
          const int buf = sdc & 1, st = sdc % nst;
          mbar_wait(&st_full[st], (sdc / nst) & 1);
          tc_fence_after();
          const uint32_t ba = smem_u32(sSt + st * BwdSmem::UNIT), da = ba + 2 * kHB;
          const uint64_t dB0 = desc_kmajor(ba, 0), dD0 = desc_kmajor(da, 0);
          const uint32_t ts = tmem_base + buf * 128;
#pragma unroll
          for (int ks = 0; ks < (NKS ? NKS : 8); ++ks)
            if (NKS || ks < nks)
              umma_f16(ts, dA0 + ((((ks >> 2) * kTB) + (ks & 3) * 32) >> 4), dB0 + ((((ks >> 2) * kHB) + (ks & 3) * 32) >> 4), idesc_s,
                       ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < (NKS ? NKS : 8); ++ks)
            if (NKS || ks < nks)
              umma_f16(ts + 64, dC0 + ((((ks >> 2) * kTB) + (ks & 3) * 32) >> 4), dD0 + ((((ks >> 2) * kHB) + (ks & 3) * 32) >> 4), idesc_s,
                       ks > 0 ? 1u : 0u);
          umma_commit(&s_full[buf]);
          tr.ev(41);
          ++sdc;
          if (++issued == p.nblk) umma_commit(&at_empty[tb]);  // all logits products of this tile are issued
        };
        issue_sd();
        if (p.nblk > 1) issue_sd();
        for (int blk = 0; blk < p.nblk; ++blk) {
          const int buf = oc & 1, st = oc % nst;
          if (blk == 0) mbar_wait(o_empty, (ti & 1) ^ 1);
          mbar_wait(&ds_full[buf], (oc >> 1) & 1);
          tc_fence_after();
          tr.ev(42);
          const uint32_t ba = smem_u32(sSt + st * BwdSmem::UNIT), da = ba + 2 * kHB;
          const uint64_t dBm = desc_mnmajor(ba, kHB, 0), dDm = desc_mnmajor(da, kHB, 0);
          const uint32_t tb_ = tmem_base + buf * 128;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // contraction over the block's 64 columns: 8 TMEM columns (16 bf16) of dS per step
            umma_f16_ts(tmem_base + kCol1, tb_ + (kk >> 1) * 32 + (kk & 1) * 8, dBm + kk * (2048 >> 4), idesc_o, (blk > 0 || kk > 0) ? 1u : 0u);
          if (COLS) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16_ts(tmem_base + kCol2, tb_ + 64 + (kk >> 1) * 32 + (kk & 1) * 8, dDm + kk * (2048 >> 4), idesc_o, (blk > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&st_empty[st]);
          tr.ev(43);
          ++oc;
          if (blk + 2 < p.nblk) issue_sd();  // overwrites the buffer whose dS / P the products above have just read (in-order pipe)
        }
        umma_commit(o_full);
        tr.ev(44);
      }
    }
  } else {
    // ----------------------------------------------------------------------------------------------- softmax / epilogue warps
    const int q = warp & 3, hf = (warp - 2) >> 2;  // TMEM lane quarter; which 32 of the block's 64 columns
    const int rl = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
    const float a2 = p.alpha * kLog2e;
    const float log2_alpha = fast_log2(p.alpha);
    float2* cs = colstat + (warp - 2) * 32;  // this warp's private copy of its 32 columns' statistics
    uint32_t bc = 0, ti = 0;
    Tracer tr;
    tr.init(2);
    const bool tracer = warp == 2 && lane == 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
      if (tracer) tr.ev(50);
      const int tt = w % p.tiles;
      const int bh = w / p.tiles;
      const int h = bh % p.heads, b = bh / p.heads;
      const int row = tt * 128 + rl;
      const bool row_ok = row < p.T;
      const float* lse_bh = p.lse2 + (long long)bh * p.n_rows;
      const float* dot_bh = p.dot + (long long)bh * p.n_rows;
      float r_lse = INFINITY, r_dot = 0.f;
      float2 nxt = make_float2(INFINITY, 0.f);
      if (COLS) {
        const int i = hf * 32 + lane;
        if (i < p.L) nxt = make_float2(lse_bh[i], p.alpha * dot_bh[i]);  // (lse, alpha dot): dS = P (alpha dP - alpha dot)
      } else if (row_ok) {
        r_lse = lse_bh[row] - log2_alpha;  // 2^(a2 S - lse + log2 alpha) = alpha P
        r_dot = dot_bh[row];
      }
      {  // residual rows of the epilogue: pull them towards L2 while the tile computes
        const BwdOut& o = (COLS && hf) ? p.o2 : p.o1;
        if (o.res && row_ok) {
          const bf16* rr = o.res + b * o.r_bs + h * o.r_hs + (long long)(row / o.row_div) * o.r_ld;
          prefetch_l2(rr);
          if (p.d > 64) prefetch_l2(rr + 64);
        }
      }
      for (int blk = 0; blk < p.nblk; ++blk, ++bc) {
        const int buf = bc & 1;
        if (COLS) {  // statistics of this block's softmax rows (this warp's 32 columns); the next block's travel meanwhile
          __syncwarp();
          cs[lane] = nxt;
          __syncwarp();
          const int i = (blk + 1) * 64 + hf * 32 + lane;
          nxt = (blk + 1 < p.nblk && i < p.L) ? make_float2(lse_bh[i], p.alpha * dot_bh[i]) : make_float2(INFINITY, 0.f);
        }
        mbar_wait(&s_full[buf], (bc >> 1) & 1);
        tc_fence_after();
        if (tracer) tr.ev(51);
        uint32_t sa[32], da[32];
        tmem_ld_32x32(lane_base + buf * 128 + hf * 32, sa);
        tmem_ld_32x32(lane_base + buf * 128 + 64 + hf * 32, da);
        tmem_ld_wait();
        uint32_t wds[16], wpp[16];
        if (COLS) {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float4 cc = reinterpret_cast<const float4*>(cs)[u];  // (lse, alpha dot) of two columns in one 16-byte broadcast load
            const float2 c0 = make_float2(cc.x, cc.y), c1 = make_float2(cc.z, cc.w);
            const float p0 = fast_exp2(fmaf(a2, __uint_as_float(sa[2 * u]), -c0.x));
            const float p1 = fast_exp2(fmaf(a2, __uint_as_float(sa[2 * u + 1]), -c1.x));
            wpp[u] = pack_bf16(p0, p1);
            wds[u] = pack_bf16(p0 * fmaf(p.alpha, __uint_as_float(da[2 * u]), -c0.y), p1 * fmaf(p.alpha, __uint_as_float(da[2 * u + 1]), -c1.y));
          }
        } else {
          const int cvalid = p.L - blk * 64 - hf * 32;  // keys of this chunk that exist
          // alpha rides in the exponent: alpha 2^x = 2^(x + log2 alpha), so dS = (alpha P) (dP - dot) costs one FADD + one FMUL
          if (cvalid >= 32) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const float p0 = fast_exp2(fmaf(a2, __uint_as_float(sa[2 * u]), -r_lse));
              const float p1 = fast_exp2(fmaf(a2, __uint_as_float(sa[2 * u + 1]), -r_lse));
              wds[u] = pack_bf16(p0 * (__uint_as_float(da[2 * u]) - r_dot), p1 * (__uint_as_float(da[2 * u + 1]) - r_dot));
            }
          } else {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const float p0 = 2 * u < cvalid ? fast_exp2(fmaf(a2, __uint_as_float(sa[2 * u]), -r_lse)) : 0.f;
              const float p1 = 2 * u + 1 < cvalid ? fast_exp2(fmaf(a2, __uint_as_float(sa[2 * u + 1]), -r_lse)) : 0.f;
              wds[u] = pack_bf16(p0 * (__uint_as_float(da[2 * u]) - r_dot), p1 * (__uint_as_float(da[2 * u + 1]) - r_dot));
            }
          }
        }
        // dS over the logits it came from, P' over dP (both already in registers); 16 TMEM columns = this warp's 32 bf16 columns
        // (each warp writes only over the columns it has itself just read: [hf*32, hf*32+16) of the S / dP regions)
        tmem_st_32x32_x16(lane_base + buf * 128 + hf * 32, wds);
        if (COLS) tmem_st_32x32_x16(lane_base + buf * 128 + 64 + hf * 32, wpp);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ds_full[buf]);
        if (tracer) tr.ev(52);
      }
      // ---- epilogue.  COLS: the two warps of a lane quarter store one output each; ROWS: one half of out1's columns each
      mbar_wait(o_full, ti & 1);
      tc_fence_after();
      if (tracer) tr.ev(53);
      if (COLS) {
        if (hf == 0) bwd_store_out(p.o1, lane_base + kCol1, 0, p.dpad, p.d, b, h, row, row_ok);
        else bwd_store_out(p.o2, lane_base + kCol2, 0, p.dpad, p.d, b, h, row, row_ok);
      } else {
        bwd_store_out(p.o1, lane_base + kCol1, hf * 64, min(p.dpad, hf * 64 + 64), p.d, b, h, row, row_ok);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      if (tracer) tr.ev(54);
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
  if (p.nb <= 4) {
    static DeviceOnce once2;
    if (once2.first()) MB_CUDA(cudaFuncSetAttribute(flash_fwd_twopass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL));
    MB_CUDA(launch_pdl(flash_fwd_twopass_kernel, dim3(grid), dim3(kThreadsF), FwdSmem::TOTAL, STREAM, tmX, tmY, tmV, p));
  } else {
    MB_CUDA(launch_pdl(flash_fwd_kernel, dim3(grid), dim3(kThreadsF), FwdSmem::TOTAL, STREAM, tmX, tmY, tmV, p));
  }
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
  // six 32 KB units of shared memory: short tiles (few blocks) double-buffer the tile operands, long tiles deepen the block ring
  // (measured with the event trace: with two stages a short tile stalls ~1.4 k cycles per block on the TMA latency of the stage that
  // the output products have just released; four stages keep the loads two blocks ahead, also across tile boundaries)
  p.ntb = 1;
  p.nst = 4;
  const long long total = (long long)p.batch * p.heads * p.tiles;
  const int grid = (int)(total < num_sms() ? total : num_sms());
  auto go = [&](auto kern) -> int {
    MB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::TOTAL));  // idempotent, cheap
    MB_CUDA(launch_pdl(kern, dim3(grid), dim3(kThreadsB), BwdSmem::TOTAL, STREAM, tmA, tmB, tmC, tmD, p));
    return 0;
  };
  const bool k6 = (p.d + 15) / 16 == 6;
  rc = a->cols ? (k6 ? go(flash_bwd_kernel<true, 6>) : go(flash_bwd_kernel<true, 0>))
               : (k6 ? go(flash_bwd_kernel<false, 6>) : go(flash_bwd_kernel<false, 0>));
  if (rc) return rc;
  return 0;
}

// debug hook (measurement only): install / remove the event-trace buffer read by tools/flash_trace.py
extern "C" int mirror_debug_flash_trace(void* buf, int64_t capacity) {
  unsigned long long* pbuf = reinterpret_cast<unsigned long long*>(buf);
  unsigned long long cap = (unsigned long long)capacity;
  MB_CUDA(cudaMemcpyToSymbol(g_trace, &pbuf, sizeof(pbuf)));
  MB_CUDA(cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap)));
  return 0;
}
