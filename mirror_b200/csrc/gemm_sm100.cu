// Batched bf16 GEMM for sm_100a: TMA -> 128B-swizzled shared memory -> tcgen05.mma (TMEM
// accumulators, double buffered) -> fused epilogue.  Persistent, warp specialised:
//   warp 0      TMA producer (one elected lane)
//   warp 1      TMEM allocator + MMA issuer (one elected lane)
//   warps 2..9  epilogue (two warps per TMEM lane quarter, each owning half of the tile's columns):
//               residual prefetch -> tcgen05.ld -> alpha/diag/bias/act/dropout/residual/beta -> 128-bit global stores
// See include/mirror_b200.h (mirror_gemm_bf16) for the contract and the reference call sites.
#include <cudaTypedefs.h>
#include <stdlib.h>

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
constexpr int kMaxAccStages = 4;  // barrier slots reserved; 128-wide tiles use all four (4 x 128 = the 512 TMEM columns)

struct KParams {
  Epi e;
  int K, batch1, batch2, tiles_m, tiles_n, split_k;
  int a_b1, a_b2, b_b1, b_b2;  // 0 where an operand is broadcast along that batch dim (batch stride 0), else 1
};

constexpr int kSinkBytes = 4096;  // per epilogue warp: one 32x32 f32 box (4 KB, 128B swizzle) or two bf16 boxes (2 KB, 64B swizzle)

// TMAS: the epilogue hands finished 32x32 chunks to TMA (bulk tensor stores out of a per-warp staging box) instead of
// storing from registers.  The 32 KB of boxes fit beside the full pipeline for BN = 128 / 256; BN = 192 gives up a stage.
template <int BN, bool TMAS = false>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = TMAS ? kEpiWarps * kSinkBytes : 0;
  static constexpr int STAGES = ((BN == 256) ? 4 : (BN == 192 ? 5 : 6)) - ((TMAS && BN == 192) ? 1 : 0);
  // accumulator stages in TMEM: the epilogue of tile i overlaps the MMAs of the following tiles; 4 for the 128-wide tile
  static constexpr int ACC_STAGES = (BN == 128) ? 4 : kAccStages;
  static constexpr int TMEM_COLS = (ACC_STAGES * BN <= 256) ? 256 : 512;
  static constexpr int BAR_BYTES = (2 * STAGES + 2 * kMaxAccStages) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
};

// ---------------------------------------------------------------------------------------------- epilogue
// A 32-row x 32-column chunk (one epilogue warp, one tcgen05.ld.32x32b.x32): thread = row, 32 consecutive columns.
// (A variant that transposed the chunk through swizzled shared memory for fully coalesced stores was measured
// 20-45 % SLOWER on every epilogue-bound shape and was dropped; see DESIGN.md.)
// ResBuf holds the residual (or the accumulate target) of one chunk; it is fetched BEFORE the accumulator is ready
// so that the global-load latency hides behind the tile's MMAs.
struct ResBuf {
  uint4 v[8];
};

// slow path (ragged N tail, unaligned views, split-K atomics, GELU).  Out of line, and it re-reads TMEM itself so that
// the hot path never has to keep the accumulator chunk addressable (= spilled to local memory).  Epi travels BY VALUE: a
// reference would let the address of the kernel parameter escape, and then every e.field of the hot path becomes a generic
// LD.E (re-issued after each asm volatile) instead of a constant-bank operand -- measured: long-scoreboard stalls all over.
__device__ __noinline__ void epilogue_chunk_scalar(const Epi e, uint32_t taddr, int b1, int b2, int row, int col0) {
  uint32_t acc[32];
  tmem_ld_32x32(taddr, acc);
  tmem_ld_wait();
  if (row >= e.M) return;
#pragma unroll 1
  for (int j = 0; j < 32; ++j)
    if (col0 + j < e.N) epi_store_scalar(e, __uint_as_float(acc[j]), b1, b2, row, col0 + j);
}

__device__ __forceinline__ void res_prefetch(const Epi& e, ResBuf& rb, int b1, int b2, int row0, int col0, int lane) {
  const int row = row0 + lane;
  if (row >= e.M) return;
  if (e.res) {
    const long long off = b2 * e.r_bs2 + b1 * e.r_bs1 + (long long)(row / e.res_row_div) * e.ldr + col0;
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

// Where finished chunks go.  tma = 0: straight from registers (store_rows_paired).  tma = 1: through this warp's staging
// boxes and a bulk tensor store -- the warp only pays conflict-free st.shared, the TMA unit writes whole lines
// asynchronously, and row/column tails are clipped by the tensor map.
struct Sink {
  const CUtensorMap* tm32;
  const CUtensorMap* tm16;
  uint32_t buf;  // shared address of this warp's kSinkBytes
  int tma;
  int flip;      // bf16-only output: alternates between two staging boxes
};

// f32 chunk: lane = row, 8 pieces of 16 B.  128B swizzle: piece j of row r lives at r*128 + ((j ^ (r & 7)) * 16).
__device__ __forceinline__ void sink_tma_f32(Sink& s, const uint4 (&pc)[8], bool both, int b1, int b2, int row0, int col0, int lane) {
  (void)both;
  if (lane == 0) bulk_wait_read<0>();  // the box (shared with the bf16 boxes) must have been read out
  __syncwarp();
  const uint32_t base = s.buf + lane * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j) st_shared_v4(base + ((j ^ (lane & 7)) << 4), pc[j]);
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_4d(s.tm32, s.buf, col0, row0, b1, b2);
    bulk_commit();
  }
}

// bf16 chunk: 4 pieces of 16 B per row.  64B swizzle: piece j of row r lives at r*64 + ((j ^ ((r >> 1) & 3)) * 16).
__device__ __forceinline__ void sink_tma_bf16(Sink& s, const uint4 (&pc)[4], bool both, int b1, int b2, int row0, int col0, int lane) {
  if (lane == 0) {
    if (both) bulk_wait_read<0>();  // the f32 box of this chunk occupies the same bytes
    else bulk_wait_read<1>();       // two boxes in rotation: one group may stay in flight
  }
  __syncwarp();
  const uint32_t box = s.buf + (s.flip ? 2048 : 0);
  s.flip ^= 1;
  const uint32_t base = box + lane * 64;
#pragma unroll
  for (int j = 0; j < 4; ++j) st_shared_v4(base + ((j ^ ((lane >> 1) & 3)) << 4), pc[j]);
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_4d(s.tm16, box, col0, row0, b1, b2);
    bulk_commit();
  }
}

// Store NP 16-byte pieces per row so that every warp-wide store instruction writes FULL 32-byte sectors: lanes 2i and
// 2i+1 (rows A = row0+2i and B = A+1) swap half of their pieces, then both write into the same row -- instruction j
// covers bytes [32j, 32j+32) of row A (first NP/2 instructions) or row B (last NP/2).  With one row per lane each
// instruction would touch 32 half-filled sectors, and the L1 -> L2 write path (not DRAM) is what saturates.
template <int NP>
__device__ __forceinline__ void store_rows_paired(uint4 (&pc)[NP], char* rowA, long long row_bytes, bool okA, bool okB, int lane) {
  const bool odd = lane & 1;
  uint4 rc[NP / 2];
#pragma unroll
  for (int j = 0; j < NP / 2; ++j) {  // even lane sends its odd pieces, odd lane its even pieces
    const uint4 snd = odd ? pc[2 * j] : pc[2 * j + 1];
    rc[j].x = __shfl_xor_sync(0xffffffffu, snd.x, 1);
    rc[j].y = __shfl_xor_sync(0xffffffffu, snd.y, 1);
    rc[j].z = __shfl_xor_sync(0xffffffffu, snd.z, 1);
    rc[j].w = __shfl_xor_sync(0xffffffffu, snd.w, 1);
  }
  char* a = rowA + (odd ? 16 : 0);
  char* b = rowA + row_bytes + (odd ? 16 : 0);
#pragma unroll
  for (int j = 0; j < NP / 2; ++j)  // row A: even lane writes its own piece 2j, odd lane the received piece 2j+1
    if (okA) *reinterpret_cast<uint4*>(a + 32 * j) = odd ? rc[j] : pc[2 * j];
#pragma unroll
  for (int j = 0; j < NP / 2; ++j)  // row B: even lane writes the received piece 2j, odd lane its own piece 2j+1
    if (okB) *reinterpret_cast<uint4*>(b + 32 * j) = odd ? pc[2 * j + 1] : rc[j];
}

__device__ __forceinline__ void epilogue_chunk_vec(const Epi& e, uint32_t taddr, const ResBuf& rb, Sink& sink, int b1, int b2,
                                                   int row0, int col0, int lane) {
  uint32_t acc[32];
  tmem_ld_32x32(taddr, acc);
  tmem_ld_wait();
  const int row = row0 + lane;
  const bool row_ok = row < e.M;
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
  if (e.act == MIRROR_ACT_RELU) {  // (GELU goes through the scalar path: only tiny GEMMs use it)
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (e.drop_p > 0.f) {
    const uint64_t base = ((uint64_t)(b2 * e.batch1 + b1) * e.M + row) * e.N + col0;
    const uint64_t seed = epoch_seed(e.drop_seed, e.drop_epoch);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = hash_u01(seed, base + j) >= e.drop_p ? v[j] * e.drop_scale : 0.f;
  }
  if (e.res && row_ok) {
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
  if (e.res2 && row_ok) {
    const uint4* rp = reinterpret_cast<const uint4*>(e.res2 + b2 * e.r_bs2 + b1 * e.r_bs1 + (long long)(row / e.res_row_div) * e.ldr + col0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 u = rp[j];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __bfloat1622float2(h[t]);
        v[j * 8 + 2 * t] += e.gamma2 * f.x;
        v[j * 8 + 2 * t + 1] += e.gamma2 * f.y;
      }
    }
  }
  const int rowA = row0 + (lane & ~1);
  const bool okA = rowA < e.M, okB = rowA + 1 < e.M;
  if (e.o32) {
    const long long off = b2 * e.c32_bs2 + b1 * e.c32_bs1 + col0;
    if (e.beta != 0.f && row_ok) {
      const float4* op = reinterpret_cast<const float4*>(e.o32 + off + (long long)row * e.ldc32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 f = e.res ? op[j] : *reinterpret_cast<const float4*>(&rb.v[j]);  // prefetched unless res took the buffer
        v[4 * j] += e.beta * f.x; v[4 * j + 1] += e.beta * f.y;
        v[4 * j + 2] += e.beta * f.z; v[4 * j + 3] += e.beta * f.w;
      }
    }
    uint4 pc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      pc[j] = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
    if (sink.tma) sink_tma_f32(sink, pc, e.o16 != nullptr, b1, b2, row0, col0, lane);
    else store_rows_paired<8>(pc, reinterpret_cast<char*>(e.o32 + off + (long long)rowA * e.ldc32), e.ldc32 * 4, okA, okB, lane);
  }
  if (e.o16) {
    uint4 pc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pc[j]);
#pragma unroll
      for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[j * 8 + 2 * t], v[j * 8 + 2 * t + 1]);
    }
    if (sink.tma) sink_tma_bf16(sink, pc, e.o32 != nullptr, b1, b2, row0, col0, lane);
    else store_rows_paired<4>(pc, reinterpret_cast<char*>(e.o16 + b2 * e.c16_bs2 + b1 * e.c16_bs1 + col0 + (long long)rowA * e.ldc16),
                              e.ldc16 * 2, okA, okB, lane);
  }
}

// Fused row-softmax epilogues (MIRROR_GEMM_ROWSTATS / SOFTMAX / ROWDOT / SOFTMAX_BWD): thread = row, this warp's BN/2
// columns are one "part" of the row.  Logits live in the base-2 domain (x2 = alpha*log2(e)*acc) so exp is one MUFU.EX2.
// MODE != 0: the mode is a compile-time constant of the instantiation (all mode tests fold away); 0: read it from e.mode
template <int BN, int MODE = 0>
__device__ __forceinline__ void epilogue_tile_softmax(const Epi& e, Sink& sink, uint32_t taddr, int b1, int b2, int row0, int cbase,
                                                      int lane, uint64_t* tfull_bar, uint32_t aphase, int part) {
  constexpr int NCH = BN / 64;
  const int mode = MODE ? MODE : e.mode;
  const int row = row0 + lane;
  const bool row_ok = row < e.M;
  float2* stats = reinterpret_cast<float2*>(e.stats) + ((long long)(b2 * e.batch1 + b1) * e.M + (row_ok ? row : 0)) * e.nparts;
  const float a2 = e.alpha * 1.4426950408889634f;
  float M2 = 0.f, invS = 0.f, dot = 0.f;
  if (row_ok && mode == MIRROR_GEMM_SOFTMAX) {  // combine the partials of pass 1 (these loads overlap the tile's MMAs)
    float m = -INFINITY;
    for (int i = 0; i < e.nparts; ++i) m = fmaxf(m, stats[i].x);
    float ssum = 0.f;
    for (int i = 0; i < e.nparts; ++i) {
      const float2 t = stats[i];
      ssum += t.y * fast_exp2(t.x - m);
    }
    M2 = m;
    invS = 1.f / ssum;
  } else if (row_ok && mode == MIRROR_GEMM_SOFTMAX_BWD) {
    for (int i = 0; i < e.nparts; ++i) dot += stats[i].x;
  } else if (row_ok && mode == MIRROR_GEMM_SOFTMAX_BWD_DOT) {
    dot = e.stats[(long long)(b2 * e.batch1 + b1) * e.M + row];
  }
  mbar_wait(tfull_bar, aphase);
  tc_fence_after();
  if (row0 >= e.M) return;
  const bf16* prow = reinterpret_cast<const bf16*>(e.res) + b2 * e.r_bs2 + b1 * e.r_bs1 + (long long)(row_ok ? row : 0) * e.ldr;
  const int rowA = row0 + (lane & ~1);
  const bool okA = rowA < e.M, okB = rowA + 1 < e.M;
  float run_m = -INFINITY, run_s = 0.f, run_dot = 0.f;
  const float nad = -e.alpha * dot;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col0 = cbase + c * 32;
    if (col0 >= e.N) continue;  // warp-uniform; N % 32 == 0, so a chunk is either complete or absent
    uint32_t acc[32];
    tmem_ld_32x32(taddr + c * 32, acc);
    tmem_ld_wait();
    float v[32];
    if (mode == 6) {  // measurement only (tools/softmax_gemm_probe.py): the cost of draining TMEM and of the tile hand-shake alone
      run_dot += __uint_as_float(acc[0]) + __uint_as_float(acc[31]);
      continue;
    }
    if (mode == MIRROR_GEMM_ROWSTATS) {
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = a2 * __uint_as_float(acc[j]);
        cm = fmaxf(cm, v[j]);
      }
      if (cm > run_m) {
        run_s *= fast_exp2(run_m - cm);
        run_m = cm;
      }
      float s4[4] = {0.f, 0.f, 0.f, 0.f};  // four chains: the adds must not serialise behind the MUFU results
#pragma unroll
      for (int j = 0; j < 32; ++j) s4[j & 3] += fast_exp2(v[j] - run_m);
      run_s += (s4[0] + s4[1]) + (s4[2] + s4[3]);
      continue;
    }
    if (mode == MIRROR_GEMM_SOFTMAX) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fast_exp2(a2 * __uint_as_float(acc[j]) - M2) * invS;
    } else {  // ROWDOT / SOFTMAX_BWD read the probabilities
      uint4 pu[4];
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) pu[j] = reinterpret_cast<const uint4*>(prow + col0)[j];
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) pu[j] = make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pu[j]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __bfloat1622float2(h[t]);
          v[j * 8 + 2 * t] = f.x;
          v[j * 8 + 2 * t + 1] = f.y;
        }
      }
      if (mode == MIRROR_GEMM_ROWDOT) {
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; ++j) d4[j & 3] += v[j] * __uint_as_float(acc[j]);
        run_dot += (d4[0] + d4[1]) + (d4[2] + d4[3]);
        continue;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= fmaf(e.alpha, __uint_as_float(acc[j]), nad);  // P * (alpha*G - alpha*dot): two ops per element
    }
    if (e.o32) {
      uint4 pc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        pc[j] = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
      if (sink.tma) sink_tma_f32(sink, pc, e.o16 != nullptr, b1, b2, row0, col0, lane);
      else store_rows_paired<8>(pc, reinterpret_cast<char*>(e.o32 + b2 * e.c32_bs2 + b1 * e.c32_bs1 + col0 + (long long)rowA * e.ldc32),
                                e.ldc32 * 4, okA, okB, lane);
    }
    if (e.o16) {
      uint4 pc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pc[j]);
#pragma unroll
        for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[j * 8 + 2 * t], v[j * 8 + 2 * t + 1]);
      }
      if (sink.tma) sink_tma_bf16(sink, pc, e.o32 != nullptr, b1, b2, row0, col0, lane);
      else store_rows_paired<4>(pc, reinterpret_cast<char*>(e.o16 + b2 * e.c16_bs2 + b1 * e.c16_bs1 + col0 + (long long)rowA * e.ldc16),
                                e.ldc16 * 2, okA, okB, lane);
    }
  }
  if (row_ok) {
    if (mode == MIRROR_GEMM_ROWSTATS) stats[part] = make_float2(run_m, run_s);
    else if (mode == MIRROR_GEMM_ROWDOT || mode == 6) stats[part] = make_float2(run_dot, 0.f);
  }
}

// One epilogue warp's share of a finished tile: the 32 rows of its TMEM lane quarter (first row row0) x BN/2 columns
// starting at cbase.
// LEAN epilogue: out_bf16 = alpha*acc + diag*I + gamma*res16 + gamma2*res2_16 through the bulk-store sink, nothing else.
// The generic chunk body executes ~230 instructions per 32x32 chunk of which 52 are the multiply / convert / store it is
// there for (the rest tests features and rebuilds addresses); the Moore-Penrose, attention-value and similarity-gradient
// products -- two thirds of all launches, all epilogue-bound -- need none of those features.  The host selects this
// instantiation only when N % 32 == 0 (no ragged chunk), so there is no scalar path either.
template <int BN>
__device__ __forceinline__ void epilogue_tile_lean(const Epi& e, uint32_t taddr, Sink& sink, int b1, int b2, int row0, int cbase,
                                                   int lane, uint64_t* tfull_bar, uint32_t aphase) {
  constexpr int NCH = BN / 64;
  const int row = row0 + lane;
  const bool row_ok = row < e.M;
  const int rrow = e.res_row_div > 1 ? row / e.res_row_div : row;  // row-broadcast residual (landmark-mean backward)
  const long long roff = b2 * e.r_bs2 + b1 * e.r_bs1 + (long long)rrow * e.ldr;
  const bf16* r1 = (e.res && row_ok) ? reinterpret_cast<const bf16*>(e.res) + roff : nullptr;
  const bf16* r2 = (e.res2 && row_ok) ? e.res2 + roff : nullptr;
  // residual of ALL this warp's chunks, requested before the accumulator is waited for: the loads do not depend on it, and with a
  // one-chunk-ahead prefetch 15 % of the kernel's stall samples sat on the first use of the residual (ncu, 384^3 chain products)
  uint4 pa[NCH][4];
  if (r1) {
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      if (cbase + c * 32 < e.N) {
#pragma unroll
        for (int j = 0; j < 4; ++j) pa[c][j] = reinterpret_cast<const uint4*>(r1 + cbase + c * 32)[j];
      }
  }
  uint4 pb[NCH][4];  // second residual (three-term Moore-Penrose backward launches), likewise
  if (r2) {
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      if (cbase + c * 32 < e.N) {
#pragma unroll
        for (int j = 0; j < 4; ++j) pb[c][j] = reinterpret_cast<const uint4*>(r2 + cbase + c * 32)[j];
      }
  }
  mbar_wait(tfull_bar, aphase);
  tc_fence_after();
  if (row0 >= e.M) return;
  const float alpha = e.alpha, gamma = e.gamma, gamma2 = e.gamma2, diag = e.diag;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col0 = cbase + c * 32;
    if (col0 >= e.N) break;  // warp-uniform
    uint32_t acc[32];
    tmem_ld_32x32(taddr + c * 32, acc);
    tmem_ld_wait();
    float v[32];
    if (r1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w[4] = {pa[c][j].x, pa[c][j].y, pa[c][j].z, pa[c][j].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {  // bf16 -> f32 is a shift / a mask
          v[j * 8 + 2 * t] = fmaf(alpha, __uint_as_float(acc[j * 8 + 2 * t]), gamma * __uint_as_float(w[t] << 16));
          v[j * 8 + 2 * t + 1] = fmaf(alpha, __uint_as_float(acc[j * 8 + 2 * t + 1]), gamma * __uint_as_float(w[t] & 0xffff0000u));
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = alpha * __uint_as_float(acc[j]);
    }
    if (r2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w[4] = {pb[c][j].x, pb[c][j].y, pb[c][j].z, pb[c][j].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          v[j * 8 + 2 * t] = fmaf(gamma2, __uint_as_float(w[t] << 16), v[j * 8 + 2 * t]);
          v[j * 8 + 2 * t + 1] = fmaf(gamma2, __uint_as_float(w[t] & 0xffff0000u), v[j * 8 + 2 * t + 1]);
        }
      }
    }
    if (diag != 0.f && row >= col0 && row < col0 + 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += (col0 + j == row) ? diag : 0.f;
    }
    uint4 pc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pc[j]);
#pragma unroll
      for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[j * 8 + 2 * t], v[j * 8 + 2 * t + 1]);
    }
    sink_tma_bf16(sink, pc, false, b1, b2, row0, col0, lane);
  }
}

// SM: the kernel instantiation serves the fused row-softmax modes ONLY (and the others never): each kernel carries one of the
// two epilogues -- with both inlined the epilogue warps lost a quarter of their issue slots to instruction-cache misses.
// EK: which ONE epilogue the instantiation carries: 0 generic, 1 the fused row-softmax modes, 2 lean (above)
template <int BN, int EK>
__device__ __forceinline__ void epilogue_tile(const Epi& e, bool fast, uint32_t taddr, Sink& sink, int b1, int b2, int row0,
                                              int cbase, int lane, uint64_t* tfull_bar, uint32_t aphase, int part) {
  constexpr int NCH = BN / 64;  // 32-column chunks per thread
  if constexpr (EK == 1) {
    epilogue_tile_softmax<BN>(e, sink, taddr, b1, b2, row0, cbase, lane, tfull_bar, aphase, part);
    return;
  }
  if constexpr (EK >= 11) {  // one softmax mode, fixed at compile time
    epilogue_tile_softmax<BN, EK - 10>(e, sink, taddr, b1, b2, row0, cbase, lane, tfull_bar, aphase, part);
    return;
  }
  if constexpr (EK == 2) {
    epilogue_tile_lean<BN>(e, taddr, sink, b1, b2, row0, cbase, lane, tfull_bar, aphase);
    return;
  }
  const bool rows_ok = row0 < e.M;  // warp-uniform
  ResBuf rb0, rb1;
  if (fast && rows_ok) {  // start the residual loads of the first two chunks while the MMAs of this tile still run
    if (cbase + 32 <= e.N) res_prefetch(e, rb0, b1, b2, row0, cbase, lane);
    if (NCH > 1 && cbase + 64 <= e.N) res_prefetch(e, rb1, b1, b2, row0, cbase + 32, lane);
  }
  mbar_wait(tfull_bar, aphase);
  tc_fence_after();
  if (!rows_ok) return;
  // chunks go in pairs (static register names for the two prefetch buffers).  The pair loop is NOT unrolled: the vector
  // chunk body is ~1000 instructions and four copies of it (BN = 256) overflow the instruction cache of the SM -- the
  // epilogue warps were losing ~a quarter of their issue slots to instruction fetch (ncu: stall_no_inst).
#pragma unroll 1
  for (int c = 0; c < NCH; c += 2) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int col0 = cbase + (c + u) * 32;
      if (c + u < NCH && col0 < e.N) {  // warp-uniform
        if (fast && col0 + 32 <= e.N) epilogue_chunk_vec(e, taddr + (c + u) * 32, u ? rb1 : rb0, sink, b1, b2, row0, col0, lane);
        else epilogue_chunk_scalar(e, taddr + (c + u) * 32, b1, b2, row0 + lane, col0);
        if (c + u + 2 < NCH && fast && col0 + 96 <= e.N) res_prefetch(e, u ? rb1 : rb0, b1, b2, row0, col0 + 64, lane);
      }
    }
  }
}

template <int BN, int A_MN, int B_MN, bool TMAS, int EK>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC32, const __grid_constant__ CUtensorMap tmC16,
                    const __grid_constant__ KParams p, const int vec_ok) {
  using C = Cfg<BN, TMAS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * C::A_BYTES;
  const uint32_t epi_stage = smem_u32(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::EPI_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + kMaxAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kMaxAccStages);

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
      for (int s = 0; s < C::ACC_STAGES; ++s) {
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
  pdl_wait();               // everything above overlapped the previous grid's tail; global memory is touched only below
  pdl_launch_dependents();  // the next grid may be scheduled as our CTAs retire (it waits for our completion itself)

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
        const int as = it % C::ACC_STAGES;
        const uint32_t aphase = (it / C::ACC_STAGES) & 1;
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
    const bool fast = vec_ok != 0 && !p.e.atomic && p.e.act != MIRROR_ACT_GELU;
    Sink sink{&tmC32, &tmC16, epi_stage + (warp - 2) * kSinkBytes, TMAS ? 1 : 0, 0};
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
      const int row0 = mb_ * BM + q * 32;
      const int as = it % C::ACC_STAGES;
      const uint32_t aphase = (it / C::ACC_STAGES) & 1;
      epilogue_tile<BN, EK>(p.e, fast, tmem_base + (uint32_t(q * 32) << 16) + as * BN + half * (BN / 2), sink, b1, b2, row0, cbase, lane,
                        &tfull[as], aphase, nb * 2 + half);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      ++it;
    }
    if (TMAS && lane == 0) bulk_wait_read<0>();  // the staging boxes must outlive their stores
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Cluster variant: CL CTAs of a thread-block cluster own CL consecutive 128-row tiles of the SAME N range.  Each CTA
// loads its own A rows and 1/CL of the B tile, which TMA *multicasts* into the shared memory of all CL CTAs, so a
// B byte is fetched from L2 once per cluster instead of once per CTA (the plain kernel is L2-bandwidth bound: a
// 128x256x64 step moves 48 KB for 4.2 MFLOP).  CL = 3 covers the 384-row Moore-Penrose matrices with one cluster.
// MMAs, TMEM and epilogue stay per CTA (cta_group::1).  Protocol differences to the plain kernel:
//   full[s]   still 1 local arrival (expect_tx of the whole stage); the bytes now arrive from CL different TMA issuers
//   empty[s]  CL arrivals: every CTA's tcgen05.commit is multicast to all CTAs, because a refill of stage s by ANY
//             producer overwrites the B slice in EVERY CTA
//   cluster barrier at start (barriers initialised before remote traffic) and at exit (no CTA leaves while peers may
//   still multicast into it)
template <int BN, int A_MN, int B_MN, int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tcgen05_cluster_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const __grid_constant__ KParams p, const int vec_ok) {
  using C = Cfg<BN>;
  static_assert((BN / CL) % 64 == 0 || !B_MN, "MN-major B is sliced in 64-column chunks");
  static_assert(BN % CL == 0, "B tile must split evenly over the cluster");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * C::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::EPI_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + kMaxAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kMaxAccStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cid = blockIdx.x / CL, nclusters = gridDim.x / CL;
  constexpr uint16_t kMask = (1u << CL) - 1;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < C::STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], CL);
      }
      for (int s = 0; s < kAccStages; ++s) {
        mbar_init(&tfull[s], 1);
        mbar_init(&tempty[s], kEpiWarps);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, C::TMEM_COLS);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item = (batch, group of CL consecutive m tiles, n tile [, k split]); p.tiles_m counts groups here
  const int tiles = p.tiles_m * p.tiles_n * p.batch1 * p.batch2;
  const int total = tiles * p.split_k;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_per = (kb_total + p.split_k - 1) / p.split_k;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cid; w < total; w += nclusters) {
        const int tile = w / p.split_k, ks = w - tile * p.split_k;
        const int kb0 = ks * kb_per, kb1 = min(kb_total, kb0 + kb_per);
        if (kb0 >= kb1) continue;
        const int nb = tile % p.tiles_n;
        int t = tile / p.tiles_n;
        const int mg = t % p.tiles_m;
        t /= p.tiles_m;
        const int b1 = t % p.batch1, b2 = t / p.batch1;
        const int m0 = (mg * CL + rank) * BM, n0 = nb * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);  // all CL consumers released this slot
          mbar_expect_tx(&full[stage], C::STAGE_BYTES);
          uint8_t* a = sA + stage * C::A_BYTES;
          uint8_t* b = sB + stage * C::B_BYTES;
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_4d(&tmA, &full[stage], a + i * (BK * 128), m0 + i * 64, kb * BK, b1 * p.a_b1, b2 * p.a_b2);
          } else {
            tma_load_4d(&tmA, &full[stage], a, kb * BK, m0, b1 * p.a_b1, b2 * p.a_b2);
          }
          // this CTA's 1/CL slice of the B tile, delivered to every CTA of the cluster
          if (B_MN) {
            constexpr int kChunks = BN / 64 / CL;
#pragma unroll
            for (int i = 0; i < kChunks; ++i) {
              const int ch = rank * kChunks + i;
              tma_load_4d_mc(&tmB, &full[stage], b + ch * (BK * 128), n0 + ch * 64, kb * BK, b1 * p.b_b1, b2 * p.b_b2, kMask);
            }
          } else {
            constexpr int kRows = BN / CL;
            tma_load_4d_mc(&tmB, &full[stage], b + rank * (kRows * 128), kb * BK, n0 + rank * kRows, b1 * p.b_b1, b2 * p.b_b2, kMask);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      constexpr uint32_t a_lbo = A_MN ? BK * 128 : 0, b_lbo = B_MN ? BK * 128 : 0;
      constexpr uint32_t a_kstep = A_MN ? (UMMA_K / 8) * 1024 : UMMA_K * 2;
      constexpr uint32_t b_kstep = B_MN ? (UMMA_K / 8) * 1024 : UMMA_K * 2;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = cid; w < total; w += nclusters) {
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
          umma_commit_mc(&empty[stage], kMask);  // this CTA is done with the slot: tell every producer of the cluster
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[as]);
        ++it;
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const bool fast = vec_ok != 0 && !p.e.atomic && p.e.act != MIRROR_ACT_GELU;
    Sink sink{nullptr, nullptr, 0u, 0, 0};
    int it = 0;
    for (int w = cid; w < total; w += nclusters) {
      const int tile = w / p.split_k, ks = w - tile * p.split_k;
      const int kb0 = ks * kb_per, kb1 = min(kb_total, kb0 + kb_per);
      if (kb0 >= kb1) continue;
      const int nb = tile % p.tiles_n;
      int t = tile / p.tiles_n;
      const int mg = t % p.tiles_m;
      t /= p.tiles_m;
      const int b1 = t % p.batch1, b2 = t / p.batch1;
      const int cbase = nb * BN + half * (BN / 2);
      const int row0 = (mg * CL + rank) * BM + q * 32;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      epilogue_tile<BN, 0>(p.e, fast, tmem_base + (uint32_t(q * 32) << 16) + as * BN + half * (BN / 2), sink, b1, b2, row0, cbase, lane,
                        &tfull[as], aphase, nb * 2 + half);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      ++it;
    }
  }

  tc_fence_before();
  __syncwarp();
  cluster_sync_all();  // peers may still multicast data / barrier arrivals into this CTA until they are all done
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}


// ------------------------------------------------------------------------------------------------------------------
// Multi-term variant: D = epilogue( sum_t A_t * B_t^T ) with up to kMaxTerms operand pairs that share M, N and the
// batch dims but may differ in K and in their layouts (majors are runtime values here).  All terms accumulate into the
// same TMEM tile, so a sum of matrix products costs ONE epilogue pass instead of one fp32 read-modify-write pass per
// product -- the Moore-Penrose backward (g_E = g_F G1^T + g_G1 E^T + E^T g_G1, g_z = g F^T + a2^T g_E,
// g_a2 = sum over the 6 iterations of g_E z^T) is built from such sums of 384^3 products.
constexpr int kMaxTerms = 6;
struct MultiMaps {
  CUtensorMap a[kMaxTerms];
  CUtensorMap b[kMaxTerms];
  CUtensorMap c32, c16;  // outputs (bulk-store epilogue)
};
struct MultiInfo {
  int nterms;
  int kb[kMaxTerms];    // k-blocks of term t
  int a_mn[kMaxTerms];  // operand majors of term t
  int b_mn[kMaxTerms];
};

template <int BN, bool TMAS, int EK>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_multi_kernel(const __grid_constant__ MultiMaps maps, const __grid_constant__ KParams p,
                          const __grid_constant__ MultiInfo mi, const int vec_ok) {
  using C = Cfg<BN, TMAS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * C::A_BYTES;
  const uint32_t epi_stage = smem_u32(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::EPI_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + kMaxAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kMaxAccStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && elect_one()) {
    for (int t = 0; t < mi.nterms; ++t) {
      tma_prefetch_desc(&maps.a[t]);
      tma_prefetch_desc(&maps.b[t]);
    }
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < C::STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int s = 0; s < C::ACC_STAGES; ++s) {
        mbar_init(&tfull[s], 1);
        mbar_init(&tempty[s], kEpiWarps);
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
  pdl_wait();  // see gemm_tcgen05_kernel
  pdl_launch_dependents();
  const int total = p.tiles_m * p.tiles_n * p.batch1 * p.batch2;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int nb = w % p.tiles_n;
        int t = w / p.tiles_n;
        const int mb_ = t % p.tiles_m;
        t /= p.tiles_m;
        const int b1 = t % p.batch1, b2 = t / p.batch1;
        const int m0 = mb_ * BM, n0 = nb * BN;
        for (int term = 0; term < mi.nterms; ++term) {
          const CUtensorMap* ma = &maps.a[term];
          const CUtensorMap* mbp = &maps.b[term];
          const int a_mn = mi.a_mn[term], b_mn = mi.b_mn[term];
          for (int kb = 0; kb < mi.kb[term]; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], C::STAGE_BYTES);
            uint8_t* a = sA + stage * C::A_BYTES;
            uint8_t* b = sB + stage * C::B_BYTES;
            if (a_mn) {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i) tma_load_4d(ma, &full[stage], a + i * (BK * 128), m0 + i * 64, kb * BK, b1, b2);
            } else {
              tma_load_4d(ma, &full[stage], a, kb * BK, m0, b1, b2);
            }
            if (b_mn) {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i) tma_load_4d(mbp, &full[stage], b + i * (BK * 128), n0 + i * 64, kb * BK, b1, b2);
            } else {
              tma_load_4d(mbp, &full[stage], b, kb * BK, n0, b1, b2);
            }
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int as = it % C::ACC_STAGES;
        const uint32_t aphase = (it / C::ACC_STAGES) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        uint32_t accumulate = 0;
        for (int term = 0; term < mi.nterms; ++term) {
          const int a_mn = mi.a_mn[term], b_mn = mi.b_mn[term];
          const uint32_t idesc = make_idesc_bf16(BM, BN, a_mn, b_mn);
          const uint32_t a_lbo = a_mn ? BK * 128 : 0, b_lbo = b_mn ? BK * 128 : 0;
          const uint32_t a_kstep = a_mn ? (UMMA_K / 8) * 1024 : UMMA_K * 2;
          const uint32_t b_kstep = b_mn ? (UMMA_K / 8) * 1024 : UMMA_K * 2;
          for (int kb = 0; kb < mi.kb[term]; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sA + stage * C::A_BYTES);
            const uint32_t b_addr = smem_u32(sB + stage * C::B_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t ad = make_smem_desc_sw128(a_addr + k * a_kstep, a_lbo, 1024);
              const uint64_t bd = make_smem_desc_sw128(b_addr + k * b_kstep, b_lbo, 1024);
              umma_f16(tmem_d, ad, bd, idesc, accumulate);
              accumulate = 1;
            }
            umma_commit(&empty[stage]);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(&tfull[as]);
        ++it;
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const bool fast = vec_ok != 0 && p.e.act != MIRROR_ACT_GELU;
    Sink sink{&maps.c32, &maps.c16, epi_stage + (warp - 2) * kSinkBytes, TMAS ? 1 : 0, 0};
    int it = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int nb = w % p.tiles_n;
      int t = w / p.tiles_n;
      const int mb_ = t % p.tiles_m;
      t /= p.tiles_m;
      const int b1 = t % p.batch1, b2 = t / p.batch1;
      const int cbase = nb * BN + half * (BN / 2);
      const int row0 = mb_ * BM + q * 32;
      const int as = it % C::ACC_STAGES;
      const uint32_t aphase = (it / C::ACC_STAGES) & 1;
      epilogue_tile<BN, EK>(p.e, fast, tmem_base + (uint32_t(q * 32) << 16) + as * BN + half * (BN / 2), sink, b1, b2, row0, cbase, lane,
                        &tfull[as], aphase, nb * 2 + half);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      ++it;
    }
    if (TMAS && lane == 0) bulk_wait_read<0>();
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

template <int BN, int A_MN, int B_MN, bool TMAS, int EK = 0>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC32, const CUtensorMap& tmC16, const KParams& p,
           int vec_ok, cudaStream_t stream) {
  using C = Cfg<BN, TMAS>;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, TMAS, EK>;
  static DeviceOnce once;
  if (once.first()) {
    MB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  const long long total = (long long)p.tiles_m * p.tiles_n * p.batch1 * p.batch2 * p.split_k;
  const int grid = (int)(total < num_sms() ? total : num_sms());
  MB_CUDA(launch_pdl(kern, dim3(grid), dim3(kThreads), C::SMEM_BYTES, stream, tmA, tmB, tmC32, tmC16, p, vec_ok));
  return 0;
}

// Output tensor map of the TMA-store epilogue: [batch2, batch1, M, N] with element strides (bs2, bs1, ld, 1), 32 x 32 boxes.
int make_output_map(CUtensorMap* map, void* ptr, bool f32, long long M, long long N, long long ld, long long bs1, int batch1,
                    long long bs2, int batch2) {
  auto enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return MB_ERR_DRIVER;
  }
  const cuuint64_t es = f32 ? 4 : 2;
  cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)batch1, (cuuint64_t)batch2};
  const cuuint64_t fallback = (cuuint64_t)ld * es;
  cuuint64_t strides[3] = {(cuuint64_t)ld * es, batch1 > 1 ? (cuuint64_t)bs1 * es : fallback, batch2 > 1 ? (cuuint64_t)bs2 * es : fallback};
  cuuint32_t box[4] = {32, 32, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (output) failed (%d): M=%lld N=%lld ld=%lld", (int)r, M, N, ld);
    return MB_ERR_DRIVER;
  }
  return 0;
}

// the bulk-store epilogue needs 16-byte aligned bases and strides (a TMA rule); everything else keeps register stores
bool tma_store_ok(const mirror_gemm_args* g) {
  auto ok = [&](const void* ptr, long long es, long long ld, long long bs1, long long bs2) {
    return !ptr || (((reinterpret_cast<uintptr_t>(ptr) & 15) == 0) && (ld * es) % 16 == 0 &&
                    (g->batch1 <= 1 || (bs1 > 0 && (bs1 * es) % 16 == 0)) && (g->batch2 <= 1 || (bs2 > 0 && (bs2 * es) % 16 == 0)));
  };
  return ok(g->out_f32, 4, g->ldc32, g->c32_bs1, g->c32_bs2) && ok(g->out_bf16, 2, g->ldc16, g->c16_bs1, g->c16_bs2);
}

// Epilogue-bound products (short K: the tile's stores, not its MMAs, set the pace) hand their output to TMA.
// MIRROR_B200_TMA_STORE=0 keeps the register stores everywhere (A/B switch).
bool use_tma_store(const mirror_gemm_args* g, const Epi* e, int vec, long long ktot) {
  static const int env = [] { const char* v = getenv("MIRROR_B200_TMA_STORE"); return v && *v ? atoi(v) : -1; }();
  if (env == 0) return false;
  static const long long kmax = [] { const char* v = getenv("MIRROR_B200_AB_TMAS_KMAX"); return v && *v ? atoll(v) : 3072LL; }();
  return (g->out_f32 || g->out_bf16) && vec && !e->atomic && e->act != MIRROR_ACT_GELU && g->split_k <= 1 && ktot <= kmax &&
         g->N % 32 == 0 && tma_store_ok(g);
}

// the lean epilogue covers: bf16 output only, alpha, diag, bf16 residual(s) read row by row -- and nothing else
// (MIRROR_B200_AB_NO_LEAN=1 keeps the generic epilogue: A/B switch)
bool lean_epilogue_ok(const mirror_gemm_args* g) {
  static const int off = [] { const char* v = getenv("MIRROR_B200_AB_NO_LEAN"); return v && *v == '1'; }();
  return !off && g->out_bf16 && !g->out_f32 && !g->bias && g->act == MIRROR_ACT_NONE && g->drop_p == 0.f && g->beta == 0.f &&
         (!g->res || g->res_is_bf16) && g->mode == MIRROR_GEMM_NORMAL && g->split_k <= 1 && g->N % 32 == 0;
}

int make_output_maps(const mirror_gemm_args* g, CUtensorMap* c32, CUtensorMap* c16) {
  int rc = 0;
  if (g->out_f32) rc = make_output_map(c32, g->out_f32, true, g->M, g->N, g->ldc32, g->c32_bs1, g->batch1, g->c32_bs2, g->batch2);
  if (!rc && g->out_bf16) rc = make_output_map(c16, g->out_bf16, false, g->M, g->N, g->ldc16, g->c16_bs1, g->batch1, g->c16_bs2, g->batch2);
  return rc;
}

template <int BN, int A_MN, int B_MN, int CL>
int launch_cluster(const CUtensorMap& tmA, const CUtensorMap& tmB, const KParams& p, int vec_ok, cudaStream_t stream) {
  using C = Cfg<BN>;
  auto kern = gemm_tcgen05_cluster_kernel<BN, A_MN, B_MN, CL>;
  static DeviceOnce once;
  if (once.first()) {
    MB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  const long long total = (long long)p.tiles_m * p.tiles_n * p.batch1 * p.batch2 * p.split_k;
  const int max_clusters = num_sms() / CL;
  const int clusters = (int)(total < max_clusters ? total : max_clusters);
  kern<<<CL * clusters, kThreads, C::SMEM_BYTES, stream>>>(tmA, tmB, p, vec_ok);  // cluster dims (CL,1,1) are compiled in
  MB_LAUNCH_CHECK();
  return 0;
}


template <int BN, bool TMAS, int EK = 0>
int launch_multi(const MultiMaps& maps, const KParams& p, const MultiInfo& mi, int vec_ok, cudaStream_t stream) {
  using C = Cfg<BN, TMAS>;
  auto kern = gemm_tcgen05_multi_kernel<BN, TMAS, EK>;
  static DeviceOnce once;
  if (once.first()) {
    MB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  const long long total = (long long)p.tiles_m * p.tiles_n * p.batch1 * p.batch2;
  const int grid = (int)(total < num_sms() ? total : num_sms());
  MB_CUDA(launch_pdl(kern, dim3(grid), dim3(kThreads), C::SMEM_BYTES, stream, maps, p, mi, vec_ok));
  return 0;
}

int fill_epi(const mirror_gemm_args* g, Epi* e) {
  MB_CHECK_ARG(g && g->a && g->b, "gemm: null operand");
  MB_CHECK_ARG(g->M > 0 && g->N > 0 && g->K > 0 && g->batch1 > 0 && g->batch2 > 0, "gemm: bad shape M=%d N=%d K=%d", g->M,
               g->N, g->K);
  MB_CHECK_ARG(g->out_f32 || g->out_bf16 || g->mode == MIRROR_GEMM_ROWSTATS || g->mode == MIRROR_GEMM_ROWDOT || g->mode == 6, "gemm: no output");
  MB_CHECK_ARG(g->beta == 0.f || g->out_f32, "gemm: beta needs out_f32");
  MB_CHECK_ARG(g->drop_p >= 0.f && g->drop_p < 1.f, "gemm: drop_p out of range");
  MB_CHECK_ARG(g->split_k <= 1 || (g->out_f32 && !g->out_bf16 && !g->bias && !g->res && !g->res2 && g->act == 0 && g->drop_p == 0.f && g->diag == 0.f),
               "gemm: split_k supports only alpha and an fp32 accumulate target");
  e->M = g->M; e->N = g->N; e->batch1 = g->batch1;
  e->alpha = g->alpha; e->diag = g->diag; e->bias = g->bias; e->act = g->act;
  e->drop_p = g->drop_p; e->drop_scale = 1.f / (1.f - g->drop_p); e->drop_seed = g->drop_seed; e->drop_epoch = drop_epoch_ptr();
  e->res = g->res; e->res_is_bf16 = g->res_is_bf16; e->gamma = g->gamma;
  e->res2 = reinterpret_cast<const bf16*>(g->res2); e->gamma2 = g->gamma2;
  e->res_row_div = g->res_row_div > 1 ? g->res_row_div : 1;
  e->mode = g->mode; e->stats = g->stats; e->nparts = 0;
  if (g->mode != MIRROR_GEMM_NORMAL) {
    MB_CHECK_ARG(g->mode >= 1 && g->mode <= 6 && g->stats && g->N % 32 == 0 && g->split_k <= 1 && g->res_row_div <= 1,
                 "gemm: softmax modes need stats, N %% 32 == 0 and no split-K");
    MB_CHECK_ARG(g->mode < MIRROR_GEMM_ROWDOT || g->mode == 6 || (g->res && g->res_is_bf16),
                 "gemm: ROWDOT / SOFTMAX_BWD read the bf16 probabilities through `res`");
    MB_CHECK_ARG(g->mode == MIRROR_GEMM_ROWSTATS || g->mode == MIRROR_GEMM_ROWDOT || g->mode == 6 || g->out_f32 || g->out_bf16, "gemm: no output");
  }
  MB_CHECK_ARG(!g->res2 || g->res, "gemm: res2 needs res (it shares its strides)");
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
    if (g->res2) ok = ok && al(g->res2, 16) && g->ldr % 8 == 0 && g->r_bs1 % 8 == 0 && g->r_bs2 % 8 == 0;
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

static int tile_n_for(int N) {
  static const int no192 = [] { const char* v = getenv("MIRROR_B200_AB_NO_BN192"); return v && *v == '1'; }();  // A/B switch
  return (N % 256 == 0 || N >= 1024) ? 256 : ((N % 192 == 0 && !no192) ? 192 : 128);
}

extern "C" int mirror_gemm_nparts(int32_t N) {
  const int bn = tile_n_for(N);
  return 2 * ((N + bn - 1) / bn);
}

extern "C" int mirror_gemm_bf16(const mirror_gemm_args* g, mirror_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KParams p;
  int rc = fill_epi(g, &p.e);
  if (rc) return rc;
  // Clusters with TMA-multicast B whenever a batch item has several 128-row tiles: CL = 3 for the 384-row
  // Moore-Penrose / landmark matrices (one cluster per matrix and N tile), CL = 2 otherwise.
  // Opt-in (MIRROR_B200_CLUSTER=1): measured on B200 it does not pay -- with clusters of 2..3 CTAs the multicast did not
  // reduce L2 traffic (B300_MICROARCH: "MC ~ UC at cluster size <= 4"), CL=2 gave +1..4 %, CL=3 halved the throughput of
  // the 384^3 GEMMs (lock-step of three CTAs); the limiter of these kernels is the L1->L2 write path of the epilogue.
  static const bool cluster_enabled = [] {
    const char* v = getenv("MIRROR_B200_CLUSTER");
    return v && v[0] && v[0] != '0';
  }();
  const int mt = (g->M + BM - 1) / BM;
  if (mt >= 2 && cluster_enabled && g->mode == MIRROR_GEMM_NORMAL) {
    const int CL = (mt == 3) ? 3 : 2;
    const int BNc = (CL == 3) ? ((g->N % 192 == 0 || g->N > 128) ? 192 : 0)
                              : ((g->N % 256 == 0 || g->N >= 1024) ? 256 : 128);
    if (BNc) {
      p.K = g->K;
      p.batch1 = g->batch1;
      p.batch2 = g->batch2;
      p.tiles_m = (mt + CL - 1) / CL;  // groups of CL row tiles
      p.tiles_n = (g->N + BNc - 1) / BNc;
      p.split_k = g->split_k > 1 ? g->split_k : 1;
      p.a_b1 = (g->batch1 > 1 && g->a_bs1 == 0) ? 0 : 1;
      p.a_b2 = (g->batch2 > 1 && g->a_bs2 == 0) ? 0 : 1;
      p.b_b1 = (g->batch1 > 1 && g->b_bs1 == 0) ? 0 : 1;
      p.b_b2 = (g->batch2 > 1 && g->b_bs2 == 0) ? 0 : 1;
      CUtensorMap tmA, tmB;
      rc = make_operand_map(&tmA, g->a, g->a_mn_major, g->M, g->K, g->lda, g->a_bs1, p.a_b1 ? g->batch1 : 1, g->a_bs2,
                            p.a_b2 ? g->batch2 : 1, BM);
      if (rc) return rc;
      rc = make_operand_map(&tmB, g->b, g->b_mn_major, g->N, g->K, g->ldb, g->b_bs1, p.b_b1 ? g->batch1 : 1, g->b_bs2,
                            p.b_b2 ? g->batch2 : 1, BNc / CL);
      if (rc) return rc;
      const int vec = epi_vec_ok(g) ? 1 : 0;
      const int key = (g->a_mn_major ? 2 : 0) | (g->b_mn_major ? 1 : 0);
#define MB_DISPATCHC(BNV, CLV)                                                 \
  switch (key) {                                                               \
    case 0: return launch_cluster<BNV, 0, 0, CLV>(tmA, tmB, p, vec, stream);   \
    case 1: return launch_cluster<BNV, 0, 1, CLV>(tmA, tmB, p, vec, stream);   \
    case 2: return launch_cluster<BNV, 1, 0, CLV>(tmA, tmB, p, vec, stream);   \
    default: return launch_cluster<BNV, 1, 1, CLV>(tmA, tmB, p, vec, stream);  \
  }
      if (CL == 3) { MB_DISPATCHC(192, 3) }
      if (BNc == 256) { MB_DISPATCHC(256, 2) }
      MB_DISPATCHC(128, 2)
#undef MB_DISPATCHC
    }
  }
  // N tile: 256 when it divides N (or N is large), 192 for N = 384-like sizes (the Nystrom landmark matrices), else 128.
  int BN = tile_n_for(g->N);
  {  // a batch-sized product (M = 64 rows) has only N / BN tiles and streams its weight through that many SMs: narrower tiles
     // double the SMs that pull on the weight (MIRROR_B200_AB_NO_NARROW=1: A/B switch)
    static const int off = [] { const char* v = getenv("MIRROR_B200_AB_NO_NARROW"); return v && *v == '1'; }();
    const long long tiles256 = (long long)((g->M + BM - 1) / BM) * ((g->N + 255) / 256) * g->batch1 * g->batch2 *
                               (g->split_k > 1 ? g->split_k : 1);
    if (!off && BN == 256 && g->mode == MIRROR_GEMM_NORMAL && g->N % 128 == 0 && tiles256 * 2 <= num_sms()) BN = 128;
  }
  p.K = g->K;
  p.batch1 = g->batch1;
  p.batch2 = g->batch2;
  p.tiles_m = (g->M + BM - 1) / BM;
  p.tiles_n = (g->N + BN - 1) / BN;
  p.e.nparts = 2 * p.tiles_n;
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
  const int key = (g->a_mn_major ? 2 : 0) | (g->b_mn_major ? 1 : 0);
  const bool tmas = use_tma_store(g, &p.e, vec, g->K);
  CUtensorMap tmC32 = tmA, tmC16 = tmA;  // placeholders when unused
  if (tmas) {
    rc = make_output_maps(g, &tmC32, &tmC16);
    if (rc) return rc;
  }
  if (g->mode != MIRROR_GEMM_NORMAL) {  // the fused softmax epilogues exist for K-major operands (q k^T-shaped products) only
    MB_CHECK_ARG(key == 0, "gemm: softmax modes need K-major A and B");
    // one instantiation per mode (the mode tests fold away); writers go through the bulk-store sink, the two statistics
    // passes have no output.  Anything else (unaligned outputs) falls back to the instantiation that reads e.mode.
#define MB_SOFTMAX(BNV)                                                                                            \
  switch (g->mode) {                                                                                                \
    case MIRROR_GEMM_ROWSTATS: return launch<BNV, 0, 0, false, 11>(tmA, tmB, tmC32, tmC16, p, vec, stream);          \
    case MIRROR_GEMM_ROWDOT: return launch<BNV, 0, 0, false, 13>(tmA, tmB, tmC32, tmC16, p, vec, stream);            \
    case MIRROR_GEMM_SOFTMAX: if (tmas) return launch<BNV, 0, 0, true, 12>(tmA, tmB, tmC32, tmC16, p, vec, stream); break;          \
    case MIRROR_GEMM_SOFTMAX_BWD: if (tmas) return launch<BNV, 0, 0, true, 14>(tmA, tmB, tmC32, tmC16, p, vec, stream); break;      \
    case MIRROR_GEMM_SOFTMAX_BWD_DOT: if (tmas) return launch<BNV, 0, 0, true, 15>(tmA, tmB, tmC32, tmC16, p, vec, stream); break;  \
    default: break;                                                                                                 \
  }                                                                                                                 \
  return tmas ? launch<BNV, 0, 0, true, 1>(tmA, tmB, tmC32, tmC16, p, vec, stream)                                   \
              : launch<BNV, 0, 0, false, 1>(tmA, tmB, tmC32, tmC16, p, vec, stream);
    if (BN == 256) { MB_SOFTMAX(256) }
    if (BN == 192) { MB_SOFTMAX(192) }
    MB_SOFTMAX(128)
#undef MB_SOFTMAX
  }
  if (tmas && lean_epilogue_ok(g)) {
#define MB_DISPATCHL(BNV)                                                           \
  switch (key) {                                                                    \
    case 0: return launch<BNV, 0, 0, true, 2>(tmA, tmB, tmC32, tmC16, p, vec, stream);   \
    case 1: return launch<BNV, 0, 1, true, 2>(tmA, tmB, tmC32, tmC16, p, vec, stream);   \
    case 2: return launch<BNV, 1, 0, true, 2>(tmA, tmB, tmC32, tmC16, p, vec, stream);   \
    default: return launch<BNV, 1, 1, true, 2>(tmA, tmB, tmC32, tmC16, p, vec, stream);  \
  }
    if (BN == 256) { MB_DISPATCHL(256) }
    if (BN == 192) { MB_DISPATCHL(192) }
    MB_DISPATCHL(128)
#undef MB_DISPATCHL
  }
#define MB_DISPATCH(BNV, TM)                                                        \
  switch (key) {                                                                    \
    case 0: return launch<BNV, 0, 0, TM>(tmA, tmB, tmC32, tmC16, p, vec, stream);   \
    case 1: return launch<BNV, 0, 1, TM>(tmA, tmB, tmC32, tmC16, p, vec, stream);   \
    case 2: return launch<BNV, 1, 0, TM>(tmA, tmB, tmC32, tmC16, p, vec, stream);   \
    default: return launch<BNV, 1, 1, TM>(tmA, tmB, tmC32, tmC16, p, vec, stream);  \
  }
  if (BN == 256) { if (tmas) { MB_DISPATCH(256, true) } MB_DISPATCH(256, false) }
  if (BN == 192) { if (tmas) { MB_DISPATCH(192, true) } MB_DISPATCH(192, false) }
  if (tmas) { MB_DISPATCH(128, true) }
  MB_DISPATCH(128, false)
#undef MB_DISPATCH
}


extern "C" int mirror_gemm_bf16_multi(const mirror_gemm_args* terms, int32_t nterms, mirror_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(terms && nterms >= 1 && nterms <= kMaxTerms, "gemm_multi: 1..%d terms", kMaxTerms);
  const mirror_gemm_args* g = &terms[0];
  KParams p;
  int rc = fill_epi(g, &p.e);
  if (rc) return rc;
  MB_CHECK_ARG(g->split_k <= 1 && g->mode == MIRROR_GEMM_NORMAL, "gemm_multi: split_k / softmax modes are not supported");
  const int BN = tile_n_for(g->N);
  p.K = g->K;
  p.batch1 = g->batch1;
  p.batch2 = g->batch2;
  p.tiles_m = (g->M + BM - 1) / BM;
  p.tiles_n = (g->N + BN - 1) / BN;
  p.split_k = 1;
  p.a_b1 = p.a_b2 = p.b_b1 = p.b_b2 = 1;
  MultiMaps maps;
  MultiInfo mi;
  mi.nterms = nterms;
  for (int t = 0; t < nterms; ++t) {
    const mirror_gemm_args* x = &terms[t];
    MB_CHECK_ARG(x->a && x->b && x->M == g->M && x->N == g->N && x->batch1 == g->batch1 && x->batch2 == g->batch2 && x->K > 0,
                 "gemm_multi: term %d does not match M/N/batch of term 0", t);
    MB_CHECK_ARG((g->batch1 == 1 || (x->a_bs1 && x->b_bs1)) && (g->batch2 == 1 || (x->a_bs2 && x->b_bs2)),
                 "gemm_multi: broadcast operands are not supported");
    mi.kb[t] = (x->K + BK - 1) / BK;
    mi.a_mn[t] = x->a_mn_major ? 1 : 0;
    mi.b_mn[t] = x->b_mn_major ? 1 : 0;
    rc = make_operand_map(&maps.a[t], x->a, x->a_mn_major, x->M, x->K, x->lda, x->a_bs1, x->batch1, x->a_bs2, x->batch2, BM);
    if (rc) return rc;
    rc = make_operand_map(&maps.b[t], x->b, x->b_mn_major, x->N, x->K, x->ldb, x->b_bs1, x->batch1, x->b_bs2, x->batch2, BN);
    if (rc) return rc;
  }
  for (int t = nterms; t < kMaxTerms; ++t) {
    mi.kb[t] = 0;
    mi.a_mn[t] = mi.b_mn[t] = 0;
    maps.a[t] = maps.a[0];
    maps.b[t] = maps.b[0];
  }
  const int vec = epi_vec_ok(g) ? 1 : 0;
  long long ktot = 0;
  for (int t = 0; t < nterms; ++t) ktot += terms[t].K;
  bool tmas = use_tma_store(g, &p.e, vec, ktot);
  maps.c32 = maps.c16 = maps.a[0];
  if (tmas) {
    rc = make_output_maps(g, &maps.c32, &maps.c16);
    if (rc) return rc;
  }
  if (tmas && lean_epilogue_ok(g)) {
    if (BN == 256) return launch_multi<256, true, 2>(maps, p, mi, vec, stream);
    if (BN == 192) return launch_multi<192, true, 2>(maps, p, mi, vec, stream);
    return launch_multi<128, true, 2>(maps, p, mi, vec, stream);
  }
  if (BN == 256) return tmas ? launch_multi<256, true>(maps, p, mi, vec, stream) : launch_multi<256, false>(maps, p, mi, vec, stream);
  if (BN == 192) return tmas ? launch_multi<192, true>(maps, p, mi, vec, stream) : launch_multi<192, false>(maps, p, mi, vec, stream);
  return tmas ? launch_multi<128, true>(maps, p, mi, vec, stream) : launch_multi<128, false>(maps, p, mi, vec, stream);
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
