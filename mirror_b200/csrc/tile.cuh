// Building blocks shared by the fused tcgen05 kernels that chain two tensor-core products through the CTA
// (contrastive.cu: S = X Y^T -> G -> dX = G Y;  flash_nystrom.cu: S = q k^T -> P -> O = P V): 2-D TMA loads,
// swizzled shared-memory tiles that threads write and tcgen05.mma reads, named barriers for the softmax warps.
#pragma once
#include <cudaTypedefs.h>

#include "common.cuh"
#include "ptx.cuh"

namespace mb {

// ---- PTX pieces not in ptx.cuh ---------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// barrier among a subset of the CTA's warps (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float fast_log2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// ---- TMEM as the A operand ("TS" form): registers -> TMEM, and an MMA whose A matrix [128 x K] lives in TMEM ----------
// thread i of the warp writes r[0..15] to columns [c, c+16) of TMEM lane (base_lane + i).  For 16-bit operands one 32-bit
// column holds two consecutive K elements, so 16 columns = 32 bf16 of A's row i.
__device__ __forceinline__ void tmem_st_32x32_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]; A is K-major by construction (a_major bit of idesc = 0); k-step kk = 8 columns
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// explicit global-space vector access (a pointer that travelled through a kernel-parameter struct compiles to generic LD / ST)
__device__ __forceinline__ uint4 ldg_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void stg_v4(void* p, const uint4& v) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

// ---- a [rows x 64] bf16 tile in the 128-byte-swizzled layout TMA writes and tcgen05.mma reads ------------------
// Row r occupies bytes [r*128, r*128+128); its 16-byte piece p lives at ((p ^ (r & 7)) << 4).  The same bytes serve as a
// K-major operand (rows = M/N, the 64 columns = K) and as an MN-major operand (rows = K, the 64 columns = M/N).
// Tiles must start on a 1024-byte boundary.
constexpr int kTileCols = 64;
__device__ __forceinline__ uint32_t tile_piece_addr(uint32_t tile_base, int row, int piece) {
  return tile_base + row * 128 + ((piece ^ (row & 7)) << 4);
}
// thread = row: store 32 consecutive columns (col0 = 0 or 32 within the 64-column tile) given as 4 pieces of 8 bf16
__device__ __forceinline__ void tile_store_32cols(uint32_t tile_base, int row, int half, const uint4 (&pc)[4]) {
#pragma unroll
  for (int p = 0; p < 4; ++p) st_shared_v4(tile_piece_addr(tile_base, row, half * 4 + p), pc[p]);
}

// K-major operand: tile = [MN rows x 64 k]; k-step kk (16 elements) advances 32 bytes inside the swizzle row
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_base, int kk) { return make_smem_desc_sw128(tile_base + kk * 32, 0, 1024); }
// MN-major operand made of 64-column chunks [K rows x 64 mn], `chunk_bytes` apart (= rows * 128); k-step kk = 16 rows
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t chunk0_base, uint32_t chunk_bytes, int kk) {
  return make_smem_desc_sw128(chunk0_base + kk * 2048, chunk_bytes, 1024);
}

// ---- host: 2-D bf16 tensor map, box = 64 columns x box_rows rows, 128B swizzle, OOB -> zeros -----------------------
inline PFN_cuTensorMapEncodeTiled_v12000 tile_get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  }
  return fn;
}
inline int make_map_2d(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  auto enc = tile_get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return MB_ERR_DRIVER;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 7)) {
    set_error("tile operand misaligned: ptr=%p ld=%lld (need 16 B / multiple of 8 elements)", ptr, ld);
    return MB_ERR_ALIGN;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d) failed (%d): rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
    return MB_ERR_DRIVER;
  }
  return 0;
}

// 4-D bf16 view [n2, n1, rows, cols] with element strides (s2, s1, ld, 1): box = 64 columns x box_rows rows of one (n2, n1)
// slice.  Coordinates of a load: (col, row, i1, i2).  Columns / rows beyond the extents read as zeros.
inline int make_map_4d(CUtensorMap* map, const void* ptr, long long cols, long long rows, long long ld, long long n1, long long s1,
                       long long n2, long long s2, int box_rows) {
  auto enc = tile_get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return MB_ERR_DRIVER;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 7) || (n1 > 1 && (s1 & 7)) || (n2 > 1 && (s2 & 7))) {
    set_error("tile operand misaligned: ptr=%p ld=%lld s1=%lld s2=%lld (need 16 B / multiples of 8 elements)", ptr, ld, s1, s2);
    return MB_ERR_ALIGN;
  }
  cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)n1, (cuuint64_t)n2};
  const cuuint64_t fb = (cuuint64_t)ld * 2;
  cuuint64_t strides[3] = {fb, n1 > 1 ? (cuuint64_t)s1 * 2 : fb, n2 > 1 ? (cuuint64_t)s2 * 2 : fb};
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d) failed (%d): cols=%lld rows=%lld ld=%lld n1=%lld s1=%lld n2=%lld s2=%lld", (int)r, cols, rows, ld,
              n1, s1, n2, s2);
    return MB_ERR_DRIVER;
  }
  return 0;
}

}  // namespace mb
