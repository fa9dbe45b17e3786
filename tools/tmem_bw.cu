// Micro-benchmark (measurement only): tcgen05.ld throughput per SM as a function of the number of reading warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tmem_bw tools/tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

template <int MODE>  // 0: ld + wait each; 1: 4 ld then wait; 2: ld + 32 ex2 (softmax-like); 3: st x16 + wait
__global__ void k(long long* out, int iters, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
  float acc = 0.f;
  uint32_t r[32];
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      tmem_ld_32x32(base + ((i * 32 + (warp >> 2) * 64) & 255), r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(r[i & 31]);
    } else if (MODE == 1) {
      uint32_t r2[32], r3[32], r4[32];
      tmem_ld_32x32(base, r); tmem_ld_32x32(base + 32, r2); tmem_ld_32x32(base + 64, r3); tmem_ld_32x32(base + 96, r4);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(r[i & 31]) + __uint_as_float(r2[i & 31]) + __uint_as_float(r3[i & 31]) + __uint_as_float(r4[i & 31]);
    } else if (MODE == 2) {
      tmem_ld_32x32(base + ((i * 32) & 255), r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int e = 0; e < 32; ++e) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(r[e]) * 0.001f)); acc += y; }
    } else {
      uint32_t w[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) w[e] = i + e;
      tmem_st_x16(base + ((i * 16) & 255), w);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 12345.f) *sink = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

int main() {
  long long* d; float* s;
  cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
  const int iters = 2000;
  const char* names[4] = {"ld.x32 + wait", "4 x ld.x32 then wait", "ld.x32 + wait + 32 ex2", "st.x16 + wait"};
  for (int mode = 0; mode < 4; ++mode)
    for (int nw : {4, 8, 16}) {
      for (int grid : {1, 148}) {
        if (mode == 0) k<0><<<grid, nw * 32>>>(d, iters, s);
        if (mode == 1) k<1><<<grid, nw * 32>>>(d, iters, s);
        if (mode == 2) k<2><<<grid, nw * 32>>>(d, iters, s);
        if (mode == 3) k<3><<<grid, nw * 32>>>(d, iters, s);
        long long h[148];
        cudaError_t e = cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        const double per_iter = (double)h[0] / iters;
        const double bytes = (mode == 1 ? 4 : 1) * (mode == 3 ? 2048.0 : 4096.0) * nw;
        printf("%-26s warps %2d grid %3d: %8.1f clk/iter  -> %7.1f B/clk/SM\n", names[mode], nw, grid, per_iter, bytes / per_iter);
      }
    }
  return 0;
}
