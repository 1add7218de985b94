// Microbenchmark: TMEM read bandwidth (tcgen05.ld 32x32b.x32) and MUFU.EX2 rate per SM, by warps per CTA.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void tmem_ld_kernel(int iters, float* out, long long* cyc) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s + (((uint32_t)(warp & 3) * 32) << 16);
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[32];
    const uint32_t addr = base + ((it + (warp >> 2)) & 7) * 32;
    asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc += __uint_as_float(r[it & 31]) ;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512u) : "memory");
}
__global__ void mufu_kernel(int iters, float* out, long long* cyc) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float a = 0; for (int i = 0; i < 8; ++i) a += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 1 << 16);
  long long h[148];
  const int iters = 4096;
  for (int warps : {1, 4, 8, 16, 32}) {
    tmem_ld_kernel<<<148, warps * 32>>>(iters, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double bytes = (double)warps * iters * 4096.0;
    printf("tmem_ld warps=%2d: %s cycles=%lld  -> %.1f B/clk/SM\n", warps, cudaGetErrorString(e), h[0], bytes / h[0]);
  }
  for (int warps : {4, 8, 16, 32}) {
    mufu_kernel<<<148, warps * 32>>>(iters, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double n = (double)warps * 32 * iters * 8;
    printf("mufu warps=%2d: %s cycles=%lld  -> %.2f ex2/clk/SM\n", warps, cudaGetErrorString(e), h[0], n / h[0]);
  }
  return 0;
}
