// Microbenchmark: TMEM read bandwidth of tcgen05.ld by shape / outstanding loads / warps per SM, and the
// MUFU.EX2 rate.  One CTA per SM (148), all of TMEM allocated.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LD_X32(r, addr) asm volatile( \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
      : "r"(addr) : "memory")
#define LD_X16(r, addr) asm volatile( \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 " \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
      : "r"(addr) : "memory")
#define LD_X8(r, addr) asm volatile( \
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) \
      : "r"(addr) : "memory")
// 16x256b.x8: 16 lanes x 8 x 256 bits = 32 registers per thread, 4 KB per warp
#define LD_16x256_X8(r, addr) asm volatile( \
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 " \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
      : "r"(addr) : "memory")
#define WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

// MODE 0: x32 + wait each; 1: two x32 in flight; 2: x16 + wait; 3: x8 + wait; 4: 16x256b.x8 + wait;
// 5: x32, every warp its own column range
template <int MODE>
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
  uint32_t accu = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[32], q[32];
    const uint32_t col = MODE == 5 ? (uint32_t)((warp >> 2) * 64 + (it & 1) * 32) : (uint32_t)(((it + (warp >> 2)) & 7) * 32);
    if (MODE == 0 || MODE == 5) { LD_X32(r, base + col); WAIT_LD(); { uint32_t x = 0; _Pragma("unroll") for (int i = 0; i < 32; ++i) x ^= r[i]; accu ^= x; } }
    if (MODE == 1) { LD_X32(r, base + col); LD_X32(q, base + ((col + 32) & 255)); WAIT_LD(); { uint32_t x = 0; _Pragma("unroll") for (int i = 0; i < 32; ++i) x ^= r[i] ^ q[i]; accu ^= x; } }
    if (MODE == 2) { LD_X16(r, base + col); WAIT_LD(); { uint32_t x = 0; _Pragma("unroll") for (int i = 0; i < 16; ++i) x ^= r[i]; accu ^= x; } }
    if (MODE == 3) { LD_X8(r, base + col); WAIT_LD(); { uint32_t x = 0; _Pragma("unroll") for (int i = 0; i < 8; ++i) x ^= r[i]; accu ^= x; } }
    if (MODE == 4) { LD_16x256_X8(r, base + col); WAIT_LD(); { uint32_t x = 0; _Pragma("unroll") for (int i = 0; i < 32; ++i) x ^= r[i]; accu ^= x; } }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(accu);
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
template <int MODE>
void run(const char* name, double bytes_per_iter, float* out, long long* cyc) {
  long long h[148];
  const int iters = 2048;
  for (int warps : {1, 2, 4, 8, 16}) {
    tmem_ld_kernel<MODE><<<148, warps * 32>>>(iters, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double bytes = (double)warps * iters * bytes_per_iter;
    printf("%-28s warps=%2d: %s cycles=%8lld  %6.1f clk/iter/warp -> %6.1f B/clk/SM\n", name, warps, cudaGetErrorString(e), h[0], (double)h[0] / iters, bytes / h[0]);
  }
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 1 << 16);
  run<0>("32x32b.x32 wait-each", 4096, out, cyc);
  run<1>("32x32b.x32 two-in-flight", 8192, out, cyc);
  run<2>("32x32b.x16 wait-each", 2048, out, cyc);
  run<3>("32x32b.x8 wait-each", 1024, out, cyc);
  run<4>("16x256b.x8 wait-each", 4096, out, cyc);
  run<5>("32x32b.x32 own-columns", 4096, out, cyc);
  long long h[148];
  for (int warps : {4, 16}) {
    mufu_kernel<<<148, warps * 32>>>(2048, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double n = (double)warps * 32 * 2048 * 8;
    printf("mufu warps=%2d: %s cycles=%lld  -> %.2f ex2/clk/SM\n", warps, cudaGetErrorString(e), h[0], n / h[0]);
  }
  return 0;
}
