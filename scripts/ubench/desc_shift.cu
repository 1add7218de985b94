// Micro-experiment: can a K-major SWIZZLE_128B tcgen05 shared-memory descriptor start on a row that is NOT the
// first row of a 1024-byte swizzle atom, and with an 8-row-group stride (SBO) other than 1024?  That is what a
// 3x3 conv needs to read its nine tap views out of ONE halo tile (pixel rows of 128 B, image rows 2048 B apart).
// A (160+ rows x 64 fp16) is written in the layout TMA produces (16-byte chunk index XOR (row & 7), absolute row
// index, 1024-aligned base); B = 64 x 64 identity, so D[m][n] must equal A[row(m)][n] with
// row(m) = shift + (m / 8) * (sbo / 128) + m % 8.  Variants: base_offset field = 0 or (shift & 7).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../lightdiffusion_b200/csrc desc_shift.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

using namespace ldx;

static constexpr int ROWS = 512;   // rows of A resident in smem (64 KB)

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t sbo, uint32_t base_off) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)(base_off & 7) << 49) | ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(128, 1) desc_shift_kernel(const __half* a_log, float* out, int shift, int sbo,
                                                            int base_off) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;                       // ROWS x 128 B
  uint8_t* sB = smem + ROWS * 128;          // 64 x 128 B identity
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < ROWS * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(a_log + r * 64 + c * 8);
  }
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    __half v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float2half(c * 8 + j == r ? 1.0f : 0.0f);
    *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 64); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(128, 64);
      const uint32_t a0 = smem_u32(sA) + shift * 128, b0 = smem_u32(sB);
      for (int k = 0; k < 4; ++k)
        umma_f16_ss(tmem, desc_sw128(a0 + k * 32, sbo, base_off), desc_sw128(b0 + k * 32, 1024, 0), idesc, k != 0);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t r[32];
  for (int c = 0; c < 64; c += 32) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[threadIdx.x * 64 + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<__half> a(ROWS * 64);
  // the value names the row (mod 256) and the 16-byte chunk; integers <= 2047 are exact in fp16
  for (int r = 0; r < ROWS; ++r)
    for (int c = 0; c < 64; ++c) a[r * 64 + c] = __float2half((float)((r % 256) + 256 * (c / 8)));
  __half* da; float* dout;
  cudaMalloc(&da, a.size() * 2);
  cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(desc_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  std::vector<float> out(128 * 64);
  const int sbos[3] = {1024, 2048, 1280};
  for (int si = 0; si < 3; ++si)
    for (int shift = 0; shift < 20; shift += (shift < 9 ? 1 : 8))
      for (int mode = 0; mode < 2; ++mode) {
        const int sbo = sbos[si], bo = mode ? (shift & 7) : 0;
        if (mode == 1 && bo == 0) continue;
        cudaMemset(dout, 0xff, 128 * 64 * 4);
        desc_shift_kernel<<<1, 128, ROWS * 128 + 64 * 128 + 2048>>>(da, dout, shift, sbo, bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sbo %d shift %d base_off %d: CUDA error %s\n", sbo, shift, bo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first_m = -1; float got = 0, want = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            const int row = shift + (m / 8) * (sbo / 128) + m % 8;
            const float w = __half2float(a[row * 64 + n]);
            if (out[m * 64 + n] != w) { if (!bad) { first_m = m * 64 + n; got = out[m * 64 + n]; want = w; } ++bad; }
          }
        printf("sbo %4d shift %2d base_off %d: %s (%d mismatches", sbo, shift, bo, bad ? "WRONG" : "ok", bad);
        if (bad) printf(", first at m=%d n=%d got %.4f want %.4f", first_m / 64, first_m % 64, got, want);
        printf(")\n");
      }
  return 0;
}
