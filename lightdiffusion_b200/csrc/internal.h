// Host-side internals shared by the .cu translation units of libldx.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>

#include "../../include/ldx.h"

namespace ldx {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define LDX_CHECK_ARG(cond, ...)   \
  do {                             \
    if (!(cond)) {                 \
      ldx::set_error(__VA_ARGS__); \
      return LDX_E_INVALID;        \
    }                              \
  } while (0)

#define LDX_CUDA(call)                                        \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) return ldx::cuda_fail(e__, #call); \
  } while (0)

inline cudaStream_t as_stream(ldx_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LDX_PDL");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}

// Launch with programmatic stream serialization (the kernel must call pdl_prologue() before it touches
// global memory written by its predecessors) and an optional thread-block cluster.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
int num_sms();          // of the calling thread's current device (cached per device)
int current_device();

static constexpr int LDX_MAX_DEVICES = 64;

// cudaFuncSetAttribute is per device: one bit per device ordinal, set after the first successful call on that
// device.  Two threads racing on the same device both set the (idempotent) attribute; nothing is torn.
template <typename Kernel>
inline int set_max_dyn_smem_once(Kernel kernel, int bytes, std::atomic<uint64_t>& done) {
  const uint64_t bit = 1ull << (current_device() & (LDX_MAX_DEVICES - 1));
  if (!(done.load(std::memory_order_acquire) & bit)) {
    LDX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.fetch_or(bit, std::memory_order_release);
  }
  return LDX_OK;
}

// ---- implicit GEMM: frozen kernel parameters (TMA descriptors encoded once)
struct KSeg {
  int32_t map, c0, dx, dy, nchunks;
};

struct alignas(64) GemmKParams {
  CUtensorMap amap[LDX_MAX_SRC];
  CUtensorMap bmap;
  CUtensorMap omap;    // output, box (32 cols, one warp's 32-pixel sub-box), SWIZZLE_64B: tail passes
  CUtensorMap omap64;  // output, box (64 cols, same sub-box), SWIZZLE_128B: full passes
  CUtensorMap idmap;   // 256 x 256 identity, box (64, bn / ctas): B operand of the fused-residual K steps
  KSeg seg[LDX_MAX_SEG];
  int32_t nseg, nk;    // nk = K steps per tile, including diag_chunks
  int32_t diag_src, diag_chunks;   // residual fused as an identity K segment: amap index, bn / 64 steps (0 = none)
  int32_t diag64;                  // 1: each identity step is an N = 64 MMA chain on its own 64 accumulator columns
  int32_t N, W, H, B;
  int32_t bw, bh, bb;
  int32_t tiles_x, tiles_y, tiles_b, n_tiles, num_tiles;
  uint32_t inv_n_tiles, inv_tiles_x, inv_tiles_xy;
  // split-K: num_tiles = nsplit * base_tiles, split = tile / base_tiles owns K steps [split * ksplit, ...) and
  // writes an fp32 partial [rows, N] at out + split * split_stride (elements)
  int32_t nsplit, base_tiles, ksplit;
  int64_t split_stride;   // ceil(2^32 / d): tile index -> coordinates without integer division
  int32_t bn, stages;
  // nsub = 2: the N tile is two bn / 2 wide accumulators fed by the same A stage (bn up to 512, CTA pairs only);
  // they fill TMEM, so the accumulator is single buffered (single_acc) and the epilogue is not overlapped
  int32_t nsub, single_acc;
  uint32_t tmem_cols, acc_stride;
  void* out;
  int64_t o_sx, o_sy, o_sb, o_off;
  const float* bias;
  const float* rowadd;
  int64_t ld_rowadd;
  const void* res;
  int64_t ldr;
  int32_t flags;
  float* rowstats_out;        // [rows][n_tiles * 2][2]
  const float* rowstats_in;   // [rows][rs_in_parts][2]
  const float* ln_gsum;
  int32_t rs_in_parts;
  float ln_eps, ln_inv_c;
  // GroupNorm statistics of the produced tensor: [B][cs_slots][N][2] per-(32-row block, column) partial sums
  float* colstats_out;
  int32_t cs_slots, cs_slot_base, cs_wps;   // cs_wps: 32-row blocks of one batch element inside a 128-row tile
  int32_t vec_ok;   // 16-byte vector epilogue allowed (direct-store path)
  int32_t tma_epi;  // 1: stage 32-column chunks in smem and TMA-store them (fp16 outputs)
  // ---- halo-tile 3x3 conv (conv_halo.cu): the conv's K axis is halo_nsrc sources x 64-channel chunks, each chunk
  // ONE (64, 10, 18, 1) box whose nine tap views the MMA reads through shifted descriptors; seg[] then holds only the
  // trailing 1x1 K segments (weight columns from 9 * halo_cp on)
  CUtensorMap hmap[2];
  int32_t halo_nsrc, halo_cp;       // halo_cp: weight columns per tap (sum of the sources' padded channel counts)
  int32_t halo_nch[2], halo_coff[2], halo_c[2];   // chunks, first weight / coefficient column, real channels per source
  int32_t na_stages, nb_stages;
  const float* gn_coef;             // [B][gn_ctot][2] (scale, shift) or null: GroupNorm (+SiLU) applied to the landed tile
  int32_t gn_ctot, gn_silu;
  int32_t w_batch_rows, bias_bstride;   // per-sample weights / bias (ldx_gemm_desc_t.w_batch_rows), 0 = shared
};

// second launch of a split-K GEMM: out = act(sum_s partial[s] + bias + rowadd + residual) as fp16
struct SplitReduce {
  const float* ws;
  int nsplit;
  int64_t split_stride;
  int M, N, W, H;
  const float* bias;
  const float* rowadd;
  int64_t ld_rowadd;
  const void* res;
  int64_t ldr;
  int silu;
  void* out;
  int64_t o_sx, o_sy, o_sb, o_off;
};

struct GemmLaunch {
  GemmKParams p;
  int grid;
  int smem_bytes;
  int ctas;      // 1, or 2 = CTA-pair kernel (cluster of 2, tcgen05 cta_group::2)
  int halo = 0;               // 1: conv_halo.cu kernel (3x3 conv from halo tiles)
  int nsplit = 1;             // > 1: split-K, two launches (GEMM into the workspace, then SplitReduce)
  size_t ws_bytes = 0;        // fp32 workspace the caller must bind with gemm_bind_workspace()
  SplitReduce red;
};
void gemm_bind_workspace(GemmLaunch* g, void* ws);

int gemm_prepare(const ldx_gemm_desc_t* d, GemmLaunch* out);
int gemm_launch(const GemmLaunch* g, cudaStream_t stream);
int conv_halo_launch(const GemmLaunch* g, cudaStream_t stream);     // conv_halo.cu
int conv_halo_smem_bytes(int bn_cta, int* na, int* nb);              // ring depths that fit; total dynamic smem
int encode_map_f16(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

int run_generic_op(const ldx_generic_op_t* op, cudaStream_t stream);

}  // namespace ldx
