// Pieces shared by the implicit-GEMM kernel (gemm.cu) and the halo-tile 3x3 conv kernel (conv_halo.cu):
// tile constants, the epilogue staging geometry and the tile index -> coordinates map.
#pragma once
#include "common.cuh"
#include "internal.h"

namespace ldx {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KiB
static constexpr int EPI_GROUPS = 2;                          // epilogue warpgroups (4 warps = 128 TMEM lanes each)
static constexpr int GEMM_THREADS = 64 + EPI_GROUPS * 128;   // warp 0 TMA, warp 1 MMA, then the epilogue groups
static constexpr int MAX_STAGES = 6;

static constexpr int EPI_CHUNK = 32;                         // output columns per TMEM load
static constexpr int EPI_PASS = 64;                          // output columns per staged pass (one fence + one TMA store)
static constexpr int EPI_SLOT_BYTES = 32 * EPI_PASS * 2;     // 4 KiB: one warp's 32 rows x 128 bytes, SWIZZLE_128B
static constexpr int EPI_NSLOT = 2;                          // staging slots per epilogue warp
static constexpr int EPI_SMEM = EPI_GROUPS * 4 * EPI_NSLOT * EPI_SLOT_BYTES;

// tile index -> (n tile, x / y / batch tile) with multiply-high instead of integer division (exact while
// index * divisor < 2^32, far beyond any launch here); the divisions cost ~1000 cycles per tile and warp.
template <int kCtas>
__device__ __forceinline__ void tile_coords(const GemmKParams& p, int tile, int cta_rank, int& nt, int& tx, int& ty,
                                            int& tb) {
  const int mq = p.n_tiles == 1 ? tile : (int)__umulhi((uint32_t)tile, p.inv_n_tiles);
  nt = tile - mq * p.n_tiles;
  const int mt = mq * kCtas + cta_rank;
  tb = p.tiles_x * p.tiles_y == 1 ? mt : (int)__umulhi((uint32_t)mt, p.inv_tiles_xy);
  const int rem = mt - tb * (p.tiles_x * p.tiles_y);
  ty = p.tiles_x == 1 ? rem : (int)__umulhi((uint32_t)rem, p.inv_tiles_x);
  tx = rem - ty * p.tiles_x;
}

}  // namespace ldx
