// 3x3 stride-1 pad-1 convolution from halo tiles, with GroupNorm (+ SiLU) applied in the operand path (sm_100a).
//
// The implicit-GEMM kernel (gemm.cu) fetches the A operand of a 3x3 conv nine times, once per tap (nine shifted
// 128-pixel boxes per 64 channels).  Here an output tile is 8 x 16 pixels and its A operand for 64 channels is ONE
// 4-D TMA box of the 10 x 18 halo tile (23 040 B, SWIZZLE_128B: one 128-byte row per pixel).  tcgen05 computes the
// swizzle from absolute shared-memory address bits (measured: scripts/ubench/desc_shift.cu), so a K-major descriptor
// may start at any pixel row and step its 8-row groups by any multiple of 16 bytes: the view of tap (ky, kx) is
//     start = tile + (ky * 10 + kx) * 128,  SBO = 10 * 128  (the next image row of the halo tile)
// and the nine taps are nine MMA chains over the same resident tile.  The operand is read 1.4x instead of 9x.
//
// Because the tile is resident and read by nine taps, a transform warpgroup can rewrite it in place between the TMA
// landing and the MMAs at 1.4x (not 9x) the work of a separate pass:
//     a = fp16( silu?( x * scale[b, c] + shift[b, c] ) )      (GroupNorm folded to per-(sample, channel) coefficients)
// which removes the GroupNorm apply pass (one read + one write of the map) in front of every ResBlock conv
// (LD.py:5220-5232, 5250-5261; VAE ResnetBlock LD.py:3562-3576).  Out-of-image pixels stay zero: the reference pads
// the NORMALISED map.
//
// Pipeline (512 threads, one CTA per SM, optionally a CTA pair with cta_group::2 MMAs over a 256-pixel tile):
//   warp 0   A producer: one halo box (or a plain 128-pixel box for 1x1 K segments / the fused residual) per item
//   warp 2   B producer: per item 9 (or 1) weight slices bn x 64, own ring
//   warps 12-15 transform: wait A landed -> rewrite -> fence.proxy.async -> arrive a_ready (on the leader CTA)
//   warp 1   MMA issuer: per item wait a_ready, per tap wait b_full, 4 x tcgen05.mma (K = 16 each), commits
//   warps 4-11 epilogue (as gemm.cu's TMA-store epilogue): TMEM -> +bias, +timestep row -> fp16 -> smem -> TMA store,
//            optional GroupNorm column statistics of the produced tile (colstats)
// Roofline: tensor pipe; algorithmic FLOPs 2 * M * N * 9 * Cin; operand bytes per tile and chunk 23 KB + 9 * bn * 128.
#include "gemm_tile.cuh"

namespace ldx {

#ifndef LDX_HALO_XF_UNROLL
#define LDX_HALO_XF_UNROLL 4      // 16-byte pieces a transform thread keeps in flight (12 per halo tile and thread)
#endif
static constexpr int HALO_W = 10, HALO_H = 18, HALO_PIX = HALO_W * HALO_H;
static constexpr int HA_STAGE_BYTES = 23 * 1024;            // 23 040 used, 1024-aligned slots
static constexpr int HALO_THREADS = 512;
static constexpr int HALO_MAX_A = 4, HALO_MAX_B = 8;
static constexpr int XF_THREADS = 128;
static constexpr int XF_UNROLL = LDX_HALO_XF_UNROLL;

struct __align__(8) HaloBarriers {
  uint64_t a_full[HALO_MAX_A];    // TMA landed (local to each CTA)
  uint64_t a_ready[HALO_MAX_A];   // transformed; on the leader CTA, one arrive per CTA of the pair
  uint64_t a_empty[HALO_MAX_A];   // MMAs that read the stage have retired (multicast commit)
  uint64_t b_full[HALO_MAX_B];
  uint64_t b_empty[HALO_MAX_B];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

int conv_halo_smem_bytes(int bn_cta, int* na, int* nb) {
  const int b_stage = bn_cta * BK * 2;
  const int budget = 227 * 1024 - EPI_SMEM - (int)sizeof(HaloBarriers) - 1024;
  int a = 3, b = (budget - a * HA_STAGE_BYTES) / b_stage;
  if (b < 4) { a = 2; b = (budget - a * HA_STAGE_BYTES) / b_stage; }
  if (b > HALO_MAX_B) b = HALO_MAX_B;
  if (b < 2) return -1;
  *na = a; *nb = b;
  return a * HA_STAGE_BYTES + b * b_stage + EPI_SMEM + (int)sizeof(HaloBarriers) + 1024;
}

// K-major SWIZZLE_128B descriptor with an explicit stride between 8-row groups
LDX_DEVICE uint64_t umma_desc_kmajor_sw128_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// GroupNorm coefficients of the eight channels a transform thread owns, and the in-place rewrite of one 16-byte piece:
// fp16(silu?(x * scale + shift)) with the same fmaf / ex2 / rcp sequence as gn_apply_kernel (bit-identical operand).
struct XfCoef {
  float sc[8], sh[8];
  LDX_DEVICE void load(const float* coef) {      // [8][2] = (scale, shift) pairs, 64-byte aligned
    const float4* cf = reinterpret_cast<const float4*>(coef);
    const float4 q0 = __ldg(cf), q1 = __ldg(cf + 1), q2 = __ldg(cf + 2), q3 = __ldg(cf + 3);
    sc[0] = q0.x; sh[0] = q0.y; sc[1] = q0.z; sh[1] = q0.w; sc[2] = q1.x; sh[2] = q1.y; sc[3] = q1.z; sh[3] = q1.w;
    sc[4] = q2.x; sh[4] = q2.y; sc[5] = q2.z; sh[5] = q2.w; sc[6] = q3.x; sh[6] = q3.y; sc[7] = q3.z; sh[7] = q3.w;
  }
  LDX_DEVICE void apply(uint4* ptr, bool do_silu) const {
    const uint4 vv = *ptr;
    const float2 f0 = unpack_h2(vv.x), f1 = unpack_h2(vv.y), f2 = unpack_h2(vv.z), f3 = unpack_h2(vv.w);
    float f[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = fmaf(f[j], sc[j], sh[j]);
      if (do_silu) f[j] = silu_f(f[j]);
    }
    uint4 o;
    o.x = pack_h2(f[0], f[1]); o.y = pack_h2(f[2], f[3]); o.z = pack_h2(f[4], f[5]); o.w = pack_h2(f[6], f[7]);
    *ptr = o;
  }
};

// The K loop of one output tile as a list of items; every role walks it with its own functor.
//   f(kind, map, c0, dx, dy, kcol, ntaps, src):  kind 0 = halo chunk (A: hmap[map]; B column of tap t: t * halo_cp + kcol),
//   1 = plain 1x1 K step (A: amap[map] shifted by (dx, dy); B column kcol), 2 = fused-residual identity step
//   (A: amap[map] columns c0.., B: rows of I).
template <typename F>
LDX_DEVICE void halo_for_each_item(const GemmKParams& p, int nt, F&& f) {
  for (int s = 0; s < p.halo_nsrc; ++s)
    for (int j = 0; j < p.halo_nch[s]; ++j) f(0, s, j * BK, 0, 0, p.halo_coff[s] + j * BK, 9, s);
  int kcol = 9 * p.halo_cp;
  for (int s = 0; s < p.nseg; ++s) {
    const KSeg sg = p.seg[s];
    for (int ch = 0; ch < sg.nchunks; ++ch, kcol += BK) f(1, sg.map, sg.c0 + ch * BK, sg.dx, sg.dy, kcol, 1, 0);
  }
  for (int dk = 0; dk < p.diag_chunks; ++dk) f(2, p.diag_src, nt * p.bn + dk * BK, 0, 0, dk * BK, 1, 0);
}

// kStats: 0 none, 1 GroupNorm column statistics of the produced tile (colstats_out), 2 LayerNorm row statistics
// (rowstats_out, for a LayerNorm folded into the consuming GEMM)
template <int kCtas, int kStats>
__global__ void __launch_bounds__(HALO_THREADS, 1) ldx_conv_halo_kernel(const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int na = p.na_stages, nb = p.nb_stages;
  const uint32_t cta_rank = kCtas == 2 ? cluster_ctarank() : 0u;
  const int worker = (int)blockIdx.x / kCtas, nworkers = (int)gridDim.x / kCtas;
  const int bn_cta = p.bn / kCtas;
  const int nsub = p.nsub;
  const int sub_rows = bn_cta / nsub;
  const uint32_t b_stage_bytes = (uint32_t)bn_cta * BK * 2;
  const uint32_t b_sub_bytes = (uint32_t)sub_rows * BK * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)na * HA_STAGE_BYTES;
  uint8_t* sC = sB + (size_t)nb * b_stage_bytes;
  HaloBarriers* bars = reinterpret_cast<HaloBarriers*>(sC + EPI_SMEM);

  if (threadIdx.x == 0) {
    for (int i = 0; i < LDX_MAX_SRC; ++i) tma_prefetch_desc(&p.amap[i]);
    tma_prefetch_desc(&p.hmap[0]);
    tma_prefetch_desc(&p.hmap[1]);
    tma_prefetch_desc(&p.bmap);
    tma_prefetch_desc(&p.idmap);
    tma_prefetch_desc(&p.omap);
    tma_prefetch_desc(&p.omap64);
    for (int s = 0; s < na; ++s) {
      mbar_init(&bars->a_full[s], 1);
      mbar_init(&bars->a_ready[s], kCtas);
      mbar_init(&bars->a_empty[s], 1);
    }
    for (int s = 0; s < nb; ++s) {
      mbar_init(&bars->b_full[s], 1);
      mbar_init(&bars->b_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], 4 * EPI_GROUPS * kCtas);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (kCtas == 2) {
      tmem_alloc_pair(&bars->tmem_base, p.tmem_cols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(&bars->tmem_base, p.tmem_cols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (kCtas == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_prologue();

  if (warp == 0) {
    // ------------------------------------------------------------------ A producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = worker; tile < p.num_tiles; tile += nworkers) {
      int nt, tx, ty, tb;
      tile_coords<kCtas>(p, tile, (int)cta_rank, nt, tx, ty, tb);
      const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb;
      halo_for_each_item(p, nt, [&](int kind, int map, int c0, int dx, int dy, int, int, int) {
        mbar_wait(&bars->a_empty[stage], phase ^ 1);
        uint8_t* dA = sA + (size_t)stage * HA_STAGE_BYTES;
        if (elect_one()) {
          if (kind == 0) {
            mbar_arrive_expect_tx(&bars->a_full[stage], HALO_PIX * 128);
            tma_load_4d(dA, &p.hmap[map], &bars->a_full[stage], c0, x0 - 1, y0 - 1, b0);
          } else {
            mbar_arrive_expect_tx(&bars->a_full[stage], A_STAGE_BYTES);
            tma_load_4d(dA, &p.amap[map], &bars->a_full[stage], c0, x0 + dx, y0 + dy, b0);
          }
        }
        __syncwarp();
        if (++stage == na) { stage = 0; phase ^= 1; }
      });
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ B producer
    int stage = 0;
    uint32_t phase = 0;
    const int sub_step = p.bn / nsub;
    for (int tile = worker; tile < p.num_tiles; tile += nworkers) {
      int nt, tx, ty, tb;
      tile_coords<kCtas>(p, tile, (int)cta_rank, nt, tx, ty, tb);
      const int n0 = nt * p.bn + (int)cta_rank * sub_rows;
      halo_for_each_item(p, nt, [&](int kind, int, int, int, int, int kcol, int ntaps, int) {
        const CUtensorMap* bm = kind == 2 ? &p.idmap : &p.bmap;
        const bool d64 = kind == 2 && p.diag64;            // 64 x 64 identity block (see gemm.cu)
        const int brow = d64 ? (int)cta_rank * (64 / kCtas) : kind == 2 ? (int)cta_rank * sub_rows : n0;
        const uint32_t bbytes = d64 ? (uint32_t)(64 / kCtas) * BK * 2 : b_stage_bytes;
        for (int t = 0; t < ntaps; ++t) {
          const int bk = kind == 0 ? t * p.halo_cp + kcol : d64 ? 0 : kcol;
          mbar_wait(&bars->b_empty[stage], phase ^ 1);
          uint8_t* dB = sB + (size_t)stage * b_stage_bytes;
          if constexpr (kCtas == 2) {
            const uint32_t lbar = mapa_shared(smem_u32(&bars->b_full[stage]), 0);
            if (elect_one()) {
              if (cta_rank == 0) mbar_arrive_expect_tx(&bars->b_full[stage], 2 * bbytes);
              tma_load_2d_pair(dB, bm, lbar, bk, brow);
              if (nsub == 2 && !d64) tma_load_2d_pair(dB + b_sub_bytes, bm, lbar, bk, brow + sub_step);
            }
          } else {
            if (elect_one()) {
              mbar_arrive_expect_tx(&bars->b_full[stage], bbytes);
              tma_load_2d(dB, bm, &bars->b_full[stage], bk, brow);
            }
          }
          __syncwarp();
          if (++stage == nb) { stage = 0; phase ^= 1; }
        }
      });
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (cta_rank == 0) {
      const uint32_t idesc = umma_idesc_f16(BM * kCtas, p.bn / nsub);
      const uint32_t idesc64 = umma_idesc_f16(BM * kCtas, 64);
      const uint32_t sub_cols = (uint32_t)(p.bn / nsub);
      int astage = 0, bstage = 0;
      uint32_t aphase = 0, bphase = 0;
      int it = 0;
      for (int tile = worker; tile < p.num_tiles; tile += nworkers, ++it) {
        const int as = p.single_acc ? 0 : (it & 1);
        const uint32_t accphase = p.single_acc ? (uint32_t)(it & 1) : (uint32_t)((it >> 1) & 1);
        mbar_wait(&bars->tmem_empty[as], accphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * p.acc_stride;
        int items_left = p.nk;                      // items of this tile (halo chunks + plain steps + identity steps)
        uint32_t acc = 0;
        halo_for_each_item(p, 0, [&](int kind, int, int, int, int, int kcol, int ntaps, int) {
          mbar_wait(&bars->a_ready[astage], aphase);
          tc_fence_after();
          --items_left;
          const uint32_t a_base = smem_u32(sA + (size_t)astage * HA_STAGE_BYTES);
          for (int t = 0; t < ntaps; ++t) {
            mbar_wait(&bars->b_full[bstage], bphase);
            tc_fence_after();
            const uint32_t a_addr = kind == 0 ? a_base + (uint32_t)((t / 3) * HALO_W + (t % 3)) * 128u : a_base;
            const uint32_t sbo = kind == 0 ? (uint32_t)HALO_W * 128u : 1024u;
            const uint32_t b_addr = smem_u32(sB + (size_t)bstage * b_stage_bytes);
            if (elect_one()) {
              if (kind == 2 && p.diag64) {
                const uint32_t dcol = d_tmem + (uint32_t)kcol;      // kcol = 64 * (identity step index)
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  const uint64_t ad = umma_desc_kmajor_sw128(a_addr + k * 32), bd = umma_desc_kmajor_sw128(b_addr + k * 32);
                  if constexpr (kCtas == 2) umma_f16_ss_pair(dcol, ad, bd, idesc64, 1u);
                  else umma_f16_ss(dcol, ad, bd, idesc64, 1u);
                }
              } else
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                const uint64_t ad = umma_desc_kmajor_sw128_sbo(a_addr + k * 32, sbo);
                const uint64_t bd = umma_desc_kmajor_sw128(b_addr + k * 32);
                if constexpr (kCtas == 2) {
                  umma_f16_ss_pair(d_tmem, ad, bd, idesc, acc | (uint32_t)k);
                  if (nsub == 2)
                    umma_f16_ss_pair(d_tmem + sub_cols, ad, umma_desc_kmajor_sw128(b_addr + b_sub_bytes + k * 32), idesc,
                                     acc | (uint32_t)k);
                } else {
                  umma_f16_ss(d_tmem, ad, bd, idesc, acc | (uint32_t)k);
                }
              }
              if constexpr (kCtas == 2) umma_commit_pair(&bars->b_empty[bstage], 3); else umma_commit(&bars->b_empty[bstage]);
              if (t == ntaps - 1) {
                if constexpr (kCtas == 2) umma_commit_pair(&bars->a_empty[astage], 3); else umma_commit(&bars->a_empty[astage]);
                if (items_left == 0) {
                  if constexpr (kCtas == 2) umma_commit_pair(&bars->tmem_full[as], 3); else umma_commit(&bars->tmem_full[as]);
                }
              }
            }
            __syncwarp();
            acc = 1;
            if (++bstage == nb) { bstage = 0; bphase ^= 1; }
          }
          if (++astage == na) { astage = 0; aphase ^= 1; }
        });
      }
    }
  } else if (warp >= 12) {
    // ------------------------------------------------------------------ transform warpgroup
    const int t = threadIdx.x - 12 * 32;         // 0..127
    const int chunk = t & 7;                     // the 16-byte (8-channel) piece of a pixel row this thread owns
    const uint32_t leader_ready0 = kCtas == 2 ? mapa_shared(smem_u32(&bars->a_ready[0]), 0) : 0u;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = worker; tile < p.num_tiles; tile += nworkers) {
      int nt, tx, ty, tb;
      tile_coords<kCtas>(p, tile, (int)cta_rank, nt, tx, ty, tb);
      const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb;
      halo_for_each_item(p, nt, [&](int kind, int map, int c0, int, int, int kcol, int, int src) {
        mbar_wait(&bars->a_full[stage], phase);
        const int cch = c0 + chunk * 8;            // first channel (inside the source) of this thread's piece
        // which items are normalised: the halo chunks, or (normalised 1x1 conv, halo_nsrc == 0) source 0's plain tiles
        const bool xf_plain = kind == 1 && map == 0 && p.halo_nsrc == 0 && cch < p.halo_c[0];
        const bool xf_halo = kind == 0 && cch < p.halo_c[src];
        if (p.gn_coef != nullptr && b0 < p.B && (xf_plain || xf_halo)) {
          XfCoef k8;
          k8.load(p.gn_coef + ((int64_t)b0 * p.gn_ctot + (xf_halo ? kcol + chunk * 8 : cch)) * 2);
          uint8_t* slot = sA + (size_t)stage * HA_STAGE_BYTES;
          const bool do_silu = p.gn_silu != 0;
          if (xf_halo) {
#pragma unroll XF_UNROLL
            for (int i = 0; i < (HALO_PIX * 8 + XF_THREADS - 1) / XF_THREADS; ++i) {
              const int pix = (t + XF_THREADS * i) >> 3;
              if (pix < HALO_PIX) {
                const int hy = pix / HALO_W, hx = pix - hy * HALO_W;
                const int iy = y0 - 1 + hy, ix = x0 - 1 + hx;
                if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)        // out-of-image pixels stay zero (conv padding)
                  k8.apply(reinterpret_cast<uint4*>(slot + pix * 128 + ((chunk ^ (pix & 7)) << 4)), do_silu);
              }
            }
          } else {                                  // the plain 128-pixel tile [16 rows][8 pixels]
#pragma unroll 4
            for (int i = 0; i < BM * 8 / XF_THREADS; ++i) {
              const int pix = (t + XF_THREADS * i) >> 3;
              if (y0 + (pix >> 3) < p.H)            // rows past the image stay zero (never stored)
                k8.apply(reinterpret_cast<uint4*>(slot + pix * 128 + ((chunk ^ (pix & 7)) << 4)), do_silu);
            }
          }
          fence_proxy_async_smem();
        }
        named_bar_sync(1, XF_THREADS);            // every transform thread has seen the stage and finished its part
        if (t == 0) {
          // (default .release.cta on the remote arrive, as for the TMEM-release arrives of the epilogue: the data it
          // orders went through fence.proxy.async above; a cluster-scope release measured +4 % on the VAE convs)
          if constexpr (kCtas == 2) mbar_arrive_cluster(leader_ready0 + (uint32_t)stage * 8u);
          else mbar_arrive(&bars->a_ready[stage]);
        }
        if (++stage == na) { stage = 0; phase ^= 1; }
      });
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: 2 groups x 4 warps (gemm.cu's
    // TMA-store epilogue: bias, per-sample row add, activation, optional column statistics)
    const int grp = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int n_out_tile = p.bn;
    const int n_out_total = p.N;
    const bool elected = lane == 0;
    uint8_t* sW = sC + (size_t)((grp * 4 + q) * EPI_NSLOT) * EPI_SLOT_BYTES;
    const int sub_x = (q * 32) % p.bw, sub_y = ((q * 32) / p.bw) % p.bh;
    const uint32_t leader_empty0 = kCtas == 2 ? mapa_shared(smem_u32(&bars->tmem_empty[0]), 0) : 0u;
    const uint32_t leader_empty1 = kCtas == 2 ? mapa_shared(smem_u32(&bars->tmem_empty[1]), 0) : 0u;
    auto release_tmem = [&](int as) {
      if constexpr (kCtas == 2) mbar_arrive_cluster(as ? leader_empty1 : leader_empty0);
      else mbar_arrive(&bars->tmem_empty[as]);
    };
    uint32_t pidx = 0;
    int it = 0;
    for (int tile = worker; tile < p.num_tiles; tile += nworkers, ++it) {
      const int as = p.single_acc ? 0 : (it & 1);
      const uint32_t aphase = p.single_acc ? (uint32_t)(it & 1) : (uint32_t)((it >> 1) & 1);
      int nt, tx, ty, tb;
      tile_coords<kCtas>(p, tile, (int)cta_rank, nt, tx, ty, tb);
      const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb;
      const int x = x0 + row % p.bw;
      const int y = y0 + (row / p.bw) % p.bh;
      const int b = b0;
      const bool valid = (x < p.W) && (y < p.H) && (b < p.B);
      const float* rowadd_row = (p.rowadd && valid) ? p.rowadd + (int64_t)b * p.ld_rowadd : nullptr;
      const int n_out0 = nt * n_out_tile;
      const uint32_t trow = tmem_base + as * p.acc_stride + ((uint32_t)(q * 32) << 16);
      const int n_here = min(n_out_tile, n_out_total - n_out0);
      const int npass = (n_here + EPI_PASS - 1) / EPI_PASS;

      mbar_wait(&bars->tmem_full[as], aphase);
      tc_fence_after();
      if (grp >= npass) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_tmem(as);
        if (kStats == 2 && valid)
          *reinterpret_cast<float2*>(p.rowstats_out + ((((int64_t)b * p.H + y) * p.W + x) * (p.n_tiles * EPI_GROUPS) + nt * EPI_GROUPS + grp) * 2) = make_float2(0.f, 0.f);
        continue;
      }
      float rs1 = 0.f, rs2 = 0.f;               // row statistics of this warp's output columns
      const int64_t m_lin = ((int64_t)b * p.H + y) * p.W + x;
      const int rs_parts = p.n_tiles * EPI_GROUPS;
      int64_t cs_base = -1;
      if constexpr (kStats == 1) {
        const int cs_slot = p.cs_slot_base + (ty * p.tiles_x + tx) * p.cs_wps + q;
        if (b0 < p.B) cs_base = ((int64_t)b0 * p.cs_slots + cs_slot) * p.N;
      }
      int last_ps = grp;
      while (last_ps + EPI_GROUPS < npass) last_ps += EPI_GROUPS;
      for (int ps = grp; ps < npass; ps += EPI_GROUPS, ++pidx) {
        uint8_t* cbuf = sW + (pidx % EPI_NSLOT) * EPI_SLOT_BYTES;
        const bool wide = n_out_tile - ps * EPI_PASS >= EPI_PASS;
        const int nhalf = wide ? 2 : 1;
        if (elected) tma_store_wait_read<EPI_NSLOT - 1>();
        __syncwarp();
        for (int half = 0; half < nhalf; ++half) {
          const int c = ps * EPI_PASS + half * EPI_CHUNK;
          const int n_glob = n_out0 + c;
          uint32_t r[32];
          float v[32];
          tmem_ld32(trow + c, r);
          float4 bb4[8];
          const bool bias_vec = n_glob + EPI_CHUNK <= p.N;
          if (bias_vec) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) bb4[j4] = __ldg(reinterpret_cast<const float4*>(p.bias + n_glob) + j4);
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (bias_vec) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              v[j4 * 4 + 0] += bb4[j4].x; v[j4 * 4 + 1] += bb4[j4].y; v[j4 * 4 + 2] += bb4[j4].z; v[j4 * 4 + 3] += bb4[j4].w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n_glob + j < p.N) v[j] += __ldg(p.bias + n_glob + j);
          }
          if (rowadd_row) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n_glob + j < p.N) v[j] += __ldg(rowadd_row + n_glob + j);
          }
          if (ps == last_ps && half == nhalf - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_tmem(as);
          }
          if (p.flags & (LDX_EPI_SILU | LDX_EPI_QUICK_GELU | LDX_EPI_LRELU | LDX_EPI_RELU)) {
            if (p.flags & (LDX_EPI_LRELU | LDX_EPI_RELU)) {
              const float slope = (p.flags & LDX_EPI_RELU) ? 0.0f : 0.2f;
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], slope * v[j]);
            } else {
              const float nk = (p.flags & LDX_EPI_QUICK_GELU) ? -1.702f * 1.4426950408889634f : -1.4426950408889634f;
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = sigmoid_gate_f(v[j], nk);
            }
          }
          if constexpr (kStats == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n_glob + j < n_out_total) { rs1 += v[j]; rs2 = fmaf(v[j], v[j], rs2); }
          }
          uint8_t* srow = wide ? cbuf + lane * 128 : cbuf + lane * 64;
          const int sw = wide ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint4 o;
            o.x = pack_h2(v[ch * 8 + 0], v[ch * 8 + 1]);
            o.y = pack_h2(v[ch * 8 + 2], v[ch * 8 + 3]);
            o.z = pack_h2(v[ch * 8 + 4], v[ch * 8 + 5]);
            o.w = pack_h2(v[ch * 8 + 6], v[ch * 8 + 7]);
            *reinterpret_cast<uint4*>(srow + (((half * 4 + ch) ^ sw) << 4)) = o;
          }
        }
        if constexpr (kStats == 1) {
          __syncwarp();
          const uint32_t vm = __ballot_sync(0xffffffffu, valid);
          const int col = n_out0 + ps * EPI_PASS + 2 * lane;
          if (cs_base >= 0 && lane < (wide ? 32 : 16) && col < n_out_total) {
            const uint32_t chunk = (uint32_t)lane >> 2, off = ((uint32_t)lane & 3u) << 2;
            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll 8
            for (int r2 = 0; r2 < 32; ++r2) {
              const uint32_t a = wide ? (uint32_t)r2 * 128u + ((chunk ^ ((uint32_t)r2 & 7u)) << 4) + off
                                      : (uint32_t)r2 * 64u + ((chunk ^ (((uint32_t)r2 >> 1) & 3u)) << 4) + off;
              const float2 f = unpack_h2(*reinterpret_cast<const uint32_t*>(cbuf + a));
              if ((vm >> r2) & 1u) {
                s0 += f.x; q0 = fmaf(f.x, f.x, q0);
                s1 += f.y; q1 = fmaf(f.y, f.y, q1);
              }
            }
            *reinterpret_cast<float4*>(p.colstats_out + (cs_base + col) * 2) = make_float4(s0, q0, s1, q1);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elected) {
          tma_store_4d(wide ? &p.omap64 : &p.omap, cbuf, n_out0 + ps * EPI_PASS, x0 + sub_x, y0 + sub_y, b0);
          tma_store_commit();
        }
      }
      if (kStats == 2 && valid)
        *reinterpret_cast<float2*>(p.rowstats_out + (m_lin * rs_parts + nt * EPI_GROUPS + grp) * 2) = make_float2(rs1, rs2);
    }
    if (elected) tma_store_wait_all();
  }

  tc_fence_before();
  if constexpr (kCtas == 2) {
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_base, p.tmem_cols);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

template <int kCtas, int kStats>
static int conv_halo_launch_t(const GemmLaunch* g, cudaStream_t stream) {
  static std::atomic<uint64_t> attr_done{0};
  int arc = set_max_dyn_smem_once(ldx_conv_halo_kernel<kCtas, kStats>, 227 * 1024, attr_done);
  if (arc) return arc;
  LDX_CUDA(launch_k(ldx_conv_halo_kernel<kCtas, kStats>, dim3(g->grid), dim3(HALO_THREADS), (size_t)g->smem_bytes, stream,
                    kCtas, g->p));
  return LDX_OK;
}

int conv_halo_launch(const GemmLaunch* g, cudaStream_t stream) {
  const int st = g->p.colstats_out ? 1 : (g->p.rowstats_out ? 2 : 0);
  if (g->ctas == 2)
    return st == 1 ? conv_halo_launch_t<2, 1>(g, stream) : st == 2 ? conv_halo_launch_t<2, 2>(g, stream) : conv_halo_launch_t<2, 0>(g, stream);
  return st == 1 ? conv_halo_launch_t<1, 1>(g, stream) : st == 2 ? conv_halo_launch_t<1, 2>(g, stream) : conv_halo_launch_t<1, 0>(g, stream);
}

}  // namespace ldx
