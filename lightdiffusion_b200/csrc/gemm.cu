// Implicit GEMM on tcgen05 / TMEM fed by TMA (sm_100a).
//
// One persistent, warp-specialised kernel serves F.linear (LD.py:2368), 1x1 conv and 3x3 conv
// (LD.py:2382) in all the variants the SD1.5 UNet / VAE decoder use (stride 1, stride 2 through
// parity-strided views, nearest-2x-upsample folded into 2x2 phase convs, channel-concat inputs).
//
//   out[b,y,x,n] = epi( sum over K segments s, channels c:  src_s[b, y+dy_s, x+dx_s, c0_s+c] * w[n, k] )
//
// * A operand: 128 output pixels x 64 channels per pipeline stage, fetched by ONE 4-D TMA box
//   (64, bw, bh, bb) on the NHWC activation; the tap shift (dx,dy) moves the box, and TMA's
//   out-of-bounds zero fill implements the conv zero padding.  The box lands in the canonical
//   K-major SWIZZLE_128B layout tcgen05.mma reads.
// * B operand: bn x 64 slice of the [N, K] weight matrix, 2-D TMA, same layout.
// * D: fp32 accumulators in TMEM, double buffered (2 x bn columns) so the epilogue of tile i
//   overlaps the main loop of tile i+1.
// * warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2..9 = epilogue, each warp an
//   independent pipeline on its 32 rows (tcgen05.ld -> bias / timestep-embedding row add / folded LayerNorm /
//   SiLU / GEGLU -> fp16 into its own SWIZZLE_128B staging slot -> its own TMA store).
// * Residual adds (x + h) are an extra K segment against a 256 x 256 identity: they ride the operand pipeline
//   and the tensor core adds them into the fp32 accumulator.
// * Split-K (few output tiles, long K): K slices per CTA-pair group, fp32 partials, splitk_reduce_kernel.
//
// Roofline: tensor-pipe bound; algorithmic FLOPs = 2 * M * N * K.
#include <stdlib.h>

#include <mutex>

#include "gemm_tile.cuh"

namespace ldx {

#ifdef LDX_GEMM_TIMING
// Variant-build instrumentation: CTA 0 stamps %globaltimer at pipeline milestones (ns).
__device__ unsigned long long g_gemm_ts[16];
__device__ __forceinline__ void ts_mark(int i) {
  if (blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_gemm_ts[i] = t;
  }
}
#define TS_MARK(i) ts_mark(i)
#define PH_DECL long long ph_t = clock64(), ph_acc[7] = {0, 0, 0, 0, 0, 0, 0}; int ph_n = 0;
#define PH(i) { const long long ph_now = clock64(); ph_acc[i] += ph_now - ph_t; ph_t = ph_now; }
#define PH_COUNT ++ph_n;
#define PH_FLUSH if (blockIdx.x == 0 && threadIdx.x == 64) { for (int i = 0; i < 7; ++i) g_gemm_ts[8 + i] = (unsigned long long)ph_acc[i]; g_gemm_ts[15] = (unsigned long long)ph_n; }
#else
#define TS_MARK(i)
#define PH_DECL
#define PH(i)
#define PH_COUNT
#define PH_FLUSH
#endif

struct __align__(8) GemmBarriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

template <bool kGeglu, bool kOutF32>
__device__ __forceinline__ void epilogue_chunk(const GemmKParams& p, uint32_t tmem_row_addr, int col, int n_glob,
                                               bool row_valid, char* out_row, const __half* res_row,
                                               const float* rowadd_row) {
  // one thread = one output row; 32 consecutive output columns starting at n_glob
  uint32_t r[32];
  float v[32];
  tmem_ld32(tmem_row_addr + col, r);
  if constexpr (kGeglu) {
    uint32_t g[32];
    tmem_ld32(tmem_row_addr + col + p.bn / 2, g);
    tmem_ld_wait();
    const int n_packed = (n_glob / (p.bn / 2)) * p.bn + (n_glob % (p.bn / 2));   // column in packed weight order
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float a = __uint_as_float(r[j]), gt = __uint_as_float(g[j]);
      if (p.bias) {
        a += __ldg(p.bias + n_packed + j);
        gt += __ldg(p.bias + n_packed + p.bn / 2 + j);
      }
      v[j] = a * gelu_erf_fast(gt);
    }
  } else {
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n_glob + j < p.N) v[j] += __ldg(p.bias + n_glob + j);
    }
    if (rowadd_row) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n_glob + j < p.N) v[j] += __ldg(rowadd_row + n_glob + j);
    }
  }
  if (!row_valid) return;
  const int n_out_total = kGeglu ? p.N / 2 : p.N;
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
  if (p.vec_ok) {
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      const int n = n_glob + g8 * 8;
      if (n >= n_out_total) break;
      if (res_row) {
        const uint4 rr = __ldg(reinterpret_cast<const uint4*>(res_row + n));
        const float2 a = unpack_h2(rr.x), b = unpack_h2(rr.y), c = unpack_h2(rr.z), d = unpack_h2(rr.w);
        v[g8 * 8 + 0] += a.x; v[g8 * 8 + 1] += a.y; v[g8 * 8 + 2] += b.x; v[g8 * 8 + 3] += b.y;
        v[g8 * 8 + 4] += c.x; v[g8 * 8 + 5] += c.y; v[g8 * 8 + 6] += d.x; v[g8 * 8 + 7] += d.y;
      }
      if constexpr (kOutF32) {
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_row) + n);
        o[0] = make_float4(v[g8 * 8 + 0], v[g8 * 8 + 1], v[g8 * 8 + 2], v[g8 * 8 + 3]);
        o[1] = make_float4(v[g8 * 8 + 4], v[g8 * 8 + 5], v[g8 * 8 + 6], v[g8 * 8 + 7]);
      } else {
        uint4 o;
        o.x = pack_h2(v[g8 * 8 + 0], v[g8 * 8 + 1]);
        o.y = pack_h2(v[g8 * 8 + 2], v[g8 * 8 + 3]);
        o.z = pack_h2(v[g8 * 8 + 4], v[g8 * 8 + 5]);
        o.w = pack_h2(v[g8 * 8 + 6], v[g8 * 8 + 7]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(out_row) + n) = o;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = n_glob + j;
      if (n < n_out_total) {
        float val = v[j];
        if (res_row) val += __half2float(res_row[n]);
        if constexpr (kOutF32) reinterpret_cast<float*>(out_row)[n] = val;
        else reinterpret_cast<__half*>(out_row)[n] = __float2half_rn(val);
      }
    }
  }
}

// kCtas = 1: one CTA per 128 x bn tile.  kCtas = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per
// 256 x bn tile: each CTA loads its own 128 A rows and HALF of the B tile, so the L2->smem operand
// traffic per FLOP drops by a third and the MMA reads B from both CTAs' shared memory.
// kEpi: epilogue features compiled in (bit 0 GEGLU, bit 1 LayerNorm-folded input rows, bit 2 row statistics
// out).  They are template parameters because the epilogue is register-bound: a single kernel carrying every
// path spilled inside the hot loop and slowed the plain GEMMs too.
constexpr int KEPI_GEGLU = 1, KEPI_ROWNORM = 2, KEPI_ROWSTATS = 4, KEPI_COLSTATS = 8;
template <int kCtas, int kEpi>
__global__ void __launch_bounds__(GEMM_THREADS, 1) ldx_gemm_kernel(const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);   // SWIZZLE_128B wants 1024-B alignment

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TS_MARK(0);
  const int stages = p.stages;
  const uint32_t cta_rank = kCtas == 2 ? cluster_ctarank() : 0u;
  const int worker = (int)blockIdx.x / kCtas, nworkers = (int)gridDim.x / kCtas;
  const int bn_cta = p.bn / kCtas;                           // B rows this CTA loads per stage
  const int nsub = p.nsub;                                   // accumulators per tile (1, or 2 halves of bn / 2 columns)
  const int sub_rows = bn_cta / nsub;                        // B rows of one accumulator held by this CTA
  const uint32_t b_stage_bytes = (uint32_t)bn_cta * BK * 2;
  const uint32_t b_sub_bytes = (uint32_t)sub_rows * BK * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)stages * A_STAGE_BYTES;
  uint8_t* sC = sB + (size_t)stages * b_stage_bytes;          // per-warp staging slots (1024-aligned)
  GemmBarriers* bars = reinterpret_cast<GemmBarriers*>(sC + EPI_SMEM);

  if (threadIdx.x == 0) {
    for (int i = 0; i < LDX_MAX_SRC; ++i) tma_prefetch_desc(&p.amap[i]);
    tma_prefetch_desc(&p.bmap);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], 4 * EPI_GROUPS * kCtas);
    }
    tma_prefetch_desc(&p.omap);
    tma_prefetch_desc(&p.omap64);
    tma_prefetch_desc(&p.idmap);
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
  if (threadIdx.x == 0) TS_MARK(1);
  pdl_prologue();   // everything above overlaps the tail of the previous kernel
  if (threadIdx.x == 0) TS_MARK(2);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (warp-uniform loop)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < p.num_tiles; tile += nworkers) {
        int nt, tx, ty, tb;
        const int split = p.nsplit > 1 ? tile / p.base_tiles : 0;
        tile_coords<kCtas>(p, tile - split * p.base_tiles, (int)cta_rank, nt, tx, ty, tb);
        const int k_lo = split * p.ksplit, k_hi = min(p.nk, k_lo + p.ksplit);   // this tile's K steps
        // rows of B this CTA fetches for accumulator j: bn0 + j * (bn / nsub) (the pair splits each accumulator's rows)
        const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb * p.bb;
        const int n0 = nt * p.bn + (int)cta_rank * sub_rows + b0 * p.w_batch_rows;   // per-sample weights: bb == 1
        const int sub_step = p.bn / nsub;
        int kcol = 0;
        // b_bytes: bytes of B this CTA loads for the step (the whole stage, or the 64-column identity block)
        auto load_step = [&](const CUtensorMap* am, int ac0, int ax, int ay, const CUtensorMap* bm, int bk, int bn0,
                             uint32_t b_bytes, bool second) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* dA = sA + (size_t)stage * A_STAGE_BYTES;
          uint8_t* dB = sB + (size_t)stage * b_stage_bytes;
          if constexpr (kCtas == 2) {
            // both CTAs' loads complete on the leader's barrier, which the leader arms for the pair
            const uint32_t lbar = mapa_shared(smem_u32(&bars->full[stage]), 0);
            if (elect_one()) {
              if (cta_rank == 0) mbar_arrive_expect_tx(&bars->full[stage], 2 * (A_STAGE_BYTES + b_bytes));
              tma_load_4d_pair(dA, am, lbar, ac0, ax, ay, b0);
              tma_load_2d_pair(dB, bm, lbar, bk, bn0);
              if (second) tma_load_2d_pair(dB + b_sub_bytes, bm, lbar, bk, bn0 + sub_step);
            }
          } else {
            if (elect_one()) {
              mbar_arrive_expect_tx(&bars->full[stage], A_STAGE_BYTES + b_bytes);
              tma_load_4d(dA, am, &bars->full[stage], ac0, ax, ay, b0);
              tma_load_2d(dB, bm, &bars->full[stage], bk, bn0);
            }
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        };
        int step = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const KSeg sg = p.seg[s];
          if (step + sg.nchunks <= k_lo || step >= k_hi) {       // segment outside this split's K range
            step += sg.nchunks;
            kcol += sg.nchunks * BK;
            continue;
          }
          for (int ch = 0; ch < sg.nchunks; ++ch, ++step, kcol += BK)
            if (step >= k_lo && step < k_hi)
              load_step(&p.amap[sg.map], sg.c0 + ch * BK, x0 + sg.dx, y0 + sg.dy, &p.bmap, kcol, n0, b_stage_bytes, nsub == 2);
        }
        // residual fused as an identity K segment (see gemm_prepare): A = the residual tile's own column
        // range, B = the matching rows of I
        // (diag64: the identity is applied per 64-column block of the tile -- A = the residual's columns of that
        // block, B = the 64 x 64 identity, D = that block's accumulator columns -- so a step costs an N = 64 MMA
        // chain instead of an N = bn one)
        for (int dk = 0; dk < p.diag_chunks; ++dk) {
          if (p.diag64)
            load_step(&p.amap[p.diag_src], nt * p.bn + dk * BK, x0, y0, &p.idmap, 0, (int)cta_rank * (64 / kCtas),
                      (uint32_t)(64 / kCtas) * BK * 2, false);
          else
            load_step(&p.amap[p.diag_src], nt * p.bn + dk * BK, x0, y0, &p.idmap, dk * BK, (int)cta_rank * sub_rows,
                      b_stage_bytes, nsub == 2);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA; warp-uniform
    // loop, one elected lane issues)
    if (cta_rank == 0) {
      const uint32_t idesc = umma_idesc_f16(BM * kCtas, p.bn / nsub);
      const uint32_t idesc64 = umma_idesc_f16(BM * kCtas, 64);
      const uint32_t sub_cols = (uint32_t)(p.bn / nsub);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = worker; tile < p.num_tiles; tile += nworkers, ++it) {
        const int as = p.single_acc ? 0 : (it & 1);
        const uint32_t aphase = p.single_acc ? (uint32_t)(it & 1) : (uint32_t)((it >> 1) & 1);
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * p.acc_stride;
        const int split = p.nsplit > 1 ? tile / p.base_tiles : 0;
        const int nk_tile = p.nsplit > 1 ? min(p.nk, (split + 1) * p.ksplit) - split * p.ksplit : p.nk;
        for (int kb = 0; kb < nk_tile; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          if (lane == 0 && it == 0 && kb == 0) TS_MARK(3);
          if (lane == 0 && it == 0 && kb == nk_tile - 1) TS_MARK(4);
          const uint32_t a_addr = smem_u32(sA + (size_t)stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sB + (size_t)stage * b_stage_bytes);
          const int dk = kb - (nk_tile - p.diag_chunks);       // >= 0: fused-residual identity step
          if (p.diag64 && dk >= 0) {
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                const uint64_t ad = umma_desc_kmajor_sw128(a_addr + k * 32), bd = umma_desc_kmajor_sw128(b_addr + k * 32);
                if constexpr (kCtas == 2) umma_f16_ss_pair(d_tmem + (uint32_t)dk * 64u, ad, bd, idesc64, 1u);
                else umma_f16_ss(d_tmem + (uint32_t)dk * 64u, ad, bd, idesc64, 1u);
              }
              if constexpr (kCtas == 2) umma_commit_pair(&bars->empty[stage], 3); else umma_commit(&bars->empty[stage]);
              if (kb == nk_tile - 1) {
                if constexpr (kCtas == 2) umma_commit_pair(&bars->tmem_full[as], 3); else umma_commit(&bars->tmem_full[as]);
              }
            }
          } else if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t ad = umma_desc_kmajor_sw128(a_addr + k * 32), bd = umma_desc_kmajor_sw128(b_addr + k * 32);
              if constexpr (kCtas == 2) {
                umma_f16_ss_pair(d_tmem, ad, bd, idesc, (uint32_t)((kb | k) != 0));
                if (nsub == 2)    // second accumulator: same A stage, the other half of the B stage
                  umma_f16_ss_pair(d_tmem + sub_cols, ad, umma_desc_kmajor_sw128(b_addr + b_sub_bytes + k * 32), idesc,
                                   (uint32_t)((kb | k) != 0));
              } else {
                umma_f16_ss(d_tmem, ad, bd, idesc, (uint32_t)((kb | k) != 0));
              }
            }
            // frees the smem stage (in both CTAs) once these MMAs have read it
            if constexpr (kCtas == 2) umma_commit_pair(&bars->empty[stage], 3); else umma_commit(&bars->empty[stage]);
            // accumulator complete -> epilogue (of both CTAs)
            if (kb == nk_tile - 1) {
              if constexpr (kCtas == 2) umma_commit_pair(&bars->tmem_full[as], 3); else umma_commit(&bars->tmem_full[as]);
            }
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: EPI_GROUPS x 4 warps.
    // A thread owns one output row; group g owns the 64-column passes ps = g, g + EPI_GROUPS, ... of a tile.
    // Per pass: two 32-column steps TMEM -> regs -> (+bias, +rowadd, activation) -> fp16 into the warp's
    // staging slot (32 rows x 128 B, SWIZZLE_128B), then ONE proxy fence and ONE TMA store of the warp's
    // 32-row sub-box.  Every warp runs its own pipeline (own slots, own bulk groups): there is no block- or
    // group-level barrier in the loop.  The residual never reaches the epilogue: it is accumulated by the
    // tensor core through an identity K segment.
    const int grp = (warp - 2) >> 2;
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    constexpr bool geglu = (kEpi & KEPI_GEGLU) != 0;
    constexpr bool rownorm = (kEpi & KEPI_ROWNORM) != 0;
    constexpr bool rowstats = (kEpi & KEPI_ROWSTATS) != 0;
    constexpr bool colstats = (kEpi & KEPI_COLSTATS) != 0;
    const bool f32 = (p.flags & LDX_EPI_OUT_F32) != 0;
    const int n_out_tile = geglu ? p.bn / 2 : p.bn;
    const int n_out_total = geglu ? p.N / 2 : p.N;
    const int osz = f32 ? 4 : 2;
    const bool elected = lane == 0;             // lane 0 of each warp issues the warp's TMA stores
    uint8_t* sW = sC + (size_t)((grp * 4 + q) * EPI_NSLOT) * EPI_SLOT_BYTES;   // this warp's staging ring
    // origin of the warp's 32 rows inside the 128-pixel tile box (bw, bh, bb are powers of two)
    const int sub_x = (q * 32) % p.bw, sub_y = ((q * 32) / p.bw) % p.bh, sub_b = (q * 32) / (p.bw * p.bh);
    const uint32_t leader_empty0 = kCtas == 2 ? mapa_shared(smem_u32(&bars->tmem_empty[0]), 0) : 0u;
    const uint32_t leader_empty1 = kCtas == 2 ? mapa_shared(smem_u32(&bars->tmem_empty[1]), 0) : 0u;
    auto release_tmem = [&](int as) {
      if constexpr (kCtas == 2) mbar_arrive_cluster(as ? leader_empty1 : leader_empty0);
      else mbar_arrive(&bars->tmem_empty[as]);
    };
    uint32_t pidx = 0;                          // this warp's running pass counter (staging ring)
    int it = 0;
    PH_DECL
    for (int tile = worker; tile < p.num_tiles; tile += nworkers, ++it) {
      const int as = p.single_acc ? 0 : (it & 1);
      const uint32_t aphase = p.single_acc ? (uint32_t)(it & 1) : (uint32_t)((it >> 1) & 1);
      int nt, tx, ty, tb;
      const int split = p.nsplit > 1 ? tile / p.base_tiles : 0;
      tile_coords<kCtas>(p, tile - split * p.base_tiles, (int)cta_rank, nt, tx, ty, tb);
      const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb * p.bb;
      const int x = x0 + row % p.bw;
      const int y = y0 + (row / p.bw) % p.bh;
      const int b = b0 + row / (p.bw * p.bh);
      const bool valid = (x < p.W) && (y < p.H) && (b < p.B);
      const float* rowadd_row = (p.rowadd && valid) ? p.rowadd + (int64_t)b * p.ld_rowadd : nullptr;
      const int n_out0 = nt * n_out_tile;
      const uint32_t trow = tmem_base + as * p.acc_stride + ((uint32_t)(q * 32) << 16);
      const int n_here = min(n_out_tile, n_out_total - n_out0);        // output columns of this tile
      const int npass = (n_here + EPI_PASS - 1) / EPI_PASS;

      PH(5)
      mbar_wait(&bars->tmem_full[as], aphase);
      tc_fence_after();
      PH(0)
      if (threadIdx.x == 64 && it == 0) TS_MARK(5);

      if (p.flags & (1 << 17)) {               // debug: epilogue only drains the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_tmem(as);
        continue;
      }
      if (p.tma_epi) {
        const int64_t m_lin = ((int64_t)b * p.H + y) * p.W + x;      // output row index
        const int rs_parts = p.n_tiles * EPI_GROUPS;
        if (grp >= npass) {                     // nothing to do for this group in a narrow tile
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_tmem(as);
          if (rowstats && valid)
            *reinterpret_cast<float2*>(p.rowstats_out + (m_lin * rs_parts + nt * EPI_GROUPS + grp) * 2) = make_float2(0.f, 0.f);
          continue;
        }
        // LayerNorm folded into this GEMM: statistics of the A row (= output row of a plain GEMM)
        float ln_mean = 0.f, ln_rstd = 1.f;
        if (rownorm && valid) {
          float s1 = 0.f, s2 = 0.f;
          const float2* st = reinterpret_cast<const float2*>(p.rowstats_in) + m_lin * p.rs_in_parts;
          for (int i = 0; i < p.rs_in_parts; ++i) {     // fixed order: bit-reproducible
            const float2 t2 = __ldg(st + i);
            s1 += t2.x;
            s2 += t2.y;
          }
          ln_mean = s1 * p.ln_inv_c;
          ln_rstd = rsqrtf(fmaxf(s2 * p.ln_inv_c - ln_mean * ln_mean, 0.f) + p.ln_eps);
        }
        const float* bias_b = p.bias + (int64_t)(p.bias_bstride ? min(b, p.B - 1) : 0) * p.bias_bstride;   // per-sample bias
        float rs1 = 0.f, rs2 = 0.f;             // row statistics of this warp's output columns
        // column statistics: the warp's 32 rows are one slot of one batch element (gemm_prepare checks bw * bh >= 32)
        int64_t cs_base = -1;
        if constexpr (colstats) {
          const int per = p.bw * p.bh;          // rows of one batch element inside the tile
          const int cs_b = b0 + (q * 32) / per;
          const int cs_slot = p.cs_slot_base + (ty * p.tiles_x + tx) * p.cs_wps + ((q * 32) % per) / 32;
          if (cs_b < p.B) cs_base = ((int64_t)cs_b * p.cs_slots + cs_slot) * p.N;
        }
        int last_ps = grp;                      // this group's last pass of the tile
        while (last_ps + EPI_GROUPS < npass) last_ps += EPI_GROUPS;
        for (int ps = grp; ps < npass; ps += EPI_GROUPS, ++pidx) {
          uint8_t* cbuf = sW + (pidx % EPI_NSLOT) * EPI_SLOT_BYTES;
          // a pass is 64 columns wide unless the tile's width leaves a 32-column tail
          const bool wide = n_out_tile - ps * EPI_PASS >= EPI_PASS;
          const int nhalf = wide ? 2 : 1;
          PH(5)
          if (elected) tma_store_wait_read<EPI_NSLOT - 1>();   // the slot's previous store has drained it
          __syncwarp();
          PH(1)
          PH_COUNT
          for (int half = 0; half < nhalf; ++half) {
            const int c = ps * EPI_PASS + half * EPI_CHUNK;     // column inside the tile
            const int n_glob = n_out0 + c;
            uint32_t r[32];
            float v[32];
            tmem_ld32(trow + c, r);
            if (geglu) {
              uint32_t g[32];
              tmem_ld32(trow + c + p.bn / 2, g);
              tmem_ld_wait();
              // p.bias is never null here (gemm_prepare substitutes a zero vector): no per-load predicates
              const float4* bv = reinterpret_cast<const float4*>(p.bias + nt * p.bn + c);
              const float4* bgp = bv + (p.bn >> 3);
              // with the LayerNorm folded in: rstd * (acc - mean * gsum[n]) + bias'[n] = acc * ln_a + (gsum[n] * ln_b + bias')
              const float ln_a = rownorm ? ln_rstd : 1.0f, ln_b = rownorm ? -ln_mean * ln_rstd : 0.0f;
              const float4* gv = reinterpret_cast<const float4*>((rownorm ? p.ln_gsum : p.bias) + nt * p.bn + c);
              const float4* ggp = gv + (p.bn >> 3);
              // packed fp32 pairs (FFMA2 / FMUL2): the same operations as the scalar form, about half the issue slots of
              // an epilogue whose limiter is the warp schedulers
              const f32x2_t lna2 = f2_pack(ln_a, ln_a), lnb2 = f2_pack(ln_b, ln_b);
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 ba = __ldg(bv + j4), bg = __ldg(bgp + j4);
                f32x2_t ba0 = f2_pack(ba.x, ba.y), ba1 = f2_pack(ba.z, ba.w), bg0 = f2_pack(bg.x, bg.y), bg1 = f2_pack(bg.z, bg.w);
                if (rownorm) {
                  const float4 ga = __ldg(gv + j4), gg = __ldg(ggp + j4);
                  ba0 = f2_fma(f2_pack(ga.x, ga.y), lnb2, ba0);
                  ba1 = f2_fma(f2_pack(ga.z, ga.w), lnb2, ba1);
                  bg0 = f2_fma(f2_pack(gg.x, gg.y), lnb2, bg0);
                  bg1 = f2_fma(f2_pack(gg.z, gg.w), lnb2, bg1);
                }
                const f32x2_t a0 = f2_fma(f2_pack(__uint_as_float(r[j4 * 4 + 0]), __uint_as_float(r[j4 * 4 + 1])), lna2, ba0);
                const f32x2_t a1 = f2_fma(f2_pack(__uint_as_float(r[j4 * 4 + 2]), __uint_as_float(r[j4 * 4 + 3])), lna2, ba1);
                const f32x2_t g0 = f2_fma(f2_pack(__uint_as_float(g[j4 * 4 + 0]), __uint_as_float(g[j4 * 4 + 1])), lna2, bg0);
                const f32x2_t g1 = f2_fma(f2_pack(__uint_as_float(g[j4 * 4 + 2]), __uint_as_float(g[j4 * 4 + 3])), lna2, bg1);
                f2_unpack(f2_mul(a0, gelu_erf_fast2(g0)), v[j4 * 4 + 0], v[j4 * 4 + 1]);
                f2_unpack(f2_mul(a1, gelu_erf_fast2(g1)), v[j4 * 4 + 2], v[j4 * 4 + 3]);
              }
            } else {
              // bias / rowadd values are fetched while the TMEM load is in flight
              float4 bb4[8];
              const bool bias_vec = n_glob + EPI_CHUNK <= p.N;      // p.bias is never null on this path
              if (bias_vec) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) bb4[j4] = __ldg(reinterpret_cast<const float4*>(bias_b + n_glob) + j4);
              }
              tmem_ld_wait();
              PH(3)
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
              if (rownorm) {                   // N % 32 == 0 is checked on the host (so bias_vec holds)
                // rstd * (acc - mean * gsum[n]) + bias[n] = acc * ln_a + (gsum[n] * ln_b + bias[n]): two FFMA2 per column pair
                const float4* gv = reinterpret_cast<const float4*>(p.ln_gsum + n_glob);
                const float ln_b = -ln_mean * ln_rstd;
                const f32x2_t lna2 = f2_pack(ln_rstd, ln_rstd), lnb2 = f2_pack(ln_b, ln_b);
                if (!bias_vec) {                // cannot happen with N % 32 == 0; keeps the registers defined
#pragma unroll
                  for (int j4 = 0; j4 < 8; ++j4) bb4[j4] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 ga = __ldg(gv + j4);
                  const f32x2_t b0 = f2_fma(f2_pack(ga.x, ga.y), lnb2, f2_pack(bb4[j4].x, bb4[j4].y));
                  const f32x2_t b1 = f2_fma(f2_pack(ga.z, ga.w), lnb2, f2_pack(bb4[j4].z, bb4[j4].w));
                  f2_unpack(f2_fma(f2_pack(v[j4 * 4 + 0], v[j4 * 4 + 1]), lna2, b0), v[j4 * 4 + 0], v[j4 * 4 + 1]);
                  f2_unpack(f2_fma(f2_pack(v[j4 * 4 + 2], v[j4 * 4 + 3]), lna2, b1), v[j4 * 4 + 2], v[j4 * 4 + 3]);
                }
              } else if (bias_vec) {
                const f32x2_t one2 = f2_pack(1.0f, 1.0f);       // v * 1 + b == v + b, one FFMA2 per pair
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  f2_unpack(f2_fma(f2_pack(v[j4 * 4 + 0], v[j4 * 4 + 1]), one2, f2_pack(bb4[j4].x, bb4[j4].y)), v[j4 * 4 + 0], v[j4 * 4 + 1]);
                  f2_unpack(f2_fma(f2_pack(v[j4 * 4 + 2], v[j4 * 4 + 3]), one2, f2_pack(bb4[j4].z, bb4[j4].w)), v[j4 * 4 + 2], v[j4 * 4 + 3]);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n_glob + j < p.N) v[j] += __ldg(bias_b + n_glob + j);
              }
              if (rowadd_row) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n_glob + j < p.N) v[j] += __ldg(rowadd_row + n_glob + j);
              }
            }
            PH(2)
            if (ps == last_ps && half == nhalf - 1) {   // this warp has read all its accumulator columns
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
            if (rowstats) {                    // row statistics for a LayerNorm folded into the consumer GEMM
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n_glob + j < n_out_total) { rs1 += v[j]; rs2 = fmaf(v[j], v[j], rs2); }
            }
            // wide: SWIZZLE_128B rows of 128 B, 16-byte chunk index XOR (row & 7);
            // tail: SWIZZLE_64B rows of 64 B, chunk index XOR ((row >> 1) & 3)
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
          if constexpr (colstats) {
            // per-column (sum, sum of squares) over the warp's 32 staged rows, read back as the fp16 values the
            // consumer will see; lane l owns columns 2l, 2l + 1 of the pass: one conflict-free 4-byte load per row
            __syncwarp();
            const uint32_t vm = __ballot_sync(0xffffffffu, valid);
            const int col = n_out0 + ps * EPI_PASS + 2 * lane;
            if (cs_base >= 0 && lane < (wide ? 32 : 16) && col < n_out_total) {
              const uint32_t chunk = (uint32_t)lane >> 2, off = ((uint32_t)lane & 3u) << 2;
              float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll 8
              for (int r = 0; r < 32; ++r) {
                const uint32_t a = wide ? (uint32_t)r * 128u + ((chunk ^ ((uint32_t)r & 7u)) << 4) + off
                                        : (uint32_t)r * 64u + ((chunk ^ (((uint32_t)r >> 1) & 3u)) << 4) + off;
                const float2 f = unpack_h2(*reinterpret_cast<const uint32_t*>(cbuf + a));
                if ((vm >> r) & 1u) {
                  s0 += f.x; q0 = fmaf(f.x, f.x, q0);
                  s1 += f.y; q1 = fmaf(f.y, f.y, q1);
                }
              }
              *reinterpret_cast<float4*>(p.colstats_out + (cs_base + col) * 2) = make_float4(s0, q0, s1, q1);
            }
          }
          PH(4)
          fence_proxy_async_smem();
          PH(6)
          __syncwarp();
          if (elected && !(p.flags & (1 << 16))) {   // (bit 16: debug, skip the store)
            tma_store_4d(wide ? &p.omap64 : &p.omap, cbuf, n_out0 + ps * EPI_PASS, x0 + sub_x, y0 + sub_y, b0 + sub_b);
            tma_store_commit();
          }
          PH(1)
        }
        if (rowstats && valid)
          *reinterpret_cast<float2*>(p.rowstats_out + (m_lin * rs_parts + nt * EPI_GROUPS + grp) * 2) = make_float2(rs1, rs2);
        continue;
      }

      // ---- direct path (fp32 outputs / unaligned outputs)
      const int64_t opix = p.o_off + (int64_t)b * p.o_sb + (int64_t)y * p.o_sy + (int64_t)x * p.o_sx +
                           (int64_t)split * p.split_stride;
      char* out_row = reinterpret_cast<char*>(p.out) + opix * osz;
      const __half* res_row =
          (p.res && valid) ? reinterpret_cast<const __half*>(p.res) + (((int64_t)b * p.H + y) * p.W + x) * p.ldr
                           : nullptr;
      for (int c = grp * 32; c < n_out_tile; c += 32 * EPI_GROUPS) {
        if (geglu) {
          if (f32) epilogue_chunk<true, true>(p, trow, c, n_out0 + c, valid, out_row, res_row, rowadd_row);
          else epilogue_chunk<true, false>(p, trow, c, n_out0 + c, valid, out_row, res_row, rowadd_row);
        } else {
          if (f32) epilogue_chunk<false, true>(p, trow, c, n_out0 + c, valid, out_row, res_row, rowadd_row);
          else epilogue_chunk<false, false>(p, trow, c, n_out0 + c, valid, out_row, res_row, rowadd_row);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) release_tmem(as);
    }
    PH_FLUSH
    if (threadIdx.x == 64) TS_MARK(6);
    if (p.tma_epi && elected) tma_store_wait_all();
    if (threadIdx.x == 64) TS_MARK(7);
  }

  tc_fence_before();
  if constexpr (kCtas == 2) {
    cluster_sync_all();       // the peer's smem / TMEM stay alive until the leader's last MMA has retired
    if (warp == 1) tmem_dealloc_pair(tmem_base, p.tmem_cols);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// Second launch of a split-K GEMM: sums the fp32 partials in split order (bit-reproducible) and applies the
// epilogue the single-launch path fuses (bias, timestep-embedding row add, residual, SiLU) -> fp16.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const SplitReduce r) {
  pdl_prologue();
  const int nv = r.N >> 2;                                   // float4 vectors per row (N % 4 == 0)
  const int64_t total = (int64_t)r.M * nv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / nv), n = (int)(i % nv) * 4;
    float4 acc = __ldg(reinterpret_cast<const float4*>(r.ws + (int64_t)m * r.N + n));
    for (int sp = 1; sp < r.nsplit; ++sp) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(r.ws + sp * r.split_stride + (int64_t)m * r.N + n));
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    const int hw = r.W * r.H;
    const int b = m / hw, y = (m % hw) / r.W, x = m % r.W;
    if (r.bias) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(r.bias + n));
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    if (r.rowadd) {
      const float* ra = r.rowadd + (int64_t)b * r.ld_rowadd + n;
      acc.x += __ldg(ra); acc.y += __ldg(ra + 1); acc.z += __ldg(ra + 2); acc.w += __ldg(ra + 3);
    }
    if (r.res) {
      const uint2 rr = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(r.res) + (int64_t)m * r.ldr + n));
      const float2 a0 = unpack_h2(rr.x), a1 = unpack_h2(rr.y);
      acc.x += a0.x; acc.y += a0.y; acc.z += a1.x; acc.w += a1.y;
    }
    if (r.silu >= 3) {            // 3 = LeakyReLU(0.2), 4 = ReLU
      const float slope = r.silu == 4 ? 0.0f : 0.2f;
      acc.x = fmaxf(acc.x, slope * acc.x); acc.y = fmaxf(acc.y, slope * acc.y);
      acc.z = fmaxf(acc.z, slope * acc.z); acc.w = fmaxf(acc.w, slope * acc.w);
    } else if (r.silu) {          // 1 = SiLU, 2 = quick_gelu
      const float nk = r.silu == 2 ? -1.702f * 1.4426950408889634f : -1.4426950408889634f;
      acc.x = sigmoid_gate_f(acc.x, nk); acc.y = sigmoid_gate_f(acc.y, nk);
      acc.z = sigmoid_gate_f(acc.z, nk); acc.w = sigmoid_gate_f(acc.w, nk);
    }
    __half* o = reinterpret_cast<__half*>(r.out) + r.o_off + (int64_t)b * r.o_sb + (int64_t)y * r.o_sy + (int64_t)x * r.o_sx + n;
    *reinterpret_cast<uint2*>(o) = make_uint2(pack_h2(acc.x, acc.y), pack_h2(acc.z, acc.w));
  }
}

// ================================================================================ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// fp16 tensor map: up to 4 dims, dim0 contiguous, SWIZZLE_128B, zero OOB fill.
int encode_map_f16_sw(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return LDX_E_NODEVICE;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) {
    set_error("TMA source pointer %p is not 16-byte aligned", ptr);
    return LDX_E_INVALID;
  }
  for (int i = 0; i + 1 < rank; ++i)
    if (gstr[i] % 16 != 0) {
      set_error("TMA stride %d = %llu bytes is not a multiple of 16", i + 1, (unsigned long long)gstr[i]);
      return LDX_E_INVALID;
    }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return LDX_E_CUDA;
  }
  return LDX_OK;
}

int encode_map_f16(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  return encode_map_f16_sw(map, ptr, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

static int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}
static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p *= 2;
  return p;
}

// Read-only device constants, one copy per device ordinal, created on first use under a mutex (the reference
// samples from a worker thread, LD.py:10453, and one process may drive several GPUs).
static std::mutex g_const_mu;

// IDENT_N x IDENT_N fp16 identity: B operand of the fused-residual K steps.  Any bn x 64 slice of a larger identity
// that a tile needs equals a slice of this one (bn <= 512 with two accumulators).
static constexpr int IDENT_N = 512;
static const void* identity256() {
  static void* dev_ptr[LDX_MAX_DEVICES] = {};
  const int dev = current_device() & (LDX_MAX_DEVICES - 1);
  std::lock_guard<std::mutex> lock(g_const_mu);
  if (!dev_ptr[dev]) {
    __half* host = (__half*)calloc((size_t)IDENT_N * IDENT_N, sizeof(__half));
    if (!host) return nullptr;
    for (int i = 0; i < IDENT_N; ++i) host[(size_t)i * IDENT_N + i] = __float2half(1.0f);
    void* ptr = nullptr;
    if (cudaMalloc(&ptr, (size_t)IDENT_N * IDENT_N * sizeof(__half)) == cudaSuccess &&
        cudaMemcpy(ptr, host, (size_t)IDENT_N * IDENT_N * sizeof(__half), cudaMemcpyHostToDevice) == cudaSuccess)
      dev_ptr[dev] = ptr;
    free(host);
  }
  return dev_ptr[dev];
}

const void* identity_matrix() { return identity256(); }      // for the other translation units (ff_fused.cu)
int identity_matrix_n() { return IDENT_N; }

// Zero fp32 vector: stands in for a missing bias so the TMA-store epilogue needs no null checks.
static const float* zeros_f32(int n) {
  static float* dev_ptr[LDX_MAX_DEVICES] = {};
  static int cap[LDX_MAX_DEVICES] = {};
  const int dev = current_device() & (LDX_MAX_DEVICES - 1);
  std::lock_guard<std::mutex> lock(g_const_mu);
  if (n > cap[dev]) {
    const int want = n > 16384 ? n : 16384;
    float* ptr = nullptr;
    if (cudaMalloc(&ptr, (size_t)want * sizeof(float)) != cudaSuccess) return nullptr;
    if (cudaMemset(ptr, 0, (size_t)want * sizeof(float)) != cudaSuccess) return nullptr;
    dev_ptr[dev] = ptr;   // an older, smaller vector stays alive for launches that already captured it
    cap[dev] = want;
  }
  return dev_ptr[dev];
}

// fp16 output rows addressable with 8-byte vector stores (split-K reduce kernel)
static bool tma_ok_out_aligned(const ldx_gemm_desc_t* d) {
  return ((reinterpret_cast<uintptr_t>(d->out) + (uint64_t)d->o_off * 2) % 8 == 0) && d->o_sx % 4 == 0 &&
         d->o_sy % 4 == 0 && d->o_sb % 4 == 0 && d->o_sx > 0;
}

// Tile configuration: N tile width, 1-CTA vs CTA-pair kernel, one or two accumulators per tile.  Cost model per
// SM in tensor-pipe clocks: per 16-wide K step the tensor pipe needs bn/2 cycles and the operand fetch (L2 -> smem)
// (128 + bn/ctas) rows x 32 bytes at an effective ~61 B/clk/SM (calibrated on B200: the 1-CTA 128x256 tile runs at
// ~0.64 of the MMA rate); a tile costs nk16 such steps, and a CTA runs ceil(tiles / workers) tiles back to back.
// Two-accumulator tiles (CTA pairs, bn in {320, ..., 512} = 2 x bn/2 columns sharing every A stage) exist for the
// wave count: N = 1280 at M = 4096 is 80 pair tiles of 256 x 256 on 74 CTA pairs -- two waves, the second 8 % full
// -- but 64 tiles of 256 x 320, ONE wave.  They fill TMEM, so their accumulator is single buffered and the epilogue
// of a tile (~12 clocks per column, measured: 65536 x 320 x 1280 + residual lost 10 % at 7) is exposed instead of
// hidden behind the next tile's main loop.
static void choose_config(int N, int m_tiles, int flags, int sms, int force_ctas, int nk16, bool allow_dual, int* bn_out,
                          int* ctas_out, int* nsub_out, bool halo = false, int max_bn_cta = 256) {
  const bool geglu = (flags & LDX_EPI_GEGLU) != 0;
  double best_cost = 1e30;
  *bn_out = 128;
  *ctas_out = 1;
  *nsub_out = 1;
  if (nk16 < 1) nk16 = 1;
  for (int ctas = 1; ctas <= 2; ++ctas) {
    if (force_ctas && ctas != force_ctas) continue;
    for (int bn = (ctas == 2 && allow_dual) ? 512 : 256; bn >= 32; bn -= 32) {
      const int nsub = bn > 256 ? 2 : 1;
      if (nsub == 2 && (bn % 64 || geglu)) continue;          // each accumulator is a multiple of 32 columns wide
      if (geglu && bn % 64) continue;
      if (ctas == 2 && bn < 64) continue;
      const long long n_tiles = (N + bn - 1) / bn;
      const long long tiles = ((m_tiles + ctas - 1) / ctas) * n_tiles;
      const long long workers = sms / ctas;
      const double waves = (double)((tiles + workers - 1) / workers);
      if (bn / ctas > max_bn_cta) continue;
      // halo-tile convs fetch their A rows once per nine taps (x 1.4 for the halo): ~20 row-equivalents per K step
      const double mma = bn / 2.0, fetch = ((halo ? 20 : 128) + bn / (double)ctas) * 0.52;
      const double per_k = mma > fetch ? mma : fetch;
      double cost = waves * (per_k + 6.0) * nk16;
      if (nsub == 2) cost = cost * 1.03 + waves * bn * 12.0;  // exposed epilogue; keep single-accumulator tiles on ties
      if (cost < best_cost - 1e-9) {
        best_cost = cost;
        *bn_out = bn;
        *ctas_out = ctas;
        *nsub_out = nsub;
      }
    }
  }
}

int gemm_prepare(const ldx_gemm_desc_t* d, GemmLaunch* g) {
  LDX_CHECK_ARG(d && g, "null descriptor");
  LDX_CHECK_ARG(d->nsrc >= 1 && d->nsrc <= LDX_MAX_SRC, "nsrc %d out of range", d->nsrc);
  // the halo-tile kernel (conv_halo.cu) also serves 1x1 convs whose input is normalised in the operand path
  // (halo_nsrc == 0, gn_coef set: SpatialTransformer.norm + proj_in, LD.py:4246-4251): its K items are then seg[] only
  const bool halo = d->halo_nsrc > 0 || d->gn_coef != nullptr;
  LDX_CHECK_ARG(d->nseg >= (d->halo_nsrc > 0 ? 0 : 1) && d->nseg <= LDX_MAX_SEG, "nseg %d out of range", d->nseg);
  if (halo) {
    LDX_CHECK_ARG(d->halo_nsrc >= 0 && d->halo_nsrc <= 2 && d->halo_nsrc <= d->nsrc, "halo_nsrc %d out of range", d->halo_nsrc);
    LDX_CHECK_ARG(d->W % 8 == 0, "halo conv: W = %d must be a multiple of 8", d->W);
    LDX_CHECK_ARG(!(d->flags & (LDX_EPI_GEGLU | LDX_EPI_ROWNORM | LDX_EPI_OUT_F32)) && !(d->rowstats_out && d->colstats_out),
                  "halo conv: fp16 output, no GEGLU / ROWNORM, at most one of rowstats_out / colstats_out");
    if (d->halo_nsrc == 0)
      LDX_CHECK_ARG(d->src[0].W == d->W && d->src[0].H == d->H && d->src[0].B == d->B && d->src[0].C % 8 == 0,
                    "normalised 1x1 conv: source 0 must cover the output domain with C %% 8 == 0");
    for (int i = 0; i < d->halo_nsrc; ++i) {
      LDX_CHECK_ARG(d->src[i].W == d->W && d->src[i].H == d->H && d->src[i].B == d->B && d->src[i].C % 8 == 0,
                    "halo conv: source %d must cover the output domain (stride 1) with C %% 8 == 0", i);
      LDX_CHECK_ARG(i == d->halo_nsrc - 1 || d->src[i].C % 64 == 0, "halo conv: source %d needs C %% 64 == 0", i);
    }
    LDX_CHECK_ARG(!d->gn_coef || (reinterpret_cast<uintptr_t>(d->gn_coef) % 16 == 0 && d->gn_ctot % 8 == 0),
                  "halo conv: gn_coef alignment");
  }
  LDX_CHECK_ARG(d->N >= 1 && d->W >= 1 && d->H >= 1 && d->B >= 1, "bad output extents");
  LDX_CHECK_ARG(d->w && d->out, "null weight/output pointer");
  GemmKParams& p = g->p;
  memset(&p, 0, sizeof(p));

  // tile geometry: 128 output pixels = bw x bh x bb
  int bw, bh, bb;
  if (halo) {
    bw = 8; bh = 16; bb = 1;                     // 8-pixel row groups: the unit a shifted tap view can step by
  } else if (d->H == 1 && d->B == 1) {
    bw = 128; bh = 1; bb = 1;
  } else {
    int div = 1;
    while (div < 128 && d->W % (div * 2) == 0) div *= 2;
    bw = div >= 8 ? div : (pow2_floor(d->W) < 128 ? pow2_floor(d->W) : 128);
    int rem = 128 / bw;
    bh = pow2_ceil(d->H) < rem ? pow2_ceil(d->H) : rem;
    bb = rem / bh;
  }
  p.bw = bw; p.bh = bh; p.bb = bb;
  p.W = d->W; p.H = d->H; p.B = d->B; p.N = d->N;
  p.tiles_x = (d->W + bw - 1) / bw;
  p.tiles_y = (d->H + bh - 1) / bh;
  p.tiles_b = (d->B + bb - 1) / bb;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;

  const int sms = num_sms();
  static int force_ctas = -1;
  if (force_ctas < 0) {
    const char* e = getenv("LDX_GEMM_CTAS");
    force_ctas = e ? atoi(e) : 0;
  }
  // TMA-store epilogue (fp16 outputs whose strides satisfy the TMA rules); otherwise direct stores
  const int osz = (d->flags & LDX_EPI_OUT_F32) ? 4 : 2;
  bool tma_ok = osz == 2 && ((reinterpret_cast<uintptr_t>(d->out) + (uint64_t)d->o_off * 2) % 16 == 0) &&
                (d->o_sx * 2) % 16 == 0 && (d->o_sy * 2) % 16 == 0 && (d->o_sb * 2) % 16 == 0 && d->o_sx > 0;
  if (d->bias && (reinterpret_cast<uintptr_t>(d->bias) % 16 != 0)) tma_ok = false;   // float4 bias loads
  // Residual: fused into the GEMM as an identity K segment (A = the residual tile's own columns, B = I),
  // so it rides the operand pipeline and the fp32 accumulator adds it exactly.
  bool fuse_res = d->residual && tma_ok && d->nsrc < LDX_MAX_SRC && (reinterpret_cast<uintptr_t>(d->residual) % 16 == 0) &&
                  (d->ldr * 2) % 16 == 0;
  if (d->residual && !fuse_res) tma_ok = false;                 // direct-store path adds the residual itself
  // Split-K: a GEMM with few output tiles and a long K (the 8x8-resolution convs: M = 1024) cannot fill the
  // GPU with output tiles, and narrow tiles re-read A once per N tile.  Use 256-wide pair tiles and give each
  // of `nsplit` CTA-pair groups a slice of K; fp32 partials go to a workspace and a second launch reduces them
  // and applies the epilogue.
  int k_chunks = 0;
  for (int s = 0; s < d->nseg; ++s) k_chunks += d->seg[s].nchunks;
  int halo_chunks = 0;                                          // 64-channel chunks of the halo sources
  for (int i = 0; i < d->halo_nsrc; ++i) halo_chunks += (d->src[i].C + 63) / 64;
  k_chunks += 9 * halo_chunks;
  int nsplit = 1;
  {
    static int split_env = -1;
    if (split_env < 0) { const char* e = getenv("LDX_GEMM_SPLITK"); split_env = e ? atoi(e) : 1; }
    const int pair_tiles = ((m_tiles + 1) / 2) * ((d->N + 255) / 256);
    const int workers2 = sms / 2;
    const bool eligible = split_env && !halo && d->bn == 0 && osz == 2 && !(d->flags & (LDX_EPI_GEGLU | LDX_EPI_ROWNORM)) &&
                          !d->rowstats_out && !d->colstats_out && d->N % 8 == 0 && d->N >= 256 && tma_ok_out_aligned(d) &&
                          (!d->residual || ((reinterpret_cast<uintptr_t>(d->residual) % 8 == 0) && d->ldr % 4 == 0)) &&
                          (!d->bias || reinterpret_cast<uintptr_t>(d->bias) % 16 == 0);
    if (eligible && pair_tiles * 2 <= workers2 && k_chunks >= 160) {   // measured: K = 5120 is faster unsplit
      nsplit = workers2 / pair_tiles;
      if (nsplit > k_chunks / 16) nsplit = k_chunks / 16;
      if (nsplit > 8) nsplit = 8;
      if (nsplit < 2) nsplit = 1;
    }
  }
  if (nsplit > 1) { tma_ok = false; fuse_res = false; }
  g->nsplit = nsplit;
  int bn = d->bn, ctas = 1, nsub = 1;
  static int dual_env = -1;
  if (dual_env < 0) { const char* e = getenv("LDX_GEMM_DUAL"); dual_env = e ? atoi(e) : 1; }
  // K steps of 16 per tile, as the kernel will run them (the fused residual adds ~bn / 64 identity steps of 64)
  const int nk16 = 4 * (k_chunks + (d->residual ? 4 : 0));
  if (nsplit > 1) {
    bn = 256;
    ctas = 2;
  } else if (bn == 0) {
    // two-accumulator tiles need the TMA-store epilogue (fp16 output): the direct-store path walks 32-column chunks
    choose_config(d->N, m_tiles, d->flags, sms, force_ctas, nk16, dual_env != 0 && tma_ok, &bn, &ctas, &nsub, halo && getenv("LDX_HALO_COST") != nullptr);
  } else {
    int bn_unused, nsub_unused;
    choose_config(d->N, m_tiles, d->flags, sms, force_ctas, nk16, false, &bn_unused, &ctas, &nsub_unused);
    if (ctas == 2 && bn < 64) ctas = 1;
    if (bn > 256) { ctas = 2; nsub = 2; }       // explicit two-accumulator tile (experiments / tests)
  }
  g->ctas = ctas;
  LDX_CHECK_ARG(bn >= 32 && bn <= (nsub == 2 ? 512 : 256) && bn % 32 == 0, "bn %d must be a multiple of 32 in [32,256]", bn);
  LDX_CHECK_ARG(nsub == 1 || (ctas == 2 && bn % 64 == 0), "two-accumulator tiles need a CTA pair and bn %% 64 == 0");
  if (d->flags & LDX_EPI_GEGLU) {
    LDX_CHECK_ARG(bn % 64 == 0 && d->N % bn == 0, "GEGLU needs bn %% 64 == 0 and N %% bn == 0 (N=%d bn=%d)", d->N, bn);
    LDX_CHECK_ARG(!d->residual && !d->rowadd, "GEGLU epilogue does not take residual/rowadd");
  }
  p.bn = bn;
  p.n_tiles = (d->N + bn - 1) / bn;
  p.base_tiles = ((m_tiles + ctas - 1) / ctas) * p.n_tiles;  // (pair-)tiles
  p.num_tiles = p.base_tiles * nsplit;
  p.nsplit = nsplit;
  auto inv32 = [](int dv) { return dv <= 1 ? 0u : (uint32_t)(((1ull << 32) + (uint64_t)dv - 1) / (uint64_t)dv); };
  // divisor 1 has no 32-bit reciprocal: tile_coords() must then see quotient = index, which umulhi(x, 0) = 0
  // would break, so those cases use the identity below
  p.inv_n_tiles = inv32(p.n_tiles);
  p.inv_tiles_x = inv32(p.tiles_x);
  p.inv_tiles_xy = inv32(p.tiles_x * p.tiles_y);
  p.nsub = nsub;
  p.single_acc = nsub == 2 ? 1 : 0;
  p.acc_stride = (uint32_t)pow2_ceil(bn < 32 ? 32 : bn);
  p.tmem_cols = p.single_acc ? (uint32_t)pow2_ceil(bn) : 2 * p.acc_stride;

  LDX_CHECK_ARG(!halo || tma_ok, "halo conv: the output needs TMA-compatible strides (fp16, 16-byte aligned)");
  int nk = 0;
  for (int s = 0; s < d->nseg; ++s) {
    const ldx_seg_t& sg = d->seg[s];
    LDX_CHECK_ARG(sg.src >= 0 && sg.src < d->nsrc && sg.nchunks >= 1, "bad segment %d", s);
    LDX_CHECK_ARG(sg.c0 % 8 == 0, "segment %d: c0 must be a multiple of 8", s);
    p.seg[s] = KSeg{sg.src, sg.c0, sg.dx, sg.dy, sg.nchunks};
    nk += sg.nchunks;
  }
  p.nseg = d->nseg;
  LDX_CHECK_ARG(d->ldw >= (int64_t)(nk + 9 * halo_chunks) * BK, "weight pitch %lld < K %d", (long long)d->ldw,
                (nk + 9 * halo_chunks) * BK);
  const int k_weights = nk + 9 * halo_chunks;  // K steps that read the weight matrix
  nk += halo_chunks;                           // pipeline items per tile: a halo chunk is one item of nine taps
  p.diag_src = 0;
  // ceil(bn / 64) identity steps: I's rows stop at bn, so the columns past bn of the last step add nothing
  p.diag_chunks = fuse_res ? (bn + 63) / 64 : 0;
  {
    // 64-column identity blocks need the accumulator's TMEM allocation to cover ceil64(bn) columns
    static int d64_env = -1;
    if (d64_env < 0) { const char* e = getenv("LDX_DIAG64"); d64_env = e ? atoi(e) : 1; }
    p.diag64 = (fuse_res && d64_env && (bn + 63) / 64 * 64 <= (int)p.acc_stride) ? 1 : 0;
  }
  nk += p.diag_chunks;
  p.nk = nk;
  p.ksplit = nsplit > 1 ? (nk + nsplit - 1) / nsplit : nk;
  if (nsplit > 1 && (nsplit - 1) * p.ksplit >= nk) {           // every split needs at least one K step
    set_error("split-K: %d splits of %d steps do not cover K = %d steps", nsplit, p.ksplit, nk);
    return LDX_E_INVALID;
  }

  const uint32_t b_stage = (uint32_t)(bn / ctas) * BK * 2;
  g->halo = halo ? 1 : 0;
  if (halo) {
    p.halo_nsrc = d->halo_nsrc;
    p.halo_cp = halo_chunks * BK;
    int coff = 0;
    for (int i = 0; i < d->halo_nsrc; ++i) {
      const ldx_src_t& s = d->src[i];
      p.halo_nch[i] = (s.C + 63) / 64;
      p.halo_coff[i] = coff;
      p.halo_c[i] = s.C;
      coff += p.halo_nch[i] * BK;
      const uint64_t dims[4] = {(uint64_t)s.C, (uint64_t)s.W, (uint64_t)s.H, (uint64_t)s.B};
      const uint64_t str[3] = {(uint64_t)s.sw * 2, (uint64_t)s.sh * 2, (uint64_t)s.sb * 2};
      const uint32_t box[4] = {(uint32_t)BK, 10u, 18u, 1u};
      int rc = encode_map_f16(&p.hmap[i], s.ptr, 4, dims, str, box);
      if (rc) return rc;
    }
    for (int i = d->halo_nsrc; i < 2; ++i) p.hmap[i] = p.hmap[0];
    if (d->halo_nsrc == 0) {                     // normalised 1x1 conv: the transform applies to source 0's K steps
      p.halo_c[0] = d->src[0].C;
      coff = (d->src[0].C + 63) / 64 * BK;
      p.halo_nch[0] = coff / BK;
    }
    p.gn_coef = d->gn_coef;
    p.gn_ctot = d->gn_ctot;
    p.gn_silu = d->gn_silu;
    {
      const int last = d->halo_nsrc > 0 ? d->halo_nsrc - 1 : 0;
      LDX_CHECK_ARG(!d->gn_coef || d->gn_ctot >= coff - (p.halo_nch[last] * BK - p.halo_c[last]),
                    "halo conv: gn_ctot %d smaller than the normalised channels", d->gn_ctot);
    }
  }
  int stages = (int)((224 * 1024 - EPI_SMEM - 2048) / (A_STAGE_BYTES + b_stage));
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  {
    static int st_env = -1;
    if (st_env < 0) { const char* e = getenv("LDX_GEMM_STAGES"); st_env = e ? atoi(e) : 0; }
    if (st_env > 0 && st_env < stages) stages = st_env;
  }
  if (stages > nk + 1) stages = nk + 1 > 2 ? nk + 1 : 2;
  p.stages = stages;
  g->smem_bytes = stages * (A_STAGE_BYTES + (int)b_stage) + EPI_SMEM + (int)sizeof(GemmBarriers) + 1024;
  if (g->smem_bytes < 120 * 1024) g->smem_bytes = 120 * 1024;   // keep one CTA per SM: it owns up to all of TMEM
  if (halo) {
    g->smem_bytes = conv_halo_smem_bytes(bn / ctas, &p.na_stages, &p.nb_stages);
    LDX_CHECK_ARG(g->smem_bytes > 0, "halo conv: bn %d does not fit the operand rings", bn);
  }
  {
    const int workers = sms / ctas;
    g->grid = ctas * (p.num_tiles < workers ? p.num_tiles : workers);
  }

  for (int i = 0; i < d->nsrc; ++i) {
    const ldx_src_t& s = d->src[i];
    LDX_CHECK_ARG(s.ptr && s.C >= 1 && s.W >= 1 && s.H >= 1 && s.B >= 1, "bad source %d", i);
    const uint64_t dims[4] = {(uint64_t)s.C, (uint64_t)s.W, (uint64_t)s.H, (uint64_t)s.B};
    const uint64_t str[3] = {(uint64_t)s.sw * 2, (uint64_t)s.sh * 2, (uint64_t)s.sb * 2};
    const uint32_t box[4] = {(uint32_t)BK, (uint32_t)bw, (uint32_t)bh, (uint32_t)bb};
    int rc = encode_map_f16(&p.amap[i], s.ptr, 4, dims, str, box);
    if (rc) return rc;
  }
  for (int i = d->nsrc; i < LDX_MAX_SRC; ++i) p.amap[i] = p.amap[0];
  p.idmap = p.bmap;
  if (fuse_res) {
    p.diag_src = d->nsrc;
    const uint64_t dims[4] = {(uint64_t)d->N, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t str[3] = {(uint64_t)d->ldr * 2, (uint64_t)d->ldr * 2 * d->W, (uint64_t)d->ldr * 2 * d->W * d->H};
    const uint32_t box[4] = {(uint32_t)BK, (uint32_t)bw, (uint32_t)bh, (uint32_t)bb};
    int rc = encode_map_f16(&p.amap[d->nsrc], d->residual, 4, dims, str, box);
    if (rc) return rc;
    const void* ident = identity256();
    LDX_CHECK_ARG(ident != nullptr, "identity matrix allocation failed");
    const uint64_t idims[2] = {IDENT_N, IDENT_N};
    const uint64_t istr[1] = {IDENT_N * 2};
    const uint32_t ibox[2] = {(uint32_t)BK, p.diag64 ? (uint32_t)(64 / ctas) : (uint32_t)(bn / (ctas * nsub))};
    rc = encode_map_f16(&p.idmap, ident, 2, idims, istr, ibox);
    if (rc) return rc;
  }
  {
    const uint64_t w_rows = d->w_batch_rows ? (uint64_t)d->w_batch_rows * (uint64_t)(d->B - 1) + (uint64_t)d->N : (uint64_t)d->N;
    const uint64_t dims[2] = {(uint64_t)k_weights * BK, w_rows};
    const uint64_t str[1] = {(uint64_t)d->ldw * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)(bn / (ctas * nsub))};
    int rc = encode_map_f16(&p.bmap, d->w, 2, dims, str, box);
    if (rc) return rc;
  }
  if (halo && d->halo_nsrc == 0) p.hmap[0] = p.hmap[1] = p.bmap;   // never dereferenced, but prefetched: keep them valid

  p.out = d->out;
  p.o_sx = d->o_sx; p.o_sy = d->o_sy; p.o_sb = d->o_sb; p.o_off = d->o_off;
  p.bias = d->bias;
  p.rowadd = d->rowadd;
  p.ld_rowadd = d->ld_rowadd;
  p.res = fuse_res ? nullptr : d->residual;    // only the direct-store epilogue still sees a residual pointer
  p.ldr = d->ldr;
  p.flags = d->flags;
  p.split_stride = 0;
  if (nsplit > 1) {
    const int64_t rows = (int64_t)d->W * d->H * d->B;
    p.split_stride = rows * d->N;
    g->ws_bytes = (size_t)nsplit * (size_t)p.split_stride * sizeof(float);
    g->red = SplitReduce{nullptr, nsplit, p.split_stride, (int)rows, d->N, d->W, d->H, d->bias, d->rowadd, d->ld_rowadd,
                         d->residual, d->ldr, (d->flags & LDX_EPI_SILU) ? 1 : ((d->flags & LDX_EPI_QUICK_GELU) ? 2 : ((d->flags & LDX_EPI_LRELU) ? 3 : ((d->flags & LDX_EPI_RELU) ? 4 : 0))), d->out,
                         d->o_sx,
                         d->o_sy > 0 ? d->o_sy : d->o_sx * d->W,
                         d->o_sb > 0 ? d->o_sb : (d->o_sy > 0 ? d->o_sy : d->o_sx * d->W) * d->H, d->o_off};
    p.out = nullptr;                        // bound by gemm_bind_workspace()
    p.o_sx = d->N; p.o_sy = (int64_t)d->N * d->W; p.o_sb = (int64_t)d->N * d->W * d->H; p.o_off = 0;
    p.bias = nullptr; p.rowadd = nullptr; p.res = nullptr;
    p.flags = LDX_EPI_OUT_F32;
  }
  {
    static int verbose = -1;
    if (verbose < 0) { const char* e = getenv("LDX_GEMM_VERBOSE"); verbose = e ? atoi(e) : 0; }
    if (verbose)
      fprintf(stderr, "[ldx_gemm] M=%d(%dx%dx%d) N=%d K=%d -> ctas=%d bn=%d stages=%d tiles=%d grid=%d tma_epi=%d\n",
              d->W * d->H * d->B, d->B, d->H, d->W, d->N, nk * BK, ctas, bn, stages, p.num_tiles, g->grid, p.tma_epi);
  }
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("LDX_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; }
    if (dbg == 1) p.flags |= (1 << 16);
    if (dbg == 2) p.flags |= (1 << 17);
    if (dbg == 3) p.num_tiles = 0;              // timing experiments: launch skeleton only (prologue, TMEM, teardown)
  }
  const int n_out = (d->flags & LDX_EPI_GEGLU) ? d->N / 2 : d->N;
  bool vec = (n_out % 8 == 0) && ((reinterpret_cast<uintptr_t>(d->out) + (uint64_t)d->o_off * osz) % 16 == 0) &&
             ((d->o_sx * osz) % 16 == 0) && ((d->o_sy * osz) % 16 == 0) && ((d->o_sb * osz) % 16 == 0);
  if (d->residual) vec = vec && (reinterpret_cast<uintptr_t>(d->residual) % 16 == 0) && ((d->ldr * 2) % 16 == 0);
  p.vec_ok = vec ? 1 : 0;
  if (nsplit > 1) p.vec_ok = 1;             // fp32 partial rows: N % 8 == 0, 256-byte aligned workspace

  auto fix = [](int64_t s, int64_t fallback) { return (uint64_t)(s > 0 ? s : fallback) * 2; };
  p.tma_epi = 0;
  p.omap = p.bmap;
  p.omap64 = p.bmap;
  if (tma_ok) {
    // one epilogue warp moves 32 of the tile's 128 pixels: a (sw, sh, sb) sub-box of the (bw, bh, bb) tile
    const int sw_ = bw < 32 ? bw : 32, sh_ = bh < 32 / sw_ ? bh : 32 / sw_, sb_ = 32 / (sw_ * sh_);
    const uint32_t box[4] = {(uint32_t)EPI_CHUNK, (uint32_t)sw_, (uint32_t)sh_, (uint32_t)sb_};
    const uint32_t box64[4] = {(uint32_t)EPI_PASS, (uint32_t)sw_, (uint32_t)sh_, (uint32_t)sb_};
    const uint64_t odims[4] = {(uint64_t)n_out, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t sx = (uint64_t)d->o_sx * 2;
    const uint64_t sy = fix(d->o_sy, d->o_sx * d->W);
    const uint64_t sb = fix(d->o_sb, (d->o_sy > 0 ? d->o_sy : d->o_sx * d->W) * d->H);
    const uint64_t ostr[3] = {sx, sy, sb};
    const char* obase = reinterpret_cast<const char*>(d->out) + d->o_off * 2;
    int rc = encode_map_f16_sw(&p.omap, obase, 4, odims, ostr, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    rc = encode_map_f16_sw(&p.omap64, obase, 4, odims, ostr, box64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    p.tma_epi = 1;
    p.rowstats_out = d->rowstats_out;
    LDX_CHECK_ARG(!d->rowstats_out || d->rowstats_out_parts == 0 || d->rowstats_out_parts == p.n_tiles * EPI_GROUPS,
                  "rowstats_out holds %d parts per row but this launch writes %d", d->rowstats_out_parts,
                  p.n_tiles * EPI_GROUPS);
    if (d->flags & LDX_EPI_ROWNORM) {
      LDX_CHECK_ARG(d->rowstats_in && d->ln_gsum && d->rowstats_in_parts >= 1 && d->rowstats_in_parts <= 256 &&
                        d->ln_c >= 1,
                    "LDX_EPI_ROWNORM needs rowstats_in / ln_gsum / parts / ln_c");
      LDX_CHECK_ARG(d->H == 1 && d->B == 1 && !d->rowadd && (reinterpret_cast<uintptr_t>(d->ln_gsum) % 16 == 0) &&
                        d->N % 32 == 0,
                    "LDX_EPI_ROWNORM: plain [M, K] x [N, K] GEMMs only");
      p.rowstats_in = d->rowstats_in;
      p.rs_in_parts = d->rowstats_in_parts;
      p.ln_gsum = d->ln_gsum;
      p.ln_eps = d->ln_eps;
      p.ln_inv_c = 1.0f / (float)d->ln_c;
    }
    if (d->w_batch_rows || d->bias_batch_stride) {
      LDX_CHECK_ARG(d->w_batch_rows >= d->N && d->bias && d->bias_batch_stride >= d->N && d->bias_batch_stride % 4 == 0 &&
                        bb == 1 && (p.tiles_x * p.tiles_y) % ctas == 0 && nsplit == 1 && !halo &&
                        !(d->flags & (LDX_EPI_GEGLU | LDX_EPI_ROWNORM)),
                    "w_batch_rows / bias_batch_stride: need both, tiles inside one batch element (got %d per tile), the plain "
                    "epilogue, no split-K, no halo sources", bb);
      p.w_batch_rows = d->w_batch_rows;
      p.bias_bstride = d->bias_batch_stride;
    }
    if (!p.bias) {
      p.bias = zeros_f32(d->N);
      LDX_CHECK_ARG(p.bias != nullptr, "zero bias allocation failed");
    }
    if (d->colstats_out) {
      const int per = bw * bh;
      LDX_CHECK_ARG(per >= 32 && !(d->flags & LDX_EPI_GEGLU) && !d->rowstats_out,
                    "colstats_out: needs >= 32 pixels per batch element in a tile, no GEGLU, no rowstats_out");
      p.cs_wps = (per < BM ? per : BM) / 32;
      const int slots = p.tiles_x * p.tiles_y * p.cs_wps;
      LDX_CHECK_ARG(d->colstats_slot_base >= 0 && d->colstats_slot_base + slots <= d->colstats_slots,
                    "colstats_out: slots [%d, %d) do not fit the %d slots of the buffer", d->colstats_slot_base,
                    d->colstats_slot_base + slots, d->colstats_slots);
      LDX_CHECK_ARG(reinterpret_cast<uintptr_t>(d->colstats_out) % 16 == 0 && d->N % 2 == 0, "colstats_out: alignment");
      p.colstats_out = d->colstats_out;
      p.cs_slots = d->colstats_slots;
      p.cs_slot_base = d->colstats_slot_base;
    }
  }
  LDX_CHECK_ARG(p.tma_epi || !d->colstats_out, "colstats_out needs the TMA-store epilogue (fp16 output, no split-K)");
  LDX_CHECK_ARG(p.tma_epi || (!d->rowstats_out && !(d->flags & LDX_EPI_ROWNORM)),
                "rowstats_out / LDX_EPI_ROWNORM need the TMA-store epilogue (fp16 output, 16-byte aligned strides)");
  return LDX_OK;
}

template <int kCtas, int kEpi>
static int gemm_launch_t(const GemmLaunch* g, cudaStream_t stream) {
  static std::atomic<uint64_t> attr_done{0};          // per kernel instantiation, one bit per device
  int arc = set_max_dyn_smem_once(ldx_gemm_kernel<kCtas, kEpi>, 227 * 1024, attr_done);
  if (arc) return arc;
  LDX_CUDA(launch_k(ldx_gemm_kernel<kCtas, kEpi>, dim3(g->grid), dim3(GEMM_THREADS), (size_t)g->smem_bytes, stream, kCtas,
                    g->p));
  return LDX_OK;
}

template <int kCtas>
static int gemm_launch_c(const GemmLaunch* g, cudaStream_t stream) {
  const int epi = ((g->p.flags & LDX_EPI_GEGLU) ? KEPI_GEGLU : 0) | ((g->p.flags & LDX_EPI_ROWNORM) ? KEPI_ROWNORM : 0) |
                  (g->p.rowstats_out ? KEPI_ROWSTATS : 0) | (g->p.colstats_out ? KEPI_COLSTATS : 0);
  switch (epi) {
    case KEPI_COLSTATS: return gemm_launch_t<kCtas, KEPI_COLSTATS>(g, stream);
    case 0: return gemm_launch_t<kCtas, 0>(g, stream);
    case KEPI_GEGLU: return gemm_launch_t<kCtas, KEPI_GEGLU>(g, stream);
    case KEPI_ROWNORM: return gemm_launch_t<kCtas, KEPI_ROWNORM>(g, stream);
    case KEPI_ROWSTATS: return gemm_launch_t<kCtas, KEPI_ROWSTATS>(g, stream);
    case KEPI_GEGLU | KEPI_ROWNORM: return gemm_launch_t<kCtas, KEPI_GEGLU | KEPI_ROWNORM>(g, stream);
    case KEPI_ROWNORM | KEPI_ROWSTATS: return gemm_launch_t<kCtas, KEPI_ROWNORM | KEPI_ROWSTATS>(g, stream);
  }
  set_error("unsupported epilogue combination %d (GEGLU cannot write row statistics)", epi);
  return LDX_E_INVALID;
}

void gemm_bind_workspace(GemmLaunch* g, void* ws) {
  g->p.out = ws;
  g->red.ws = reinterpret_cast<const float*>(ws);
}

int gemm_launch(const GemmLaunch* g, cudaStream_t stream) {
  if (g->nsplit > 1 && !g->p.out) {
    set_error("split-K GEMM launched without a workspace");
    return LDX_E_INVALID;
  }
  if (g->halo) return conv_halo_launch(g, stream);
  int rc = g->ctas == 2 ? gemm_launch_c<2>(g, stream) : gemm_launch_c<1>(g, stream);
  if (rc || g->nsplit == 1) return rc;
  const int64_t vecs = (int64_t)g->red.M * (g->red.N >> 2);
  int blocks = (int)((vecs + 255) / 256);
  const int cap = num_sms() * 8;
  if (blocks > cap) blocks = cap;
  LDX_CUDA(launch_k(splitk_reduce_kernel, dim3(blocks), dim3(256), 0, stream, 1, g->red));
  return LDX_OK;
}

// ================================================================================ weight packing kernels
__global__ void pack_conv3x3_kernel(const __half* __restrict__ w, __half* __restrict__ out, int cout, int cin, int cp) {
  const int64_t total = (int64_t)cout * 9 * cp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cp);
    const int t = (int)((i / cp) % 9);
    const int o = (int)(i / ((int64_t)cp * 9));
    out[i] = c < cin ? w[((int64_t)o * cin + c) * 9 + t] : __float2half(0.f);
  }
}

// phase p = py*2+px, tap (a,b): rows R(py,a): py=0: a=0->{0}, a=1->{1,2}; py=1: a=0->{0,1}, a=1->{2}
__global__ void pack_upconv3x3_kernel(const __half* __restrict__ w, __half* __restrict__ out, int cout, int cin,
                                      int cp) {
  const int64_t total = (int64_t)4 * cout * 4 * cp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cp);
    const int tap = (int)((i / cp) % 4);
    const int o = (int)((i / ((int64_t)cp * 4)) % cout);
    const int ph = (int)(i / ((int64_t)cp * 4 * cout));
    float acc = 0.f;
    if (c < cin) {
      const int py = ph >> 1, px = ph & 1, a = tap >> 1, b = tap & 1;
      const int ky0 = py == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2);
      const int ky1 = py == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
      const int kx0 = px == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2);
      const int kx1 = px == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
      for (int ky = ky0; ky <= ky1; ++ky)
        for (int kx = kx0; kx <= kx1; ++kx) acc += __half2float(w[((int64_t)o * cin + c) * 9 + ky * 3 + kx]);
    }
    out[i] = __float2half_rn(acc);
  }
}

__global__ void pack_geglu_kernel(const __half* __restrict__ w, const float* __restrict__ bias,
                                  __half* __restrict__ wout, float* __restrict__ bout, int two_f, int k, int half_bn) {
  const int f = two_f / 2;
  const int64_t total = (int64_t)two_f * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % k);
    const int r = (int)(i / k);                 // packed row
    const int tile = r / (2 * half_bn), within = r % (2 * half_bn);
    const int src = within < half_bn ? tile * half_bn + within : f + tile * half_bn + (within - half_bn);
    wout[i] = w[(int64_t)src * k + kk];
    if (kk == 0 && bias && bout) bout[r] = bias[src];
  }
}

}  // namespace ldx

using namespace ldx;

#ifdef LDX_GEMM_TIMING
extern "C" int ldx_debug_gemm_ts(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_gemm_ts, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int ldx_gemm_rowstats_parts(const ldx_gemm_desc_t* d) {
  GemmLaunch g;
  if (!d) return -1;
  // plan the launch exactly as it will run with rowstats_out attached (a GEMM that writes row statistics is
  // never split-K), whether or not the caller has bound the buffer yet
  ldx_gemm_desc_t q = *d;
  if (!q.rowstats_out) q.rowstats_out = reinterpret_cast<float*>(static_cast<uintptr_t>(16));
  q.rowstats_out_parts = 0;
  if (gemm_prepare(&q, &g) != LDX_OK) return -1;
  if (!g.p.tma_epi) {
    set_error("rowstats_out needs the TMA-store epilogue (fp16 output, 16-byte aligned strides)");
    return -1;
  }
  return g.p.n_tiles * EPI_GROUPS;
}

extern "C" int ldx_gemm_colstats_slots(const ldx_gemm_desc_t* d) {
  GemmLaunch g;
  if (!d) return -1;
  ldx_gemm_desc_t q = *d;                      // plan the launch exactly as it will run with colstats_out attached
  q.colstats_out = reinterpret_cast<float*>(static_cast<uintptr_t>(16));
  q.colstats_slots = 1 << 30;
  q.colstats_slot_base = 0;
  q.rowstats_out = nullptr;
  if (gemm_prepare(&q, &g) != LDX_OK || !g.p.colstats_out) return -1;
  return g.p.tiles_x * g.p.tiles_y * g.p.cs_wps;
}

extern "C" int ldx_gemm(const ldx_gemm_desc_t* d, ldx_stream_t stream) {
  GemmLaunch g;
  int rc = gemm_prepare(d, &g);
  if (rc) return rc;
  if (g.nsplit == 1) return gemm_launch(&g, as_stream(stream));
  void* ws = nullptr;                       // one-shot call: stream-ordered workspace
  LDX_CUDA(cudaMallocAsync(&ws, g.ws_bytes, as_stream(stream)));
  gemm_bind_workspace(&g, ws);
  rc = gemm_launch(&g, as_stream(stream));
  cudaFreeAsync(ws, as_stream(stream));
  return rc;
}

extern "C" int ldx_pack_conv3x3(const void* w, void* out, int cout, int cin, ldx_stream_t stream) {
  LDX_CHECK_ARG(w && out && cout > 0 && cin > 0, "bad pack_conv3x3 args");
  const int cp = (cin + 63) / 64 * 64;
  pack_conv3x3_kernel<<<num_sms() * 4, 256, 0, as_stream(stream)>>>((const __half*)w, (__half*)out, cout, cin, cp);
  LDX_CUDA(cudaGetLastError());
  return LDX_OK;
}

extern "C" int ldx_pack_upconv3x3(const void* w, void* out, int cout, int cin, ldx_stream_t stream) {
  LDX_CHECK_ARG(w && out && cout > 0 && cin > 0, "bad pack_upconv3x3 args");
  const int cp = (cin + 63) / 64 * 64;
  pack_upconv3x3_kernel<<<num_sms() * 4, 256, 0, as_stream(stream)>>>((const __half*)w, (__half*)out, cout, cin, cp);
  LDX_CUDA(cudaGetLastError());
  return LDX_OK;
}

extern "C" int ldx_pack_geglu(const void* w, const float* bias, void* wout, float* bout, int two_f, int k, int half_bn,
                              ldx_stream_t stream) {
  LDX_CHECK_ARG(w && wout && two_f > 0 && k > 0 && half_bn > 0 && (two_f / 2) % half_bn == 0, "bad pack_geglu args");
  pack_geglu_kernel<<<num_sms() * 4, 256, 0, as_stream(stream)>>>((const __half*)w, bias, (__half*)wout, bout, two_f, k,
                                                                 half_bn);
  LDX_CUDA(cudaGetLastError());
  return LDX_OK;
}
