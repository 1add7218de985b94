// Flash-style attention on tcgen05 / TMEM (sm_100a) for optimized_attention (LD.py:3966-3988):
//   o[b, i, h*d:(h+1)*d] = softmax_j( q_i . k_j / sqrt(d) ) v_j      (no mask, dropout 0)
// with head dims 40 / 80 / 160 (SD1.5: 8 heads at 320 / 640 / 1280 channels) plus 64 / 128.
//
// One CTA = NG groups of 128 query rows of one (batch, head); keys/values are streamed ONCE per CTA in
// tiles of 64 and shared by all groups (for 128-row CTAs the K/V re-reads from L2 were the bound at
// d = 40; NG = 4 cuts them 4x).
//   warp 0        : TMA producer (Q per group, K/V rings)
//   warp 1        : V patcher (row sums on the tensor core, see AttnCfg::ONES)
//   warps 2..NG+1 : tcgen05.mma issuers, one per group (S_g = Q_g K^T into TMEM, O_g += P_g V); the groups
//                   run independently apart from the shared K/V ring, so one group's exponentials overlap
//                   another's TMEM loads / P stores / barrier round trips
//   warps NG+2..  : online softmax, 4 warps per group, one thread per query row, at HALF-tile (32 key)
//                   granularity: S from TMEM -> exp2 -> P (fp16) into smem in the K-major SWIZZLE_128B
//                   layout tcgen05.mma reads; lazy rescale of O in TMEM
// Head slices are addressed by a 4-D TMA view (d, heads, N, B) of the [B, N, heads*d] tensor, so the
// box's columns beyond d are out of bounds and arrive as zeros (K padding for d = 40).
// V is consumed straight from its row-major tile through an MN-major smem descriptor.
// Roofline: 4*Nq*Nk*d*heads FLOPs per batch element against the tensor peak.  For d <= 80 the practical
// bound is per score, not per FLOP: each of the Nq*Nk scores costs one MUFU.EX2 (16/clk/SM, measured by
// scripts/ubench/tmem_bw.cu) plus ~4-6 issue slots of fp32 work (scale/max/pack/convert), and the four
// schedulers of an SM run at 60-70 % issue utilisation in the ncu capture (profiles/).  At d = 40 the MUFU
// floor alone is ~480 us for the 16 x 8 x 4096^2 SD1.5 shape; the kernel runs at ~730 us.  Evaluating a
// share of the exponentials on the FMA pipe (POLY_MOD) adds instructions and measured no gain.
#include "common.cuh"
#include "internal.h"

namespace ldx {

int encode_map_f16(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

static constexpr int BQ = 128;
static constexpr int BKV = 64;
static constexpr int ATT_MAX_NG = 4;
static constexpr int ATT_MAX_ST = 6;        // K / V ring depth bound
#ifndef LDX_ATT_DBG
#define LDX_ATT_DBG 0           // timing experiments only: 1 = no exponentials
#endif
#ifndef LDX_ATT_NG_SMALL
#define LDX_ATT_NG_SMALL 4      // query groups per CTA for d <= 64
#endif
#ifndef LDX_ATT_ONES
#define LDX_ATT_ONES 1           // row sums on the tensor core where P V has a spare column (d = 40): ~4 %
#endif
#ifndef LDX_ATT_SLEEP_NS
#define LDX_ATT_SLEEP_NS 64       // poll back-off of the producer / patcher / MMA-issuer warps
#endif
// The four softmax warps that share an SM sub-partition (one per query group) contend for its MUFU unit: started in
// phase they all run their 64 exponentials together (4 x 512 clocks of MUFU work) and then all do their MUFU-free work
// (TMEM load, row maximum, barriers: ~1.2 k clocks) together, leaving the unit idle.  With LDX_ATT_TOKENS = T > 0 the
// exponentials of (tile j, group g) may start only when those of the T-th previous (tile, group) in round-robin order
// have been issued, which staggers the groups so that at most T warps per sub-partition are in their exponentials.
// Measured (d = 40, 16 x 8 x 4096^2, scripts/attn_timeline.py on the -DLDX_ATT_TIMING build): a softmax warp spends
// 2.74 k clocks per 64-key tile -- 1.54 k in the exponential / pack / store phase (0.51 k of MUFU work of its own, the
// rest queueing behind its three neighbours), 1.2 k in TMEM load, row maximum and barriers.  T = 2 shortens the
// exponential phase to 1.24 k but adds 0.34 k of token wait: the same 2.74 k, 700 vs 675 us; T = 1 serialises (898 us).
// So the stagger is not what the kernel lacks: per tile round a CTA also moves ~256 KB through shared memory (Q re-read
// by every S = Q K^T, K / V / P operand reads, P stores) against 128 B/clk = 2.0 k of those 2.74 k clocks, MUFU is 75 %
// busy and the schedulers 52 %: three resources within a third of each other.  Kept as a knob, default off.
#ifndef LDX_ATT_TOKENS
#define LDX_ATT_TOKENS 0
#endif
static_assert(LDX_ATT_TOKENS >= 0 && LDX_ATT_TOKENS <= 2, "with three tokens a group can lap its successor's barrier phase");
// Bit i set: pair i of each 32-key half tile (the NG = 4 kernels, d <= 64) takes its two exponentials from the FMA pipe as packed
// fp32 (ex2_poly2, 5 issue slots per exponential) instead of MUFU.EX2 (the unit this kernel queues on: 75 % busy).  Which
// pairs matters as much as how many -- the mask changes ptxas's schedule of the whole tile: measured at 16 x 8 x 4096^2,
// d = 40 (scripts/gpu_poly_sweep.sh, one box, us): 0 -> 643; 0x4144 (4 of 16 pairs) -> 609; 0x4244 610; 0x4124 613; 0x4924
// (5 pairs) 620; 0x0144 (3 pairs) 632; 0x4344 643; 0x2492 662; 0x4140 697; 0xAAAA (8 pairs) 746.
// OFF by default: the 5 % it gains on the kernel timed alone disappears in the pipeline, which runs at the 1 kW power cap --
// the extra FMA-pipe instructions cost energy and the SM clock drops from 1750 to 1720 MHz (scripts/gpu_ab9.sh, same box,
// bench conditions: 24.33 / 24.34 images/s off, 24.31 / 24.29 on; UNet graph replay 14.73 vs 14.75 ms).
#ifndef LDX_ATT_POLY_MASK
#define LDX_ATT_POLY_MASK 0
#endif
#ifndef LDX_ATT_POLY_MOD
#define LDX_ATT_POLY_MOD 0      // every POLY_MOD-th exponential runs on the FMA pipe (0 = all on MUFU); measured: no gain
#endif

#ifdef LDX_ATT_TIMING
// variant-build instrumentation: cycles per phase of the first softmax warp of CTA (0, 0, 0), summed over its K/V tiles
__device__ unsigned long long g_att_ts[16];
#define AT_DECL long long at_t = clock64(), at_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define AT(i) { const long long at_now = clock64(); at_acc[i] += at_now - at_t; at_t = at_now; }
#define AT_FLUSH(n) if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == C::HELPERS && lane == 0) { for (int i = 0; i < 8; ++i) g_att_ts[i] = (unsigned long long)at_acc[i]; g_att_ts[8] = (unsigned long long)(n); }
#else
#define AT_DECL
#define AT(i)
#define AT_FLUSH(n)
#endif

template <int D, int NG_>
struct AttnCfg {
  static constexpr int DK = (D + 15) / 16;          // 16-wide K steps of Q K^T
  static constexpr int DCH = (D + 63) / 64;         // 64-column smem chunks per row
  static constexpr int DN = DK * 16;                // UMMA N of P V
  static constexpr int NG = NG_;                    // 128-row query groups per CTA
  // Register rebalancing (setmaxnreg) for the multi-group configuration: the helper warps (producer, V patcher, NG MMA
  // issuers, padded to two whole warpgroups) drop to 32 registers and the softmax warpgroups rise from the launch-time
  // 80 to 104, so that a softmax thread holds a whole 64-key S row: one TMEM round trip per tile, S released to the
  // next Q K^T before any exponential, one running-max test per tile, no spills.
#ifndef LDX_ATT_REBALANCE
#define LDX_ATT_REBALANCE 1
#endif
#ifndef LDX_ATT_HELPER_REGS
#define LDX_ATT_HELPER_REGS 40     // 256 * helper + 512 * softmax <= 768 * 80 (the launch-time allocation)
#define LDX_ATT_SOFTMAX_REGS 96
#endif
  static constexpr bool REBAL = (NG_ > 1) && (LDX_ATT_REBALANCE != 0);
#ifndef LDX_ATT_WHOLE_NG1
#define LDX_ATT_WHOLE_NG1 1      // single-group configurations too: 87.6 -> 80.2 us at d = 80, 18.6 -> 16.4 at d = 160
#endif
  static constexpr bool WHOLE = REBAL || (NG_ == 1 && LDX_ATT_WHOLE_NG1 != 0);   // whole-tile softmax variant
  static constexpr int HELPERS = REBAL ? 8 : 2 + NG_;     // warps before the first softmax warp
  static constexpr int THREADS = 32 * (HELPERS + 4 * NG);  // producer, V patcher, NG MMA issuers, 4 softmax warps / group
  // Row sums on the tensor core: when P V has a spare accumulator column (DN > D) the V tile's column D is
  // set to 1.0 after it lands, so O[:, D] accumulates sum_j P_ij and the softmax warps skip 64 FADDs a tile.
  static constexpr bool ONES = (DN > D) && (LDX_ATT_ONES != 0);
  // single-group CTAs (short K/V: cross-attention) stay small so that several fit on an SM
#ifndef LDX_ATT_KST_NG
#define LDX_ATT_KST_NG 2        // K / V ring depth and P buffers of the multi-group (d <= 64) configuration
#define LDX_ATT_VST_NG 2
#define LDX_ATT_PB_NG 2
#endif
  static constexpr int KST = (DCH == 1 && NG > 1) ? LDX_ATT_KST_NG : ((DCH == 1 && NG == 1) ? 1 : 2);
  static constexpr int VST = (DCH == 1 && NG > 1) ? LDX_ATT_VST_NG : 1;
  static constexpr int Q_BYTES = DCH * BQ * 128;    // per group
  static constexpr int KV_BYTES = DCH * BKV * 128;
  static constexpr int PB = (DCH == 1 && NG > 1) ? LDX_ATT_PB_NG : 1;   // P tiles per group: with one, P(j) waits for P V(j-1) to drain
  static constexpr int P_BYTES = BQ * BKV * 2;      // per group and buffer
  static constexpr int TMEM_NEED = NG * (BKV + DN);
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
  static constexpr int SMEM = NG * (Q_BYTES + PB * P_BYTES) + (KST + VST) * KV_BYTES + 768 + 1024;
  static constexpr int BY_TMEM = 512 / TMEM_COLS;
  static constexpr int BY_SMEM = (227 * 1024) / (SMEM + 1024);
  static constexpr int BY_THREADS = 2048 / THREADS;
  static constexpr int B0 = BY_TMEM < BY_SMEM ? BY_TMEM : BY_SMEM;
  static constexpr int BY_REGS = 65536 / (80 * ((THREADS + 127) / 128 * 128));   // at least 80 registers per thread
  static constexpr int B1a = B0 < BY_THREADS ? B0 : BY_THREADS;
  static constexpr int B1 = B1a < BY_REGS ? B1a : BY_REGS;
  static constexpr int MIN_BLOCKS = B1 < 1 ? 1 : B1;
  static_assert(TMEM_NEED <= 512, "TMEM budget");
  static_assert(SMEM <= 227 * 1024, "smem budget");
  static_assert(KST <= ATT_MAX_ST && VST <= ATT_MAX_ST, "ring depth");
};

struct AttnBarriers {           // must fit the 768 bytes reserved in AttnCfg::SMEM
  uint64_t q_full[ATT_MAX_NG], s_full[ATT_MAX_NG], s_free[ATT_MAX_NG], p_full[ATT_MAX_NG][2], pv_done[ATT_MAX_NG][2];
  uint64_t k_full[ATT_MAX_ST], k_empty[ATT_MAX_ST], v_full[ATT_MAX_ST], v_empty[ATT_MAX_ST], v_ready[ATT_MAX_ST];
  uint64_t tok[4][ATT_MAX_NG];  // MUFU turn-taking among the softmax warps that share a scheduler (see LDX_ATT_TOKENS)
  uint32_t tmem_base;
};
static_assert(sizeof(AttnBarriers) <= 768, "AttnBarriers must fit its reserved shared memory");

struct alignas(64) AttnParams {
  CUtensorMap qmap, kmap, vmap;
  __half* o;
  int64_t ldo;
  int32_t B, Nq, Nk, heads;
  float scale_log2;   // log2(e) / sqrt(d)
  int32_t causal;
};

// 2^x for x <= 0 on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax polynomial, max relative
// error 7.5e-5 -- below fp16 rounding of P), so a share of the exponentials does not queue on MUFU.
LDX_DEVICE float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = __fadd_rd(x, 12582912.0f);            // floor(x) in the low mantissa bits
  const float fr = x - (t - 12582912.0f);               // [0, 1)
  float p = fmaf(0.07802422344684601f, fr, 0.22606723010540009f);
  p = fmaf(p, fr, 0.695833683013916f);
  p = fmaf(p, fr, 0.9999251961708069f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// The same on a pair of values with packed fp32 (FFMA2): 2 FMNMX + 6 FFMA2 + 2 LEA per pair = 5 issue slots per exponential
// (10 for the scalar form) against one issue slot + 8 MUFU-pipe clocks per warp for ex2.approx.
LDX_DEVICE void ex2_poly2(float x0, float x1, float& r0, float& r1) {
  x0 = fmaxf(x0, -125.0f);
  x1 = fmaxf(x1, -125.0f);
  const f32x2_t x = f2_pack(x0, x1), one = f2_pack(1.0f, 1.0f);
  f32x2_t t;                                              // x + 1.5 * 2^23 rounded down: floor(x) in the low mantissa bits
  asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(x), "l"(one), "l"(f2_pack(12582912.0f, 12582912.0f)));
  const f32x2_t fl = f2_fma(t, one, f2_pack(-12582912.0f, -12582912.0f));       // floor(x), exact
  const f32x2_t fr = f2_fma(fl, f2_pack(-1.0f, -1.0f), x);                      // [0, 1)
  f32x2_t p = f2_fma(f2_pack(0.07802422344684601f, 0.07802422344684601f), fr, f2_pack(0.22606723010540009f, 0.22606723010540009f));
  p = f2_fma(p, fr, f2_pack(0.695833683013916f, 0.695833683013916f));
  p = f2_fma(p, fr, f2_pack(0.9999251961708069f, 0.9999251961708069f));
  float p0, p1, t0, t1;
  f2_unpack(p, p0, p1);
  f2_unpack(t, t0, t1);
  r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// CAUSAL (the CLIP text tower's mask, LD.py:4443-4449: -inf above the diagonal) is a compile-time variant so
// that the unmasked kernels keep their instruction stream; key j contributes to query i only when j <= i.
template <int D, int NG_, bool CAUSAL = false>
__global__ void __launch_bounds__(AttnCfg<D, NG_>::THREADS, AttnCfg<D, NG_>::MIN_BLOCKS) ldx_attention_kernel(
    const __grid_constant__ AttnParams p) {
  using C = AttnCfg<D, NG_>;
  constexpr int NG = C::NG;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;                                 // NG x Q_BYTES
  uint8_t* sK = sQ + NG * C::Q_BYTES;
  uint8_t* sV = sK + C::KST * C::KV_BYTES;
  uint8_t* sP = sV + C::VST * C::KV_BYTES;            // NG x PB x P_BYTES
  AttnBarriers* bars = reinterpret_cast<AttnBarriers*>(sP + NG * C::PB * C::P_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_cta = blockIdx.x * (NG * BQ), h = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.Nk + BKV - 1) / BKV;
  const int rows_left = p.Nq - q_cta;
  const int ng_act = rows_left >= NG * BQ ? NG : (rows_left + BQ - 1) / BQ;   // groups with at least one query row

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.qmap);
    tma_prefetch_desc(&p.kmap);
    tma_prefetch_desc(&p.vmap);
    for (int g = 0; g < ATT_MAX_NG; ++g) {
      mbar_init(&bars->q_full[g], 1);
      mbar_init(&bars->s_full[g], 1);
      mbar_init(&bars->s_free[g], 4);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars->p_full[g][i], 4);
        mbar_init(&bars->pv_done[g][i], 1);
      }
      for (int qq = 0; qq < 4; ++qq) mbar_init(&bars->tok[qq][g], 1);
    }
    for (int i = 0; i < ATT_MAX_ST; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->k_empty[i], ng_act);    // one tcgen05.commit per active group
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], ng_act);
      mbar_init(&bars->v_ready[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&bars->tmem_base, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t tS = tmem_base;                      // group g: 64 fp32 columns at g * 64
  const uint32_t tO = tmem_base + NG * BKV;           // group g: DN fp32 columns at g * DN
  pdl_prologue();

  // With register rebalancing the roles split at warpgroup granularity first, so that each setmaxnreg dominates the
  // code of its side (ptxas budgets registers per region): warps 0-7 helpers, 8.. softmax.
  if (warp < C::HELPERS) {
  if constexpr (C::REBAL) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LDX_ATT_HELPER_REGS));
  if (warp == 0) {
    // ---- TMA producer: warp-uniform loop, one elected lane issues.  K runs one tile ahead of V: the K
    // buffer frees as soon as every group's S = Q K^T has been issued, the V buffer only after the softmax.
    if (elect_one()) {
      for (int g = 0; g < ng_act; ++g) {
        mbar_arrive_expect_tx(&bars->q_full[g], C::Q_BYTES);
        for (int c = 0; c < C::DCH; ++c)
          tma_load_4d(sQ + g * C::Q_BYTES + c * (BQ * 128), &p.qmap, &bars->q_full[g], c * 64, h, q_cta + g * BQ, b);
      }
    }
    __syncwarp();
    auto load_k = [&](int t) {
      const int ks = t % C::KST;
      mbar_wait_relaxed(&bars->k_empty[ks], ((t / C::KST) & 1) ^ 1, LDX_ATT_SLEEP_NS);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->k_full[ks], C::KV_BYTES);
        for (int c = 0; c < C::DCH; ++c)
          tma_load_4d(sK + ks * C::KV_BYTES + c * (BKV * 128), &p.kmap, &bars->k_full[ks], c * 64, h, t * BKV, b);
      }
      __syncwarp();
    };
    for (int t = 0; t < C::KST - 1 && t < nkv; ++t) load_k(t);
    for (int t = 0; t < nkv; ++t) {
      if (t + C::KST - 1 < nkv) load_k(t + C::KST - 1);
      const int vs = t % C::VST;
      mbar_wait_relaxed(&bars->v_empty[vs], ((t / C::VST) & 1) ^ 1, LDX_ATT_SLEEP_NS);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->v_full[vs], C::KV_BYTES);
        for (int c = 0; c < C::DCH; ++c)
          tma_load_4d(sV + vs * C::KV_BYTES + c * (BKV * 128), &p.vmap, &bars->v_full[vs], c * 64, h, t * BKV, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- V patcher: column D of every landed V tile becomes 1.0 (see AttnCfg::ONES)
    if constexpr (C::ONES) {
      const __half one = __float2half(1.0f);
      for (int t = 0; t < nkv; ++t) {
        const int vs = t % C::VST;
        mbar_wait_relaxed(&bars->v_full[vs], (t / C::VST) & 1, LDX_ATT_SLEEP_NS);
        uint8_t* tile = sV + vs * C::KV_BYTES + (D / 64) * (BKV * 128);
#pragma unroll
        for (int rr = 0; rr < BKV; rr += 32) {
          const int r = rr + lane;
          *reinterpret_cast<__half*>(tile + r * 128 + ((((D % 64) / 8) ^ (r & 7)) << 4) + (D % 8) * 2) = one;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->v_ready[vs]);
      }
    }
  } else if (warp < 2 + NG) {
    // ---- MMA issuers, one warp per query group (warp-uniform control flow, one elected lane issues
    // tcgen05.mma / commit).  S_g(t+1) is issued as soon as the softmax has read S_g(t) out of TMEM, so it
    // overlaps the exponentials of tile t; a K/V ring slot is released by the LAST group's commit (the
    // slot's "empty" barrier counts one tcgen05.commit per active group).
    const int g = warp - 2;
    if (g < ng_act) {
      const uint32_t idesc_s = umma_idesc_f16(BQ, BKV, 0, 0);
      const uint32_t idesc_o = umma_idesc_f16(BQ, C::DN, 0, 1);       // B = V is MN-major
      const uint32_t q_addr = smem_u32(sQ + g * C::Q_BYTES), p_addr = smem_u32(sP + g * C::PB * C::P_BYTES);
      const uint32_t tSg = tS + g * BKV, tOg = tO + g * C::DN;
      auto mma_s = [&](int t) {
        const int ks = t % C::KST;
        mbar_wait_relaxed(&bars->k_full[ks], (t / C::KST) & 1, LDX_ATT_SLEEP_NS);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + ks * C::KV_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < C::DK; ++s) {
            const uint32_t qa = q_addr + (s / 4) * (BQ * 128) + (s % 4) * 32;
            const uint32_t ka = k_addr + (s / 4) * (BKV * 128) + (s % 4) * 32;
            umma_f16_ss(tSg, umma_desc_kmajor_sw128(qa), umma_desc_kmajor_sw128(ka), idesc_s, (uint32_t)(s != 0));
          }
          umma_commit(&bars->s_full[g]);
          umma_commit(&bars->k_empty[ks]);
        }
        __syncwarp();
      };
      mbar_wait(&bars->q_full[g], 0);
      mma_s(0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) {
          mbar_wait_relaxed(&bars->s_free[g], j & 1, LDX_ATT_SLEEP_NS);
          tc_fence_after();
          mma_s(j + 1);
        }
        const int vs = j % C::VST, pb = j % C::PB;
        mbar_wait_relaxed(&bars->p_full[g][pb], (j / C::PB) & 1, LDX_ATT_SLEEP_NS);
        mbar_wait_relaxed(C::ONES ? &bars->v_ready[vs] : &bars->v_full[vs], (j / C::VST) & 1, LDX_ATT_SLEEP_NS);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV + vs * C::KV_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < BKV / 16; ++s) {
            umma_f16_ss(tOg, umma_desc_kmajor_sw128(p_addr + pb * C::P_BYTES + s * 32),
                        umma_desc_mnmajor_sw128(v_addr + s * 2048, BKV * 128), idesc_o, (uint32_t)((j | s) != 0));
          }
          umma_commit(&bars->pv_done[g][pb]);
          umma_commit(&bars->v_empty[vs]);
        }
        __syncwarp();
      }
    }
  }
  } else {
  if constexpr (C::REBAL) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(LDX_ATT_SOFTMAX_REGS));
  if ((warp - C::HELPERS) / 4 < ng_act) {
    // ------------------------------------------------------------------ softmax warps (thread = query row)
    // Online softmax, one 64-key tile at a time with the whole S row in registers (WHOLE, the default everywhere).
    // The HALF-tile (32 key) variant below it is the fallback for LDX_ATT_REBALANCE=0 / LDX_ATT_WHOLE_NG1=0: it
    // keeps 32 live score registers and fits the 80-register cap of the multi-group configuration without setmaxnreg.
    const int g = (warp - C::HELPERS) / 4;
    const int qd = warp & 3;                    // TMEM lane quarter this warp may access
    const int row = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tSg = tS + g * BKV + lane_off, tOg = tO + g * C::DN + lane_off;
    uint64_t* s_full = &bars->s_full[g];
    uint64_t* s_free = &bars->s_free[g];
    uint64_t* p_full = &bars->p_full[g][0];
    uint64_t* pv_done = &bars->pv_done[g][0];
    uint8_t* prow0 = sP + g * C::PB * C::P_BYTES + row * 128;
    const float c = p.scale_log2;
    float m_used = -INFINITY;   // row max the exponentials are relative to (raw score units)
    float l = 0.f;
    if constexpr (C::WHOLE) {
      // ---- whole-tile variant: S row in registers, S released at once, one max test per tile.  With PREFETCH the
      // first half of S(j + 1) is requested from TMEM in the middle of tile j's exponentials (its registers are dead by
      // then) and the second half right after P(j) is published, so most of the TMEM round trip is off the chain.
#ifndef LDX_ATT_PREFETCH
#define LDX_ATT_PREFETCH 0       // measured: 705 vs 675 us at d = 40 (the mid-tile wait for S(j + 1) stalls the in-order warp)
#endif
      constexpr bool PF = LDX_ATT_PREFETCH != 0;
      // barrier addresses in the shared window, computed once (see smem_u32_pinned)
      const uint32_t a_sfull = smem_u32_pinned(s_full), a_sfree = smem_u32_pinned(s_free);
      const uint32_t a_pfull = smem_u32_pinned(p_full), a_pvdone = smem_u32_pinned(pv_done);
      uint32_t ra[32], rb[32];
      auto release_s = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(a_sfree);   // S is in registers: the accumulator may be reused
      };
      if constexpr (PF) {
        mbar_wait_a(a_sfull, 0);
        tc_fence_after();
        tmem_ld32(tSg, ra);
        tmem_ld32(tSg + 32, rb);
        tmem_ld_wait();
        release_s();
      }
      AT_DECL
      for (int j = 0; j < nkv; ++j) {
        const int pb = j % C::PB;
        uint8_t* prow = prow0 + pb * C::P_BYTES;
        const int kv_left = p.Nk - j * BKV;
        AT(7)
        if constexpr (!PF) {
          mbar_wait_a(a_sfull, j & 1);
          AT(0)
          tc_fence_after();
          tmem_ld32(tSg, ra);
          tmem_ld32(tSg + 32, rb);
          tmem_ld_wait();
          release_s();
          AT(1)
        }
        const uint32_t ninf = 0xff800000u;
        if (kv_left < BKV) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i >= kv_left) ra[i] = ninf;
            if (32 + i >= kv_left) rb[i] = ninf;
          }
        }
        if constexpr (CAUSAL) {
          const int k0 = j * BKV, qi = q_cta + g * BQ + row;
          if (k0 + 63 > qi) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (k0 + i > qi) ra[i] = ninf;
              if (k0 + 32 + i > qi) rb[i] = ninf;
            }
          }
        }
        float mx0 = fmaxf(__uint_as_float(ra[0]), __uint_as_float(ra[1]));
        float mx1 = fmaxf(__uint_as_float(ra[2]), __uint_as_float(ra[3]));
#pragma unroll
        for (int i = 4; i < 32; i += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(ra[i]), __uint_as_float(ra[i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(ra[i + 2]), __uint_as_float(ra[i + 3])));
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(rb[i]), __uint_as_float(rb[i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(rb[i + 2]), __uint_as_float(rb[i + 3])));
        }
        const float m_t = fmaxf(mx0, mx1);
        const bool grow = (m_t - m_used) * c > 8.0f;
        if (j == 0) {
          m_used = m_t;
        } else if (__any_sync(0xffffffffu, grow)) {
          const float m_new = fmaxf(m_used, m_t);
          const float alpha = ex2_approx((m_used - m_new) * c);
          m_used = m_new;
          if constexpr (!C::ONES) l *= alpha;
          mbar_wait_a(a_pvdone + 8 * ((j - 1) % C::PB), ((j - 1) / C::PB) & 1);     // O is quiescent
          tc_fence_after();
          // four columns at a time: the 64 live score registers leave little room in this (rare) path
#pragma unroll 1
          for (int cc = 0; cc < C::DN / 4; ++cc) {
            uint32_t o0, o1, o2, o3;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3) : "r"(tOg + cc * 4) : "memory");
            tmem_ld_wait();
            o0 = __float_as_uint(__uint_as_float(o0) * alpha);
            o1 = __float_as_uint(__uint_as_float(o1) * alpha);
            o2 = __float_as_uint(__uint_as_float(o2) * alpha);
            o3 = __float_as_uint(__uint_as_float(o3) * alpha);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%4], {%0, %1, %2, %3};"
                         :: "r"(o0), "r"(o1), "r"(o2), "r"(o3), "r"(tOg + cc * 4) : "memory");
          }
          tmem_st_wait();
        }
        const float mc = m_used * c;
        AT(2)
        if (j >= C::PB) mbar_wait_a(a_pvdone + 8 * pb, ((j - C::PB) / C::PB) & 1);   // P V(j - PB) has read this buffer
        AT(3)
        if constexpr (LDX_ATT_TOKENS > 0 && NG > 1) {
          if (ng_act == NG && (j > 0 || g >= LDX_ATT_TOKENS)) {
            const int wcount = g >= LDX_ATT_TOKENS ? j : j - 1;
            mbar_wait(&bars->tok[qd][g], (uint32_t)(wcount & 1));
          }
        }
        AT(6)
        float l0 = 0.f, l1 = 0.f;
        // 64 keys = 8 16-byte chunks; chunk index XOR (row % 8) = SWIZZLE_128B
        auto half = [&](const uint32_t (&r)[32], int ch0) {
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int e = ch * 8 + 2 * i;
              float x0, x1;                     // one FFMA2 for the pair (bit-identical to two fmaf)
              f2_unpack(f2_fma(f2_pack(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), f2_pack(c, c), f2_pack(-mc, -mc)), x0, x1);
              float p0, p1;
              if (LDX_ATT_DBG == 1) {
                p0 = x0;
                p1 = x1;
              } else if ((LDX_ATT_POLY_MASK >> (ch * 4 + i)) & 1) {
                ex2_poly2(x0, x1, p0, p1);
              } else {
                p0 = (LDX_ATT_POLY_MOD > 0 && e % (LDX_ATT_POLY_MOD > 0 ? LDX_ATT_POLY_MOD : 1) == LDX_ATT_POLY_MOD - 1)
                         ? ex2_poly(x0) : ex2_approx(x0);
                p1 = (LDX_ATT_POLY_MOD > 0 && (e + 1) % (LDX_ATT_POLY_MOD > 0 ? LDX_ATT_POLY_MOD : 1) == LDX_ATT_POLY_MOD - 1)
                         ? ex2_poly(x1) : ex2_approx(x1);
              }
              if constexpr (!C::ONES)
                f2_unpack(f2_fma(f2_pack(p0, p1), f2_pack(1.0f, 1.0f), f2_pack(l0, l1)), l0, l1);   // l0 += p0, l1 += p1
              pk[i] = pack_h2(p0, p1);
            }
            *reinterpret_cast<uint4*>(prow + (((ch0 + ch) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        };
        half(ra, 0);
        const bool more = j + 1 < nkv;
        if constexpr (PF) {
          if (more) {                            // S(j + 1) was issued when S(j) was released: normally long complete
            mbar_wait_a(a_sfull, (j + 1) & 1);
            tc_fence_after();
            tmem_ld32(tSg, ra);
          }
        }
        half(rb, 4);
        if constexpr (LDX_ATT_TOKENS > 0 && NG > 1) {
          if (ng_act == NG && lane == 0) mbar_arrive(&bars->tok[qd][(g + LDX_ATT_TOKENS) % NG]);
        }
        l += l0 + l1;
        AT(4)
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(a_pfull + 8 * pb);
        AT(5)
        if constexpr (PF) {
          if (more) {
            tmem_ld32(tSg + 32, rb);
            tmem_ld_wait();
            release_s();
          }
        }
      }
      AT_FLUSH(nkv)
    } else
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int pb = j % C::PB;
      uint8_t* prow = prow0 + pb * C::P_BYTES;
      const int kv_left = p.Nk - j * BKV;      // valid keys in this tile (>= 1)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t r[32];
        tmem_ld32(tSg + hf * 32, r);
        tmem_ld_wait();
        if (hf == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free);   // S(j) has been read: the accumulator may be reused
        }
        float s[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(r[i]);
        if (kv_left < BKV) {                     // only the last tile of a ragged Nk needs key masking
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (hf * 32 + i >= kv_left) s[i] = -INFINITY;
        }
        if constexpr (CAUSAL) {
          const int k0 = j * BKV + hf * 32, qi = q_cta + g * BQ + row;
          if (k0 + 31 > qi) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (k0 + i > qi) s[i] = -INFINITY;
          }
        }
        float mx0 = fmaxf(s[0], s[1]), mx1 = fmaxf(s[2], s[3]);
#pragma unroll
        for (int i = 4; i < 32; i += 4) {
          mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 1]));
          mx1 = fmaxf(mx1, fmaxf(s[i + 2], s[i + 3]));
        }
        const float m_h = fmaxf(mx0, mx1);
        // lazy rescale: only when some row of this warp grew its max by more than 2^8 after scaling
        const bool grow = (m_h - m_used) * c > 8.0f;
        if (j == 0 && hf == 0) {
          m_used = m_h;
        } else if (__any_sync(0xffffffffu, grow)) {
          const float m_new = fmaxf(m_used, m_h);
          const float alpha = ex2_approx((m_used - m_new) * c);
          m_used = m_new;
          if constexpr (!C::ONES) l *= alpha;
          if (j > 0) {
            mbar_wait(pv_done + (j - 1) % C::PB, ((j - 1) / C::PB) & 1);     // O is quiescent
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < C::DN / 16; ++cc) {
              uint32_t o[16];
              tmem_ld16(tOg + cc * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              asm volatile(
                  "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
                  "%12, %13, %14, %15};" ::"r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]),
                  "r"(o[7]), "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]),
                  "r"(o[15]), "r"(tOg + cc * 16)
                  : "memory");
            }
            tmem_st_wait();
          }
          if (hf == 1) {                          // the first half of this tile's P row is already in smem
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              uint4* q4 = reinterpret_cast<uint4*>(prow + ((ch ^ (row & 7)) << 4));
              uint4 u = *q4;
              float2 a0 = unpack_h2(u.x), a1 = unpack_h2(u.y), a2 = unpack_h2(u.z), a3 = unpack_h2(u.w);
              u.x = pack_h2(a0.x * alpha, a0.y * alpha);
              u.y = pack_h2(a1.x * alpha, a1.y * alpha);
              u.z = pack_h2(a2.x * alpha, a2.y * alpha);
              u.w = pack_h2(a3.x * alpha, a3.y * alpha);
              *q4 = u;
            }
          }
        }
        const float mc = m_used * c;
        if (hf == 0 && j >= C::PB) mbar_wait(pv_done + pb, ((j - C::PB) / C::PB) & 1);   // P V(j - PB) has read this buffer
        float l0 = 0.f, l1 = 0.f;
        // 32 keys = 4 16-byte chunks; chunk index XOR (row % 8) is the SWIZZLE_128B pattern
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = ch * 8 + 2 * i;
            float x0, x1;
            f2_unpack(f2_fma(f2_pack(s[e], s[e + 1]), f2_pack(c, c), f2_pack(-mc, -mc)), x0, x1);
            float p0, p1;
            if (LDX_ATT_DBG == 1) {
              p0 = x0;
              p1 = x1;
            } else {
              if (LDX_ATT_POLY_MOD > 0 && e % LDX_ATT_POLY_MOD == LDX_ATT_POLY_MOD - 1) p0 = ex2_poly(x0);
              else p0 = ex2_approx(x0);
              if (LDX_ATT_POLY_MOD > 0 && (e + 1) % LDX_ATT_POLY_MOD == LDX_ATT_POLY_MOD - 1) p1 = ex2_poly(x1);
              else p1 = ex2_approx(x1);
            }
            if constexpr (!C::ONES) {
              l0 += p0;
              l1 += p1;
            }
            pk[i] = pack_h2(p0, p1);
          }
          const int chunk = (hf * 4 + ch) ^ (row & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        l += l0 + l1;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + pb);
    }
    // ------------------------------------------------------------------ finalise: O / l -> global
    mbar_wait(pv_done + (nkv - 1) % C::PB, ((nkv - 1) / C::PB) & 1);
    tc_fence_after();
    if constexpr (C::ONES) {                    // the row sum is accumulator column D
      uint32_t lv;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(lv) : "r"(tOg + D) : "memory");
      tmem_ld_wait();
      l = __uint_as_float(lv);
    }
    const float inv_l = 1.0f / l;
    const int qrow = q_cta + g * BQ + row;
    __half* orow = p.o + ((int64_t)b * p.Nq + qrow) * p.ldo + h * D;
#pragma unroll
    for (int cc = 0; cc < C::DN / 16; ++cc) {
      uint32_t o[16];
      tmem_ld16(tOg + cc * 16, o);
      tmem_ld_wait();
      if (qrow < p.Nq) {
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          const int col = cc * 16 + g8 * 8;
          if (col < D) {
            uint4 u;
            u.x = pack_h2(__uint_as_float(o[g8 * 8 + 0]) * inv_l, __uint_as_float(o[g8 * 8 + 1]) * inv_l);
            u.y = pack_h2(__uint_as_float(o[g8 * 8 + 2]) * inv_l, __uint_as_float(o[g8 * 8 + 3]) * inv_l);
            u.z = pack_h2(__uint_as_float(o[g8 * 8 + 4]) * inv_l, __uint_as_float(o[g8 * 8 + 5]) * inv_l);
            u.w = pack_h2(__uint_as_float(o[g8 * 8 + 6]) * inv_l, __uint_as_float(o[g8 * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + col) = u;
          }
        }
      }
    }
  }

  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

struct AttnLaunch {
  AttnParams p;
  dim3 grid;
  int d, ng;
  void* xa = nullptr;          // short-context cross-attention: the dedicated kernel's launch record (xattn.cu)
};

// xattn.cu
bool xattn_eligible(int Nk, int d, int heads, int causal);
int xattn_prepare(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                  int B, int Nq, int Nk, int heads, int d, void** handle);
int xattn_run(void* handle, cudaStream_t st);
void xattn_free(void* handle);

// Query groups per CTA: 4 for the long self-attention streams of narrow heads (K/V read once per 512 queries),
// 1 when K/V is short (cross-attention over 77 tokens: more, smaller CTAs hide the per-CTA latency chain).
static int attention_groups(int d, int Nk) {
  static int short_ng = -1;                     // experiments: LDX_ATT_NG_SHORT=4 runs short K/V (cross-attention) multi-group too
  if (short_ng < 0) { const char* e = getenv("LDX_ATT_NG_SHORT"); short_ng = e ? atoi(e) : 1; }
  if (d <= 64 && Nk <= 256 && short_ng == LDX_ATT_NG_SMALL) return LDX_ATT_NG_SMALL;
  return (d <= 64 && Nk > 256) ? LDX_ATT_NG_SMALL : 1;
}

static int attention_prepare(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                             void* o, int64_t ldo, int B, int Nq, int Nk, int heads, int d, AttnLaunch* L, int causal = 0) {
  LDX_CHECK_ARG(q && k && v && o, "attention: null pointer");
  LDX_CHECK_ARG(d == 40 || d == 64 || d == 80 || d == 128 || d == 160, "attention: head dim %d unsupported", d);
  LDX_CHECK_ARG(B >= 1 && Nq >= 1 && Nk >= 1 && heads >= 1, "attention: bad extents");
  LDX_CHECK_ARG(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0, "attention: output must be 16-byte aligned");
  memset(&L->p, 0, sizeof(L->p));
  L->xa = nullptr;
  if (xattn_eligible(Nk, d, heads, causal)) {
    L->d = d;
    return xattn_prepare(q, ldq, k, ldk, v, ldv, o, ldo, B, Nq, Nk, heads, d, &L->xa);
  }
  auto enc = [&](CUtensorMap* m, const void* ptr, int64_t ld, int N, int rows) {
    const uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)N, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)ld * 2 * (uint64_t)N};
    const uint32_t box[4] = {64, 1, (uint32_t)rows, 1};
    return encode_map_f16(m, ptr, 4, dims, str, box);
  };
  int rc;
  if ((rc = enc(&L->p.qmap, q, ldq, Nq, BQ))) return rc;
  if ((rc = enc(&L->p.kmap, k, ldk, Nk, BKV))) return rc;
  if ((rc = enc(&L->p.vmap, v, ldv, Nk, BKV))) return rc;
  L->p.o = (__half*)o;
  L->p.ldo = ldo;
  L->p.B = B; L->p.Nq = Nq; L->p.Nk = Nk; L->p.heads = heads;
  L->p.scale_log2 = 1.4426950408889634f / sqrtf((float)d);
  L->p.causal = causal;
  LDX_CHECK_ARG(!causal || Nq == Nk, "attention: the causal mask needs Nq == Nk");
  L->d = d;
  L->ng = attention_groups(d, Nk);
  L->grid = dim3((Nq + L->ng * BQ - 1) / (L->ng * BQ), heads, B);
  return LDX_OK;
}

template <int D, int NG_, bool CAUSAL = false>
static int attention_launch_t(const AttnLaunch* L, cudaStream_t st) {
  using C = AttnCfg<D, NG_>;
  static std::atomic<uint64_t> attr_done{0};          // per kernel instantiation, one bit per device
  int arc = set_max_dyn_smem_once(ldx_attention_kernel<D, NG_, CAUSAL>, C::SMEM, attr_done);
  if (arc) return arc;
  LDX_CUDA(launch_k(ldx_attention_kernel<D, NG_, CAUSAL>, L->grid, dim3(C::THREADS), (size_t)C::SMEM, st, 1, L->p));
  return LDX_OK;
}

static int attention_launch(const AttnLaunch* L, cudaStream_t st) {
  if (L->xa) return xattn_run(L->xa, st);
  if (L->p.causal) {
    if (L->d == 64 && L->ng == 1) return attention_launch_t<64, 1, true>(L, st);
    set_error("attention: the causal variant exists for d = 64 and Nk <= 256 (CLIP text tower), got d = %d, Nk = %d", L->d,
              L->p.Nk);
    return LDX_E_INVALID;
  }
  switch (L->d) {
    case 40: return L->ng == 1 ? attention_launch_t<40, 1>(L, st) : attention_launch_t<40, LDX_ATT_NG_SMALL>(L, st);
    case 64: return L->ng == 1 ? attention_launch_t<64, 1>(L, st) : attention_launch_t<64, LDX_ATT_NG_SMALL>(L, st);
    case 80: return attention_launch_t<80, 1>(L, st);
    case 128: return attention_launch_t<128, 1>(L, st);
    case 160: return attention_launch_t<160, 1>(L, st);
  }
  set_error("attention: head dim %d unsupported", L->d);
  return LDX_E_INVALID;
}

// program support (api.cu)
int attention_prepare_op(const ldx_generic_op_t* op, void** handle) {
  AttnLaunch* L = new AttnLaunch;
  // p: q, k, v, o ; l: ldq, ldk, ldv, ldo ; i: B, Nq, Nk, heads, d, causal
  int rc = attention_prepare(op->p[0], op->l[0], op->p[1], op->l[1], op->p[2], op->l[2], (void*)op->p[3], op->l[3],
                             op->i[0], op->i[1], op->i[2], op->i[3], op->i[4], L, op->i[5]);
  if (rc) {
    delete L;
    return rc;
  }
  *handle = L;
  return LDX_OK;
}
int attention_run_handle(void* handle, cudaStream_t st) { return attention_launch((AttnLaunch*)handle, st); }
void attention_free_handle(void* handle) {
  AttnLaunch* L = (AttnLaunch*)handle;
  if (L->xa) xattn_free(L->xa);
  delete L;
}

}  // namespace ldx

using namespace ldx;

#ifdef LDX_ATT_TIMING
extern "C" int ldx_debug_att_ts(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, ldx::g_att_ts, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int ldx_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                             void* o, int64_t ldo, int B, int Nq, int Nk, int heads, int d, ldx_stream_t stream) {
  AttnLaunch L;
  int rc = attention_prepare(q, ldq, k, ldk, v, ldv, o, ldo, B, Nq, Nk, heads, d, &L);
  if (rc) return rc;
  rc = attention_launch(&L, as_stream(stream));
  if (L.xa) xattn_free(L.xa);
  return rc;
}

extern "C" int ldx_attention_causal(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                    void* o, int64_t ldo, int B, int N, int heads, int d, ldx_stream_t stream) {
  AttnLaunch L;
  int rc = attention_prepare(q, ldq, k, ldk, v, ldv, o, ldo, B, N, N, heads, d, &L, 1);
  if (rc) return rc;
  return attention_launch(&L, as_stream(stream));
}
