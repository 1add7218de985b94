// Cross-attention over a short context (Nk <= 80: the 77 CLIP tokens of CrossAttention.forward with context,
// LD.py:4035-4060 through optimized_attention, LD.py:3966-3988) on tcgen05 / TMEM (sm_100a).
//
//   o[b, i, h*d:(h+1)*d] = softmax_j( q_i . k_j / sqrt(d) ) v_j,   j < Nk <= 80,  d in {40, 80}
//
// The general flash kernel (attention.cu) runs this shape as 8192 small CTAs, each a serial latency chain (load ->
// S -> softmax -> P V over TWO 64-key tiles of which the second holds 13 keys -> read-out): 62 us at 16 x 8 x 4096 x 77
// against 13 us of HBM time for reading Q and writing O.  Here the whole context is ONE 80-key tile (single-pass
// softmax, no running maximum, no O rescale), CTAs are persistent, and a CTA pipelines its (batch, query tile, head)
// units: two softmax warpgroups alternate units, each with its own S / P / O buffers, while one MMA warp and one TMA
// warp run ahead.
//   warp 0      : TMA producer: Q (128 x d) and K (80 x d) of a unit into the QK ring, V (80 x d) into the V ring
//   warp 1      : MMA issuer: S(u) = Q K^T (M 128, N 80, K ceil16(d)) as soon as the buffer of u's warpgroup is read;
//                 O(u) = P V (M 128, N ceil16(d), K 80; V row-major through an MN-major descriptor) when P(u) is stored
//   warps 4-7, 8-11 : softmax warpgroups (thread = query row): S row (80 fp32) from TMEM -> mask keys >= Nk -> max ->
//                 exp2 -> fp16 P row into the K-major SWIZZLE_128B tile (64 + 16 keys) -> arrive; then O from TMEM,
//                 times 1 / row sum, fp16, 16-byte stores to the [B, Nq, heads * d] output
// Rows / keys past Nq / Nk arrive as zeros through TMA's out-of-bounds fill (4-D views (d, heads, N, B)).
// Roofline: HBM (Q read + O write, 2 * B * Nq * heads * d * 2 bytes) with the MUFU floor B * heads * Nq * 80 / 16 per
// clock and SM right behind it.
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace ldx {

#ifndef LDX_XA_DBG
#define LDX_XA_DBG 0      // timing experiments: 1 no exponentials, 2 no output stores, 4 no key masking, 8 no P stores
#endif
#ifdef LDX_XA_TIMING
// variant-build instrumentation: cycles per phase, summed over the units of CTA 0 (softmax warp 4 / the MMA warp)
__device__ unsigned long long g_xa_ts[32];
#define XT_DECL long long xt_t = clock64(), xt_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define XT(i) { const long long xt_now = clock64(); xt_acc[i] += xt_now - xt_t; xt_t = xt_now; }
#define XT_FLUSH(base, n) if (blockIdx.x == 0 && lane == 0) { for (int i = 0; i < 8; ++i) g_xa_ts[(base) + i] = (unsigned long long)xt_acc[i]; g_xa_ts[(base) + 8] = (unsigned long long)(n); }
#else
#define XT_DECL
#define XT(i)
#define XT_FLUSH(base, n)
#endif
static constexpr int XA_BQ = 128, XA_BK = 80;
static constexpr int XA_KV_CHUNK = XA_BK * 128;      // 80 rows x 128 B
static constexpr int XA_P_BYTES = 2 * XA_BQ * 128;   // keys 0..63 | keys 64..79 (16 of the 64 columns used)
static constexpr int XA_MAX_ST = 4, XA_MAX_WG = 3;

template <int D>
struct XaCfg {
  static constexpr int DK = (D + 15) / 16;
  static constexpr int DCH = (D + 63) / 64;
  static constexpr int DN = DK * 16;
  static constexpr int QK_BYTES = DCH * (XA_BQ * 128 + XA_KV_CHUNK);
  static constexpr int V_BYTES = DCH * XA_KV_CHUNK;
  // softmax warpgroups (each with its own S / P / O buffers) and ring depths.  Measured with the phase counters below
  // (16 x 8 x 4096 x 77, d = 40, two warpgroups, one MMA warp): a warpgroup's unit is 1.4 k clocks of softmax, 1.2 k
  // waiting for P V and 0.7 k of read-out, and the single MMA warp spent 2.0 k clocks per unit issuing and waiting --
  // hence three warpgroups where shared memory allows and separate issuers for S and P V.  A Q tile comes from HBM
  // (~2 us under load), so the Q/K ring runs three units ahead.
  static constexpr int NWG = D <= 64 ? 3 : 2;
  static constexpr int NQK = D <= 64 ? 3 : 2;
  static constexpr int NV = D <= 64 ? 3 : 2;
  static constexpr int THREADS = 128 + NWG * 128;
  static constexpr int SMEM = NQK * QK_BYTES + NV * V_BYTES + NWG * XA_P_BYTES + 512 + 1024;
  static constexpr int S_STRIDE = 96;                 // TMEM columns per S buffer (80 used)
  static constexpr int O_STRIDE = DN <= 64 ? 64 : 96;
  static_assert(NWG * (S_STRIDE + O_STRIDE) <= 512, "TMEM budget");
  static_assert(SMEM <= 227 * 1024, "smem budget");
  static_assert(32 * D * 2 <= 2 * 4096 && 16 * D * 2 <= 4096, "output staging must fit the warp's own P rows");
};

struct XaBarriers {
  uint64_t qk_full[XA_MAX_ST], qk_empty[XA_MAX_ST], v_full[XA_MAX_ST], v_empty[XA_MAX_ST];
  uint64_t s_full[XA_MAX_WG], s_free[XA_MAX_WG], p_full[XA_MAX_WG], o_full[XA_MAX_WG];
  uint32_t tmem_base;
};

struct alignas(64) XaParams {
  CUtensorMap qmap, kmap, vmap;     // (d, heads, N, B) views; boxes (64, 1, 128, 1) and (64, 1, 80, 1)
  __half* o;
  int64_t ldo;
  int32_t B, Nq, Nk, heads, qtiles, units;
  float scale_log2;
};

template <int D>
__global__ void __launch_bounds__(XaCfg<D>::THREADS, 1) ldx_xattn_kernel(const __grid_constant__ XaParams p) {
  using C = XaCfg<D>;
  constexpr int NWG = C::NWG;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQK = smem;                                   // NQK x [Q chunks | K chunks]
  uint8_t* sV = sQK + C::NQK * C::QK_BYTES;
  uint8_t* sP = sV + C::NV * C::V_BYTES;                 // NWG x XA_P_BYTES
  XaBarriers* bars = reinterpret_cast<XaBarriers*>(sP + NWG * XA_P_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this CTA's units: u = blockIdx.x + k * gridDim.x, k = 0 .. nu - 1; unit k belongs to warpgroup k % NWG
  const int nu = ((int)blockIdx.x < p.units) ? (p.units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.qmap);
    tma_prefetch_desc(&p.kmap);
    tma_prefetch_desc(&p.vmap);
    for (int i = 0; i < XA_MAX_ST; ++i) {
      mbar_init(&bars->qk_full[i], 1);
      mbar_init(&bars->qk_empty[i], 1);
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], 1);
    }
    for (int i = 0; i < XA_MAX_WG; ++i) {
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->s_free[i], 4);
      mbar_init(&bars->p_full[i], 4);
      mbar_init(&bars->o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t tS = tmem_base, tO = tmem_base + NWG * C::S_STRIDE;
  pdl_prologue();

  auto unit_coords = [&](int k, int& b, int& q0, int& h) {
    const int u = (int)blockIdx.x + k * (int)gridDim.x;
    h = u % p.heads;
    const int t = u / p.heads;
    q0 = (t % p.qtiles) * XA_BQ;
    b = t / p.qtiles;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer of Q and K
    for (int k = 0; k < nu; ++k) {
      int b, q0, h;
      unit_coords(k, b, q0, h);
      const int qs = k % C::NQK;
      mbar_wait(&bars->qk_empty[qs], ((k / C::NQK) & 1) ^ 1);
      if (elect_one()) {
        uint8_t* dq = sQK + qs * C::QK_BYTES;
        uint8_t* dk = dq + C::DCH * (XA_BQ * 128);
        mbar_arrive_expect_tx(&bars->qk_full[qs], C::QK_BYTES);
        for (int c = 0; c < C::DCH; ++c) {
          tma_load_4d(dq + c * (XA_BQ * 128), &p.qmap, &bars->qk_full[qs], c * 64, h, q0, b);
          tma_load_4d(dk + c * XA_KV_CHUNK, &p.kmap, &bars->qk_full[qs], c * 64, h, 0, b);
        }
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ TMA producer of V (its own warp: a V slot frees
    // only after P V, long after the Q/K slot of the same unit, and must not hold the Q/K stream back)
    for (int k = 0; k < nu; ++k) {
      int b, q0, h;
      unit_coords(k, b, q0, h);
      const int vs = k % C::NV;
      mbar_wait(&bars->v_empty[vs], ((k / C::NV) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->v_full[vs], C::V_BYTES);
        for (int c = 0; c < C::DCH; ++c)
          tma_load_4d(sV + vs * C::V_BYTES + c * XA_KV_CHUNK, &p.vmap, &bars->v_full[vs], c * 64, h, 0, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer of S = Q K^T
    const uint32_t idesc_s = umma_idesc_f16(XA_BQ, XA_BK, 0, 0);
    XT_DECL
    for (int k = 0; k < nu; ++k) {
      const int w = k % NWG, qs = k % C::NQK;
      XT(7)
      mbar_wait(&bars->qk_full[qs], (k / C::NQK) & 1);
      XT(0)
      if (k >= NWG) mbar_wait(&bars->s_free[w], ((k / NWG) - 1) & 1);   // the warpgroup has read S(k - NWG)
      XT(1)
      tc_fence_after();
      const uint32_t q_addr = smem_u32(sQK + qs * C::QK_BYTES), k_addr = q_addr + C::DCH * (XA_BQ * 128);
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < C::DK; ++s) {
          const uint32_t qa = q_addr + (s / 4) * (XA_BQ * 128) + (s % 4) * 32;
          const uint32_t ka = k_addr + (s / 4) * XA_KV_CHUNK + (s % 4) * 32;
          umma_f16_ss(tS + w * C::S_STRIDE, umma_desc_kmajor_sw128(qa), umma_desc_kmajor_sw128(ka), idesc_s, (uint32_t)(s != 0));
        }
        umma_commit(&bars->s_full[w]);
        umma_commit(&bars->qk_empty[qs]);
      }
      __syncwarp();
      XT(2)
    }
    XT_FLUSH(16, nu)
  } else if (warp == 3) {
    // ------------------------------------------------------------------ MMA issuer of O = P V
    const uint32_t idesc_o = umma_idesc_f16(XA_BQ, C::DN, 0, 1);        // B = V is MN-major
    for (int k = 0; k < nu; ++k) {
      const int w = k % NWG, vs = k % C::NV;
      mbar_wait(&bars->p_full[w], (k / NWG) & 1);
      mbar_wait(&bars->v_full[vs], (k / C::NV) & 1);
      tc_fence_after();
      const uint32_t p_addr = smem_u32(sP + w * XA_P_BYTES), v_addr = smem_u32(sV + vs * C::V_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < XA_BK / 16; ++s) {
          const uint32_t pa = s < 4 ? p_addr + s * 32 : p_addr + XA_BQ * 128;
          umma_f16_ss(tO + w * C::O_STRIDE, umma_desc_kmajor_sw128(pa), umma_desc_mnmajor_sw128(v_addr + s * 2048, XA_KV_CHUNK),
                      idesc_o, (uint32_t)(s != 0));
        }
        umma_commit(&bars->o_full[w]);
        umma_commit(&bars->v_empty[vs]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int w = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tSw = tS + w * C::S_STRIDE + lane_off, tOw = tO + w * C::O_STRIDE + lane_off;
    uint8_t* prow0 = sP + w * XA_P_BYTES + row * 128;
    uint8_t* prow1 = prow0 + XA_BQ * 128;
    // output staging: the warp's own 32 P rows (4 KB in each of the two P chunks), free between P V(k) and P(k + NWG)
    uint8_t* st0 = sP + w * XA_P_BYTES + qd * 4096;
    uint8_t* st1 = st0 + XA_BQ * 128;
    const float c = p.scale_log2;
    const uint32_t ninf = 0xff800000u;
    XT_DECL
    for (int k = w; k < nu; k += NWG) {
      int b, q0, h;
      unit_coords(k, b, q0, h);
      const uint32_t ph = (uint32_t)((k / NWG) & 1);
      uint32_t ra[32], rb[32], rc[16];
      XT(7)
      mbar_wait(&bars->s_full[w], ph);
      XT(0)
      tc_fence_after();
      tmem_ld32(tSw, ra);
      tmem_ld32(tSw + 32, rb);
      tmem_ld16(tSw + 64, rc);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_free[w]);
      XT(1)
      if (!(LDX_XA_DBG & 4) && p.Nk < XA_BK) {
        if (p.Nk < 64) {                       // general short context
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i >= p.Nk) ra[i] = ninf;
            if (32 + i >= p.Nk) rb[i] = ninf;
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (64 + i >= p.Nk) rc[i] = ninf;
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
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(rc[i]), __uint_as_float(rc[i + 1])));
        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(rc[i + 2]), __uint_as_float(rc[i + 3])));
      }
      const float mc = fmaxf(mx0, mx1) * c;
      float l0 = 0.f, l1 = 0.f;
      // a P row: 8 + 2 chunks of 8 keys (16 bytes); chunk index XOR (row % 8) = SWIZZLE_128B
      auto chunks = [&](const uint32_t* r, int nch, uint8_t* prow, int ch0) {
#pragma unroll
        for (int ch = 0; ch < nch; ++ch) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = ch * 8 + 2 * i;
            float p0, p1;                       // one FFMA2 for the pair (bit-identical to two fmaf)
            f2_unpack(f2_fma(f2_pack(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), f2_pack(c, c), f2_pack(-mc, -mc)), p0, p1);
            if (!(LDX_XA_DBG & 1)) { p0 = ex2_approx(p0); p1 = ex2_approx(p1); }
            f2_unpack(f2_fma(f2_pack(p0, p1), f2_pack(1.0f, 1.0f), f2_pack(l0, l1)), l0, l1);   // l0 += p0, l1 += p1
            pk[i] = pack_h2(p0, p1);
          }
          *reinterpret_cast<uint4*>(prow + (((ch0 + ch) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      };
      chunks(ra, 4, prow0, 0);
      chunks(rb, 4, prow0, 4);
      chunks(rc, 2, prow1, 0);
      XT(2)
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[w]);
      // the row sum is taken from the fp32 exponentials; P V uses their fp16 roundings (as the flash kernel does)
      const float inv_l = 1.0f / (l0 + l1);
      XT(3)
      mbar_wait(&bars->o_full[w], ph);
      XT(4)
      tc_fence_after();
      // O row: all of it in flight, one wait; then through the staging block so that a store instruction covers whole
      // output rows (d = 40: row-per-lane stores of 80 bytes at a pitch of 640 cost 13 of 53 us)
      uint32_t o[C::DN];
#pragma unroll
      for (int cc = 0; cc < C::DN / 16; ++cc) tmem_ld16(tOw + cc * 16, *reinterpret_cast<uint32_t(*)[16]>(&o[cc * 16]));
      tmem_ld_wait();
      tc_fence_before();      // O(k) is out of TMEM before this warpgroup's next p_full lets P V(k + NWG) overwrite it
      uint8_t* strow = lane < 16 ? st0 + lane * (D * 2) : st1 + (lane - 16) * (D * 2);
#pragma unroll
      for (int col = 0; col < D; col += 8) {
        uint4 u;
        u.x = pack_h2(__uint_as_float(o[col + 0]) * inv_l, __uint_as_float(o[col + 1]) * inv_l);
        u.y = pack_h2(__uint_as_float(o[col + 2]) * inv_l, __uint_as_float(o[col + 3]) * inv_l);
        u.z = pack_h2(__uint_as_float(o[col + 4]) * inv_l, __uint_as_float(o[col + 5]) * inv_l);
        u.w = pack_h2(__uint_as_float(o[col + 6]) * inv_l, __uint_as_float(o[col + 7]) * inv_l);
        *reinterpret_cast<uint4*>(strow + col * 2) = u;
      }
      __syncwarp();
      // 32 rows x (D / 8) 16-byte pieces, consecutive lanes on consecutive pieces of a row
      constexpr int PIECES = D / 8;
      __half* obase = p.o + ((int64_t)b * p.Nq + q0 + qd * 32) * p.ldo + h * D;
      const int rows_ok = p.Nq - (q0 + qd * 32);
#pragma unroll
      for (int i = 0; i < PIECES; ++i) {
        const int pc = i * 32 + lane, r = pc / PIECES, part = pc - r * PIECES;
        const uint8_t* src = r < 16 ? st0 + r * (D * 2) + part * 16 : st1 + (r - 16) * (D * 2) + part * 16;
        if (r < rows_ok && !(LDX_XA_DBG & 2)) *reinterpret_cast<uint4*>(obase + (int64_t)r * p.ldo + part * 8) = *reinterpret_cast<const uint4*>(src);
      }
      __syncwarp();           // the staging block is this warp's own P rows: next written by P(k + NWG)
      XT(5)
    }
    if (warp == 4) { XT_FLUSH(0, nu) }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

struct XaLaunch {
  XaParams p;
  int d, grid;
};

bool xattn_eligible(int Nk, int d, int heads, int causal) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("LDX_XATTN"); on = e ? atoi(e) : 1; }
  return on && !causal && Nk >= 1 && Nk <= XA_BK && (d == 40 || d == 80) && heads >= 1;
}

int xattn_prepare(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                  int B, int Nq, int Nk, int heads, int d, void** handle) {
  XaLaunch* L = new XaLaunch;
  memset(&L->p, 0, sizeof(L->p));
  auto enc = [&](CUtensorMap* m, const void* ptr, int64_t ld, int N, int rows) {
    const uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)N, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)ld * 2 * (uint64_t)N};
    const uint32_t box[4] = {64, 1, (uint32_t)rows, 1};
    return encode_map_f16(m, ptr, 4, dims, str, box);
  };
  int rc;
  if ((rc = enc(&L->p.qmap, q, ldq, Nq, XA_BQ)) || (rc = enc(&L->p.kmap, k, ldk, Nk, XA_BK)) ||
      (rc = enc(&L->p.vmap, v, ldv, Nk, XA_BK))) {
    delete L;
    return rc;
  }
  L->p.o = (__half*)o;
  L->p.ldo = ldo;
  L->p.B = B; L->p.Nq = Nq; L->p.Nk = Nk; L->p.heads = heads;
  L->p.qtiles = (Nq + XA_BQ - 1) / XA_BQ;
  L->p.units = B * L->p.qtiles * heads;
  L->p.scale_log2 = 1.4426950408889634f / sqrtf((float)d);
  L->d = d;
  const int sms = num_sms();
  L->grid = L->p.units < sms ? L->p.units : sms;
  *handle = L;
  return LDX_OK;
}

template <int D>
static int xattn_launch_t(const XaLaunch* L, cudaStream_t st) {
  using C = XaCfg<D>;
  static std::atomic<uint64_t> attr_done{0};
  int arc = set_max_dyn_smem_once(ldx_xattn_kernel<D>, C::SMEM, attr_done);
  if (arc) return arc;
  LDX_CUDA(launch_k(ldx_xattn_kernel<D>, dim3(L->grid), dim3(C::THREADS), (size_t)C::SMEM, st, 1, L->p));
  return LDX_OK;
}

int xattn_run(void* handle, cudaStream_t st) {
  const XaLaunch* L = (const XaLaunch*)handle;
  return L->d == 40 ? xattn_launch_t<40>(L, st) : xattn_launch_t<80>(L, st);
}

void xattn_free(void* handle) { delete (XaLaunch*)handle; }

}  // namespace ldx

#ifdef LDX_XA_TIMING
extern "C" int ldx_debug_xattn_ts(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, ldx::g_xa_ts, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : 1;
}
#endif
