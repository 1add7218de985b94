// Flash-style attention on tcgen05 / TMEM (sm_100a) for optimized_attention (LD.py:3966-3988):
//   o[b, i, h*d:(h+1)*d] = softmax_j( q_i . k_j / sqrt(d) ) v_j      (no mask, dropout 0)
// with head dims 40 / 80 / 160 (SD1.5: 8 heads at 320 / 640 / 1280 channels) plus 64 / 128.
//
// One CTA = 128 query rows of one (batch, head); keys/values are streamed in tiles of 64.
//   warp 0   : TMA producer (Q once, K double buffered, V)
//   warp 1   : tcgen05.mma issuer (one thread): S = Q K^T (fp32 in TMEM), O += P V (fp32 in TMEM)
//   warps 2-5: online softmax, one thread per query row: S from TMEM -> exp2 -> P (fp16) into smem in
//              the K-major SWIZZLE_128B layout tcgen05.mma reads; lazy rescale of O in TMEM
// Head slices are addressed by a 4-D TMA view (d, heads, N, B) of the [B, N, heads*d] tensor, so the
// box's columns beyond d are out of bounds and arrive as zeros (K padding for d = 40).
// V is consumed straight from its row-major tile through an MN-major smem descriptor.
// Several CTAs share an SM (TMEM: 64 + ceil16(d) columns each), so one CTA's softmax overlaps
// another's MMAs.  Roofline: 4*Nq*Nk*d*heads FLOPs per batch element against the tensor peak; for
// d = 40 the exp throughput (MUFU) is the practical bound.
#include "common.cuh"
#include "internal.h"

namespace ldx {

int encode_map_f16(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

static constexpr int ATT_THREADS = 192;
static constexpr int BQ = 128;
static constexpr int BKV = 64;
#ifndef LDX_ATT_KST
#define LDX_ATT_KST 1
#endif
#ifndef LDX_ATT_VST
#define LDX_ATT_VST 1
#endif

template <int D>
struct AttnCfg {
  static constexpr int DK = (D + 15) / 16;          // 16-wide K steps of Q K^T
  static constexpr int DCH = (D + 63) / 64;         // 64-column smem chunks per row
  static constexpr int DN = DK * 16;                // UMMA N of P V
  static constexpr int KST = DCH == 1 ? LDX_ATT_KST : 2;
  static constexpr int VST = DCH == 1 ? LDX_ATT_VST : 1;
  static constexpr int Q_BYTES = DCH * BQ * 128;
  static constexpr int KV_BYTES = DCH * BKV * 128;
  static constexpr int P_BYTES = BQ * BKV * 2;
  static constexpr int TMEM_COLS = (64 + DN) <= 128 ? 128 : 256;
  static constexpr int SMEM = Q_BYTES + (KST + VST) * KV_BYTES + P_BYTES + 256 + 1024;
  static constexpr int MIN_BLOCKS = SMEM <= 55 * 1024 ? 4 : (SMEM <= 72 * 1024 ? 3 : (SMEM <= 110 * 1024 ? 2 : 1));
};

struct AttnBarriers {
  uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full, s_free, p_full, o_done;
  uint32_t tmem_base;
};

struct alignas(64) AttnParams {
  CUtensorMap qmap, kmap, vmap;
  __half* o;
  int64_t ldo;
  int32_t B, Nq, Nk, heads;
  float scale_log2;   // log2(e) / sqrt(d)
};

template <int D>
__global__ void __launch_bounds__(ATT_THREADS, AttnCfg<D>::MIN_BLOCKS) ldx_attention_kernel(
    const __grid_constant__ AttnParams p) {
  using C = AttnCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + C::Q_BYTES;
  uint8_t* sV = sK + C::KST * C::KV_BYTES;
  uint8_t* sP = sV + C::VST * C::KV_BYTES;
  AttnBarriers* bars = reinterpret_cast<AttnBarriers*>(sP + C::P_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.Nk + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.qmap);
    tma_prefetch_desc(&p.kmap);
    tma_prefetch_desc(&p.vmap);
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->k_empty[i], 1);
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], 1);
    }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->s_free, 4);
    mbar_init(&bars->p_full, 4);
    mbar_init(&bars->o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t tS = tmem_base;          // 64 fp32 columns
  const uint32_t tO = tmem_base + 64;     // DN fp32 columns
  pdl_prologue();

  if (warp == 0) {
    // ---- TMA producer: warp-uniform loop, one elected lane issues
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars->q_full, C::Q_BYTES);
      for (int c = 0; c < C::DCH; ++c) tma_load_4d(sQ + c * (BQ * 128), &p.qmap, &bars->q_full, c * 64, h, q0, b);
    }
    __syncwarp();
    for (int j = 0; j < nkv; ++j) {
      const int ks = j % C::KST, vs = j % C::VST;
      mbar_wait(&bars->k_empty[ks], ((j / C::KST) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->k_full[ks], C::KV_BYTES);
        for (int c = 0; c < C::DCH; ++c)
          tma_load_4d(sK + ks * C::KV_BYTES + c * (BKV * 128), &p.kmap, &bars->k_full[ks], c * 64, h, j * BKV, b);
      }
      __syncwarp();
      mbar_wait(&bars->v_empty[vs], ((j / C::VST) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->v_full[vs], C::KV_BYTES);
        for (int c = 0; c < C::DCH; ++c)
          tma_load_4d(sV + vs * C::KV_BYTES + c * (BKV * 128), &p.vmap, &bars->v_full[vs], c * 64, h, j * BKV, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- MMA issuer: warp-uniform control flow, one elected lane issues tcgen05.mma / commit
    const uint32_t idesc_s = umma_idesc_f16(BQ, BKV, 0, 0);
    const uint32_t idesc_o = umma_idesc_f16(BQ, C::DN, 0, 1);       // B = V is MN-major
    const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
    auto mma_s = [&](int j) {
      const int ks = j % C::KST;
      mbar_wait(&bars->k_full[ks], (j / C::KST) & 1);
      tc_fence_after();
      const uint32_t k_addr = smem_u32(sK + ks * C::KV_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < C::DK; ++s) {
          const uint32_t qa = q_addr + (s / 4) * (BQ * 128) + (s % 4) * 32;
          const uint32_t ka = k_addr + (s / 4) * (BKV * 128) + (s % 4) * 32;
          umma_f16_ss(tS, umma_desc_kmajor_sw128(qa), umma_desc_kmajor_sw128(ka), idesc_s, (uint32_t)(s != 0));
        }
        umma_commit(&bars->k_empty[ks]);
        umma_commit(&bars->s_full);
      }
      __syncwarp();
    };
    mbar_wait(&bars->q_full, 0);
    mma_s(0);
    for (int j = 0; j < nkv; ++j) {
      if (j + 1 < nkv) {
        mbar_wait(&bars->s_free, j & 1);
        tc_fence_after();
        mma_s(j + 1);
      }
      const int vs = j % C::VST;
      mbar_wait(&bars->p_full, j & 1);
      mbar_wait(&bars->v_full[vs], (j / C::VST) & 1);
      tc_fence_after();
      const uint32_t v_addr = smem_u32(sV + vs * C::KV_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < BKV / 16; ++s) {
          umma_f16_ss(tO, umma_desc_kmajor_sw128(p_addr + s * 32),
                      umma_desc_mnmajor_sw128(v_addr + s * 2048, BKV * 128), idesc_o, (uint32_t)((j | s) != 0));
        }
        umma_commit(&bars->v_empty[vs]);
        umma_commit(&bars->o_done);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const float c = p.scale_log2;
    float m_used = -INFINITY;   // row max the exponentials are relative to (raw score units)
    float l = 0.f;
    uint8_t* prow = sP + row * 128;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&bars->s_full, j & 1);
      tc_fence_after();
      const int kv_left = p.Nk - j * BKV;   // valid keys in this tile (>= 1)
      const bool full_tile = kv_left >= BKV;   // only the last tile of a ragged Nk needs key masking
      uint32_t r[32];
      float m_tile = -INFINITY;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        tmem_ld32(tS + lane_off + hf * 32, r);
        tmem_ld_wait();
        if (full_tile) {
#pragma unroll
          for (int i = 0; i < 32; ++i) m_tile = fmaxf(m_tile, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (hf * 32 + i < kv_left) m_tile = fmaxf(m_tile, __uint_as_float(r[i]));
        }
      }
      if (j > 0) {
        mbar_wait(&bars->o_done, (j - 1) & 1);   // P buffer and O are free again
        tc_fence_after();
      }
      // lazy rescale: only when some row of this warp grew its max by more than 2^8 after scaling
      const bool grow = (m_tile - m_used) * c > 8.0f;
      if (j == 0) {
        m_used = m_tile;
      } else if (__any_sync(0xffffffffu, grow)) {
        const float m_new = fmaxf(m_used, m_tile);
        const float alpha = ex2_approx((m_used - m_new) * c);
        m_used = m_new;
        l *= alpha;
#pragma unroll
        for (int cc = 0; cc < C::DN / 16; ++cc) {
          uint32_t o[16];
          tmem_ld16(tO + lane_off + cc * 16, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          asm volatile(
              "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
              "%13, %14, %15};" ::"r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]),
              "r"(o[7]), "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15]),
              "r"(tO + lane_off + cc * 16)
              : "memory");
        }
        tmem_st_wait();
      }
      const float mc = m_used * c;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        tmem_ld32(tS + lane_off + hf * 32, r);
        tmem_ld_wait();
        if (hf == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->s_free);   // S may be overwritten by the next Q K^T
        }
        uint32_t pk[16];
        if (full_tile) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = ex2_approx(__uint_as_float(r[2 * i]) * c - mc);
            const float p1 = ex2_approx(__uint_as_float(r[2 * i + 1]) * c - mc);
            l0 += p0;
            l1 += p1;
            pk[i] = pack_h2(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0 = ex2_approx(__uint_as_float(r[2 * i]) * c - mc);
            float p1 = ex2_approx(__uint_as_float(r[2 * i + 1]) * c - mc);
            if (hf * 32 + 2 * i >= kv_left) p0 = 0.f;
            if (hf * 32 + 2 * i + 1 >= kv_left) p1 = 0.f;
            l0 += p0;
            l1 += p1;
            pk[i] = pack_h2(p0, p1);
          }
        }
        // 32 keys = 4 16-byte chunks; chunk index XOR (row % 8) is the SWIZZLE_128B pattern
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = (hf * 4 + ch) ^ (row & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) =
              make_uint4(pk[ch * 4 + 0], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        }
      }
      l += l0 + l1;
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full);
    }
    // ------------------------------------------------------------------ finalise: O / l -> global
    mbar_wait(&bars->o_done, (nkv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int qrow = q0 + row;
    __half* orow = p.o + ((int64_t)b * p.Nq + qrow) * p.ldo + h * D;
#pragma unroll
    for (int cc = 0; cc < C::DN / 16; ++cc) {
      uint32_t o[16];
      tmem_ld16(tO + lane_off + cc * 16, o);
      tmem_ld_wait();
      if (qrow < p.Nq) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int col = cc * 16 + g * 8;
          if (col < D) {
            uint4 u;
            u.x = pack_h2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
            u.y = pack_h2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
            u.z = pack_h2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
            u.w = pack_h2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + col) = u;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

struct AttnLaunch {
  AttnParams p;
  dim3 grid;
  int d;
};

static int attention_prepare(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                             void* o, int64_t ldo, int B, int Nq, int Nk, int heads, int d, AttnLaunch* L) {
  LDX_CHECK_ARG(q && k && v && o, "attention: null pointer");
  LDX_CHECK_ARG(d == 40 || d == 64 || d == 80 || d == 128 || d == 160, "attention: head dim %d unsupported", d);
  LDX_CHECK_ARG(B >= 1 && Nq >= 1 && Nk >= 1 && heads >= 1, "attention: bad extents");
  LDX_CHECK_ARG(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0, "attention: output must be 16-byte aligned");
  memset(&L->p, 0, sizeof(L->p));
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
  L->grid = dim3((Nq + BQ - 1) / BQ, heads, B);
  L->d = d;
  return LDX_OK;
}

template <int D>
static int attention_launch_t(const AttnLaunch* L, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    LDX_CUDA(cudaFuncSetAttribute(ldx_attention_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  AttnCfg<D>::SMEM));
    attr = true;
  }
  LDX_CUDA(launch_k(ldx_attention_kernel<D>, L->grid, dim3(ATT_THREADS), (size_t)AttnCfg<D>::SMEM, st, 1, L->p));
  return LDX_OK;
}

static int attention_launch(const AttnLaunch* L, cudaStream_t st) {
  switch (L->d) {
    case 40: return attention_launch_t<40>(L, st);
    case 64: return attention_launch_t<64>(L, st);
    case 80: return attention_launch_t<80>(L, st);
    case 128: return attention_launch_t<128>(L, st);
    case 160: return attention_launch_t<160>(L, st);
  }
  set_error("attention: head dim %d unsupported", L->d);
  return LDX_E_INVALID;
}

// program support (api.cu)
int attention_prepare_op(const ldx_generic_op_t* op, void** handle) {
  AttnLaunch* L = new AttnLaunch;
  // p: q, k, v, o ; l: ldq, ldk, ldv, ldo ; i: B, Nq, Nk, heads, d
  int rc = attention_prepare(op->p[0], op->l[0], op->p[1], op->l[1], op->p[2], op->l[2], (void*)op->p[3], op->l[3],
                             op->i[0], op->i[1], op->i[2], op->i[3], op->i[4], L);
  if (rc) {
    delete L;
    return rc;
  }
  *handle = L;
  return LDX_OK;
}
int attention_run_handle(void* handle, cudaStream_t st) { return attention_launch((AttnLaunch*)handle, st); }
void attention_free_handle(void* handle) { delete (AttnLaunch*)handle; }

}  // namespace ldx

using namespace ldx;

extern "C" int ldx_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                             void* o, int64_t ldo, int B, int Nq, int Nk, int heads, int d, ldx_stream_t stream) {
  AttnLaunch L;
  int rc = attention_prepare(q, ldq, k, ldk, v, ldv, o, ldo, B, Nq, Nk, heads, d, &L);
  if (rc) return rc;
  return attention_launch(&L, as_stream(stream));
}
