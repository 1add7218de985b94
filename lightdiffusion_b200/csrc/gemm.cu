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
// * warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warps 2..5 = epilogue
//   (tcgen05.ld -> bias / timestep-embedding row add / residual / SiLU / GEGLU -> fp16 stores).
//
// Roofline: tensor-pipe bound; algorithmic FLOPs = 2 * M * N * K.
#include "common.cuh"
#include "internal.h"

namespace ldx {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KiB
static constexpr int GEMM_THREADS = 192;
static constexpr int MAX_STAGES = 8;

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
      v[j] = a * gelu_erf_f(gt);
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
  if (p.flags & LDX_EPI_SILU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
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

__global__ void __launch_bounds__(GEMM_THREADS, 1) ldx_gemm_kernel(const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);   // SWIZZLE_128B wants 1024-B alignment

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  const uint32_t b_stage_bytes = (uint32_t)p.bn * BK * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)stages * A_STAGE_BYTES;
  GemmBarriers* bars = reinterpret_cast<GemmBarriers*>(sB + (size_t)stages * b_stage_bytes);

  if (threadIdx.x == 0) {
    for (int i = 0; i < LDX_MAX_SRC; ++i) tma_prefetch_desc(&p.amap[i]);
    tma_prefetch_desc(&p.bmap);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
        const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tb = mt / (p.tiles_x * p.tiles_y);
        const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb * p.bb, n0 = nt * p.bn;
        int kcol = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const KSeg sg = p.seg[s];
          for (int ch = 0; ch < sg.nchunks; ++ch) {
            mbar_wait(&bars->empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&bars->full[stage], A_STAGE_BYTES + b_stage_bytes);
            tma_load_4d(sA + (size_t)stage * A_STAGE_BYTES, &p.amap[sg.map], &bars->full[stage], sg.c0 + ch * BK,
                        x0 + sg.dx, y0 + sg.dy, b0);
            tma_load_2d(sB + (size_t)stage * b_stage_bytes, &p.bmap, &bars->full[stage], kcol, n0);
            kcol += BK;
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(BM, p.bn);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * p.acc_stride;
        for (int kb = 0; kb < p.nk; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + (size_t)stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sB + (size_t)stage * b_stage_bytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_f16_ss(d_tmem, umma_desc_kmajor_sw128(a_addr + k * 32), umma_desc_kmajor_sw128(b_addr + k * 32),
                        idesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit(&bars->empty[stage]);     // frees the smem stage once these MMAs have read it
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bars->tmem_full[as]);      // accumulator complete -> epilogue
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const bool geglu = (p.flags & LDX_EPI_GEGLU) != 0;
    const bool f32 = (p.flags & LDX_EPI_OUT_F32) != 0;
    const int n_out_tile = geglu ? p.bn / 2 : p.bn;
    const int osz = f32 ? 4 : 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
      const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tb = mt / (p.tiles_x * p.tiles_y);
      const int x = tx * p.bw + row % p.bw;
      const int y = ty * p.bh + (row / p.bw) % p.bh;
      const int b = tb * p.bb + row / (p.bw * p.bh);
      const bool valid = (x < p.W) && (y < p.H) && (b < p.B);
      const int64_t opix = p.o_off + (int64_t)b * p.o_sb + (int64_t)y * p.o_sy + (int64_t)x * p.o_sx;
      char* out_row = reinterpret_cast<char*>(p.out) + opix * osz;
      const __half* res_row =
          (p.res && valid) ? reinterpret_cast<const __half*>(p.res) + (((int64_t)b * p.H + y) * p.W + x) * p.ldr
                           : nullptr;
      const float* rowadd_row = (p.rowadd && valid) ? p.rowadd + (int64_t)b * p.ld_rowadd : nullptr;
      const int n_out0 = nt * n_out_tile;

      mbar_wait(&bars->tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t trow = tmem_base + as * p.acc_stride + ((uint32_t)(q * 32) << 16);
      for (int c = 0; c < n_out_tile; c += 32) {
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
      if (lane == 0) mbar_arrive(&bars->tmem_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
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
int encode_map_f16(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
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
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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

static int choose_bn(int N, int m_tiles, int flags, int sms) {
  const bool geglu = (flags & LDX_EPI_GEGLU) != 0;
  int best = 0;
  double best_cost = 1e30;
  for (int bn = 256; bn >= 32; bn -= 32) {
    if (geglu && bn % 64) continue;
    const int n_tiles = (N + bn - 1) / bn;
    const double waves = (double)(((long long)m_tiles * n_tiles + sms - 1) / sms);
    // cycles per 16-wide K step: tensor pipe bn/2, smem operand reads (128 + bn) * 32 B at 128 B/clk
    const double per_k = (bn / 2.0 > (128 + bn) / 4.0) ? bn / 2.0 : (128 + bn) / 4.0;
    const double cost = waves * (per_k + 6.0);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

int gemm_prepare(const ldx_gemm_desc_t* d, GemmLaunch* g) {
  LDX_CHECK_ARG(d && g, "null descriptor");
  LDX_CHECK_ARG(d->nsrc >= 1 && d->nsrc <= LDX_MAX_SRC, "nsrc %d out of range", d->nsrc);
  LDX_CHECK_ARG(d->nseg >= 1 && d->nseg <= LDX_MAX_SEG, "nseg %d out of range", d->nseg);
  LDX_CHECK_ARG(d->N >= 1 && d->W >= 1 && d->H >= 1 && d->B >= 1, "bad output extents");
  LDX_CHECK_ARG(d->w && d->out, "null weight/output pointer");
  GemmKParams& p = g->p;
  memset(&p, 0, sizeof(p));

  // tile geometry: 128 output pixels = bw x bh x bb
  int bw, bh, bb;
  if (d->H == 1 && d->B == 1) {
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
  int bn = d->bn ? d->bn : choose_bn(d->N, m_tiles, d->flags, sms);
  LDX_CHECK_ARG(bn >= 32 && bn <= 256 && bn % 32 == 0, "bn %d must be a multiple of 32 in [32,256]", bn);
  if (d->flags & LDX_EPI_GEGLU) {
    LDX_CHECK_ARG(bn % 64 == 0 && d->N % bn == 0, "GEGLU needs bn %% 64 == 0 and N %% bn == 0 (N=%d bn=%d)", d->N, bn);
    LDX_CHECK_ARG(!d->residual && !d->rowadd, "GEGLU epilogue does not take residual/rowadd");
  }
  p.bn = bn;
  p.n_tiles = (d->N + bn - 1) / bn;
  p.num_tiles = m_tiles * p.n_tiles;
  p.acc_stride = (uint32_t)pow2_ceil(bn < 32 ? 32 : bn);
  p.tmem_cols = 2 * p.acc_stride;

  int nk = 0;
  for (int s = 0; s < d->nseg; ++s) {
    const ldx_seg_t& sg = d->seg[s];
    LDX_CHECK_ARG(sg.src >= 0 && sg.src < d->nsrc && sg.nchunks >= 1, "bad segment %d", s);
    LDX_CHECK_ARG(sg.c0 % 8 == 0, "segment %d: c0 must be a multiple of 8", s);
    p.seg[s] = KSeg{sg.src, sg.c0, sg.dx, sg.dy, sg.nchunks};
    nk += sg.nchunks;
  }
  p.nseg = d->nseg;
  p.nk = nk;
  LDX_CHECK_ARG(d->ldw >= (int64_t)nk * BK, "weight pitch %lld < K %d", (long long)d->ldw, nk * BK);

  const uint32_t b_stage = (uint32_t)bn * BK * 2;
  int stages = (int)((200 * 1024) / (A_STAGE_BYTES + b_stage));
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages > nk + 1) stages = nk + 1 > 2 ? nk + 1 : 2;
  p.stages = stages;
  g->smem_bytes = stages * (A_STAGE_BYTES + (int)b_stage) + (int)sizeof(GemmBarriers) + 1024;
  if (g->smem_bytes < 120 * 1024) g->smem_bytes = 120 * 1024;   // keep one CTA per SM: it owns up to all of TMEM
  g->grid = p.num_tiles < sms ? p.num_tiles : sms;

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
  {
    const uint64_t dims[2] = {(uint64_t)nk * BK, (uint64_t)d->N};
    const uint64_t str[1] = {(uint64_t)d->ldw * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)bn};
    int rc = encode_map_f16(&p.bmap, d->w, 2, dims, str, box);
    if (rc) return rc;
  }

  p.out = d->out;
  p.o_sx = d->o_sx; p.o_sy = d->o_sy; p.o_sb = d->o_sb; p.o_off = d->o_off;
  p.bias = d->bias;
  p.rowadd = d->rowadd;
  p.ld_rowadd = d->ld_rowadd;
  p.res = d->residual;
  p.ldr = d->ldr;
  p.flags = d->flags;
  const int osz = (d->flags & LDX_EPI_OUT_F32) ? 4 : 2;
  const int n_out = (d->flags & LDX_EPI_GEGLU) ? d->N / 2 : d->N;
  bool vec = (n_out % 8 == 0) && ((reinterpret_cast<uintptr_t>(d->out) + (uint64_t)d->o_off * osz) % 16 == 0) &&
             ((d->o_sx * osz) % 16 == 0) && ((d->o_sy * osz) % 16 == 0) && ((d->o_sb * osz) % 16 == 0);
  if (d->residual) vec = vec && (reinterpret_cast<uintptr_t>(d->residual) % 16 == 0) && ((d->ldr * 2) % 16 == 0);
  p.vec_ok = vec ? 1 : 0;
  return LDX_OK;
}

int gemm_launch(const GemmLaunch* g, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    LDX_CUDA(cudaFuncSetAttribute(ldx_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  ldx_gemm_kernel<<<g->grid, GEMM_THREADS, g->smem_bytes, stream>>>(g->p);
  LDX_CUDA(cudaGetLastError());
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

extern "C" int ldx_gemm(const ldx_gemm_desc_t* d, ldx_stream_t stream) {
  GemmLaunch g;
  int rc = gemm_prepare(d, &g);
  if (rc) return rc;
  return gemm_launch(&g, as_stream(stream));
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
