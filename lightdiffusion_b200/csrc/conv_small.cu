// 3x3 stride-1 pad-1 convolution with a handful of output channels (Cout <= 5): the VAE decoder's conv_out
// 128 -> 3 at full image resolution (LD.py:3880-3882, fused with VAE.process_output, LD.py:6296-6298) and the
// UNet's out conv 320 -> 4 (LD.py:5655-5661).  These are HBM-bound: 2 * 9 * Cin * Cout FLOPs per Cin * 2 input
// bytes, and the implicit-GEMM kernel (gemm.cu) spends nine shifted TMA boxes per input chunk on an N = 3 tile.
//
// Here every input pixel is fetched ONCE and the tensor core produces, per INPUT pixel, the 9 * Cout partial
// products it contributes to its nine neighbours (a 1x1 "scatter" GEMM):
//     P[r, tap * Cout + o] = sum_c x[r, c] * w[o, c, tap]             r = input pixel of a 16 x 16 halo tile
//     out[p, o]            = bias[o] + sum_tap P[p + off(tap), tap * Cout + o]
// * A operand: ONE 4-D TMA box (64 channels, 16, 16, 1) per 64-channel chunk = 256 rows in the K-major
//   SWIZZLE_128B layout; rows 0..127 and 128..255 feed two M = 128 tcgen05.mma chains.  Out-of-image pixels
//   arrive as zeros (TMA out-of-bounds fill) = the conv's zero padding.
// * B operand: the repacked weights [NTP = 9 * Cout rounded up to 16][Cin], loaded once per CTA.
// * D: 2 x NTP fp32 columns in TMEM, double buffered across tiles.
// * Epilogue (8 warps, thread = input pixel): TMEM -> registers -> fp32 partials in shared memory -> named
//   barrier -> thread = OUTPUT pixel gathers its nine taps, adds the bias and stores fp16 NHWC and / or the
//   post-processed fp32 image.  The 14 x 14 interior of a tile is output; the halo is re-read by the neighbour
//   tiles (x neighbours run concurrently, y neighbours ~40 tiles later: L2 hits).
// Roofline: HBM; algorithmic bytes = B * H * W * (Cin * 2 + output bytes), read once.
#include "common.cuh"
#include "internal.h"

namespace ldx {

int encode_map_f16(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

static constexpr int CS_TW = 16, CS_TH = 16;                 // halo tile (input pixels)
static constexpr int CS_ROWS = CS_TW * CS_TH;                // 256 = two M = 128 accumulators
static constexpr int CS_A_BYTES = CS_ROWS * 128;             // one 64-channel chunk of a tile: 32 KiB
static constexpr int CS_THREADS = 64 + 256;                  // TMA warp, MMA warp, 8 epilogue warps
static constexpr int CS_MAX_STAGES = 6;
static constexpr int CS_MAX_COUT = 5;                        // 9 * 5 = 45 <= 48 partial columns

struct alignas(64) ConvSmallParams {
  CUtensorMap amap;      // x: (Cin, W, H, B), box (64, 16, 16, 1)
  CUtensorMap wmap;      // packed weights: (K64, NTP), box (64, NTP)
  const float* bias;     // [Cout] or null
  __half* out16;         // fp16 NHWC, pixel pitch ldo16 (or null)
  float* out_img;        // fp32 NHWC [B, H, W, Cout] after the post transform (or null)
  int64_t ldo16;
  int32_t B, H, W, Cout, NT, NTP, SS, nchunk, stages, post;
  int32_t tiles_x, tiles_y, num_tiles;
  uint32_t tmem_cols;
};

struct __align__(8) ConvSmallBarriers {
  uint64_t full[CS_MAX_STAGES], empty[CS_MAX_STAGES], tmem_full[2], tmem_empty[2], w_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(CS_THREADS, 1) ldx_conv3x3_small_kernel(const __grid_constant__ ConvSmallParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = p.stages;
  const uint32_t w_chunk_bytes = (uint32_t)p.NTP * 128;
  uint8_t* sA = smem;                                               // stages x 32 KiB
  uint8_t* sW = sA + (size_t)stages * CS_A_BYTES;                   // nchunk x NTP x 128 B
  float* sS = reinterpret_cast<float*>(sW + (size_t)p.nchunk * w_chunk_bytes);   // 2 x 256 x SS partials
  ConvSmallBarriers* bars = reinterpret_cast<ConvSmallBarriers*>(sS + 2 * CS_ROWS * p.SS);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.amap);
    tma_prefetch_desc(&p.wmap);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], 8);
    }
    mbar_init(&bars->w_full, 1);
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
  pdl_prologue();

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars->w_full, (uint32_t)p.nchunk * w_chunk_bytes);
      for (int c = 0; c < p.nchunk; ++c) tma_load_2d(sW + (size_t)c * w_chunk_bytes, &p.wmap, &bars->w_full, c * 64, 0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int ix0 = tx * (CS_TW - 2) - 1, iy0 = ty * (CS_TH - 2) - 1;
      for (int c = 0; c < p.nchunk; ++c) {
        mbar_wait(&bars->empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->full[stage], CS_A_BYTES);
          tma_load_4d(sA + (size_t)stage * CS_A_BYTES, &p.amap, &bars->full[stage], c * 64, ix0, iy0, b);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected lane)
    const uint32_t idesc = umma_idesc_f16(128, p.NTP);
    mbar_wait(&bars->w_full, 0);
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      mbar_wait(&bars->tmem_empty[as], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(as * 2) * p.NTP;
      for (int c = 0; c < p.nchunk; ++c) {
        mbar_wait(&bars->full[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + (size_t)stage * CS_A_BYTES);
        const uint32_t b_addr = smem_u32(sW + (size_t)c * w_chunk_bytes);
        if (elect_one()) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(d0 + hf * p.NTP, umma_desc_kmajor_sw128(a_addr + hf * (128 * 128) + k * 32),
                          umma_desc_kmajor_sw128(b_addr + k * 32), idesc, (uint32_t)((c | k) != 0));
          }
          umma_commit(&bars->empty[stage]);
          if (c == p.nchunk - 1) umma_commit(&bars->tmem_full[as]);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3, hf = (warp - 2) >> 2;
    const int r = hf * 128 + q * 32 + lane;                 // this thread's pixel of the halo tile
    const int tx_in = r % CS_TW, ty_in = r / CS_TW;
    const bool interior = tx_in >= 1 && tx_in <= CS_TW - 2 && ty_in >= 1 && ty_in <= CS_TH - 2;
    const int SS = p.SS, NT = p.NT, Cout = p.Cout;
    float bias[CS_MAX_COUT];
#pragma unroll
    for (int o = 0; o < CS_MAX_COUT; ++o) bias[o] = (p.bias && o < Cout) ? __ldg(p.bias + o) : 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int ox = tx * (CS_TW - 2) - 1 + tx_in, oy = ty * (CS_TH - 2) - 1 + ty_in;
      float* S = sS + (size_t)as * CS_ROWS * SS;
      mbar_wait(&bars->tmem_full[as], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 2 + hf) * p.NTP;
      float* srow = S + (size_t)r * SS;
      for (int c0 = 0; c0 < p.NTP; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < NT) srow[c0 + j] = __uint_as_float(v[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[as]);
      named_bar_sync(1, 256);            // all partials of this tile are in shared memory
      if (interior && ox < p.W && oy < p.H) {
        float acc[CS_MAX_COUT];
#pragma unroll
        for (int o = 0; o < CS_MAX_COUT; ++o) acc[o] = bias[o];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float* src = S + (size_t)(r + (ky - 1) * CS_TW + (kx - 1)) * SS + (ky * 3 + kx) * Cout;
#pragma unroll
            for (int o = 0; o < CS_MAX_COUT; ++o)
              if (o < Cout) acc[o] += src[o];
          }
        }
        const int64_t pix = ((int64_t)b * p.H + oy) * p.W + ox;
#pragma unroll
        for (int o = 0; o < CS_MAX_COUT; ++o) {
          if (o < Cout) {
            const __half hv = __float2half_rn(acc[o]);       // the reference's conv output dtype
            if (p.out16) p.out16[pix * p.ldo16 + o] = hv;
            if (p.out_img) {
              float v = __half2float(hv);
              if (p.post == 1) v = fminf(fmaxf((v + 1.0f) / 2.0f, 0.0f), 1.0f);
              p.out_img[pix * Cout + o] = v;
            }
          }
        }
      }
      // S[as] is rewritten two tiles from now; every thread passes the NEXT tile's barrier before that
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// weights [Cout, Cin, 3, 3] fp16 -> [NTP][K64] fp16, row tap * Cout + o, zero padded
__global__ void pack_conv3x3_small_kernel(const __half* __restrict__ w, __half* __restrict__ out, int cout, int cin,
                                          int ntp, int k64) {
  const int total = ntp * k64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % k64, n = i / k64;
    const int tap = n / cout, o = n % cout;
    out[i] = (tap < 9 && c < cin) ? w[((int64_t)o * cin + c) * 9 + tap] : __float2half(0.f);
  }
}

struct ConvSmallLaunch {
  ConvSmallParams p;
  int grid, smem_bytes;
};

static int conv_small_ntp(int cout) { return (9 * cout + 15) / 16 * 16; }

static int conv_small_prepare(const void* x, int B, int H, int W, int Cin, int64_t ldx_, const void* wp,
                              const float* bias, int Cout, void* out16, int64_t ldo16, float* out_img, int post,
                              ConvSmallLaunch* L) {
  LDX_CHECK_ARG(x && wp && (out16 || out_img), "conv3x3_small: null pointer");
  LDX_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && Cin >= 8 && Cin % 8 == 0, "conv3x3_small: bad extents (Cin %% 8)");
  LDX_CHECK_ARG(Cout >= 1 && Cout <= CS_MAX_COUT, "conv3x3_small: Cout %d not in [1, %d]", Cout, CS_MAX_COUT);
  LDX_CHECK_ARG(ldx_ >= Cin && ldx_ % 8 == 0, "conv3x3_small: input pixel pitch %lld", (long long)ldx_);
  LDX_CHECK_ARG(!out16 || ldo16 >= Cout, "conv3x3_small: output pitch %lld < Cout", (long long)ldo16);
  LDX_CHECK_ARG(post == 0 || post == 1, "conv3x3_small: post %d", post);
  ConvSmallParams& p = L->p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W; p.Cout = Cout;
  p.NT = 9 * Cout;
  p.NTP = conv_small_ntp(Cout);
  p.SS = (p.NT & 1) ? p.NT : p.NT + 1;            // odd row pitch: the 32 rows of a warp hit 32 banks
  p.nchunk = (Cin + 63) / 64;
  p.post = post;
  p.bias = bias;
  p.out16 = (__half*)out16;
  p.ldo16 = ldo16;
  p.out_img = out_img;
  p.tiles_x = (W + CS_TW - 3) / (CS_TW - 2);
  p.tiles_y = (H + CS_TH - 3) / (CS_TH - 2);
  p.num_tiles = p.tiles_x * p.tiles_y * B;
  p.tmem_cols = 4 * p.NTP <= 128 ? 128 : 256;
  const int fixed = p.nchunk * p.NTP * 128 + 2 * CS_ROWS * p.SS * 4 + (int)sizeof(ConvSmallBarriers) + 1024;
  int stages = (226 * 1024 - fixed) / CS_A_BYTES;
  if (stages > CS_MAX_STAGES) stages = CS_MAX_STAGES;
  LDX_CHECK_ARG(stages >= 2, "conv3x3_small: Cin %d too large for the resident weight tile", Cin);
  p.stages = stages;
  L->smem_bytes = stages * CS_A_BYTES + fixed;
  const int sms = num_sms();
  L->grid = p.num_tiles < sms ? p.num_tiles : sms;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)ldx_ * 2, (uint64_t)ldx_ * 2 * W, (uint64_t)ldx_ * 2 * W * H};
    const uint32_t box[4] = {64, CS_TW, CS_TH, 1};
    int rc = encode_map_f16(&p.amap, x, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t k64 = (uint64_t)p.nchunk * 64;
    const uint64_t dims[2] = {k64, (uint64_t)p.NTP};
    const uint64_t str[1] = {k64 * 2};
    const uint32_t box[2] = {64, (uint32_t)p.NTP};
    int rc = encode_map_f16(&p.wmap, wp, 2, dims, str, box);
    if (rc) return rc;
  }
  return LDX_OK;
}

static int conv_small_launch(const ConvSmallLaunch* L, cudaStream_t st) {
  static std::atomic<uint64_t> attr_done{0};
  int arc = set_max_dyn_smem_once(ldx_conv3x3_small_kernel, 227 * 1024, attr_done);
  if (arc) return arc;
  LDX_CUDA(launch_k(ldx_conv3x3_small_kernel, dim3(L->grid), dim3(CS_THREADS), (size_t)L->smem_bytes, st, 1, L->p));
  return LDX_OK;
}

// program support (api.cu).  p: x, w_packed, bias, out16, out_img ; i: B, H, W, Cin, Cout, post ; l: ldx, ldo16
int conv_small_prepare_op(const ldx_generic_op_t* op, void** handle) {
  ConvSmallLaunch* L = new ConvSmallLaunch;
  int rc = conv_small_prepare(op->p[0], op->i[0], op->i[1], op->i[2], op->i[3], op->l[0], op->p[1],
                              (const float*)op->p[2], op->i[4], (void*)op->p[3], op->l[1], (float*)op->p[4], op->i[5], L);
  if (rc) {
    delete L;
    return rc;
  }
  *handle = L;
  return LDX_OK;
}
int conv_small_run_handle(void* handle, cudaStream_t st) { return conv_small_launch((ConvSmallLaunch*)handle, st); }
void conv_small_free_handle(void* handle) { delete (ConvSmallLaunch*)handle; }

}  // namespace ldx

using namespace ldx;

extern "C" int ldx_conv3x3_small_packed_elems(int cout, int cin) {
  if (cout < 1 || cout > CS_MAX_COUT || cin < 1) return -1;
  return conv_small_ntp(cout) * ((cin + 63) / 64 * 64);
}

extern "C" int ldx_pack_conv3x3_small(const void* w, void* out, int cout, int cin, ldx_stream_t stream) {
  LDX_CHECK_ARG(w && out && cout >= 1 && cout <= CS_MAX_COUT && cin >= 1, "bad pack_conv3x3_small args");
  const int ntp = conv_small_ntp(cout), k64 = (cin + 63) / 64 * 64;
  pack_conv3x3_small_kernel<<<(ntp * k64 + 255) / 256, 256, 0, as_stream(stream)>>>((const __half*)w, (__half*)out, cout,
                                                                                  cin, ntp, k64);
  LDX_CUDA(cudaGetLastError());
  return LDX_OK;
}

extern "C" int ldx_conv3x3_small(const void* x, int B, int H, int W, int Cin, int64_t ldx_pix, const void* w_packed,
                                 const float* bias, int Cout, void* out16, int64_t ldo16, float* out_img, int post,
                                 ldx_stream_t stream) {
  ConvSmallLaunch L;
  int rc = conv_small_prepare(x, B, H, W, Cin, ldx_pix, w_packed, bias, Cout, out16, ldo16, out_img, post, &L);
  if (rc) return rc;
  return conv_small_launch(&L, as_stream(stream));
}
