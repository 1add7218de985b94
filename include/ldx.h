/* ldx.h — C ABI of libldx.so, the B200 (sm_100a) kernels behind the reference's operator surface.
 *
 * The reference (Aatrick/LightDiffusion, one file: LightDiffusion.py = "LD.py") has no native code
 * and no FFI; every FLOP on its hot path is a torch library call made from four late-bound Python
 * seams.  This header is what a Python shim binds with ctypes to replace those
 * calls; each entry point cites the reference call site it stands in for.
 *
 * Conventions
 *   - plain pointers are DEVICE pointers unless the name ends in _host; no torch types cross the ABI
 *   - activations are fp16, channels-last ("NHWC": [B, H, W, C], pixel pitch given in elements)
 *   - weights are fp16 [N, K] row-major (torch Linear layout); 3x3 conv weights are packed
 *     [Cout][tap = ky*3+kx][ceil64(Cin)] by ldx_pack_conv3x3
 *   - every call enqueues on the given cudaStream_t (void*; NULL = legacy default stream),
 *     returns 0 on success and a negative LDX_E_* otherwise; ldx_last_error_string() explains
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails
 */
#ifndef LDX_H_
#define LDX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDX_OK 0
#define LDX_E_INVALID (-1)   /* bad argument / unsupported shape */
#define LDX_E_CUDA (-2)      /* CUDA runtime or driver error */
#define LDX_E_NODEVICE (-3)  /* no sm_100 device */

typedef void* ldx_stream_t;

int ldx_version(void);
/* sizeof() of the ABI structs, for binding self-checks: 0 gemm_desc, 1 generic_op, 2 src, 3 seg. */
int ldx_sizeof(int what);
const char* ldx_last_error_string(void);
/* 0 if a CUDA device with compute capability 10.x is present. */
int ldx_device_check(void);

/* ------------------------------------------------------------------ implicit GEMM (tcgen05 + TMA)
 * One kernel serves F.linear (LD.py:2368), 1x1 and 3x3 F.conv2d (LD.py:2382; stride 1, stride 2,
 * nearest-2x-upsampled input, channel-concatenated input) as
 *     out[b,y,x,n] = epilogue( sum_seg sum_c src[seg.src][b, y+seg.dy, x+seg.dx, seg.c0+c] * w[n, k] )
 * where the K axis of w is the concatenation of the segments' channel ranges, each padded to 64.
 * Out-of-range (y+dy, x+dx, c) read as zero (conv zero padding).
 */
typedef struct {
  const void* ptr;          /* fp16 */
  int32_t C, W, H, B;       /* extents seen by the kernel */
  int64_t sw, sh, sb;       /* strides in elements for W, H, B (channel stride is 1) */
} ldx_src_t;

typedef struct {
  int32_t src;              /* index into ldx_gemm_desc_t.src */
  int32_t c0;               /* first channel */
  int32_t dx, dy;           /* spatial shift of this tap */
  int32_t nchunks;          /* number of 64-channel K chunks */
} ldx_seg_t;

#define LDX_MAX_SRC 4
#define LDX_MAX_SEG 16

#define LDX_EPI_SILU 1      /* out = silu(acc + bias)            (time_embed, LD.py:5380-5388) */
#define LDX_EPI_GEGLU 2     /* out[:, j] = a * gelu_erf(gate)     (GEGLU.forward, LD.py:4513-4515);
                               weight rows must be interleaved by ldx_pack_geglu; out has N/2 columns */
#define LDX_EPI_OUT_F32 4   /* store fp32 instead of fp16 */
#define LDX_EPI_LRELU 32    /* out = max(x, 0.2 x), x = acc + bias (+ residual)  (ESRGAN conv_block act, LD.py:6767-6770) */
#define LDX_EPI_RELU 64     /* out = max(x, 0), x = acc + bias (+ residual)  (TAESD Block, LD.py:695-703) */
#define LDX_EPI_QUICK_GELU 16 /* out = x * sigmoid(1.702 x), x = acc + bias  (CLIPMLP quick_gelu, LD.py:4296-4321) */
#define LDX_EPI_ROWNORM 8   /* LayerNorm of the A rows folded into the GEMM (F.layer_norm LD.py:2403 followed by
                               F.linear LD.py:2368): with w' = w * gamma (caller-folded), gsum[n] = sum_k w'[n,k] and
                               bias' = bias + w @ beta,  out = rstd_r * (acc - mean_r * gsum[n]) + bias'[n], where
                               (mean_r, rstd_r) come from the row statistics the producing GEMM left in rowstats_in */

typedef struct {
  ldx_src_t src[LDX_MAX_SRC];
  int32_t nsrc;
  ldx_seg_t seg[LDX_MAX_SEG];
  int32_t nseg;
  const void* w;            /* fp16 [N, ldw] */
  int64_t ldw;
  int32_t N;
  int32_t W, H, B;          /* output domain; a plain GEMM with M rows uses W=M, H=1, B=1 */
  void* out;
  int64_t o_sx, o_sy, o_sb, o_off;   /* out element address = o_off + b*o_sb + y*o_sy + x*o_sx + n */
  const float* bias;        /* [N] fp32 or NULL */
  const float* rowadd;      /* [B, ld_rowadd] fp32 or NULL: added per batch index (timestep embedding
                               broadcast h + emb_out[:, :, None, None], LD.py:5279-5285) */
  int64_t ld_rowadd;
  const void* residual;     /* fp16 [(b*H+y)*W+x, ldr] or NULL (x + h, LD.py:5287; x += n, LD.py:4142) */
  int64_t ldr;
  int32_t flags;            /* LDX_EPI_* */
  int32_t bn;               /* N tile (multiple of 32, <= 256); 0 = choose */
  float* rowstats_out;      /* NULL, or [rows, ldx_gemm_rowstats_parts(d), 2] fp32: per output row the (sum, sum of
                               squares) of this GEMM's outputs, one pair per column part (summed by the consumer in a
                               fixed order: bit-reproducible).  Needs fp16 outputs with TMA-compatible strides. */
  const float* rowstats_in; /* LDX_EPI_ROWNORM: [rows, rowstats_in_parts, 2] written by the GEMM that produced A */
  int32_t rowstats_in_parts;
  float ln_eps;             /* LDX_EPI_ROWNORM: LayerNorm eps */
  const float* ln_gsum;     /* LDX_EPI_ROWNORM: [N] fp32 (packed like the bias for LDX_EPI_GEGLU) */
  int32_t ln_c;             /* LDX_EPI_ROWNORM: normalised width C (the A row length before K padding) */
  float* colstats_out;      /* NULL, or [B, colstats_slots, N, 2] fp32: for every 32-row block ("slot") of a batch
                               element and every output column the (sum, sum of squares) of the fp16 outputs, taken
                               from the epilogue's staging tile.  ldx_group_norm_cs reduces them to GroupNorm
                               statistics, so the normalisation needs no statistics pass over the tensor.  TMA-store
                               epilogue only (fp16 output), not with split-K / GEGLU; see ldx_gemm_colstats_slots */
  int32_t colstats_slots;   /* slots per batch element in the buffer */
  int32_t colstats_slot_base; /* first slot this launch writes (the four phase GEMMs of an upsample conv share one) */
  int32_t rowstats_out_parts; /* 0 = unchecked; else the [rows, parts, 2] extent of rowstats_out: the launch is
                               rejected unless it equals ldx_gemm_rowstats_parts(d) */
  /* Halo-tile form of a 3x3 stride-1 pad-1 convolution (F.conv2d, LD.py:2382) over the channel concatenation of
   * src[0 .. halo_nsrc-1] (th.cat([h, hsp], 1), LD.py:5749): per 64-channel chunk ONE (10 x 18)-pixel halo tile is
   * fetched and the tensor core reads its nine tap views in place, so the operand is read ~1.4x instead of 9x.
   * The weight matrix is the usual ldx_pack_conv3x3 layout over the concatenated channels ([N][9][halo_cp],
   * halo_cp = sum of ceil64(C_i); every source but the last needs C_i % 64 == 0); seg[0 .. nseg) (nseg may be 0)
   * then lists only extra 1x1 K segments (a ResBlock's skip_connection, LD.py:5267) whose weight columns follow at
   * 9 * halo_cp.  Needs W % 8 == 0, fp16 output with TMA-compatible strides; 0 = off. */
  int32_t halo_nsrc;
  /* GroupNorm (+ SiLU) folded into that operand path (ResBlock in_layers / out_layers, LD.py:5220-5261; VAE
   * ResnetBlock, LD.py:3562-3576): NULL, or [B, gn_ctot, 2] fp32 (scale, shift) per (sample, concatenated channel)
   * from ldx_group_norm_coef; the landed tile is rewritten in shared memory as fp16(silu?(x * scale + shift)) before the
   * MMA reads it (out-of-image pixels stay zero: the conv pads the NORMALISED map).  halo_nsrc > 0 only. */
  const float* gn_coef;
  int32_t gn_ctot;
  int32_t gn_silu;
  /* Per-sample weights and bias (0 = shared): batch element b reads rows [b * w_batch_rows, b * w_batch_rows + N) of w and
   * bias + b * bias_batch_stride.  This is how a GroupNorm in front of a 1x1 conv (SpatialTransformer.norm -> proj_in,
   * LD.py:4242-4250) is folded into the conv: ldx_gn_fold_weights writes w'[b] = w * scale[b] and bias'[b], and the conv
   * reads the un-normalised map.  Needs tiles that do not straddle batch elements (H * W a multiple of the 128 / 256-row
   * tile), the fp16 TMA-store epilogue, no split-K, no halo sources. */
  int32_t w_batch_rows;
  int32_t bias_batch_stride;
} ldx_gemm_desc_t;

int ldx_gemm(const ldx_gemm_desc_t* d, ldx_stream_t stream);
/* Number of column parts the GEMM described by d writes per row into rowstats_out (depends on the tile
 * configuration the library picks for d); < 0 on error. */
int ldx_gemm_rowstats_parts(const ldx_gemm_desc_t* d);
/* Number of 32-row slots per batch element the GEMM described by d writes into colstats_out (depends on the tile
 * geometry), or -1 when this launch cannot produce column statistics (direct-store epilogue, split-K, maps with
 * fewer than 32 pixels per batch element). */
int ldx_gemm_colstats_slots(const ldx_gemm_desc_t* d);

/* One-time weight layout transforms (ldx_pack_weights_* of SURVEY.md §8b). */
/* torch conv weight [Cout, Cin, 3, 3] fp16 -> [Cout, 9 * ceil64(Cin)] fp16 (tap-major, zero padded). */
int ldx_pack_conv3x3(const void* w_oihw, void* w_packed, int cout, int cin, ldx_stream_t stream);
/* Same, for "nearest 2x upsample then conv3x3" (Upsample1.forward LD.py:5142-5152, Upsample.forward
 * LD.py:3507-3511): folds the 3x3 taps onto the low-resolution grid -> 4 output-parity phases of
 * 2x2 taps: [4 phases][Cout][4 * ceil64(Cin)] fp16, summed in fp32. */
int ldx_pack_upconv3x3(const void* w_oihw, void* w_packed, int cout, int cin, ldx_stream_t stream);
/* 3x3 stride-1 pad-1 convolution with 1..5 output channels: the VAE decoder's conv_out 128 -> 3 (LD.py:3880-3882)
 * and the UNet's out conv 320 -> 4 (LD.py:5655-5661), through F.conv2d / _conv_forward (LD.py:2382).  HBM bound: every
 * input pixel is read once (16 x 16 halo tiles by TMA), the tensor core produces the 9 * Cout per-tap partial
 * products of each INPUT pixel and the epilogue gathers the nine taps of each output pixel.
 *   x        fp16 NHWC [B, H, W, Cin], pixel pitch ldx_pix elements (16-byte aligned, Cin % 8 == 0)
 *   w_packed ldx_pack_conv3x3_small(w_oihw): [ceil16(9 * Cout)][ceil64(Cin)] fp16, row tap * Cout + o
 *   out16    fp16 NHWC, pixel pitch ldo16 (first Cout channels written), or NULL
 *   out_img  fp32 NHWC [B, H, W, Cout], or NULL; post = 1 applies VAE.process_output to the fp16-rounded result:
 *            clamp((v + 1) / 2, 0, 1) (LD.py:6296-6298, 6377-6380); post = 0 stores it as is */
int ldx_conv3x3_small(const void* x, int B, int H, int W, int Cin, int64_t ldx_pix, const void* w_packed,
                      const float* bias, int Cout, void* out16, int64_t ldo16, float* out_img, int post,
                      ldx_stream_t stream);
int ldx_pack_conv3x3_small(const void* w_oihw, void* w_packed, int cout, int cin, ldx_stream_t stream);
/* number of fp16 elements ldx_pack_conv3x3_small writes (-1: unsupported Cout) */
int ldx_conv3x3_small_packed_elems(int cout, int cin);
/* The feed-forward of BasicTransformerBlock in one launch (x + ff(norm3(x)), LD.py:4155-4162; FeedForward / GEGLU,
 * LD.py:4508-4536) for C = 320:  out = x + W2 . (a * gelu_erf(g)) + b2 with [a | g] = LayerNorm(x) . W1^T + b1.  The 4C-wide
 * intermediate stays on chip (GEGLU epilogue -> shared memory -> operand of the second GEMM).
 *   x          fp16 [M, ldx]  (also the residual)            out   fp16 [M, ldo]
 *   w1_packed  ldx_pack_geglu(w1 * gamma, ..., half_bn = 64): [8C, ldw1] fp16, LayerNorm folded like LDX_EPI_ROWNORM
 *   bias1, gsum1  fp32 [8C] in the same packed order (bias + w1 @ beta, rowsum(w1 * gamma))
 *   w2 fp16 [C, ldw2] (F = 4C columns), bias2 fp32 [C]
 *   rowstats   [M, parts, 2] fp32 (sum, sum of squares) of x's rows, written by the GEMM that produced x */
int ldx_ff_geglu_fused(const void* x, int64_t ldx_, int M, const void* w1_packed, int64_t ldw1, const float* bias1,
                       const float* gsum1, const void* w2, int64_t ldw2, const float* bias2, const float* rowstats,
                       int parts, float eps, int C, int F, void* out, int64_t ldo, ldx_stream_t stream);
/* GroupNorm folded into the 1x1 conv / linear that consumes it (SpatialTransformer.norm + proj_in, LD.py:4242-4250):
 *   w_out[b, n, c]  = fp16(w[n, c] * scale[b, c])                               (coef[b, c] = (scale, shift) from ldx_group_norm_coef)
 *   bias_out[b, n]  = bias[n] + sum_c (w[n, c] * beta[c] - mean[b, g(c)] * w_out[b, n, c])
 * so that  w_out[b] . x + bias_out[b]  ==  w . (x * scale + shift) + bias  with the rounding of w_out cancelled against the
 * group mean (mean is recovered from beta, shift and scale).  w [N, ldw] fp16, w_out [B, N, ldw] fp16, C % 8 == 0,
 * C <= 2048, groups = 32.  One CTA per (32 output rows, sample); fixed summation order. */
int ldx_gn_fold_weights(const void* w, int64_t ldw, const float* bias, const float* beta, const float* coef, int B, int N,
                        int C, void* w_out, float* bias_out, ldx_stream_t stream);
/* GEGLU proj weight [2*F, K] (+ bias[2*F]) -> rows interleaved in blocks of `half_bn` so that one
 * output tile holds matching value/gate columns. */
int ldx_pack_geglu(const void* w, const float* bias, void* w_packed, float* bias_packed, int two_f, int k,
                   int half_bn, ldx_stream_t stream);

/* ------------------------------------------------------------------ normalisation (HBM bound)
 * GroupNorm(32, C) over NHWC input, optionally over the channel concatenation of two tensors
 * (th.cat([h, hsp], 1) LD.py:5749 followed by ResBlock in_layers, LD.py:5220-5232), optionally
 * followed by SiLU.  Replaces torch.nn.GroupNorm.forward (LD.py:2390-2395, 3931-3939) + nn.SiLU.
 * stats: fp32 workspace of B * LDX_GN_STATS_FLOATS floats (per-slab partial sums).  */
#define LDX_GN_STATS_FLOATS 4096
int ldx_group_norm(const void* x0, int c0, int64_t ld0, const void* x1, int c1, int64_t ld1,
                   int B, int HW, float eps, const float* gamma, const float* beta, int silu,
                   void* y, int64_t ldy, float* stats, ldx_stream_t stream);
/* The same normalisation with the statistics taken from the column partials the producing GEMM(s) wrote
 * (ldx_gemm_desc_t.colstats_out, [B, slots, C, 2] per source): a reduction over B * slots * C pairs replaces the
 * statistics pass over the tensor, so GroupNorm moves its algorithmic bytes (one read, one write).  Maps small enough
 * for the single-launch path ignore the partials. */
int ldx_group_norm_cs(const void* x0, int c0, int64_t ld0, const void* x1, int c1, int64_t ld1, int B, int HW,
                      float eps, const float* gamma, const float* beta, int silu, void* y, int64_t ldy,
                      float* stats, const float* colstats0, int slots0, const float* colstats1, int slots1,
                      ldx_stream_t stream);
/* GroupNorm statistics only: coef[b, c] = (rstd_g * gamma_c, beta_c - mean_g * rstd_g * gamma_c) as fp32 [B, C, 2]
 * for a consumer that applies the normalisation itself (ldx_gemm_desc_t.gn_coef).  Same statistics, summation order
 * and rounding as ldx_group_norm / ldx_group_norm_cs (colstats0 may be NULL).  One launch: the last CTA of each sample
 * (atomic ticket in `tickets`, B zero-initialised uint32 that the call leaves zeroed) finalises. */
int ldx_group_norm_coef(const void* x0, int c0, int64_t ld0, const void* x1, int c1, int64_t ld1, int B, int HW,
                        float eps, const float* gamma, const float* beta, float* stats, uint32_t* tickets,
                        float* coef, const float* colstats0, int slots0, const float* colstats1, int slots1,
                        ldx_stream_t stream);
/* LayerNorm over the last dim (F.layer_norm, LD.py:2403), eps 1e-5, fp16 in/out. */
int ldx_layer_norm(const void* x, int64_t ldx_, int rows, int C, float eps, const float* gamma,
                   const float* beta, void* y, int64_t ldy, ldx_stream_t stream);

/* ------------------------------------------------------------------ attention (tcgen05 flash)
 * optimized_attention(q, k, v, heads) (LD.py:3966-3988): softmax(q k^T / sqrt(d)) v per head, no
 * mask.  q [B, Nq, heads*d], k/v [B, Nk, heads*d] fp16 with row pitches ldq/ldk/ldv/ldo (elements).
 * d in {40, 64, 80, 128, 160}. */
int ldx_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                  void* o, int64_t ldo, int B, int Nq, int Nk, int heads, int d, ldx_stream_t stream);

/* The same with the CLIP text tower's causal mask (CLIPTextModel_.forward, LD.py:4443-4449: -inf above the
 * diagonal, i.e. key j attends to query i only for j <= i), self-attention over N tokens; d = 64, N <= 256. */
int ldx_attention_causal(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                         void* o, int64_t ldo, int B, int N, int heads, int d, ldx_stream_t stream);

/* ------------------------------------------------------------------ data movement / small ops */
/* x fp32 NCHW [B,C,H,W] * (1/sqrt(sigma_b^2+1)) -> fp16 NHWC [B,H,W,Cpad], extra channels zero
 * (EPS.calculate_input LD.py:1259-1261 + the .to(dtype) of apply_model LD.py:5842). sigma may be NULL. */
int ldx_nchw_to_nhwc_scaled(const float* x, const float* sigma, void* y, int B, int C, int H, int W,
                            int Cpad, ldx_stream_t stream);
/* timestep = argmin_i |log sigma - log_sigmas[i]| (ModelSamplingDiscrete.timestep LD.py:1336-1339),
 * then [cos(t f) | sin(t f)] as fp16 [B, dim] (timestep_embedding LD.py:803-812). */
int ldx_timestep_embedding(const float* sigma, const float* log_sigmas, int n_sigmas, void* emb, int B,
                           int dim, ldx_stream_t stream);
/* nearest-neighbour 2x upsample, NHWC fp16 (F.interpolate mode="nearest", LD.py:5149). */
int ldx_upsample_nearest2x(const void* x, void* y, int B, int H, int W, int C, ldx_stream_t stream);
/* eps fp16 NHWC [B,H,W,ldc] (first C channels) -> denoised fp32 NCHW = x - eps * sigma_b
 * (EPS.calculate_denoised LD.py:1263-1265, .float() LD.py:5859). */
int ldx_eps_to_denoised(const void* eps, int64_t lde, const float* x, const float* sigma, float* den,
                        int B, int C, int H, int W, ldx_stream_t stream);
/* Same, fused with classifier-free guidance: eps holds the [uncond..., cond...] batch of 2b samples,
 * x/sigma are the b-sample latents; out = u + (c - u) * scale on the denoised predictions
 * (calc_cond_batch chunk LD.py:2569 + cfg_function LD.py:2594-2606). */
int ldx_cfg_denoised(const void* eps, int64_t lde, const float* x, const float* sigma, float scale,
                     float* den, int b, int C, int H, int W, ldx_stream_t stream);
/* The whole tail of one guided DPM++ 2M step in ONE launch (sampling_function LD.py:2569-2606 followed by the
 * update of sample_dpmpp_2m_sde at eta = 0, LD.py:1209-1236): eps holds the [uncond..., cond...] batch of 2b
 * samples (fp16 NHWC, pitch lde); den = u + (c - u) * cfg_scale on the denoised predictions x - eps * sigma;
 *   x <- c_x * x + c_d * den [+ c_o * (den - old_denoised) when have_old]     (x <- den when last != 0)
 *   old_denoised <- den ; denoised_out <- den (optional, may be NULL: callbacks / previews)
 *   sigma_io[0..b) <- sigma_next (optional: the sigma vector the next UNet evaluation reads)
 * x, old_denoised, denoised_out: fp32 NCHW [b, C, H, W], in place. */
int ldx_cfg_dpmpp2m_step(const void* eps, int64_t lde, float* x, float* old_denoised, float* denoised_out,
                         float sigma, float cfg_scale, float c_x, float c_d, float c_o, int have_old, int last,
                         float sigma_next, float* sigma_io, int b, int C, int H, int W, ldx_stream_t stream);
/* cfg_function on two fp32 predictions (LD.py:2605). */
int ldx_cfg_combine(const float* cond, const float* uncond, float scale, float* out, int64_t n,
                    ldx_stream_t stream);
/* decoded fp16 NHWC [B,H,W,ldc] (first 3 channels) -> clamp((v+1)/2, 0, 1) fp32 NHWC [B,H,W,3]
 * (VAE.process_output + movedim, LD.py:6296-6298, 6377-6380). */
int ldx_vae_postprocess(const void* dec, int64_t ldc, float* img, int64_t npix, ldx_stream_t stream);
/* Same conversion quantised to uint8 NHWC [B,H,W,3] the way the reference saves / gathers images
 * (np.clip(255 * img, 0, 255).astype(uint8), SaveImage LD.py:6750-6751): 4x fewer bytes for the final gather. */
int ldx_vae_postprocess_u8(const void* dec, int64_t ldc, void* img, int64_t npix, ldx_stream_t stream);
/* VAE.encode input side: pixels fp32 NHWC [npix, 3] in [0,1] -> 2x - 1, fp16 NHWC [npix, 8] with channels 3..7
 * zero (VAE.process_input + movedim + cast, LD.py:6295, 6385, 6402-6405). out must be 16-byte aligned. */
int ldx_vae_preprocess(const float* pixels, void* out, int64_t npix, ldx_stream_t stream);
/* DiagonalGaussianDistribution.sample (LD.py:166-179): moments fp32 NHWC [b*hw, 2*zc] (mean | logvar), noise
 * fp32 NCHW [b, zc, hw] or NULL (posterior mean) -> z = mean + exp(0.5 * clamp(logvar, -30, 20)) * noise,
 * fp32 NCHW. */
int ldx_vae_sample_posterior(const float* moments, const float* noise, float* z, int b, int zc, int64_t hw,
                             ldx_stream_t stream);
/* One axis of bislerp (LD.py:429-518, the hires-fix latent upscale of LatentUpscale, LD.py:6639-6654): for the
 * fp32 tensor viewed as [n, c, outer, len_old, inner], out[n, :, o, j, i] = slerp(in[n, :, o, idx1[j], i],
 * in[n, :, o, idx2[j], i], ratio[j]) over the channel vectors, j < len_new, with the reference's edge cases (zero
 * norms, dot > 1 - 1e-5 -> first vector, dot < -1 + 1e-5 -> lerp).  ratio / idx1 / idx2 are device arrays of
 * len_new entries (generate_bilinear_data, LD.py:464-485).  1 <= c <= 16. */
int ldx_bislerp_axis(const float* in, float* out, const float* ratio, const int32_t* idx1, const int32_t* idx2,
                     int n, int c, int64_t outer, int len_old, int len_new, int64_t inner, ldx_stream_t stream);
/* CLIPEmbeddings.forward (LD.py:4409-4410): out[r, :] = token_embedding[tokens[r], :] + position_embedding[r % T, :];
 * fp32 tables [vocab, C] / [T, C], int32 token ids (the caller guarantees 0 <= id < vocab), fp16 out [rows, ldo]. */
int ldx_clip_embed(const int32_t* tokens, const float* tok_emb, const float* pos_emb, void* out, int64_t ldo,
                   int rows, int T, int C, ldx_stream_t stream);
/* out[r, :] = softmax(scale * s[r, :]): fp32 scores in, fp16 probabilities out (the VAE AttnBlock's
 * single-head d=512 attention runs as GEMM / softmax / GEMM, LD.py:3591-3602). */
int ldx_softmax_rows(const float* s, int64_t lds, void* out, int64_t ldo, int64_t rows, int n, float scale,
                     ldx_stream_t stream);

/* ------------------------------------------------------------------ sampler updates (HBM bound)
 * One DPM++ 2M step (sample_dpmpp_2m_sde with eta=0, LD.py:1209-1236), in place on x:
 *   x = c_x * x + c_d * denoised + c_o * (denoised - old_denoised)        (c_o = 0 on the first step)
 * The host computes c_* from the sigmas exactly as the reference does. n = element count (fp32). */
int ldx_dpmpp2m_step(float* x, const float* denoised, const float* old_denoised, float c_x, float c_d,
                     float c_o, int64_t n, ldx_stream_t stream);
/* One Euler-ancestral step (LD.py:929-936), in place on x:
 *   x = x + (x - denoised) / sigma * (sigma_down - sigma) + noise * noise_scale   (noise may be NULL) */
int ldx_euler_ancestral_step(float* x, const float* denoised, const float* noise, float sigma,
                             float sigma_down, float noise_scale, int64_t n, ldx_stream_t stream);

/* DPM-Solver-12/23 with adaptive step size (sample_dpm_adaptive / DPMSolver, LD.py:976-1170), fp32 state:
 *   ldx_dpm_eps:      eps = (x - denoised) / sigma                               (DPMSolver.eps, LD.py:990-999)
 *   ldx_dpm_combine:  out = x - c1 * eps - c2 * (eps2 - eps)   (eps2 NULL: out = x - c1 * eps)
 *                     = u1 / u2 / x_2 / x_3 of dpm_solver_2_step and dpm_solver_3_step (LD.py:1001-1038) with the
 *                     host-computed scalars; products and differences are rounded separately, as in the reference
 *   ldx_dpm_error_norm: err = || (x_low - x_high) / max(atol, rtol * max(|x_low|, |x_prev|)) ||_2 / sqrt(n)
 *                     (LD.py:1094-1095) into the device float `err`; `workspace` holds LDX_DPM_ERROR_WS_DOUBLES
 *                     doubles; two launches, fixed summation order. */
#define LDX_DPM_ERROR_WS_DOUBLES 2048
int ldx_dpm_eps(const float* x, const float* denoised, float sigma, float* eps, int64_t n, ldx_stream_t stream);
int ldx_dpm_combine(const float* x, const float* eps, const float* eps2, float c1, float c2, float* out, int64_t n,
                    ldx_stream_t stream);
int ldx_dpm_error_norm(const float* x_low, const float* x_high, const float* x_prev, float atol, float rtol,
                       int64_t n, void* workspace, float* err, ldx_stream_t stream);

/* ------------------------------------------------------------------ recorded programs
 * A program is an ordered list of the calls above with all arguments frozen (TMA descriptors built
 * once).  ldx_program_run replays it on a stream; it is what a CUDA graph captures. */
enum {
  LDX_OP_GEMM = 1, LDX_OP_GROUP_NORM = 2, LDX_OP_LAYER_NORM = 3, LDX_OP_ATTENTION = 4,
  LDX_OP_NCHW_TO_NHWC = 5, LDX_OP_TIMESTEP_EMB = 6, LDX_OP_UPSAMPLE2X = 7, LDX_OP_EPS_TO_DENOISED = 8,
  LDX_OP_CFG_DENOISED = 9, LDX_OP_VAE_POST = 10, LDX_OP_SOFTMAX = 11,
  /* LDX_OP_ATTENTION takes i[5] = 1 for the causal variant (ldx_attention_causal) */
  LDX_OP_GROUP_NORM_COEF = 17, /* ldx_group_norm_coef.  p: x0, x1, gamma, beta, coef, stats, colstats0, colstats1, tickets ;
                                  i: c0, c1, B, HW, -, slots0, slots1 ; l: ld0, ld1 ; f: eps */
  LDX_OP_GN_FOLD_WEIGHTS = 19, /* ldx_gn_fold_weights.  p: w, bias, beta, coef, w_out, bias_out ; l: ldw ; i: B, N, C */
  LDX_OP_FF_FUSED = 18,      /* ldx_ff_geglu_fused.  p: x, w1_packed, bias1, gsum1, w2, bias2, rowstats, out ;
                                l: ldx, ldw1, ldw2, ldo ; i: M, C, F, parts ; f: eps */
  LDX_OP_CONV_SMALL = 16,    /* ldx_conv3x3_small.  p: x, w_packed, bias, out16, out_img ; i: B, H, W, Cin, Cout, post ;
                                l: pixel pitch of x, pixel pitch of out16 */
  LDX_OP_CLIP_EMBED = 15,    /* p: tokens, tok_emb, pos_emb, out ; l: ldo ; i: rows, T, C */
  LDX_OP_VAE_PRE = 13,       /* p: pixels, out ; l: npix */
  LDX_OP_VAE_SAMPLE = 14,    /* p: moments, noise, z ; l: n, hw ; i: zc */
  LDX_OP_DUP = 12            /* p: src, dst ; l: bytes.  dst[0:bytes] = dst[bytes:2*bytes] = src (device copies): the
                                [uncond, cond] batch of calc_cond_batch (torch.cat([x] * 2), LD.py:2545-2551) */
};
typedef struct {
  int32_t kind;
  int32_t i[16];
  int64_t l[8];
  float f[4];
  const void* p[10];
} ldx_generic_op_t;          /* argument slots for the non-GEMM ops, in the order of the C signature */

typedef struct ldx_program ldx_program_t;
ldx_program_t* ldx_program_create(void);
int ldx_program_add_gemm(ldx_program_t* prog, const ldx_gemm_desc_t* d);
int ldx_program_add_op(ldx_program_t* prog, const ldx_generic_op_t* op);
int ldx_program_num_launches(const ldx_program_t* prog);
int ldx_program_num_ops(const ldx_program_t* prog);
int ldx_program_run(ldx_program_t* prog, ldx_stream_t stream);
/* Replay with a CUDA event between ops; writes each op's device time (ms) to ms_out[n_ops]. Blocking. */
int ldx_program_run_timed(ldx_program_t* prog, ldx_stream_t stream, float* ms_out, int n_ops);
void ldx_program_destroy(ldx_program_t* prog);

#ifdef __cplusplus
}
#endif
#endif /* LDX_H_ */
