// HBM-bound kernels of the SD1.5 hot path: GroupNorm(+SiLU) over NHWC (optionally over a channel
// concat of two tensors), LayerNorm, layout conversion with EPS input scaling, timestep embedding,
// nearest upsample, eps -> denoised (+ CFG), VAE post-process, sampler updates.
// All are vectorised to 16-byte accesses and sized to a multiple of the SM count.
// Roofline: HBM; algorithmic bytes are stated per kernel.
#include "common.cuh"
#include "internal.h"

namespace ldx {

// ----------------------------------------------------------------------------- GroupNorm
// (Two one-pass variants were measured and dropped.  A cluster per (batch, 4-group channel slice) holding its
// pixels in shared memory: 55 vs 43 us at 16x4096x320 -- a 40-channel slice is an 80-byte piece of every
// 640-byte pixel row, so each cluster drags partial DRAM bursts that its neighbours re-fetch.  A cluster of 8
// per whole sample with full rows: 39 vs 29 us at 16x1024x640 -- 128 CTAs of 164 KB each run one per SM and
// expose their load latency.)
// Pass 1 (stats): per (b, slab) partial sums per group.  bytes: read B*HW*C*2.
// Pass 2 (apply): y = [silu]((x - mean) * rstd * gamma + beta).  bytes: read + write B*HW*C*2.
// Each thread owns a fixed 8-channel vector, so its per-channel accumulators / affine params stay
// in registers; the channel->group reduction is done in a fixed order (bit-reproducible).
#ifndef LDX_GN_STATS_U16
#define LDX_GN_STATS_U16 1     // 16-deep first loop: 2-4 % on the 64x64-level maps (34.7 -> 33.9 us at 16 x 4096 x 320)
#endif
struct GnArgs {
  const __half* x0; const __half* x1;
  int c0, c1; int64_t ld0, ld1;
  int HW, C, cpg, nvec, P, pix_per_slab, nslab;
  float eps; int silu;
  const float* gamma; const float* beta;
  __half* y; int64_t ldy;
  float* partial;   // [B][nslab][32][2]
  // statistics-only form (ldx_group_norm_coef): the last CTA of sample b (atomic ticket) reduces the slab partials in
  // gn_apply_kernel's order and writes coef[b][c] = (rstd * gamma, beta - mean * rstd * gamma) for a consumer that
  // normalises on its own (the halo-tile conv's operand transform, conv_halo.cu)
  float* coef;      // [B][C][2] or null
  unsigned* ticket; // [B], zero before and after the launch
};

// Every thread of one CTA (>= 256 threads): slab partials -> (mean, rstd) per group -> coefficients of sample b.
// Same summation order and arithmetic as the prologue of gn_apply_kernel (bit-identical normalisation).
__device__ __forceinline__ void gn_write_coef(const float* partial, int b, int nslab, int HW, int cpg, int C, float eps,
                                              const float* gamma, const float* beta, float* coef) {
  __shared__ float c_mean[32], c_rstd[32];
  const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
  double s = 0.0, ss = 0.0;
  if (g < 32) {
    // (up to 8 slabs per thread: the loads go out together, the sums keep their order)
    for (int k0 = sub; k0 < nslab; k0 += 64) {
      float2 pr[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + 8 * u;
        pr[u] = k < nslab ? __ldcg(reinterpret_cast<const float2*>(partial + (((size_t)b * nslab + k) * 32 + g) * 2)) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        s += (double)pr[u].x;
        ss += (double)pr[u].y;
      }
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if (g < 32 && sub == 0) {
    const double n = (double)HW * cpg;
    const double mean = s / n;
    double var = ss / n - mean * mean;
    if (var < 0.0) var = 0.0;
    c_mean[g] = (float)mean;
    c_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int gg = c / cpg;
    const float w = __ldg(gamma + c) * c_rstd[gg];
    *reinterpret_cast<float2*>(coef + ((size_t)b * C + c) * 2) = make_float2(w, fmaf(-c_mean[gg], w, __ldg(beta + c)));
  }
}

// Ticket protocol around it: call with all threads after the CTA's partial has been stored.
__device__ __forceinline__ void gn_finish_coef(const float* partial, unsigned* ticket, int b, int nslab, int HW, int cpg, int C,
                                               float eps, const float* gamma, const float* beta, float* coef) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(ticket + b, 1u) == (unsigned)(nslab - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  gn_write_coef(partial, b, nslab, HW, cpg, C, eps, gamma, beta, coef);
  if (threadIdx.x == 0) ticket[b] = 0u;
}

__device__ __forceinline__ uint4 gn_load(const GnArgs& a, int b, int pix, int vec) {
  const int c = vec * 8;
  if (c < a.c0) return __ldg(reinterpret_cast<const uint4*>(a.x0 + ((int64_t)b * a.HW + pix) * a.ld0 + c));
  return __ldg(reinterpret_cast<const uint4*>(a.x1 + ((int64_t)b * a.HW + pix) * a.ld1 + (c - a.c0)));
}

__global__ void gn_stats_kernel(const GnArgs a) {
  extern __shared__ float sm[];          // [P][C][2]
  pdl_prologue();
  const int b = blockIdx.x, slab = blockIdx.y;
  const int vec = threadIdx.x % a.nvec, pl = threadIdx.x / a.nvec;
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
  const int p_begin = slab * a.pix_per_slab;
  const int p_end = min(a.HW, p_begin + a.pix_per_slab);
  if (pl < a.P) {
    int pix = p_begin + pl;
    // 16 independent 16-byte loads in flight per round, the ragged last round predicated (zeros add nothing): a slab is
    // usually 8 - 16 pixels per thread, so the whole slab is ONE memory round trip.  (The earlier form ran the remainder as
    // 8-deep and then single loads: a thread with 11 pixels paid four dependent DRAM latencies.)  Same per-thread summation
    // order as before: bit-identical statistics.
    for (; pix < p_end; pix += 16 * a.P) {
      uint4 v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u)
        v[u] = pix + u * a.P < p_end ? gn_load(a, b, pix + u * a.P, vec) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float2 f0 = unpack_h2(v[u].x), f1 = unpack_h2(v[u].y), f2 = unpack_h2(v[u].z), f3 = unpack_h2(v[u].w);
        const float f[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += f[j]; ss[j] += f[j] * f[j]; }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sm[((size_t)pl * a.C + vec * 8 + j) * 2 + 0] = s[j];
      sm[((size_t)pl * a.C + vec * 8 + j) * 2 + 1] = ss[j];
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
    float acc = 0.f;
    for (int pl2 = 0; pl2 < a.P; ++pl2)
      for (int c = g * a.cpg; c < (g + 1) * a.cpg; ++c) acc += sm[((size_t)pl2 * a.C + c) * 2 + which];
    a.partial[(((size_t)b * a.nslab + slab) * 32 + g) * 2 + which] = acc;
  }
  if (a.coef) gn_finish_coef(a.partial, a.ticket, b, a.nslab, a.HW, a.cpg, a.C, a.eps, a.gamma, a.beta, a.coef);
}

__global__ void __launch_bounds__(1024, 1) gn_apply_kernel(const GnArgs a, int nslab_apply, int pix_per_slab_apply) {
  __shared__ float s_mean[32], s_rstd[32];
  const int b = blockIdx.x, slab = blockIdx.y;
  const int vec = threadIdx.x % a.nvec, pl = threadIdx.x / a.nvec;
  const bool active = pl < a.P;
  // affine parameters do not depend on the previous kernel: fetch them before the grid dependency resolves
  float gm[8], bt[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gm[j] = active ? __ldg(a.gamma + vec * 8 + j) : 0.f;
    bt[j] = active ? __ldg(a.beta + vec * 8 + j) : 0.f;
  }
  pdl_prologue();
  const int p_begin = slab * pix_per_slab_apply;
  const int p_end = min(a.HW, p_begin + pix_per_slab_apply);
  // first batch of activations in flight while the statistics are reduced
  constexpr int U = 4;
  uint4 v[U];
  int pix = p_begin + pl;
  const bool first_full = active && pix + (U - 1) * a.P < p_end;
  if (first_full) {
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = gn_load(a, b, pix + u * a.P, vec);
  }
  {
    // 8 threads per group sum the slab partials in a fixed order (bit-reproducible), in double
    const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
    double s = 0.0, ss = 0.0;
    if (g < 32) {
      for (int k = sub; k < a.nslab; k += 8) {
        const float2 pr = __ldg(reinterpret_cast<const float2*>(a.partial + (((size_t)b * a.nslab + k) * 32 + g) * 2));
        s += (double)pr.x;
        ss += (double)pr.y;
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (g < 32 && sub == 0) {
      const double n = (double)a.HW * a.cpg;
      const double mean = s / n;
      double var = ss / n - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean[g] = (float)mean;
      s_rstd[g] = (float)(1.0 / sqrt(var + (double)a.eps));
    }
  }
  __syncthreads();
  if (!active) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = vec * 8 + j;
    const int g = c / a.cpg;
    const float w = gm[j] * s_rstd[g];
    sc[j] = w;
    sh[j] = fmaf(-s_mean[g], w, bt[j]);
  }
  auto apply_store = [&](const uint4& vv, int px) {
    const float2 f0 = unpack_h2(vv.x), f1 = unpack_h2(vv.y), f2 = unpack_h2(vv.z), f3 = unpack_h2(vv.w);
    float f[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = fmaf(f[j], sc[j], sh[j]);
      if (a.silu) f[j] = silu_f(f[j]);
    }
    uint4 o;
    o.x = pack_h2(f[0], f[1]); o.y = pack_h2(f[2], f[3]); o.z = pack_h2(f[4], f[5]); o.w = pack_h2(f[6], f[7]);
    *reinterpret_cast<uint4*>(a.y + ((int64_t)b * a.HW + px) * a.ldy + vec * 8) = o;
  };
  if (first_full) {
    // software pipeline: the next batch is requested before the current one is transformed and stored
    for (;;) {
      const int nxt = pix + U * a.P;
      const bool more = nxt + (U - 1) * a.P < p_end;
      uint4 w[U];
      if (more) {
#pragma unroll
        for (int u = 0; u < U; ++u) w[u] = gn_load(a, b, nxt + u * a.P, vec);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) apply_store(v[u], pix + u * a.P);
      pix = nxt;
      if (!more) break;
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = w[u];
    }
  }
  for (; pix < p_end; pix += a.P) apply_store(gn_load(a, b, pix, vec), pix);
}

// Small maps (8x8 / 16x16 levels): one CTA per (batch, group) keeps its HW x cpg slice in registers and
// does stats + apply in a single launch.  Requires cpg % 8 == 0 and HW*cpg/8 <= 256*GN_SMALL_VPT vectors.
static constexpr int GN_SMALL_VPT = 8;
__global__ void __launch_bounds__(256) gn_small_kernel(const GnArgs a) {
  __shared__ float red[2][8];
  __shared__ float s_stat[2];
  pdl_prologue();
  const int b = blockIdx.x >> 5, g = blockIdx.x & 31;
  const int vpg = a.cpg >> 3;                 // 16-byte vectors per group per pixel
  const int nv = a.HW * vpg;
  uint4 v[GN_SMALL_VPT];
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int u = 0; u < GN_SMALL_VPT; ++u) {
    const int idx = threadIdx.x + u * 256;
    if (idx < nv) {
      v[u] = gn_load(a, b, idx / vpg, g * vpg + idx % vpg);
      const float2 f0 = unpack_h2(v[u].x), f1 = unpack_h2(v[u].y), f2 = unpack_h2(v[u].z), f3 = unpack_h2(v[u].w);
      s += ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
      ss += ((f0.x * f0.x + f0.y * f0.y) + (f1.x * f1.x + f1.y * f1.y)) +
            ((f2.x * f2.x + f2.y * f2.y) + (f3.x * f3.x + f3.y * f3.y));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tss = 0.0;
    for (int w = 0; w < 8; ++w) { ts += (double)red[0][w]; tss += (double)red[1][w]; }
    const double n = (double)a.HW * a.cpg;
    const double mean = ts / n;
    double var = tss / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_stat[0] = (float)mean;
    s_stat[1] = (float)(1.0 / sqrt(var + (double)a.eps));
  }
  __syncthreads();
  const float mean = s_stat[0], rstd = s_stat[1];
  if (a.coef) {                                   // statistics-only form: this CTA's group of channels
    for (int c = g * a.cpg + threadIdx.x; c < (g + 1) * a.cpg; c += 256) {
      const float w = __ldg(a.gamma + c) * rstd;
      *reinterpret_cast<float2*>(a.coef + ((size_t)b * a.C + c) * 2) = make_float2(w, fmaf(-mean, w, __ldg(a.beta + c)));
    }
  }
  if (!a.y) return;
#pragma unroll
  for (int u = 0; u < GN_SMALL_VPT; ++u) {
    const int idx = threadIdx.x + u * 256;
    if (idx < nv) {
      const int pix = idx / vpg, vec = g * vpg + idx % vpg;
      const float2 f0 = unpack_h2(v[u].x), f1 = unpack_h2(v[u].y), f2 = unpack_h2(v[u].z), f3 = unpack_h2(v[u].w);
      float f[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma + vec * 8)), g1 = __ldg(reinterpret_cast<const float4*>(a.gamma + vec * 8) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.beta + vec * 8)), b1 = __ldg(reinterpret_cast<const float4*>(a.beta + vec * 8) + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float w = gg[j] * rstd;
        f[j] = fmaf(f[j], w, fmaf(-mean, w, bb[j]));
        if (a.silu) f[j] = silu_f(f[j]);
      }
      uint4 o;
      o.x = pack_h2(f[0], f[1]); o.y = pack_h2(f[2], f[3]); o.z = pack_h2(f[4], f[5]); o.w = pack_h2(f[6], f[7]);
      *reinterpret_cast<uint4*>(a.y + ((int64_t)b * a.HW + pix) * a.ldy + vec * 8) = o;
    }
  }
}

// GroupNorm statistics from the column partials the producing GEMMs left behind (ldx_gemm_desc_t.colstats_out,
// [B][slots][C][2] per source): the same reduction as gn_stats_kernel, over 1/32 of the rows and without touching
// the tensor.  CTA (b, slab) owns a contiguous range of slots; a thread owns one channel of the (virtual) concat
// [x0 | x1] and one of P slot lanes, so a warp reads 256 contiguous bytes; per-group sums are formed in a fixed
// order and land where gn_apply_kernel expects slab partials.  bytes = B * slots * C * 8.
struct GnCsArgs {
  const float* cs0; const float* cs1;
  int c0, c1, slots0, slots1, C, cpg, P, nslab;
  float* partial;   // [B][nslab][32][2]
  // statistics-only form, as in GnArgs
  float* coef; unsigned* ticket;
  int HW; float eps; const float* gamma; const float* beta;
};

__global__ void __launch_bounds__(1024) gn_colstats_reduce_kernel(const GnCsArgs a) {
  extern __shared__ float sm[];          // [P][C][2]
  pdl_prologue();
  const int b = blockIdx.x, slab = blockIdx.y;
  const int c = threadIdx.x % a.C, pl = threadIdx.x / a.C;
  if (pl < a.P) {
    const bool second = c >= a.c0;
    const int cw = second ? a.c1 : a.c0, slots = second ? a.slots1 : a.slots0, cc = second ? c - a.c0 : c;
    const float2* base = reinterpret_cast<const float2*>(second ? a.cs1 : a.cs0) + (int64_t)b * slots * cw + cc;
    const int per = (slots + a.nslab - 1) / a.nslab;
    const int s_begin = slab * per, s_end = min(slots, s_begin + per);
    float s = 0.f, ss = 0.f;
    int sl = s_begin + pl;
    for (; sl + 3 * a.P < s_end; sl += 4 * a.P) {          // four independent loads in flight
      const float2 v0 = __ldg(base + (int64_t)sl * cw), v1 = __ldg(base + (int64_t)(sl + a.P) * cw);
      const float2 v2 = __ldg(base + (int64_t)(sl + 2 * a.P) * cw), v3 = __ldg(base + (int64_t)(sl + 3 * a.P) * cw);
      s += (v0.x + v1.x) + (v2.x + v3.x);
      ss += (v0.y + v1.y) + (v2.y + v3.y);
    }
    for (; sl < s_end; sl += a.P) {
      const float2 v = __ldg(base + (int64_t)sl * cw);
      s += v.x;
      ss += v.y;
    }
    sm[((size_t)pl * a.C + c) * 2 + 0] = s;
    sm[((size_t)pl * a.C + c) * 2 + 1] = ss;
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
    double acc = 0.0;
    for (int pl2 = 0; pl2 < a.P; ++pl2)
      for (int ch = g * a.cpg; ch < (g + 1) * a.cpg; ++ch) acc += (double)sm[((size_t)pl2 * a.C + ch) * 2 + which];
    a.partial[(((size_t)b * a.nslab + slab) * 32 + g) * 2 + which] = (float)acc;
  }
  if (a.coef) gn_finish_coef(a.partial, a.ticket, b, a.nslab, a.HW, a.cpg, a.C, a.eps, a.gamma, a.beta, a.coef);
}

int group_norm_num_launches(int C, int HW) {
  const int cpg = C / 32;
  return (cpg % 8 == 0 && (int64_t)HW * (cpg / 8) <= 256 * GN_SMALL_VPT) ? 1 : 2;
}

static int group_norm_impl(const void* x0, int c0, int64_t ld0, const void* x1, int c1, int64_t ld1, int B, int HW,
                           float eps, const float* gamma, const float* beta, int silu, void* y, int64_t ldy,
                           float* stats, cudaStream_t stream, const float* cs0 = nullptr, int slots0 = 0,
                           const float* cs1 = nullptr, int slots1 = 0, float* coef = nullptr, unsigned* ticket = nullptr) {
  const int C = c0 + (x1 ? c1 : 0);
  LDX_CHECK_ARG(x0 && (y || coef) && stats && gamma && beta, "group_norm: null pointer");
  LDX_CHECK_ARG(!coef || (ticket && !y && reinterpret_cast<uintptr_t>(coef) % 8 == 0),
                "group_norm_coef: needs tickets, an 8-byte aligned coefficient buffer and no y");
  LDX_CHECK_ARG(C % 32 == 0 && c0 % 8 == 0 && C % 8 == 0, "group_norm: C=%d (c0=%d) must be a multiple of 32 / 8", C, c0);
  LDX_CHECK_ARG(ld0 % 8 == 0 && ldy % 8 == 0 && (!x1 || ld1 % 8 == 0), "group_norm: pitches must be multiples of 8");
  LDX_CHECK_ARG(C <= 8192, "group_norm: C too large");
  GnArgs a;
  a.x0 = (const __half*)x0; a.x1 = (const __half*)x1;
  a.c0 = c0; a.c1 = c1; a.ld0 = ld0; a.ld1 = ld1;
  a.HW = HW; a.C = C; a.cpg = C / 32; a.nvec = C / 8;
  a.coef = coef; a.ticket = ticket;
  if (a.cpg % 8 == 0 && (int64_t)HW * (a.cpg / 8) <= 256 * GN_SMALL_VPT &&
      reinterpret_cast<uintptr_t>(gamma) % 16 == 0 && reinterpret_cast<uintptr_t>(beta) % 16 == 0) {
    a.P = 1; a.pix_per_slab = HW; a.nslab = 1;
    a.eps = eps; a.silu = silu; a.gamma = gamma; a.beta = beta;
    a.y = (__half*)y; a.ldy = ldy; a.partial = stats;
    LDX_CUDA(launch_k(gn_small_kernel, dim3(B * 32), dim3(256), 0, stream, 1, a));
    return LDX_OK;
  }
  a.P = a.nvec >= 256 ? 1 : 256 / a.nvec;
  if (a.P > HW) a.P = HW;
  int threads = ((a.nvec * a.P + 31) / 32) * 32;
  if (threads < 256) threads = 256;   // the reductions use the first 64 (stats) / 256 (apply) threads
  LDX_CHECK_ARG(threads <= 1024, "group_norm: C too large for one block");
  // stats workspace holds [B][nslab][32][2]; nslab bounded so that it fits LDX_GN_MAX_SLABS
  const int sms = num_sms();
  // one full wave of resident CTAs: more slabs than that only add a second, partly filled wave (measured:
  // 43 -> 35 us at 16 x 4096 x 320 against 4x / 8x oversubscription)
  const size_t smem_stats = (size_t)a.P * C * 2 * sizeof(float);
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_stats_kernel, threads, smem_stats) != cudaSuccess || occ < 1) occ = 2;
  int nslab = occ * sms / B;
  const int max_slabs = (HW + a.P - 1) / a.P;
  if (nslab > max_slabs) nslab = max_slabs;
  if (nslab > 64) nslab = 64;
  if (nslab < 1) nslab = 1;
  int pps = (HW + nslab - 1) / nslab;
  pps = (pps + a.P - 1) / a.P * a.P;
  nslab = (HW + pps - 1) / pps;
  a.pix_per_slab = pps; a.nslab = nslab;
  a.eps = eps; a.silu = silu; a.gamma = gamma; a.beta = beta;
  a.y = (__half*)y; a.ldy = ldy; a.partial = stats;
  if (cs0 && (!x1 || cs1) && C <= 1024) {
    // statistics from the producers' column partials: a reduction over B * slots * C pairs instead of a pass over x
    LDX_CHECK_ARG(slots0 >= 1 && (!x1 || slots1 >= 1), "group_norm: colstats slots");
    GnCsArgs f;
    f.cs0 = cs0; f.cs1 = x1 ? cs1 : cs0;
    f.c0 = c0; f.c1 = x1 ? c1 : 0; f.slots0 = slots0; f.slots1 = x1 ? slots1 : slots0;
    f.C = C; f.cpg = a.cpg;
    const int max_slots = slots0 > f.slots1 ? slots0 : f.slots1;
    f.P = 1024 / C;
    if (f.P > max_slots) f.P = max_slots;
    if (f.P < 1) f.P = 1;
    int cs_threads = ((C * f.P + 31) / 32) * 32;
    if (cs_threads < 256) cs_threads = 256;                 // the reductions use the first 64 / 256 threads
    f.coef = coef; f.ticket = ticket; f.HW = HW; f.eps = eps; f.gamma = gamma; f.beta = beta;
    int ns = 2 * sms / B;                                   // ~two CTAs per SM in flight
    const int max_ns = (max_slots + f.P - 1) / f.P;
    if (ns > max_ns) ns = max_ns;
    if (ns > 64) ns = 64;
    if (ns < 1) ns = 1;
    f.nslab = ns;
    f.partial = stats;
    a.nslab = ns;
    LDX_CUDA(launch_k(gn_colstats_reduce_kernel, dim3(B, ns), dim3(cs_threads), (size_t)f.P * C * 2 * sizeof(float), stream,
                      1, f));
  } else {
    const size_t smem = (size_t)a.P * C * 2 * sizeof(float);
    static std::atomic<uint64_t> attr_done{0};
    int arc = set_max_dyn_smem_once(gn_stats_kernel, 128 * 1024, attr_done);
    if (arc) return arc;
    LDX_CHECK_ARG(smem <= 128 * 1024, "group_norm: smem");
    LDX_CUDA(launch_k(gn_stats_kernel, dim3(B, nslab), dim3(threads), smem, stream, 1, a));
  }
  if (coef) return LDX_OK;                                  // statistics-only form: the consumer applies them
  // apply: more slabs for bandwidth
  int occ2 = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, gn_apply_kernel, threads, 0) != cudaSuccess || occ2 < 1) occ2 = 4;
  int nslab2 = occ2 * sms / B;
  if (nslab2 > max_slabs) nslab2 = max_slabs;
  if (nslab2 < 1) nslab2 = 1;
  int pps2 = (HW + nslab2 - 1) / nslab2;
  pps2 = (pps2 + a.P - 1) / a.P * a.P;
  nslab2 = (HW + pps2 - 1) / pps2;
  LDX_CUDA(launch_k(gn_apply_kernel, dim3(B, nslab2), dim3(threads), 0, stream, 1, a, nslab2, pps2));
  return LDX_OK;
}

// ----------------------------------------------------------------------------- LayerNorm
// One warp per row at a time (grid-stride over rows), row held in registers (two-pass variance); the
// affine parameters of a lane's columns are loaded once and reused for every row. bytes: r+w rows*C*2.
template <int VPL>
__global__ void __launch_bounds__(256) layer_norm_kernel(const __half* __restrict__ x, int64_t ldx_, int rows, int C,
                                                         float eps, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, __half* __restrict__ y,
                                                         int64_t ldy) {
  const int lane = threadIdx.x & 31;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int nvec = C >> 3;
  pdl_prologue();
  constexpr bool kParamsInRegs = VPL <= 2;   // wider rows re-read the (L1-resident) parameters as float4
  constexpr int PV = kParamsInRegs ? VPL : 1;
  float g[PV][8], bt[PV][8];
  if constexpr (kParamsInRegs) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + i * 32;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        g[i][j] = v < nvec ? __ldg(gamma + v * 8 + j) : 0.f;
        bt[i][j] = v < nvec ? __ldg(beta + v * 8 + j) : 0.f;
      }
    }
  }
  const float inv_c = 1.0f / (float)C;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps_total) {
    const __half* xr = x + (int64_t)r * ldx_;
    uint4 u[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + i * 32;
      u[i] = v < nvec ? __ldg(reinterpret_cast<const uint4*>(xr + v * 8)) : make_uint4(0, 0, 0, 0);
    }
    float f[VPL][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float2 a = unpack_h2(u[i].x), b = unpack_h2(u[i].y), c = unpack_h2(u[i].z), d = unpack_h2(u[i].w);
      f[i][0] = a.x; f[i][1] = a.y; f[i][2] = b.x; f[i][3] = b.y; f[i][4] = c.x; f[i][5] = c.y; f[i][6] = d.x; f[i][7] = d.y;
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += f[i][j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + i * 32 < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = f[i][j] - mean; var += d * d; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var * inv_c + eps);
    __half* yr = y + (int64_t)r * ldy;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float o[8];
        if constexpr (kParamsInRegs) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = (f[i][j] - mean) * rstd * g[i][j] + bt[i][j];
        } else {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8) + 1);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8) + 1);
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = (f[i][j] - mean) * rstd * gg[j] + bb[j];
        }
        uint4 w;
        w.x = pack_h2(o[0], o[1]); w.y = pack_h2(o[2], o[3]); w.z = pack_h2(o[4], o[5]); w.w = pack_h2(o[6], o[7]);
        *reinterpret_cast<uint4*>(yr + v * 8) = w;
      }
    }
  }
}

static int layer_norm_impl(const void* x, int64_t ldx_, int rows, int C, float eps, const float* gamma,
                           const float* beta, void* y, int64_t ldy, cudaStream_t stream) {
  LDX_CHECK_ARG(x && y && gamma && beta, "layer_norm: null pointer");
  LDX_CHECK_ARG(C % 8 == 0 && C <= 2048 && ldx_ % 8 == 0 && ldy % 8 == 0, "layer_norm: C=%d unsupported", C);
  LDX_CHECK_ARG(reinterpret_cast<uintptr_t>(gamma) % 16 == 0 && reinterpret_cast<uintptr_t>(beta) % 16 == 0,
                "layer_norm: gamma/beta must be 16-byte aligned");
  const int threads = 256;
  int blocks = (rows * 32 + threads - 1) / threads;
  const int cap = num_sms() * 8;          // persistent: every warp walks several rows
  if (blocks > cap) blocks = cap;
  const int vpl = (C / 8 + 31) / 32;
  const __half* xx = (const __half*)x;
  __half* yy = (__half*)y;
  if (vpl <= 2) LDX_CUDA(launch_k(layer_norm_kernel<2>, dim3(blocks), dim3(threads), 0, stream, 1, xx, ldx_, rows, C, eps, gamma, beta, yy, ldy));
  else if (vpl <= 5) LDX_CUDA(launch_k(layer_norm_kernel<5>, dim3(blocks), dim3(threads), 0, stream, 1, xx, ldx_, rows, C, eps, gamma, beta, yy, ldy));
  else LDX_CUDA(launch_k(layer_norm_kernel<8>, dim3(blocks), dim3(threads), 0, stream, 1, xx, ldx_, rows, C, eps, gamma, beta, yy, ldy));
  return LDX_OK;
}

// ----------------------------------------------------------------------------- layout / small ops
// mode 0: plain (optionally scaled by 1/sqrt(sigma^2+1)); 1: TAESD decoder Clamp, tanh(x / 3) * 3 (LD.py:691-693);
// 2: TAESD encoder input, x * 0.5 + 0.5 (LD.py:759)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ sigma,
                                    __half* __restrict__ y, int B, int C, int HW, int Cpad, int mode) {
  const int64_t total = (int64_t)B * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int pix = (int)(i % HW);
    __half* o = y + i * Cpad;
    for (int c = 0; c < Cpad; ++c) {
      float v = 0.f;
      if (c < C) {
        const float xv = x[((int64_t)b * C + c) * HW + pix];
        v = sigma ? xv / sqrtf(sigma[b] * sigma[b] + 1.0f) : xv;
        if (mode == 1) v = tanhf(v / 3.0f) * 3.0f;
        else if (mode == 2) v = v * 0.5f + 0.5f;
      }
      o[c] = __float2half_rn(v);
    }
  }
}

__global__ void timestep_embedding_kernel(const float* __restrict__ sigma, const float* __restrict__ log_sigmas,
                                          int n_sigmas, __half* __restrict__ emb, int dim) {
  __shared__ float s_best[32];
  __shared__ int s_idx[32];
  __shared__ int s_t;
  const int b = blockIdx.x;
  const float ls = logf(sigma[b]);
  float best = 3.4e38f;
  int idx = 0x7fffffff;
  for (int i = threadIdx.x; i < n_sigmas; i += blockDim.x) {
    const float d = fabsf(ls - log_sigmas[i]);
    if (d < best || (d == best && i < idx)) { best = d; idx = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ob < best || (ob == best && oi < idx)) { best = ob; idx = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_best[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    for (int w = 1; w < nw; ++w)
      if (s_best[w] < best || (s_best[w] == best && s_idx[w] < idx)) { best = s_best[w]; idx = s_idx[w]; }
    s_t = idx;
  }
  __syncthreads();
  const float t = (float)s_t;
  const int half_dim = dim / 2;
  for (int i = threadIdx.x; i < half_dim; i += blockDim.x) {
    const float freq = expf(-9.210340371976184f * (float)i / (float)half_dim);   // ln(10000)
    const float arg = t * freq;
    emb[(int64_t)b * dim + i] = __float2half_rn(cosf(arg));
    emb[(int64_t)b * dim + half_dim + i] = __float2half_rn(sinf(arg));
  }
}

__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W, int cvec) {
  const int64_t total = (int64_t)B * (2 * H) * (2 * W) * cvec;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    int64_t r = i / cvec;
    const int X = (int)(r % (2 * W)); r /= (2 * W);
    const int Y = (int)(r % (2 * H));
    const int b = (int)(r / (2 * H));
    y[i] = __ldg(x + (((int64_t)b * H + (Y >> 1)) * W + (X >> 1)) * cvec + v);
  }
}

// eps NHWC fp16 -> denoised NCHW fp32 = x - eps * sigma
__global__ void eps_to_denoised_kernel(const __half* __restrict__ eps, int64_t lde, const float* __restrict__ x,
                                       const float* __restrict__ sigma, float* __restrict__ den, int B, int C, int HW) {
  const int64_t total = (int64_t)B * C * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW);
    const int c = (int)((i / HW) % C);
    const int b = (int)(i / ((int64_t)HW * C));
    const float e = __half2float(eps[((int64_t)b * HW + pix) * lde + c]);
    den[i] = x[i] - e * sigma[b];
  }
}

// eps holds [uncond(b) ; cond(b)] ; out = du + (dc - du) * scale with d* = x - eps* * sigma
__global__ void cfg_denoised_kernel(const __half* __restrict__ eps, int64_t lde, const float* __restrict__ x,
                                    const float* __restrict__ sigma, float scale, float* __restrict__ den, int nb,
                                    int C, int HW) {
  const int64_t total = (int64_t)nb * C * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW);
    const int c = (int)((i / HW) % C);
    const int b = (int)(i / ((int64_t)HW * C));
    const float eu = __half2float(eps[((int64_t)b * HW + pix) * lde + c]);
    const float ec = __half2float(eps[((int64_t)(b + nb) * HW + pix) * lde + c]);
    const float xv = x[i], s = sigma[b];
    const float du = xv - eu * s, dc = xv - ec * s;
    den[i] = du + (dc - du) * scale;
  }
}

// One launch for the tail of a DPM++ 2M step under classifier-free guidance (sampling_function LD.py:2605 +
// sample_dpmpp_2m_sde at eta = 0, LD.py:1209-1236): den = du + (dc - du) * scale with d* = x - eps* * sigma;
// x <- cx * x + cd * den [+ co * (den - old)] (x <- den on the last step); old <- den; and the next step's sigma is
// written to the UNet plan's input vector.  sigma arrives as a scalar (no thread reads sigma_io, so the write
// is race-free).  bytes per latent element: 2 x 2 (eps halves, fp16) + 4 (x) + 4 (old) read, 4 + 4 written.
__global__ void cfg_dpmpp2m_step_kernel(const __half* __restrict__ eps, int64_t lde, float* __restrict__ x,
                                        float* __restrict__ old, float* __restrict__ den_out, float sigma, float scale,
                                        float cx, float cd, float co, int have_old, int last, float sigma_next,
                                        float* __restrict__ sigma_io, int nb, int C, int HW) {
  const int64_t total = (int64_t)nb * C * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW);
    const int c = (int)((i / HW) % C);
    const int b = (int)(i / ((int64_t)HW * C));
    const float eu = __half2float(eps[((int64_t)b * HW + pix) * lde + c]);
    const float ec = __half2float(eps[((int64_t)(b + nb) * HW + pix) * lde + c]);
    const float xv = x[i];
    const float du = xv - eu * sigma, dc = xv - ec * sigma;
    const float d = du + (dc - du) * scale;
    float v = d;
    if (!last) {
      v = cx * xv + cd * d;
      if (have_old) v = v + co * (d - old[i]);
    }
    x[i] = v;
    old[i] = d;
    if (den_out) den_out[i] = d;
  }
  if (sigma_io && blockIdx.x == 0 && threadIdx.x < nb) sigma_io[threadIdx.x] = sigma_next;
}

__global__ void cfg_combine_kernel(const float* __restrict__ c, const float* __restrict__ u, float scale,
                                   float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = u[i] + (c[i] - u[i]) * scale;
}

__global__ void vae_post_kernel(const __half* __restrict__ dec, int64_t ldc, float* __restrict__ img, int64_t npix) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix * 3; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / 3;
    const int c = (int)(i % 3);
    const float v = (__half2float(dec[pix * ldc + c]) + 1.0f) / 2.0f;
    img[i] = fminf(fmaxf(v, 0.0f), 1.0f);
  }
}

// same, quantised the way the reference saves images (255 * img, clipped, uint8; LD.py:6616-6620 SaveImage)
__global__ void vae_post_u8_kernel(const __half* __restrict__ dec, int64_t ldc, uint8_t* __restrict__ img, int64_t npix) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix * 3; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / 3;
    const int c = (int)(i % 3);
    const float v = fminf(fmaxf((__half2float(dec[pix * ldc + c]) + 1.0f) / 2.0f, 0.0f), 1.0f);
    img[i] = (uint8_t)fminf(fmaxf(255.0f * v, 0.0f), 255.0f);          // numpy astype(uint8) truncates
  }
}

// VAE.encode input side (LD.py:6385, 6295, 6402-6405): pixels fp32 NHWC [npix, 3] in [0,1] -> 2x - 1 as fp16
// NHWC [npix, 8] (channels 3..7 zero: the 3x3 conv_in reads 16-byte channel groups)
__global__ void vae_pre_kernel(const float* __restrict__ pix, __half* __restrict__ out, int64_t npix) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    const float r = pix[i * 3], g = pix[i * 3 + 1], b = pix[i * 3 + 2];
    const __half2 a = __floats2half2_rn(r * 2.0f - 1.0f, g * 2.0f - 1.0f), c = __floats2half2_rn(b * 2.0f - 1.0f, 0.f);
    uint4 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&c);
    u.z = 0u; u.w = 0u;
    reinterpret_cast<uint4*>(out)[i] = u;
  }
}

// DiagonalGaussianDistribution (LD.py:166-179): moments fp32 NHWC [B*hw, 2*zc] (mean | logvar) and noise fp32 NCHW
// -> z = mean + exp(0.5 * clamp(logvar, -30, 20)) * noise, fp32 NCHW [B, zc, h, w]; noise == NULL gives the mean
__global__ void vae_sample_kernel(const float* __restrict__ mom, const float* __restrict__ noise, float* __restrict__ z,
                                  int64_t n, int zc, int64_t hw) {
  // i indexes z (NCHW)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i % hw, bc = i / hw;
    const int c = (int)(bc % zc);
    const int64_t b = bc / zc;
    const float* m = mom + (b * hw + pix) * (2 * zc);
    float v = m[c];
    if (noise) v += expf(0.5f * fminf(fmaxf(m[zc + c], -30.0f), 20.0f)) * noise[i];
    z[i] = v;
  }
}

// One axis of bislerp (LD.py:429-518): spherical interpolation of the C-channel vectors at positions i1[j], i2[j]
// of the `len_old` axis with ratio r[j], for the tensor viewed as [n, C, outer, len, inner] (fp32).  The width pass
// has inner = 1, the height pass outer = 1.  One thread per output vector; C <= BISLERP_MAX_C in registers.
static constexpr int BISLERP_MAX_C = 16;
__global__ void bislerp_axis_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ ratio,
                                    const int* __restrict__ i1, const int* __restrict__ i2, int64_t total, int C,
                                    int64_t outer, int len_old, int len_new, int64_t inner) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t in_ = t % inner;
    const int j = (int)((t / inner) % len_new);
    const int64_t o = (t / (inner * len_new)) % outer;
    const int64_t n = t / (inner * len_new * outer);
    const int64_t cs_in = outer * len_old * inner, cs_out = outer * len_new * inner;
    const float* a = in + n * C * cs_in + (o * len_old + i1[j]) * inner + in_;
    const float* b = in + n * C * cs_in + (o * len_old + i2[j]) * inner + in_;
    float* d = out + n * C * cs_out + (o * len_new + j) * inner + in_;
    const float r = ratio[j];
    float v1[BISLERP_MAX_C], v2[BISLERP_MAX_C];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < BISLERP_MAX_C; ++c)
      if (c < C) {
        v1[c] = a[c * cs_in];
        v2[c] = b[c * cs_in];
        s1 = fmaf(v1[c], v1[c], s1);
        s2 = fmaf(v2[c], v2[c], s2);
      }
    const float n1 = sqrtf(s1), n2 = sqrtf(s2);
    // zero vectors normalise to zero (LD.py:444-446)
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < BISLERP_MAX_C; ++c)
      if (c < C) dot += (n1 == 0.f ? 0.f : v1[c] / n1) * (n2 == 0.f ? 0.f : v2[c] / n2);
    if (dot > 1.0f - 1e-5f) {                       // same direction: the first vector
#pragma unroll
      for (int c = 0; c < BISLERP_MAX_C; ++c)
        if (c < C) d[c * cs_out] = v1[c];
    } else if (dot < 1e-5f - 1.0f) {                // polar opposites: plain lerp
#pragma unroll
      for (int c = 0; c < BISLERP_MAX_C; ++c)
        if (c < C) d[c * cs_out] = v1[c] * (1.0f - r) + v2[c] * r;
    } else {
      const float omega = acosf(dot), so = sinf(omega);
      const float w1 = sinf((1.0f - r) * omega) / so, w2 = sinf(r * omega) / so;
      const float nm = n1 * (1.0f - r) + n2 * r;
#pragma unroll
      for (int c = 0; c < BISLERP_MAX_C; ++c)
        if (c < C) {
          const float u1 = n1 == 0.f ? 0.f : v1[c] / n1, u2 = n2 == 0.f ? 0.f : v2[c] / n2;
          d[c * cs_out] = (w1 * u1 + w2 * u2) * nm;
        }
    }
  }
}

// CLIPEmbeddings.forward (LD.py:4409-4410): out[r, :] = token_embedding[tokens[r], :] + position_embedding[r % T, :],
// fp32 tables (the reference keeps them fp32, LD.py:4423), fp16 out; one block per row
__global__ void clip_embed_kernel(const int* __restrict__ tokens, const float* __restrict__ tok_emb,
                                  const float* __restrict__ pos_emb, __half* __restrict__ out, int64_t ldo, int T, int C) {
  const int r = blockIdx.x;
  const float* te = tok_emb + (int64_t)tokens[r] * C;
  const float* pe = pos_emb + (int64_t)(r % T) * C;
  __half* o = out + (int64_t)r * ldo;
  for (int c = threadIdx.x; c < C; c += blockDim.x) o[c] = __float2half_rn(te[c] + pe[c]);
}

// ---- DPM-Solver-12/23 adaptive (DPMSolver, LD.py:976-1116): every state update has the form
//   out = x - c1 * eps - c2 * (e2 - eps)
// with host scalars c1, c2 (u1: c2 = 0).  The products and differences are rounded separately, as the reference's
// chain of torch kernels does, so that the step-size controller sees the same error norms.
__global__ void dpm_eps_kernel(const float* __restrict__ x, const float* __restrict__ den, float inv_sigma,
                               float* __restrict__ eps, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    eps[i] = __fmul_rn(__fsub_rn(x[i], den[i]), inv_sigma);
}
__global__ void dpm_combine_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ e2,
                                   float c1, float c2, float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float e = eps[i];
    float v = __fsub_rn(x[i], __fmul_rn(c1, e));
    if (e2) v = __fsub_rn(v, __fmul_rn(c2, __fsub_rn(e2[i], e)));
    out[i] = v;
  }
}
// partial[b] = sum over the block's elements of ((lo - hi) / max(atol, rtol * max(|lo|, |prev|)))^2   (LD.py:1094-1095)
__global__ void __launch_bounds__(256) dpm_error_partial_kernel(const float* __restrict__ lo, const float* __restrict__ hi,
                                                                const float* __restrict__ prev, float atol, float rtol,
                                                                int64_t n, double* __restrict__ partial) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float l = lo[i];
    const float delta = fmaxf(atol, __fmul_rn(rtol, fmaxf(fabsf(l), fabsf(prev[i]))));
    const float q = __fdiv_rn(__fsub_rn(l, hi[i]), delta);
    acc += (double)q * (double)q;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}
// err = sqrt(sum of the partials) / sqrt(n): fixed summation order, so the result is run-to-run deterministic
__global__ void dpm_error_final_kernel(const double* __restrict__ partial, int nb, double inv_sqrt_n, float* __restrict__ err) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int b = 0; b < nb; ++b) t += partial[b];
    *err = (float)(sqrt(t) * inv_sqrt_n);
  }
}

// row softmax for n = NV * 1024 (the VAE AttnBlock's 4096 / 16384 tokens): the row lives in registers, so the
// fp32 scores are read once (16-byte loads) and the fp16 probabilities written once (8-byte stores)
template <int NV>
__global__ void __launch_bounds__(256) softmax_rows_reg_kernel(const float* __restrict__ s, int64_t lds,
                                                               __half* __restrict__ out, int64_t ldo, float scale_log2) {
  __shared__ float red[2][8];
  const float4* row = reinterpret_cast<const float4*>(s + (int64_t)blockIdx.x * lds);
  uint2* orow = reinterpret_cast<uint2*>(out + (int64_t)blockIdx.x * ldo);
  float4 v[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = __ldcs(row + j * 256 + threadIdx.x);
  float m = -3.4e38f;
#pragma unroll
  for (int j = 0; j < NV; ++j) m = fmaxf(fmaxf(m, fmaxf(v[j].x, v[j].y)), fmaxf(v[j].z, v[j].w));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0][0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[0][w]);
  const float off = m * scale_log2;     // scale > 0: max(scale * s) = scale * max(s)
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j].x = exp2f(fmaf(v[j].x, scale_log2, -off));
    v[j].y = exp2f(fmaf(v[j].y, scale_log2, -off));
    v[j].z = exp2f(fmaf(v[j].z, scale_log2, -off));
    v[j].w = exp2f(fmaf(v[j].w, scale_log2, -off));
    sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = red[1][0];
#pragma unroll
  for (int w = 1; w < 8; ++w) sum += red[1][w];
  const float inv = 1.0f / sum;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const __half2 a = __floats2half2_rn(v[j].x * inv, v[j].y * inv), b = __floats2half2_rn(v[j].z * inv, v[j].w * inv);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&b);
    orow[j * 256 + threadIdx.x] = u;
  }
}

// row softmax(scale * s): fp32 scores in, fp16 probabilities out; one block per row
__global__ void softmax_rows_kernel(const float* __restrict__ s, int64_t lds, __half* __restrict__ out, int64_t ldo,
                                    int n, float scale) {
  __shared__ float red[32];
  const float* row = s + (int64_t)blockIdx.x * lds;
  __half* orow = out + (int64_t)blockIdx.x * ldo;
  float m = -3.4e38f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, row[i] * scale);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -3.4e38f;
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  m = red[0];
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sum += expf(row[i] * scale - m);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  const float inv = 1.0f / red[0];
  for (int i = threadIdx.x; i < n; i += blockDim.x) orow[i] = __float2half_rn(expf(row[i] * scale - m) * inv);
}

// ----------------------------------------------------------------------------- sampler updates
// bytes per element: dpmpp2m 16 (read x, d, old; write x), euler-a 16 (read x, d, noise; write x)
__global__ void dpmpp2m_step_kernel(float* __restrict__ x, const float* __restrict__ d, const float* __restrict__ od,
                                    float cx, float cd, float co, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // reference order (LD.py:1214-1226): x = cx*x + cd*d ; x = x + co*(d - old)
    float v = cx * x[i] + cd * d[i];
    if (od) v = v + co * (d[i] - od[i]);
    x[i] = v;
  }
}

__global__ void euler_ancestral_step_kernel(float* __restrict__ x, const float* __restrict__ d,
                                            const float* __restrict__ noise, float sigma, float dt, float ns,
                                            int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // reference order (LD.py:931-936): d = (x - denoised)/sigma ; x = x + d*dt ; x = x + noise*s_noise*sigma_up
    const float xv = x[i];
    float v = xv + ((xv - d[i]) / sigma) * dt;
    if (noise) v = v + noise[i] * ns;
    x[i] = v;
  }
}

// GroupNorm folded into the 1x1 conv that consumes it (ldx_gn_fold_weights, include/ldx.h): CTA (32 output rows, sample b),
// one warp per output row at a time.  w_out[b][n][c] = fp16(w[n][c] * scale[b][c]); bias_out[b][n] = bias[n] + sum_c (w[n][c] * beta[c] -
// mean[b][g(c)] * w_out[b][n][c]) with the lanes' partial sums combined in a fixed butterfly order (bit-reproducible).  The
// group mean comes back out of the coefficients, (beta - shift) / scale, from the channel of the group with the largest
// |scale| (a group whose scales are all zero contributes nothing whatever the mean).
static constexpr int GN_FOLD_ROWS = 32;
__global__ void __launch_bounds__(256) gn_fold_weights_kernel(const __half* __restrict__ w, int64_t ldw,
                                                              const float* __restrict__ bias, const float* __restrict__ beta,
                                                              const float* __restrict__ coef, int N, int C,
                                                              __half* __restrict__ w_out, float* __restrict__ bias_out) {
  extern __shared__ float fsm[];          // scale[C], beta[C], shift[C] -> mean of the channel's group [C]
  float* s_scale = fsm;
  float* s_beta = fsm + C;
  float* s_mean = fsm + 2 * C;
  __shared__ float g_mean[32];
  pdl_prologue();
  const int b = blockIdx.y, cpg = C / 32;
  const float2* cf = reinterpret_cast<const float2*>(coef) + (size_t)b * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float2 q = __ldg(&cf[c]);
    s_scale[c] = q.x;
    s_mean[c] = q.y;                      // shift, until the group means are known
    s_beta[c] = __ldg(beta + c);
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int g = threadIdx.x;
    float best = 0.f, mean = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const float sc = s_scale[c];
      if (fabsf(sc) > best) { best = fabsf(sc); mean = (s_beta[c] - s_mean[c]) / sc; }
    }
    g_mean[g] = mean;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_mean[c] = g_mean[c / cpg];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // GN_FOLD_ROWS output rows per CTA, warp w takes rows w, w + 8, ...: the per-sample set-up above is shared by all of them
  for (int n = blockIdx.x * GN_FOLD_ROWS + warp; n < min(N, (int)(blockIdx.x + 1) * GN_FOLD_ROWS); n += 8) {
    const uint4* wr = reinterpret_cast<const uint4*>(w + (size_t)n * ldw);
    uint4* wo = reinterpret_cast<uint4*>(w_out + ((size_t)b * N + n) * ldw);
    float acc = 0.f;
    for (int v = lane; v < C / 8; v += 32) {
      const uint4 q = __ldg(wr + v);
      const float2 f0 = unpack_h2(q.x), f1 = unpack_h2(q.y), f2 = unpack_h2(q.z), f3 = unpack_h2(q.w);
      const float f[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = v * 8 + 2 * j;
        const uint32_t pk = pack_h2(f[2 * j] * s_scale[c], f[2 * j + 1] * s_scale[c + 1]);
        const float2 r = unpack_h2(pk);                     // the ROUNDED folded weights, as the GEMM will read them
        acc = fmaf(f[2 * j], s_beta[c], acc);
        acc = fmaf(-s_mean[c], r.x, acc);
        acc = fmaf(f[2 * j + 1], s_beta[c + 1], acc);
        acc = fmaf(-s_mean[c + 1], r.y, acc);
        o[j] = pk;
      }
      wo[v] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    // K padding (ldw > C columns, zero in w) stays zero in w_out
    for (int v = C / 8 + lane; v < (int)(ldw / 8); v += 32) wo[v] = __ldg(wr + v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) bias_out[(size_t)b * N + n] = (bias ? __ldg(bias + n) : 0.f) + acc;
  }
}

static int gn_fold_weights_impl(const void* w, int64_t ldw, const float* bias, const float* beta, const float* coef, int B,
                                int N, int C, void* w_out, float* bias_out, cudaStream_t st) {
  LDX_CHECK_ARG(w && beta && coef && w_out && bias_out, "gn_fold_weights: null pointer");
  LDX_CHECK_ARG(B >= 1 && N >= 1 && C >= 32 && C % 32 == 0 && C % 8 == 0 && C <= 2048 && ldw >= C && ldw % 8 == 0,
                "gn_fold_weights: bad extents (B=%d N=%d C=%d ldw=%lld)", B, N, C, (long long)ldw);
  LDX_CHECK_ARG(reinterpret_cast<uintptr_t>(w) % 16 == 0 && reinterpret_cast<uintptr_t>(w_out) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(coef) % 8 == 0,
                "gn_fold_weights: alignment");
  LDX_CUDA(launch_k(gn_fold_weights_kernel, dim3((N + GN_FOLD_ROWS - 1) / GN_FOLD_ROWS, B), dim3(256), (size_t)3 * C * sizeof(float), st, 1,
                    (const __half*)w, ldw, bias, beta, coef, N, C, (__half*)w_out, bias_out));
  return LDX_OK;
}

static inline int ew_blocks(int64_t n, int threads = 256) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int run_generic_op(const ldx_generic_op_t* op, cudaStream_t st) {
  const int32_t* i = op->i;
  const int64_t* l = op->l;
  const float* f = op->f;
  const void* const* p = op->p;
  switch (op->kind) {
    case LDX_OP_GROUP_NORM:
      // p: x0, x1, gamma, beta, y, stats, colstats0, colstats1 ; i: c0, c1, B, HW, silu, slots0, slots1 ;
      // l: ld0, ld1, ldy ; f: eps
      return group_norm_impl(p[0], i[0], l[0], p[1], i[1], l[1], i[2], i[3], f[0], (const float*)p[2],
                             (const float*)p[3], i[4], (void*)p[4], l[2], (float*)p[5], st, (const float*)p[6], i[5],
                             (const float*)p[7], i[6]);
    case LDX_OP_GROUP_NORM_COEF:
      // p: x0, x1, gamma, beta, coef, stats, colstats0, colstats1, tickets ; i: c0, c1, B, HW, -, slots0, slots1 ;
      // l: ld0, ld1 ; f: eps
      return group_norm_impl(p[0], i[0], l[0], p[1], i[1], l[1], i[2], i[3], f[0], (const float*)p[2],
                             (const float*)p[3], 0, nullptr, 0, (float*)p[5], st, (const float*)p[6], i[5],
                             (const float*)p[7], i[6], (float*)p[4], (unsigned*)p[8]);
    case LDX_OP_GN_FOLD_WEIGHTS:
      // p: w, bias, beta, coef, w_out, bias_out ; l: ldw ; i: B, N, C
      return gn_fold_weights_impl(p[0], l[0], (const float*)p[1], (const float*)p[2], (const float*)p[3], i[0], i[1], i[2],
                                  (void*)p[4], (float*)p[5], st);
    case LDX_OP_LAYER_NORM:
      // p: x, gamma, beta, y ; i: rows, C ; l: ldx, ldy ; f: eps
      return layer_norm_impl(p[0], l[0], i[0], i[1], f[0], (const float*)p[1], (const float*)p[2], (void*)p[3], l[1], st);
    case LDX_OP_NCHW_TO_NHWC: {
      // p: x, sigma, y ; i: B, C, H, W, Cpad, mode
      const int64_t n = (int64_t)i[0] * i[2] * i[3];
      nchw_to_nhwc_kernel<<<ew_blocks(n), 256, 0, st>>>((const float*)p[0], (const float*)p[1], (__half*)p[2], i[0],
                                                        i[1], i[2] * i[3], i[4], i[5]);
      LDX_CUDA(cudaGetLastError());
      return LDX_OK;
    }
    case LDX_OP_TIMESTEP_EMB:
      // p: sigma, log_sigmas, emb ; i: n_sigmas, B, dim
      timestep_embedding_kernel<<<i[1], 256, 0, st>>>((const float*)p[0], (const float*)p[1], i[0], (__half*)p[2], i[2]);
      LDX_CUDA(cudaGetLastError());
      return LDX_OK;
    case LDX_OP_UPSAMPLE2X: {
      // p: x, y ; i: B, H, W, C
      LDX_CHECK_ARG(i[3] % 8 == 0, "upsample: C must be a multiple of 8");
      const int64_t n = (int64_t)i[0] * 4 * i[1] * i[2] * (i[3] / 8);
      upsample2x_kernel<<<ew_blocks(n), 256, 0, st>>>((const uint4*)p[0], (uint4*)p[1], i[0], i[1], i[2], i[3] / 8);
      LDX_CUDA(cudaGetLastError());
      return LDX_OK;
    }
    case LDX_OP_EPS_TO_DENOISED: {
      // p: eps, x, sigma, den ; i: B, C, H, W ; l: lde
      const int64_t n = (int64_t)i[0] * i[1] * i[2] * i[3];
      eps_to_denoised_kernel<<<ew_blocks(n), 256, 0, st>>>((const __half*)p[0], l[0], (const float*)p[1],
                                                           (const float*)p[2], (float*)p[3], i[0], i[1], i[2] * i[3]);
      LDX_CUDA(cudaGetLastError());
      return LDX_OK;
    }
    case LDX_OP_CFG_DENOISED: {
      // p: eps, x, sigma, den ; i: b, C, H, W ; l: lde ; f: scale
      const int64_t n = (int64_t)i[0] * i[1] * i[2] * i[3];
      cfg_denoised_kernel<<<ew_blocks(n), 256, 0, st>>>((const __half*)p[0], l[0], (const float*)p[1],
                                                        (const float*)p[2], f[0], (float*)p[3], i[0], i[1], i[2] * i[3]);
      LDX_CUDA(cudaGetLastError());
      return LDX_OK;
    }
    case LDX_OP_VAE_POST:
      // p: dec, img ; l: ldc, npix
      vae_post_kernel<<<ew_blocks(l[1] * 3), 256, 0, st>>>((const __half*)p[0], l[0], (float*)p[1], l[1]);
      LDX_CUDA(cudaGetLastError());
      return LDX_OK;
    case LDX_OP_VAE_PRE:
      // p: pixels, out ; l: npix
      vae_pre_kernel<<<ew_blocks(l[0]), 256, 0, st>>>((const float*)p[0], (__half*)p[1], l[0]);
      LDX_CUDA(cudaGetLastError());
      return LDX_OK;
    case LDX_OP_VAE_SAMPLE:
      // p: moments, noise, z ; l: n, hw ; i: zc
      vae_sample_kernel<<<ew_blocks(l[0]), 256, 0, st>>>((const float*)p[0], (const float*)p[1], (float*)p[2], l[0], i[0],
                                                          l[1]);
      LDX_CUDA(cudaGetLastError());
      return LDX_OK;
    case LDX_OP_CLIP_EMBED:
      // p: tokens, tok_emb, pos_emb, out ; l: ldo ; i: rows, T, C
      clip_embed_kernel<<<(unsigned)i[0], 256, 0, st>>>((const int*)p[0], (const float*)p[1], (const float*)p[2],
                                                       (__half*)p[3], l[0], i[1], i[2]);
      LDX_CUDA(cudaGetLastError());
      return LDX_OK;
    case LDX_OP_SOFTMAX:
      // p: s (fp32), out (fp16) ; l: lds, ldo, rows ; i: n ; f: scale
      {
        const float* sp = (const float*)p[0];
        __half* op = (__half*)p[1];
        const bool vec = f[0] > 0.f && l[0] % 4 == 0 && l[1] % 4 == 0 && ((uintptr_t)sp & 15) == 0 && ((uintptr_t)op & 7) == 0;
        const float sl2 = f[0] * 1.4426950408889634f;
        const unsigned rows = (unsigned)l[2];
        if (vec && i[0] == 16384) softmax_rows_reg_kernel<16><<<rows, 256, 0, st>>>(sp, l[0], op, l[1], sl2);
        else if (vec && i[0] == 8192) softmax_rows_reg_kernel<8><<<rows, 256, 0, st>>>(sp, l[0], op, l[1], sl2);
        else if (vec && i[0] == 4096) softmax_rows_reg_kernel<4><<<rows, 256, 0, st>>>(sp, l[0], op, l[1], sl2);
        else if (vec && i[0] == 2048) softmax_rows_reg_kernel<2><<<rows, 256, 0, st>>>(sp, l[0], op, l[1], sl2);
        else if (vec && i[0] == 1024) softmax_rows_reg_kernel<1><<<rows, 256, 0, st>>>(sp, l[0], op, l[1], sl2);
        else softmax_rows_kernel<<<rows, 256, 0, st>>>(sp, l[0], op, l[1], i[0], f[0]);
      }
      LDX_CUDA(cudaGetLastError());
      return LDX_OK;
    case LDX_OP_DUP:
      // p: src, dst ; l: bytes
      LDX_CUDA(cudaMemcpyAsync((void*)p[1], p[0], (size_t)l[0], cudaMemcpyDeviceToDevice, st));
      LDX_CUDA(cudaMemcpyAsync((char*)p[1] + l[0], p[0], (size_t)l[0], cudaMemcpyDeviceToDevice, st));
      return LDX_OK;
    default:
      set_error("unknown op kind %d", op->kind);
      return LDX_E_INVALID;
  }
}

}  // namespace ldx

using namespace ldx;

extern "C" int ldx_group_norm(const void* x0, int c0, int64_t ld0, const void* x1, int c1, int64_t ld1, int B, int HW,
                              float eps, const float* gamma, const float* beta, int silu, void* y, int64_t ldy,
                              float* stats, ldx_stream_t stream) {
  return group_norm_impl(x0, c0, ld0, x1, c1, ld1, B, HW, eps, gamma, beta, silu, y, ldy, stats, as_stream(stream));
}

extern "C" int ldx_group_norm_cs(const void* x0, int c0, int64_t ld0, const void* x1, int c1, int64_t ld1, int B, int HW,
                                 float eps, const float* gamma, const float* beta, int silu, void* y, int64_t ldy,
                                 float* stats, const float* colstats0, int slots0, const float* colstats1, int slots1,
                                 ldx_stream_t stream) {
  LDX_CHECK_ARG(colstats0 && slots0 >= 1 && (!x1 || (colstats1 && slots1 >= 1)), "group_norm_cs: missing column statistics");
  return group_norm_impl(x0, c0, ld0, x1, c1, ld1, B, HW, eps, gamma, beta, silu, y, ldy, stats, as_stream(stream),
                         colstats0, slots0, colstats1, slots1);
}

extern "C" int ldx_group_norm_coef(const void* x0, int c0, int64_t ld0, const void* x1, int c1, int64_t ld1, int B, int HW,
                                   float eps, const float* gamma, const float* beta, float* stats, uint32_t* tickets,
                                   float* coef, const float* colstats0, int slots0, const float* colstats1, int slots1,
                                   ldx_stream_t stream) {
  LDX_CHECK_ARG(coef && tickets, "group_norm_coef: null coefficient / ticket buffer");
  return group_norm_impl(x0, c0, ld0, x1, c1, ld1, B, HW, eps, gamma, beta, 0, nullptr, 0, stats, as_stream(stream),
                         colstats0, slots0, colstats1, slots1, coef, tickets);
}

extern "C" int ldx_gn_fold_weights(const void* w, int64_t ldw, const float* bias, const float* beta, const float* coef, int B,
                                   int N, int C, void* w_out, float* bias_out, ldx_stream_t stream) {
  return gn_fold_weights_impl(w, ldw, bias, beta, coef, B, N, C, w_out, bias_out, as_stream(stream));
}

extern "C" int ldx_layer_norm(const void* x, int64_t ldx_, int rows, int C, float eps, const float* gamma,
                              const float* beta, void* y, int64_t ldy, ldx_stream_t stream) {
  return layer_norm_impl(x, ldx_, rows, C, eps, gamma, beta, y, ldy, as_stream(stream));
}

static ldx_generic_op_t make_op(int kind) {
  ldx_generic_op_t op;
  memset(&op, 0, sizeof(op));
  op.kind = kind;
  return op;
}

extern "C" int ldx_nchw_to_nhwc_scaled(const float* x, const float* sigma, void* y, int B, int C, int H, int W, int Cpad,
                                       ldx_stream_t stream) {
  LDX_CHECK_ARG(x && y && Cpad >= C, "nchw_to_nhwc: bad args");
  ldx_generic_op_t op = make_op(LDX_OP_NCHW_TO_NHWC);
  op.p[0] = x; op.p[1] = sigma; op.p[2] = y;
  op.i[0] = B; op.i[1] = C; op.i[2] = H; op.i[3] = W; op.i[4] = Cpad;
  return run_generic_op(&op, as_stream(stream));
}

extern "C" int ldx_timestep_embedding(const float* sigma, const float* log_sigmas, int n_sigmas, void* emb, int B,
                                      int dim, ldx_stream_t stream) {
  LDX_CHECK_ARG(sigma && log_sigmas && emb && dim % 2 == 0, "timestep_embedding: bad args");
  ldx_generic_op_t op = make_op(LDX_OP_TIMESTEP_EMB);
  op.p[0] = sigma; op.p[1] = log_sigmas; op.p[2] = emb;
  op.i[0] = n_sigmas; op.i[1] = B; op.i[2] = dim;
  return run_generic_op(&op, as_stream(stream));
}

extern "C" int ldx_upsample_nearest2x(const void* x, void* y, int B, int H, int W, int C, ldx_stream_t stream) {
  LDX_CHECK_ARG(x && y, "upsample: null");
  ldx_generic_op_t op = make_op(LDX_OP_UPSAMPLE2X);
  op.p[0] = x; op.p[1] = y;
  op.i[0] = B; op.i[1] = H; op.i[2] = W; op.i[3] = C;
  return run_generic_op(&op, as_stream(stream));
}

extern "C" int ldx_eps_to_denoised(const void* eps, int64_t lde, const float* x, const float* sigma, float* den, int B,
                                   int C, int H, int W, ldx_stream_t stream) {
  LDX_CHECK_ARG(eps && x && sigma && den, "eps_to_denoised: null");
  ldx_generic_op_t op = make_op(LDX_OP_EPS_TO_DENOISED);
  op.p[0] = eps; op.p[1] = x; op.p[2] = sigma; op.p[3] = den;
  op.i[0] = B; op.i[1] = C; op.i[2] = H; op.i[3] = W; op.l[0] = lde;
  return run_generic_op(&op, as_stream(stream));
}

extern "C" int ldx_cfg_denoised(const void* eps, int64_t lde, const float* x, const float* sigma, float scale,
                                float* den, int b, int C, int H, int W, ldx_stream_t stream) {
  LDX_CHECK_ARG(eps && x && sigma && den, "cfg_denoised: null");
  ldx_generic_op_t op = make_op(LDX_OP_CFG_DENOISED);
  op.p[0] = eps; op.p[1] = x; op.p[2] = sigma; op.p[3] = den;
  op.i[0] = b; op.i[1] = C; op.i[2] = H; op.i[3] = W; op.l[0] = lde; op.f[0] = scale;
  return run_generic_op(&op, as_stream(stream));
}

extern "C" int ldx_cfg_dpmpp2m_step(const void* eps, int64_t lde, float* x, float* old_denoised, float* denoised_out,
                                    float sigma, float cfg_scale, float c_x, float c_d, float c_o, int have_old,
                                    int last, float sigma_next, float* sigma_io, int b, int C, int H, int W,
                                    ldx_stream_t stream) {
  LDX_CHECK_ARG(eps && x && old_denoised, "cfg_dpmpp2m_step: null");
  LDX_CHECK_ARG(b >= 1 && b <= 256 && C >= 1 && H >= 1 && W >= 1 && lde >= C, "cfg_dpmpp2m_step: bad extents");
  const int64_t n = (int64_t)b * C * H * W;
  cfg_dpmpp2m_step_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(
      (const __half*)eps, lde, x, old_denoised, denoised_out, sigma, cfg_scale, c_x, c_d, c_o, have_old, last,
      sigma_next, sigma_io, b, C, H * W);
  LDX_CUDA(cudaGetLastError());
  return LDX_OK;
}

extern "C" int ldx_cfg_combine(const float* cond, const float* uncond, float scale, float* out, int64_t n,
                               ldx_stream_t stream) {
  LDX_CHECK_ARG(cond && uncond && out, "cfg_combine: null");
  cfg_combine_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(cond, uncond, scale, out, n);
  LDX_CUDA(cudaGetLastError());
  return LDX_OK;
}

extern "C" int ldx_vae_postprocess(const void* dec, int64_t ldc, float* img, int64_t npix, ldx_stream_t stream) {
  LDX_CHECK_ARG(dec && img, "vae_postprocess: null");
  ldx_generic_op_t op = make_op(LDX_OP_VAE_POST);
  op.p[0] = dec; op.p[1] = img; op.l[0] = ldc; op.l[1] = npix;
  return run_generic_op(&op, as_stream(stream));
}

extern "C" int ldx_vae_postprocess_u8(const void* dec, int64_t ldc, void* img, int64_t npix, ldx_stream_t stream) {
  LDX_CHECK_ARG(dec && img && npix > 0 && ldc >= 3, "vae_postprocess_u8: null / empty");
  vae_post_u8_kernel<<<ew_blocks(npix * 3), 256, 0, as_stream(stream)>>>((const __half*)dec, ldc, (uint8_t*)img, npix);
  LDX_CUDA(cudaGetLastError());
  return LDX_OK;
}

extern "C" int ldx_vae_preprocess(const float* pixels, void* out, int64_t npix, ldx_stream_t stream) {
  LDX_CHECK_ARG(pixels && out && npix > 0, "vae_preprocess: null / empty");
  LDX_CHECK_ARG(((uintptr_t)out & 15) == 0, "vae_preprocess: out must be 16-byte aligned");
  ldx_generic_op_t op = make_op(LDX_OP_VAE_PRE);
  op.p[0] = pixels; op.p[1] = out; op.l[0] = npix;
  return run_generic_op(&op, as_stream(stream));
}

extern "C" int ldx_vae_sample_posterior(const float* moments, const float* noise, float* z, int b, int zc, int64_t hw,
                                        ldx_stream_t stream) {
  LDX_CHECK_ARG(moments && z && b > 0 && zc > 0 && hw > 0, "vae_sample_posterior: null / empty");
  ldx_generic_op_t op = make_op(LDX_OP_VAE_SAMPLE);
  op.p[0] = moments; op.p[1] = noise; op.p[2] = z; op.l[0] = (int64_t)b * zc * hw; op.l[1] = hw; op.i[0] = zc;
  return run_generic_op(&op, as_stream(stream));
}

extern "C" int ldx_bislerp_axis(const float* in, float* out, const float* ratio, const int32_t* idx1, const int32_t* idx2,
                                int n, int c, int64_t outer, int len_old, int len_new, int64_t inner, ldx_stream_t stream) {
  LDX_CHECK_ARG(in && out && ratio && idx1 && idx2, "bislerp_axis: null");
  LDX_CHECK_ARG(n > 0 && c > 0 && c <= BISLERP_MAX_C && outer > 0 && len_old > 0 && len_new > 0 && inner > 0,
                "bislerp_axis: bad shape (1 <= C <= 16)");
  const int64_t total = (int64_t)n * outer * len_new * inner;
  bislerp_axis_kernel<<<ew_blocks(total), 256, 0, as_stream(stream)>>>(in, out, ratio, idx1, idx2, total, c, outer, len_old,
                                                                      len_new, inner);
  LDX_CUDA(cudaGetLastError());
  return LDX_OK;
}

extern "C" int ldx_clip_embed(const int32_t* tokens, const float* tok_emb, const float* pos_emb, void* out, int64_t ldo,
                              int rows, int T, int C, ldx_stream_t stream) {
  LDX_CHECK_ARG(tokens && tok_emb && pos_emb && out, "clip_embed: null");
  LDX_CHECK_ARG(rows > 0 && T > 0 && C > 0 && ldo >= C, "clip_embed: bad shape");
  ldx_generic_op_t op = make_op(LDX_OP_CLIP_EMBED);
  op.p[0] = tokens; op.p[1] = tok_emb; op.p[2] = pos_emb; op.p[3] = out; op.l[0] = ldo;
  op.i[0] = rows; op.i[1] = T; op.i[2] = C;
  return run_generic_op(&op, as_stream(stream));
}

extern "C" int ldx_dpm_eps(const float* x, const float* denoised, float sigma, float* eps, int64_t n, ldx_stream_t stream) {
  LDX_CHECK_ARG(x && denoised && eps && n > 0, "dpm_eps: null / empty");
  LDX_CHECK_ARG(sigma > 0.f, "dpm_eps: sigma must be positive");
  dpm_eps_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(x, denoised, 1.0f / sigma, eps, n);
  LDX_CUDA(cudaGetLastError());
  return LDX_OK;
}

extern "C" int ldx_dpm_combine(const float* x, const float* eps, const float* eps2, float c1, float c2, float* out,
                               int64_t n, ldx_stream_t stream) {
  LDX_CHECK_ARG(x && eps && out && n > 0, "dpm_combine: null / empty");
  dpm_combine_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(x, eps, eps2, c1, c2, out, n);
  LDX_CUDA(cudaGetLastError());
  return LDX_OK;
}

extern "C" int ldx_dpm_error_norm(const float* x_low, const float* x_high, const float* x_prev, float atol, float rtol,
                                  int64_t n, void* workspace, float* err, ldx_stream_t stream) {
  LDX_CHECK_ARG(x_low && x_high && x_prev && workspace && err && n > 0, "dpm_error_norm: null / empty");
  int nb = ew_blocks(n);
  if (nb > LDX_DPM_ERROR_WS_DOUBLES) nb = LDX_DPM_ERROR_WS_DOUBLES;
  dpm_error_partial_kernel<<<nb, 256, 0, as_stream(stream)>>>(x_low, x_high, x_prev, atol, rtol, n, (double*)workspace);
  dpm_error_final_kernel<<<1, 32, 0, as_stream(stream)>>>((const double*)workspace, nb, 1.0 / sqrt((double)n), err);
  LDX_CUDA(cudaGetLastError());
  return LDX_OK;
}

extern "C" int ldx_softmax_rows(const float* s, int64_t lds, void* out, int64_t ldo, int64_t rows, int n, float scale,
                                ldx_stream_t stream) {
  LDX_CHECK_ARG(s && out, "softmax: null");
  ldx_generic_op_t op = make_op(LDX_OP_SOFTMAX);
  op.p[0] = s; op.p[1] = out; op.l[0] = lds; op.l[1] = ldo; op.l[2] = rows; op.i[0] = n; op.f[0] = scale;
  return run_generic_op(&op, as_stream(stream));
}

extern "C" int ldx_dpmpp2m_step(float* x, const float* denoised, const float* old_denoised, float c_x, float c_d,
                                float c_o, int64_t n, ldx_stream_t stream) {
  LDX_CHECK_ARG(x && denoised, "dpmpp2m_step: null");
  dpmpp2m_step_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(x, denoised, old_denoised, c_x, c_d, c_o, n);
  LDX_CUDA(cudaGetLastError());
  return LDX_OK;
}

extern "C" int ldx_euler_ancestral_step(float* x, const float* denoised, const float* noise, float sigma,
                                        float sigma_down, float noise_scale, int64_t n, ldx_stream_t stream) {
  LDX_CHECK_ARG(x && denoised, "euler_ancestral_step: null");
  euler_ancestral_step_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(x, denoised, noise, sigma,
                                                                           sigma_down - sigma, noise_scale, n);
  LDX_CUDA(cudaGetLastError());
  return LDX_OK;
}
