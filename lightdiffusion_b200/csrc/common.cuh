// Shared device-side primitives for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM wrappers.
// Hand-written inline PTX; no CUTLASS dependency.  Descriptor bit layouts follow the PTX ISA
// "tcgen05 shared memory descriptor" / "instruction descriptor" tables.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ldx {

#define LDX_DEVICE __device__ __forceinline__

LDX_DEVICE uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a fully converged warp (the same lane every time).  Keeping the surrounding control flow
// warp-uniform and predicating only the TMA / MMA issue on this lets the compiler hold descriptors in
// uniform registers; inside an `if (lane == 0)` region it emits a per-instruction R2UR waterfall loop
// (~100 cycles per tcgen05.mma), which makes the single MMA thread the bottleneck.
LDX_DEVICE bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// Programmatic dependent launch: let the next kernel in the stream start its prologue now, then wait
// until every kernel this one depends on has completed and flushed (no-op for a normal launch).
LDX_DEVICE void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ----------------------------------------------------------------------------- mbarrier
LDX_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
LDX_DEVICE void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
LDX_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
LDX_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a pipeline bug traps (clean CUDA error) instead of hanging the GPU.
#ifndef LDX_WAIT_LIMIT_CYCLES
#define LDX_WAIT_LIMIT_CYCLES (1ll << 32)
#endif
#ifndef LDX_TRYWAIT_HINT_NS
#define LDX_TRYWAIT_HINT_NS 0
#endif
LDX_DEVICE bool mbar_try_wait(uint32_t addr, uint32_t parity) {
  uint32_t done;
#if LDX_TRYWAIT_HINT_NS > 0
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(addr), "r"(parity), "r"((uint32_t)LDX_TRYWAIT_HINT_NS)
      : "memory");
#else
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(addr), "r"(parity)
      : "memory");
#endif
  return done != 0;
}
// The same on a precomputed shared-window address.  smem_u32_pinned() hides the conversion from the optimiser,
// which otherwise rematerialises the whole address computation (8-10 instructions) at every barrier operation of
// a register-tight loop.
LDX_DEVICE uint32_t smem_u32_pinned(const void* p) {
  uint32_t a;
  asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(a) : "l"(p));
  return a;
}
LDX_DEVICE void mbar_arrive_a(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
LDX_DEVICE void mbar_wait_a(uint32_t addr, uint32_t parity) {
  if (mbar_try_wait(addr, parity)) return;
  long long t0 = 0;
  for (uint32_t it = 1;; ++it) {
    if (mbar_try_wait(addr, parity)) break;
    if ((it & 1023u) == 0u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > LDX_WAIT_LIMIT_CYCLES) __trap();
    }
  }
}
LDX_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try_wait(addr, parity)) return;      // common case: the phase has already completed
  long long t0 = 0;
  for (uint32_t it = 1;; ++it) {
    if (mbar_try_wait(addr, parity)) break;
    if ((it & 1023u) == 0u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > LDX_WAIT_LIMIT_CYCLES) __trap();
    }
  }
}

// Same, for service warps whose wake-up latency is not critical: back off between polls so the spin loop
// does not take issue slots from the compute warps sharing the scheduler.
LDX_DEVICE void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t sleep_ns) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try_wait(addr, parity)) return;
  long long t0 = 0;
  for (uint32_t it = 1;; ++it) {
    __nanosleep(sleep_ns);
    if (mbar_try_wait(addr, parity)) break;
    if ((it & 1023u) == 0u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > LDX_WAIT_LIMIT_CYCLES) __trap();
    }
  }
}

// generic-proxy writes to smem -> visible to the async proxy (TMA / tcgen05.mma operand reads)
LDX_DEVICE void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ----------------------------------------------------------------------------- TMA
LDX_DEVICE void tma_prefetch_desc(const void* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
LDX_DEVICE void tma_load_2d(void* dst, const void* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
LDX_DEVICE void tma_load_4d(void* dst, const void* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---- CTA-pair (cta_group::2) variants: the loads of both CTAs signal the LEADER CTA's mbarrier
LDX_DEVICE void tma_load_2d_pair(void* dst, const void* map, uint32_t leader_bar_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar_addr), "r"(c0), "r"(c1)
      : "memory");
}
LDX_DEVICE void tma_load_4d_pair(void* dst, const void* map, uint32_t leader_bar_addr, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
LDX_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
LDX_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank` of the cluster
LDX_DEVICE uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on a peer CTA's mbarrier.  Default semantics (.release.cta): the data this orders is TMEM, which the
// tcgen05 fences cover; a cluster-scope release made every arrive wait ~1000 cycles for the thread's
// outstanding global traffic.
LDX_DEVICE void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// smem -> global tensor store (bulk async group); OOB parts of the box are clipped
LDX_DEVICE void tma_store_4d(const void* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
LDX_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
LDX_DEVICE void tma_store_wait_read() {   // at most N committed groups still reading their smem source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
LDX_DEVICE void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
LDX_DEVICE void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------- tcgen05 / TMEM
LDX_DEVICE void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
LDX_DEVICE void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
LDX_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
LDX_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
LDX_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers fp16/bf16 inputs with fp32 accumulate.
LDX_DEVICE void umma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all tcgen05 ops previously issued by this thread have completed
LDX_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- cta_group::2: one MMA spans the CTA pair (M = 256: 128 rows in each CTA's TMEM); issued by the leader
LDX_DEVICE void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
LDX_DEVICE void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
LDX_DEVICE void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
LDX_DEVICE void umma_f16_ss_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at the same smem offset in every CTA of `mask` once prior MMAs have completed
LDX_DEVICE void umma_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// TMEM -> registers: warp w reads lanes 32*(w%4).., thread = lane (row), 32 consecutive fp32 columns
LDX_DEVICE void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
LDX_DEVICE void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
LDX_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// registers -> TMEM (same lane/column mapping as tmem_ld32)
LDX_DEVICE void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};" ::"r"(r[0]),
      "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
      "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),
      "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr)
      : "memory");
}
LDX_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (64-bit):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4,
//   [46,48) version = 1 on sm_100, [61,64) layout: 0 none, 2 = SWIZZLE_128B, 4 = 64B, 6 = 32B.
// K-major SW128 tile [rows][64 fp16]: rows are 128 B apart, 8-row groups 1024 B apart (SBO); LBO unused (=1).
LDX_DEVICE uint64_t umma_desc_kmajor_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major SW128 tile [k rows][64 fp16 of MN]: 64 MN elements contiguous (128 B), k rows 128 B apart,
// 8-row k groups 1024 B apart (SBO), next 64-wide MN block `lbo_bytes` away (LBO).
LDX_DEVICE uint64_t umma_desc_mnmajor_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor (32-bit) for kind::f16: [4,6) D fmt (1 = f32), [7,10) A fmt (0 f16, 1 bf16),
// [10,13) B fmt, [15] A major (0 = K), [16] B major (0 = K, 1 = MN), [17,23) N >> 3, [24,29) M >> 4.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ----------------------------------------------------------------------------- small math
// single MUFU.EX2 (exp2f() expands to range fix-ups around it; softmax arguments are <= ~8)
LDX_DEVICE float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// x * sigmoid(x) with one MUFU.EX2 and one MUFU.RCP (an IEEE division costs ~10 more issue slots per
// element, which made the GroupNorm+SiLU pass issue-bound instead of bandwidth-bound)
LDX_DEVICE float silu_f(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + ex2_approx(-1.4426950408889634f * x)));
  return x * r;
}
// x * sigmoid(k x): k = 1 is SiLU, k = 1.702 the CLIP text tower's quick_gelu (LD.py:4296-4299)
LDX_DEVICE float sigmoid_gate_f(float x, float neg_k_log2e) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + ex2_approx(neg_k_log2e * x)));
  return x * r;
}
LDX_DEVICE float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// exact-erf GELU, x * Phi(x), as x * sigmoid(q(x)) with q = logit(Phi) fitted by an odd degree-9 polynomial
// (max abs error 3.7e-6 over all x, below fp16 output rounding wherever |gelu| > 1e-2; the tails saturate
// through exp2 -> 0 / inf).  6 FMA-pipe ops + MUFU.EX2 + MUFU.RCP: the Abramowitz-Stegun erf it replaces cost
// 20 + 2 and made the GEGLU epilogue the bottleneck of the feed-forward projection.
LDX_DEVICE float gelu_erf_fast(float x) {
  const float x2 = x * x;
  // coefficients of q(x) / x, pre-multiplied by -log2(e)
  float p = fmaf(-3.2919935e-06f, x2, 8.9313635e-05f);
  p = fmaf(p, x2, 3.5481673e-04f);
  p = fmaf(p, x2, -1.0521783e-01f);
  p = fmaf(p, x2, -2.3020485e+00f);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + ex2_approx(x * p)));
  return x * r;
}

// Packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: two IEEE fp32 results per issue slot, bit-identical to the scalar ops).
// Epilogues whose limiter is the warp scheduler's issue rate, not the FMA lanes, halve their arithmetic instruction count.
typedef unsigned long long f32x2_t;
LDX_DEVICE f32x2_t f2_pack(float a, float b) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
LDX_DEVICE void f2_unpack(f32x2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
#ifndef LDX_PACKED_F32
#define LDX_PACKED_F32 1      // 0: the same helpers through scalar fmaf / mul / add (A/B builds; results are bit-identical)
#endif
#if LDX_PACKED_F32
LDX_DEVICE f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
LDX_DEVICE f32x2_t f2_mul(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
#else
LDX_DEVICE f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
  float a0, a1, b0, b1, c0, c1;
  f2_unpack(a, a0, a1); f2_unpack(b, b0, b1); f2_unpack(c, c0, c1);
  return f2_pack(__fmaf_rn(a0, b0, c0), __fmaf_rn(a1, b1, c1));
}
LDX_DEVICE f32x2_t f2_mul(f32x2_t a, f32x2_t b) {
  float a0, a1, b0, b1;
  f2_unpack(a, a0, a1); f2_unpack(b, b0, b1);
  return f2_pack(__fmul_rn(a0, b0), __fmul_rn(a1, b1));
}
#endif
// gelu_erf_fast on a pair: the same operations in the same order (bit-identical results), 8 packed + 4 MUFU instructions
LDX_DEVICE f32x2_t gelu_erf_fast2(f32x2_t x) {
  const f32x2_t x2 = f2_mul(x, x);
  f32x2_t p = f2_fma(f2_pack(-3.2919935e-06f, -3.2919935e-06f), x2, f2_pack(8.9313635e-05f, 8.9313635e-05f));
  p = f2_fma(p, x2, f2_pack(3.5481673e-04f, 3.5481673e-04f));
  p = f2_fma(p, x2, f2_pack(-1.0521783e-01f, -1.0521783e-01f));
  p = f2_fma(p, x2, f2_pack(-2.3020485e+00f, -2.3020485e+00f));
  float t0, t1;
  f2_unpack(f2_mul(x, p), t0, t1);
  float d0, d1;
  f2_unpack(f2_fma(f2_pack(ex2_approx(t0), ex2_approx(t1)), f2_pack(1.0f, 1.0f), f2_pack(1.0f, 1.0f)), d0, d1);   // e * 1 + 1 == 1 + e
  float r0, r1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
  return f2_mul(x, f2_pack(r0, r1));
}

LDX_DEVICE uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
LDX_DEVICE float2 unpack_h2(uint32_t v) {
  __half2 h = *reinterpret_cast<__half2*>(&v);
  return __half22float2(h);
}

}  // namespace ldx
