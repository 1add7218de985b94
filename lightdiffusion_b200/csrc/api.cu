// C-ABI glue: error reporting, device check and recorded programs (ldx_program_*).
#include <string.h>

#include <vector>

#include "internal.h"

namespace ldx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return LDX_E_CUDA;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev;
}

int num_sms() {
  static std::atomic<int> cache[LDX_MAX_DEVICES];      // zero-initialised; one slot per device ordinal
  const int dev = current_device();
  std::atomic<int>& slot = cache[dev & (LDX_MAX_DEVICES - 1)];
  int n = slot.load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    slot.store(n, std::memory_order_relaxed);
  }
  return n;
}

int group_norm_num_launches(int C, int HW);
int attention_prepare_op(const ldx_generic_op_t* op, void** handle);
int attention_run_handle(void* handle, cudaStream_t st);
void attention_free_handle(void* handle);
int conv_small_prepare_op(const ldx_generic_op_t* op, void** handle);
int conv_small_run_handle(void* handle, cudaStream_t st);
void conv_small_free_handle(void* handle);
int ff_fused_prepare_op(const ldx_generic_op_t* op, void** handle);
int ff_fused_run_handle(void* handle, cudaStream_t st);
void ff_fused_free_handle(void* handle);

struct ProgOp {
  int kind;
  GemmLaunch* gemm = nullptr;
  void* attn = nullptr;
  void* conv_small = nullptr;
  void* ff = nullptr;
  ldx_generic_op_t gen;
  int launches = 1;
  void* ws = nullptr;          // split-K workspace (cudaMalloc)
};

}  // namespace ldx

using namespace ldx;

struct ldx_program {
  std::vector<ProgOp> ops;
};

extern "C" int ldx_version(void) { return 1; }

extern "C" int ldx_sizeof(int what) {
  switch (what) {
    case 0: return (int)sizeof(ldx_gemm_desc_t);
    case 1: return (int)sizeof(ldx_generic_op_t);
    case 2: return (int)sizeof(ldx_src_t);
    case 3: return (int)sizeof(ldx_seg_t);
  }
  return -1;
}

extern "C" const char* ldx_last_error_string(void) { return g_err; }

extern "C" int ldx_device_check(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device (%s)", e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
    return LDX_E_NODEVICE;
  }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    set_error("device compute capability %d.x is not sm_100", major);
    return LDX_E_NODEVICE;
  }
  return LDX_OK;
}

extern "C" ldx_program_t* ldx_program_create(void) { return new ldx_program; }

extern "C" int ldx_program_add_gemm(ldx_program_t* prog, const ldx_gemm_desc_t* d) {
  LDX_CHECK_ARG(prog && d, "program_add_gemm: null");
  GemmLaunch* g = new GemmLaunch;
  int rc = gemm_prepare(d, g);
  if (rc) {
    delete g;
    return rc;
  }
  ProgOp op;
  op.kind = LDX_OP_GEMM;
  op.gemm = g;
  if (g->nsplit > 1) {                      // split-K: the program owns the fp32 partial-sum workspace
    void* ws = nullptr;
    if (cudaMalloc(&ws, g->ws_bytes) != cudaSuccess) {
      delete g;
      set_error("split-K workspace allocation of %zu bytes failed", g->ws_bytes);
      return LDX_E_CUDA;
    }
    gemm_bind_workspace(g, ws);
    op.ws = ws;
    op.launches = 2;
  }
  prog->ops.push_back(op);
  return LDX_OK;
}

extern "C" int ldx_program_add_op(ldx_program_t* prog, const ldx_generic_op_t* gop) {
  LDX_CHECK_ARG(prog && gop, "program_add_op: null");
  ProgOp op;
  op.kind = gop->kind;
  op.gen = *gop;
  if (gop->kind == LDX_OP_ATTENTION) {
    int rc = attention_prepare_op(gop, &op.attn);
    if (rc) return rc;
  } else if (gop->kind == LDX_OP_CONV_SMALL) {
    int rc = conv_small_prepare_op(gop, &op.conv_small);
    if (rc) return rc;
  } else if (gop->kind == LDX_OP_FF_FUSED) {
    int rc = ff_fused_prepare_op(gop, &op.ff);
    if (rc) return rc;
  } else if (gop->kind == LDX_OP_GROUP_NORM) {
    op.launches = group_norm_num_launches(gop->i[0] + gop->i[1], gop->i[3]);
  } else if (gop->kind == LDX_OP_DUP) {
    op.launches = 0;             // two device-to-device copies, no kernel of ours
  } else if (gop->kind == LDX_OP_GEMM) {
    set_error("use ldx_program_add_gemm for GEMM ops");
    return LDX_E_INVALID;
  }
  prog->ops.push_back(op);
  return LDX_OK;
}

extern "C" int ldx_program_num_launches(const ldx_program_t* prog) {
  if (!prog) return 0;
  int n = 0;
  for (const ProgOp& op : prog->ops) n += op.launches;
  return n;
}

extern "C" int ldx_program_run(ldx_program_t* prog, ldx_stream_t stream) {
  LDX_CHECK_ARG(prog, "program_run: null");
  cudaStream_t st = as_stream(stream);
  for (size_t i = 0; i < prog->ops.size(); ++i) {
    const ProgOp& op = prog->ops[i];
    int rc;
    if (op.kind == LDX_OP_GEMM) rc = gemm_launch(op.gemm, st);
    else if (op.kind == LDX_OP_ATTENTION) rc = attention_run_handle(op.attn, st);
    else if (op.kind == LDX_OP_CONV_SMALL) rc = conv_small_run_handle(op.conv_small, st);
    else if (op.kind == LDX_OP_FF_FUSED) rc = ff_fused_run_handle(op.ff, st);
    else rc = run_generic_op(&op.gen, st);
    if (rc) return rc;
  }
  return LDX_OK;
}

extern "C" int ldx_program_num_ops(const ldx_program_t* prog) { return prog ? (int)prog->ops.size() : 0; }

// Replays the program with a CUDA event between consecutive ops (on `stream`), synchronises, and
// returns the device time of every op in milliseconds.  Used by bench.py for the per-kernel roofline.
extern "C" int ldx_program_run_timed(ldx_program_t* prog, ldx_stream_t stream, float* ms_out, int n) {
  LDX_CHECK_ARG(prog && ms_out && n == (int)prog->ops.size(), "program_run_timed: bad args");
  cudaStream_t st = as_stream(stream);
  std::vector<cudaEvent_t> ev(n + 1, nullptr);
  int rc = LDX_OK;
  for (int i = 0; i <= n && rc == LDX_OK; ++i) {
    cudaError_t e = cudaEventCreate(&ev[i]);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaEventCreate");
  }
  if (rc == LDX_OK) {
    cudaError_t e = cudaEventRecord(ev[0], st);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaEventRecord");
  }
  for (int i = 0; i < n && rc == LDX_OK; ++i) {
    const ProgOp& op = prog->ops[i];
    if (op.kind == LDX_OP_GEMM) rc = gemm_launch(op.gemm, st);
    else if (op.kind == LDX_OP_ATTENTION) rc = attention_run_handle(op.attn, st);
    else if (op.kind == LDX_OP_CONV_SMALL) rc = conv_small_run_handle(op.conv_small, st);
    else if (op.kind == LDX_OP_FF_FUSED) rc = ff_fused_run_handle(op.ff, st);
    else rc = run_generic_op(&op.gen, st);
    if (rc == LDX_OK && cudaEventRecord(ev[i + 1], st) != cudaSuccess) rc = LDX_E_CUDA;
  }
  if (rc == LDX_OK) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
  }
  if (rc == LDX_OK)
    for (int i = 0; i < n; ++i) cudaEventElapsedTime(&ms_out[i], ev[i], ev[i + 1]);
  for (int i = 0; i <= n; ++i)
    if (ev[i]) cudaEventDestroy(ev[i]);
  return rc;
}

extern "C" void ldx_program_destroy(ldx_program_t* prog) {
  if (!prog) return;
  for (ProgOp& op : prog->ops) {
    delete op.gemm;
    if (op.ws) cudaFree(op.ws);
    if (op.attn) attention_free_handle(op.attn);
    if (op.conv_small) conv_small_free_handle(op.conv_small);
    if (op.ff) ff_fused_free_handle(op.ff);
  }
  delete prog;
}
