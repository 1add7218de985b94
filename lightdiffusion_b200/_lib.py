"""ctypes binding of libldx.so (include/ldx.h).  There is no fallback: a missing library or a
failing kernel raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LDX_LIB_PATH") or os.path.join(_HERE, "libldx.so")   # override: kernel-variant experiments

LDX_MAX_SRC = 4
LDX_MAX_SEG = 16

EPI_SILU = 1
EPI_GEGLU = 2
EPI_OUT_F32 = 4
EPI_ROWNORM = 8
EPI_QUICK_GELU = 16
EPI_LRELU = 32
EPI_RELU = 64
DPM_ERROR_WS_DOUBLES = 2048

OP_GEMM, OP_GROUP_NORM, OP_LAYER_NORM, OP_ATTENTION = 1, 2, 3, 4
OP_NCHW_TO_NHWC, OP_TIMESTEP_EMB, OP_UPSAMPLE2X, OP_EPS_TO_DENOISED = 5, 6, 7, 8
OP_CFG_DENOISED, OP_VAE_POST, OP_SOFTMAX = 9, 10, 11
OP_DUP, OP_VAE_PRE, OP_VAE_SAMPLE, OP_CLIP_EMBED, OP_CONV_SMALL, OP_GROUP_NORM_COEF = 12, 13, 14, 15, 16, 17
OP_FF_FUSED = 18
OP_GN_FOLD_WEIGHTS = 19

GN_STATS_FLOATS_PER_BATCH = 64 * 32 * 2      # ldx_group_norm workspace: B * this many floats


class LdxError(RuntimeError):
    pass


class Src(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("C", C.c_int32), ("W", C.c_int32), ("H", C.c_int32), ("B", C.c_int32),
                ("sw", C.c_int64), ("sh", C.c_int64), ("sb", C.c_int64)]


class Seg(C.Structure):
    _fields_ = [("src", C.c_int32), ("c0", C.c_int32), ("dx", C.c_int32), ("dy", C.c_int32), ("nchunks", C.c_int32)]


class GemmDesc(C.Structure):
    _fields_ = [("src", Src * LDX_MAX_SRC), ("nsrc", C.c_int32),
                ("seg", Seg * LDX_MAX_SEG), ("nseg", C.c_int32),
                ("w", C.c_void_p), ("ldw", C.c_int64), ("N", C.c_int32),
                ("W", C.c_int32), ("H", C.c_int32), ("B", C.c_int32),
                ("out", C.c_void_p),
                ("o_sx", C.c_int64), ("o_sy", C.c_int64), ("o_sb", C.c_int64), ("o_off", C.c_int64),
                ("bias", C.c_void_p), ("rowadd", C.c_void_p), ("ld_rowadd", C.c_int64),
                ("residual", C.c_void_p), ("ldr", C.c_int64),
                ("flags", C.c_int32), ("bn", C.c_int32),
                ("rowstats_out", C.c_void_p), ("rowstats_in", C.c_void_p), ("rowstats_in_parts", C.c_int32),
                ("ln_eps", C.c_float), ("ln_gsum", C.c_void_p), ("ln_c", C.c_int32),
                ("colstats_out", C.c_void_p), ("colstats_slots", C.c_int32), ("colstats_slot_base", C.c_int32),
                ("rowstats_out_parts", C.c_int32),
                ("halo_nsrc", C.c_int32), ("gn_coef", C.c_void_p), ("gn_ctot", C.c_int32), ("gn_silu", C.c_int32),
                ("w_batch_rows", C.c_int32), ("bias_batch_stride", C.c_int32)]


class GenericOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("i", C.c_int32 * 16), ("l", C.c_int64 * 8), ("f", C.c_float * 4),
                ("p", C.c_void_p * 10)]


_lib = None

_SIGS = {
    "ldx_version": (C.c_int, []),
    "ldx_sizeof": (C.c_int, [C.c_int]),
    "ldx_last_error_string": (C.c_char_p, []),
    "ldx_device_check": (C.c_int, []),
    "ldx_gemm": (C.c_int, [C.POINTER(GemmDesc), C.c_void_p]),
    "ldx_gemm_rowstats_parts": (C.c_int, [C.POINTER(GemmDesc)]),
    "ldx_gemm_colstats_slots": (C.c_int, [C.POINTER(GemmDesc)]),
    "ldx_pack_conv3x3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ldx_pack_upconv3x3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ldx_conv3x3_small": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]),
    "ldx_pack_conv3x3_small": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ldx_conv3x3_small_packed_elems": (C.c_int, [C.c_int, C.c_int]),
    "ldx_pack_geglu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ldx_group_norm": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int,
                                 C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                 C.c_void_p]),
    "ldx_group_norm_cs": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int,
                                    C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "ldx_group_norm_coef": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int,
                                      C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "ldx_ff_geglu_fused": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p,
                                     C.c_int64, C.c_void_p]),
    "ldx_gn_fold_weights": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldx_layer_norm": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int64, C.c_void_p]),
    "ldx_attention": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ldx_nchw_to_nhwc_scaled": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_void_p]),
    "ldx_timestep_embedding": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ldx_upsample_nearest2x": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ldx_eps_to_denoised": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p]),
    "ldx_cfg_denoised": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ldx_cfg_dpmpp2m_step": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                       C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_float, C.c_void_p,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ldx_cfg_combine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int64, C.c_void_p]),
    "ldx_vae_preprocess": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ldx_vae_sample_posterior": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p]),
    "ldx_bislerp_axis": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64,
                                   C.c_int, C.c_int, C.c_int64, C.c_void_p]),
    "ldx_clip_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p]),
    "ldx_attention_causal": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                       C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ldx_dpm_eps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int64, C.c_void_p]),
    "ldx_dpm_combine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_int64,
                                  C.c_void_p]),
    "ldx_dpm_error_norm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "ldx_vae_postprocess": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "ldx_vae_postprocess_u8": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "ldx_softmax_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_float,
                                   C.c_void_p]),
    "ldx_dpmpp2m_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int64,
                                   C.c_void_p]),
    "ldx_euler_ancestral_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                           C.c_int64, C.c_void_p]),
    "ldx_program_create": (C.c_void_p, []),
    "ldx_program_add_gemm": (C.c_int, [C.c_void_p, C.POINTER(GemmDesc)]),
    "ldx_program_add_op": (C.c_int, [C.c_void_p, C.POINTER(GenericOp)]),
    "ldx_program_num_launches": (C.c_int, [C.c_void_p]),
    "ldx_program_run": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ldx_program_num_ops": (C.c_int, [C.c_void_p]),
    "ldx_program_run_timed": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    "ldx_program_destroy": (None, [C.c_void_p]),
}

EXPORTS = tuple(_SIGS)


def lib():
    """Load libldx.so (once). Raises LdxError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LdxError(f"{LIB_PATH} is missing: build it with lightdiffusion_b200/csrc/build.sh "
                           "(or __graft_entry__.build()); there is no CPU fallback")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        for what, st in enumerate((GemmDesc, GenericOp, Src, Seg)):
            if l.ldx_sizeof(what) != C.sizeof(st):
                raise LdxError(f"ABI mismatch: sizeof({st.__name__}) {C.sizeof(st)} != {l.ldx_sizeof(what)}")
        _lib = l
    return _lib


def check(rc: int, what: str = "ldx"):
    if rc != 0:
        msg = lib().ldx_last_error_string()
        raise LdxError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def require_device():
    check(lib().ldx_device_check(), "ldx_device_check")
