"""ctypes binding of libmsgm_b200.so (include/msgm_b200.h).  PyTorch is used only for device memory / streams.

There is no CPU fallback: importing this module is cheap, but the first call that needs a kernel raises
``RuntimeError`` if the library was not built or no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# MSGM_LIB_VARIANT=<name> loads libmsgm_b200_<name>.so instead (kernel A/B experiments, tools/tc_variants.py)
LIB_PATH = os.path.join(HERE, "libmsgm_b200" + ("_" + os.environ["MSGM_LIB_VARIANT"] if os.environ.get("MSGM_LIB_VARIANT")
                                                else "") + ".so")

SDE_SGM, SDE_MSGM_DENSE, SDE_MSGM_SPARSE = 0, 1, 2
SCHEME_EM, SCHEME_HEUN, SCHEME_RK4 = 0, 1, 2
PREC_FP32, PREC_F16TC = 0, 1
ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE = -1, -2, -3, -4


class SdeDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dim", C.c_int32), ("beta_min", C.c_float), ("beta_delta", C.c_float),
                ("T", C.c_float), ("G", C.c_void_p), ("L_G", C.c_void_p)]


class MlpDesc(C.Structure):
    _fields_ = [("input_dim", C.c_int32), ("premodule", C.c_int32), ("W", C.c_void_p * 4), ("b", C.c_void_p * 4)]


class SampleArgs(C.Structure):
    _fields_ = [("scheme", C.c_int32), ("num_steps", C.c_int32), ("lmbd", C.c_float),
                ("norm_correction", C.c_int32), ("include_t0", C.c_int32), ("forward_only", C.c_int32),
                ("precision", C.c_int32), ("T_", C.c_float), ("ts", C.c_void_p), ("noise", C.c_void_p),
                ("seed", C.c_uint64), ("particle_offset", C.c_uint64), ("traj", C.c_void_p),
                ("keep_step", C.c_void_p), ("keep_out", C.c_void_p), ("T_rows", C.c_void_p)]


class StepClock(C.Structure):
    _fields_ = [("clock", C.c_void_p), ("s_table", C.c_void_p), ("s_next", C.c_void_p), ("traj", C.c_void_p),
                ("keep_step", C.c_void_p), ("keep_out", C.c_void_p), ("include_t0", C.c_int32), ("reserved", C.c_int32)]


class GemmProblem(C.Structure):
    _fields_ = [("A", C.c_void_p * 2), ("B", C.c_void_p * 2), ("C", C.c_void_p), ("stride_a", C.c_int64 * 2),
                ("stride_b", C.c_int64 * 2), ("stride_c", C.c_int64), ("lda", C.c_int32 * 2), ("ldb", C.c_int32 * 2),
                ("ldc", C.c_int32), ("trans_a", C.c_int32 * 2), ("trans_b", C.c_int32 * 2), ("nseg", C.c_int32),
                ("accumulate", C.c_int32), ("alpha", C.c_float), ("reserved", C.c_int32)]


class Conv1dDesc(C.Structure):
    _fields_ = [("x1", C.c_void_p), ("x2", C.c_void_p), ("W", C.c_void_p), ("bias", C.c_void_p), ("E", C.c_void_p),
                ("out", C.c_void_p)] + [(n, C.c_int32) for n in ("B", "C1", "C2", "Cemb", "Cout", "K", "stride", "pad",
                                                                 "Lin", "Lout", "gelu")]


class Conv2dDesc(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("x1", "x2", "W", "bias", "ebias", "res", "stats", "gamma", "beta", "out")] + \
               [(n, C.c_int32) for n in ("B", "C1", "C2", "Cout", "K", "stride", "up", "Hs", "Ws", "G", "prologue")]


class Conv2dTcDesc(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("x1", "x2", "wimg", "bias", "ebias", "res", "ss", "out")] + \
               [(n, C.c_int32) for n in ("B", "C1", "C2", "Cout", "K", "stride", "up", "Hs", "Ws", "prologue", "fast")]


class Conv1dTcDesc(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("x1", "x2", "wimg", "bias", "E", "out")] + \
               [(n, C.c_int32) for n in ("B", "C1", "C2", "Cout", "K", "stride", "Lin", "gelu", "fast")]


class Conv1dTcpDesc(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("x1", "x2", "wimg", "bias", "E", "out_planes", "out_f32")] + \
               [(n, C.c_int32) for n in ("B", "C1", "C2", "Cout", "K", "Lin", "Lout", "gelu", "transposed", "fast")]


class EmbFoldMultiDesc(C.Structure):
    _fields_ = [("W", C.c_void_p * 16), ("E", C.c_void_p * 16), ("Cw", C.c_int32 * 16), ("Coff", C.c_int32 * 16),
                ("Cout", C.c_int32 * 16), ("K", C.c_int32 * 16), ("n", C.c_int32), ("Cemb", C.c_int32), ("B", C.c_int32),
                ("emb", C.c_void_p)]


_lib = None
_lock = threading.Lock()
_ctx = {}

# every symbol include/msgm_b200.h declares (tests/test_abi.py checks the .so exports all of them)
SYMBOLS = ["msgm_abi_version", "msgm_last_error", "msgm_create", "msgm_destroy", "msgm_launch_count", "msgm_async_error", "msgm_row_norm_stats", "msgm_survival_counts", "msgm_moments", "msgm_adam_step", "msgm_p2p_create", "msgm_p2p_connect", "msgm_p2p_disconnect", "msgm_p2p_destroy",
           "msgm_p2p_allreduce_adam", "msgm_pair_act", "msgm_amax", "msgm_amax2", "msgm_pow2_scale", "msgm_rows_bias_add", "msgm_channel_sums", "msgm_tap_sums_1d",
           "msgm_conv_wgrad", "msgm_tc_range_scale", "msgm_conv_tc_pack_dgrad", "msgm_conv_wgrad_tc_ok", "msgm_conv_wgrad_tc_scratch_bytes", "msgm_conv_wgrad_tc", "msgm_gemm_f32", "msgm_premodule_pair", "msgm_sparse_ssm_loss", "msgm_bgemm_f32", "msgm_gemm_group_f32", "msgm_gn_pair", "msgm_softmax_pair",
           "msgm_sincos_pair", "msgm_resample2",
           "msgm_sample_mlp", "msgm_noise_forward", "msgm_ssm_prepare", "msgm_mlp_forward", "msgm_debug_flags", "msgm_debug_counters",
           "msgm_ssm_scratch_bytes", "msgm_ssm_mlp_forward", "msgm_ssm_mlp_backward", "msgm_ssm_tc_scratch_bytes",
           "msgm_ssm_mlp_fwd_bwd_tc",
           "msgm_stage_update", "msgm_stage_update_clocked", "msgm_philox_normal_clocked", "msgm_clock_advance", "msgm_row_norm",
           "msgm_philox_normal", "msgm_latent_sample", "msgm_mmd_sums", "msgm_kde_logpdf",
           "msgm_conv1d", "msgm_emb_fold", "msgm_convt1d_k4s2", "msgm_embed_mlp", "msgm_normalize_log_radius",
           "msgm_conv2d", "msgm_gn_stats", "msgm_emb_proj", "msgm_sincos_embed_mlp", "msgm_attention", "msgm_vort_pre",
           "msgm_vort_post", "msgm_conv2d_tc", "msgm_conv2d_tc_pack_bytes", "msgm_conv2d_tc_pack", "msgm_gn_scale_shift",
           "msgm_attention_tc_supported", "msgm_attention_tc", "msgm_attention_proj_tc_supported", "msgm_attention_proj_tc", "msgm_conv1d_tc", "msgm_conv1d_tc_pack_bytes",
           "msgm_conv1d_tc_pack", "msgm_convt1d_tc_pack", "msgm_convt1d_tc", "msgm_emb_proj_multi",
           "msgm_emb_fold_multi", "msgm_embed_mlp2", "msgm_planes_bytes", "msgm_conv1d_tcp", "msgm_planes_pack", "msgm_planes_unpack", "msgm_conv1d_first_planes"]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.isfile(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -m sdeflow_light_b200.build` "
                        "(sdeflow_light_b200 has no CPU / eager fallback)")
                L = C.CDLL(LIB_PATH)
                L.msgm_last_error.restype = C.c_char_p
                L.msgm_launch_count.restype = C.c_int64
                L.msgm_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
                L.msgm_destroy.argtypes = [C.c_void_p]
                L.msgm_launch_count.argtypes = [C.c_void_p]
                L.msgm_debug_flags.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
                L.msgm_async_error.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
                L.msgm_ssm_scratch_bytes.restype = C.c_uint64
                L.msgm_ssm_scratch_bytes.argtypes = [C.c_int64]
                L.msgm_ssm_mlp_forward.argtypes = [C.c_void_p, C.POINTER(SdeDesc), C.POINTER(MlpDesc)] + [C.c_void_p] * 5 + \
                    [C.c_int64, C.c_void_p]
                L.msgm_ssm_mlp_backward.argtypes = [C.c_void_p, C.POINTER(SdeDesc), C.POINTER(MlpDesc)] + [C.c_void_p] * 6 + \
                    [C.c_int64, C.c_void_p]
                L.msgm_ssm_tc_scratch_bytes.restype = C.c_uint64
                L.msgm_ssm_tc_scratch_bytes.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64]
                L.msgm_ssm_mlp_fwd_bwd_tc.argtypes = [C.c_void_p, C.POINTER(SdeDesc), C.POINTER(MlpDesc)] + [C.c_void_p] * 7 + \
                    [C.c_float, C.c_int64, C.c_void_p]
                L.msgm_stage_update.argtypes = [C.c_void_p, C.POINTER(SdeDesc), C.c_int32, C.c_int32, C.c_float, C.c_int32,
                                                C.c_int32, C.c_float, C.c_float] + [C.c_void_p] * 6 + [C.c_int64, C.c_void_p]
                L.msgm_stage_update_clocked.argtypes = [C.c_void_p, C.POINTER(SdeDesc), C.c_int32, C.c_int32, C.c_float, C.c_int32,
                                                        C.c_int32, C.POINTER(StepClock), C.c_float] + [C.c_void_p] * 6 + \
                    [C.c_int64, C.c_void_p]
                L.msgm_philox_normal_clocked.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64,
                                                         C.c_uint64, C.POINTER(StepClock), C.c_int32, C.c_void_p, C.c_void_p]
                L.msgm_clock_advance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
                L.msgm_row_norm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]
                L.msgm_philox_normal.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64,
                                                 C.c_uint64, C.c_uint32, C.c_void_p]
                L.msgm_latent_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p]
                L.msgm_mmd_sums.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                            C.c_void_p]
                L.msgm_conv1d.argtypes = [C.c_void_p, C.POINTER(Conv1dDesc), C.c_void_p]
                L.msgm_emb_fold_multi.argtypes = [C.c_void_p, C.POINTER(EmbFoldMultiDesc), C.c_void_p]
                L.msgm_embed_mlp2.argtypes = [C.c_void_p] * 12 + [C.c_int32, C.c_int32, C.c_void_p]
                L.msgm_planes_bytes.restype = C.c_int64
                L.msgm_planes_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32]
                L.msgm_conv1d_tcp.argtypes = [C.c_void_p, C.POINTER(Conv1dTcpDesc), C.c_void_p]
                L.msgm_planes_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
                L.msgm_planes_unpack.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
                L.msgm_conv1d_first_planes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                                       C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
                L.msgm_emb_fold.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 6 + [C.c_void_p]
                L.msgm_convt1d_k4s2.argtypes = [C.c_void_p] * 5 + [C.c_int32] * 5 + [C.c_void_p]
                L.msgm_embed_mlp.argtypes = [C.c_void_p] * 7 + [C.c_int32] * 3 + [C.c_void_p]
                L.msgm_normalize_log_radius.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 2 + [C.c_void_p]
                L.msgm_conv2d.argtypes = [C.c_void_p, C.POINTER(Conv2dDesc), C.c_void_p]
                L.msgm_conv2d_tc.argtypes = [C.c_void_p, C.POINTER(Conv2dTcDesc), C.c_void_p]
                L.msgm_conv2d_tc_pack_bytes.restype = C.c_int64
                L.msgm_conv2d_tc_pack_bytes.argtypes = [C.c_int32] * 3
                L.msgm_conv2d_tc_pack.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int32] * 3 + [C.c_void_p] * 2
                L.msgm_conv1d_tc.argtypes = [C.c_void_p, C.POINTER(Conv1dTcDesc), C.c_void_p]
                L.msgm_conv1d_tc_pack_bytes.restype = C.c_int64
                L.msgm_conv1d_tc_pack_bytes.argtypes = [C.c_int32] * 3
                L.msgm_conv1d_tc_pack.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p] * 2
                L.msgm_convt1d_tc_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
                L.msgm_convt1d_tc.argtypes = [C.c_void_p] * 5 + [C.c_int32] * 6 + [C.c_void_p]
                L.msgm_emb_proj_multi.argtypes = [C.c_void_p] * 5 + [C.c_int32] * 4 + [C.POINTER(C.c_int32), C.c_void_p]
                L.msgm_gn_scale_shift.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p] + [C.c_int32] * 4 + \
                    [C.c_void_p] * 4
                L.msgm_gn_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p] * 2
                L.msgm_emb_proj.argtypes = [C.c_void_p] * 5 + [C.c_int32] * 3 + [C.c_void_p]
                L.msgm_sincos_embed_mlp.argtypes = [C.c_void_p] * 7 + [C.c_int32] * 4 + [C.c_void_p]
                L.msgm_attention.argtypes = [C.c_void_p] * 3 + [C.c_int32] * 3 + [C.c_void_p]
                L.msgm_attention_tc.argtypes = [C.c_void_p] * 3 + [C.c_int32] * 3 + [C.c_void_p]
                L.msgm_attention_tc_supported.argtypes = [C.c_int32] * 2
                L.msgm_attention_proj_tc_supported.argtypes = [C.c_int32] * 2
                L.msgm_attention_proj_tc.argtypes = [C.c_void_p] * 6 + [C.c_int32] * 3 + [C.c_void_p]
                L.msgm_vort_pre.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 5 + [C.c_void_p]
                L.msgm_vort_post.argtypes = [C.c_void_p] * 3 + [C.c_int32] * 4 + [C.c_void_p]
                L.msgm_noise_forward.argtypes = [C.c_void_p, C.POINTER(SdeDesc), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64, C.c_void_p]
                L.msgm_ssm_prepare.argtypes = [C.c_void_p, C.POINTER(SdeDesc)] + [C.c_void_p] * 4 + [C.c_int32, C.c_void_p,
                                               C.c_float, C.c_int32, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int64,
                                               C.c_void_p]
                L.msgm_kde_logpdf.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_int32,
                                              C.c_void_p]
                L.msgm_row_norm_stats.argtypes = [C.c_void_p] * 5 + [C.c_int32, C.c_int64, C.c_void_p]
                L.msgm_survival_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                                                   C.c_void_p, C.c_void_p]
                L.msgm_moments.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
                L.msgm_adam_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64] + [C.c_void_p] * 5 + \
                    [C.c_float] * 4 + [C.c_void_p]
                L.msgm_p2p_create.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_char_p]
                L.msgm_p2p_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
                L.msgm_p2p_destroy.argtypes = [C.c_void_p]
                L.msgm_p2p_disconnect.argtypes = [C.c_void_p]
                L.msgm_p2p_allreduce_adam.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64] + \
                    [C.c_void_p] * 5 + [C.c_float] * 3 + [C.c_void_p]
                L.msgm_pair_act.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_int32, C.c_void_p]
                L.msgm_amax.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
                L.msgm_pow2_scale.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
                L.msgm_rows_bias_add.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.c_int32, C.c_int64, C.c_void_p]
                L.msgm_channel_sums.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.c_int32, C.c_int64, C.c_void_p]
                L.msgm_tap_sums_1d.argtypes = [C.c_void_p] * 3 + [C.c_int64] + [C.c_int32] * 6 + [C.c_void_p]
                L.msgm_conv_wgrad.argtypes = [C.c_void_p] * 5 + [C.c_int32] * 15 + [C.c_void_p]
                L.msgm_tc_range_scale.argtypes = [C.c_void_p, C.c_void_p]
                L.msgm_amax2.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
                L.msgm_conv_tc_pack_dgrad.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p, C.c_void_p]
                L.msgm_conv_wgrad_tc_ok.argtypes = [C.c_int32] * 11
                L.msgm_conv_wgrad_tc_scratch_bytes.restype = C.c_uint64
                L.msgm_conv_wgrad_tc_scratch_bytes.argtypes = [C.c_void_p] + [C.c_int32] * 10
                L.msgm_conv_wgrad_tc.argtypes = [C.c_void_p] * 8 + [C.c_int32] * 14 + [C.c_void_p]
                L.msgm_gemm_f32.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 9 + [C.c_void_p]
                L.msgm_premodule_pair.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_int32, C.c_float, C.c_void_p]
                L.msgm_sparse_ssm_loss.argtypes = [C.c_void_p, C.POINTER(SdeDesc)] + [C.c_void_p] * 6 + [C.c_int64, C.c_void_p]
                L.msgm_bgemm_f32.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 6 + [C.c_int64] * 3 + [C.c_int32] * 3 + \
                    [C.c_float, C.c_int32, C.c_void_p]
                L.msgm_gemm_group_f32.argtypes = [C.c_void_p, C.POINTER(GemmProblem)] + [C.c_int32] * 5 + [C.c_void_p]
                L.msgm_gn_pair.argtypes = [C.c_void_p] * 9 + [C.c_int32] * 4 + [C.c_void_p]
                L.msgm_softmax_pair.argtypes = [C.c_void_p] * 7 + [C.c_int64, C.c_int32, C.c_void_p]
                L.msgm_sincos_pair.argtypes = [C.c_void_p] * 3 + [C.c_int32] * 2 + [C.c_void_p]
                L.msgm_resample2.argtypes = [C.c_void_p] * 3 + [C.c_int64] + [C.c_int32] * 3 + [C.c_void_p]
                L.msgm_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int]
                L.msgm_sample_mlp.argtypes = [C.c_void_p, C.POINTER(SdeDesc), C.POINTER(MlpDesc),
                                              C.POINTER(SampleArgs), C.c_void_p, C.c_int64, C.c_void_p]
                L.msgm_mlp_forward.argtypes = [C.c_void_p, C.POINTER(MlpDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_int64, C.c_void_p]
                _lib = L
    return _lib


def check(rc: int):
    """Translate a negative msgm_status into the exception the reference would raise for the same misuse."""
    if rc == 0:
        return
    msg = lib().msgm_last_error().decode()
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def ctx(device) -> C.c_void_p:
    """Per-device library context (created on first use)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"sdeflow_light_b200 runs on CUDA (sm_100a) only; got device '{device}'. "
                           "There is no CPU fallback.")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _ctx:
        h = C.c_void_p()
        check(lib().msgm_create(C.byref(h), idx))
        _ctx[idx] = h
    return _ctx[idx]


# Weight epoch: parameters updated INSIDE a replayed CUDA graph (train.GraphedSsmStep) do not bump ``Tensor._version``,
# so every cache derived from weights (packed tensor-core images, stacked embedding weights, captured inference graphs)
# also keys on this counter, which the trainer bumps after each replay.
_weight_epoch = 0


def weight_epoch() -> int:
    return _weight_epoch


def bump_weight_epoch() -> None:
    global _weight_epoch
    _weight_epoch += 1


def launch_count(device=None) -> int:
    if device is None:
        return sum(int(lib().msgm_launch_count(h)) for h in _ctx.values())
    return int(lib().msgm_launch_count(ctx(device)))


def debug_flags(device) -> int:
    out = C.c_int32(0)
    check(lib().msgm_debug_flags(ctx(device), C.byref(out)))
    return int(out.value)


_ASYNC_MSG = {3: "p2p gradient all-reduce: a peer rank never delivered its gradient (timeout)", 1: "a bounded mbarrier wait inside a tensor-core kernel timed out (the launch gave up; its output is "
                   "undefined)", 2: "tensor-core kernel: shared-memory / TMEM base assumption violated"}


def check_async(device=None) -> None:
    """Raise if a tensor-core kernel that has already run reported an error (no synchronisation: msgm_async_error reads a
    mapped host word).  Called by the shims on entry and after device->host copies, so a timed-out launch surfaces as a
    RuntimeError instead of silently wrong numbers."""
    handles = _ctx.values() if device is None else [ctx(device)]
    for h in handles:
        out = C.c_int32(0)
        check(lib().msgm_async_error(h, C.byref(out)))
        if out.value:
            raise RuntimeError("libmsgm_b200: " + _ASYNC_MSG.get(int(out.value), f"kernel error code {out.value}"))


def debug_counters(device):
    out = (C.c_int64 * 24)()
    check(lib().msgm_debug_counters(ctx(device), out, 24))
    return list(out)


def stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device-contiguous tensor expected"
    return C.c_void_p(t.data_ptr())


def host_float(owner, name: str) -> float:
    """``float(owner.<name>)`` without a device synchronisation per call: the value of a 1-element CUDA tensor (the
    horizon ``T``) is read back once and cached on its owner until the tensor is replaced or written in place."""
    t = getattr(owner, name)
    if not torch.is_tensor(t):
        return float(t)
    if not t.is_cuda:
        return float(t.item())
    key = (id(t), t.data_ptr(), t._version)
    cache = owner.__dict__.get("_hf_" + name)
    if cache is None or cache[0] != key:
        cache = (key, float(t.item()))
        owner.__dict__["_hf_" + name] = cache
    return cache[1]


def f32c(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()
