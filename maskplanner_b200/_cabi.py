"""ctypes binding of libmaskplanner_b200.so (the C ABI declared in include/maskplanner_b200.h).

There is NO fallback: if the library is missing (and cannot be built because nvcc is absent) or a
call fails, this module raises.  Tensors are passed as raw device pointers plus sizes/strides; the
current torch CUDA stream is passed as the stream argument, so work is enqueued exactly where
stock torch ops would enqueue it.
"""
import ctypes
import os

import torch

from . import build as _build

_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_int64
_F = ctypes.c_float
_D = ctypes.c_double

# name -> (restype, argtypes); must list every symbol of include/maskplanner_b200.h
SIGNATURES = {
    "mpb_version": (_I, []),
    "mpb_last_error_string": (ctypes.c_char_p, []),
    "mpb_device_sm_count": (_I, []),
    "mpb_device_arch": (_I, []),
    "mpb_peak_fp32_threads": (_L, [_I]),
    "mpb_peak_fp32_ffma": (_I, [_I, _I, _I, _P, _P]),
    "mpb_fps_workspace_bytes": (_L, [_I, _I]),
    "mpb_fps_f32": (_I, [_P, _L, _L, _L, _I, _I, _P, _I, _P, _P, _P]),
    "mpb_square_distance_f32": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "mpb_ball_query_f32": (_I, [_P, _L, _L, _L, _P, _L, _L, _L, _I, _I, _I, _F, _I, _P, _P]),
    "mpb_knn_group_f32": (_I, [_P, _L, _L, _L, _P, _L, _L, _L, _I, _I, _I, _I, _P, _P, _P]),
    "mpb_index_points_f32": (_I, [_P, _L, _L, _L, _I, _I, _I, _P, _L, _P, _P]),
    "mpb_index_points_bwd_f32": (_I, [_P, _P, _I, _I, _I, _L, _P, _P]),
    "mpb_group_points_f32": (_I, [_P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "mpb_group_points_bwd_f32": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "mpb_chamfer_nn_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "mpb_chamfer_nn_bwd_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mpb_knn_points_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _I, _P, _P, _P]),
    "mpb_knn_points_bwd_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _P]),
    "mpb_padded_lengths_f32": (_I, [_P, _I, _I, _I, _F, _P, _P, _P]),
    "mpb_group_points_bf16": (_I, [_P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "mpb_group_points_bwd_bf16": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _P, _P]),
    "mpb_sa_gemm_stat_partials": (_I, [_I, _I, _I, _I, _I, _I]),
    "mpb_sa_gemm_tn": (_I, [_I, _P, _P, _P, _P, _I, _I, _I, _P, _P, _I, _P, _I, _P, _P, _P, _P]),
    "mpb_sa_gemm_wgrad_workspace": (_L, [_I, _I, _I, _I, _I]),
    "mpb_sa_gemm_wgrad": (_I, [_I, _P, _P, _I, _I, _I, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "mpb_sa_gemm_tn_pool": (_I, [_I, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P, _I, _P, _P, _P, _P]),
    "mpb_sa_gemm_wgrad_pool": (_I, [_I, _P, _P, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "mpb_sa_gemm_wgrad_reduce": (_I, [_I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _P, _P]),
    "mpb_rng_advance": (_I, [_P, _P]),
    "mpb_head_act_fwd": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _F, _F, _I, _F, ctypes.c_uint64, _P, _I, _P, _P, _P, _P, _P, _P, _P]),
    "mpb_head_act_bwd": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _I, _F, ctypes.c_uint64, _P, _I, _P, _P, _P, _P, _P]),
    "mpb_head_to_feature_major": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "mpb_head_to_batch_major": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "mpb_head_pose_out_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _P, _P]),
    "mpb_head_pose_out_bwd": (_I, [_P, _P, _P, _I, _I, _I, _F, _P, _P, _P, _P, _P]),
    "mpb_bn_stat_partials": (_I, [_L, _I]),
    "mpb_bn_colstats": (_I, [_I, _P, _L, _I, _P, _I, _P]),
    "mpb_bn_finalize_f32": (_I, [_P, _I, _I, _I, _L, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P]),
    "mpb_bn_relu": (_I, [_I, _P, _P, _P, _L, _I, _P, _P]),
    "mpb_bn_relu_max": (_I, [_I, _P, _P, _P, _L, _I, _I, _P, _P, _P, _P]),
    "mpb_bn_bwd_stats": (_I, [_I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _L, _I, _P, _I, _P, _P]),
    "mpb_bn_bwd_finalize_f32": (_I, [_P, _I, _I, _I, _L, _P, _P, _P, _P, _P, _P, _P, _L, _P, _P]),
    "mpb_pack_weight_bf16": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "mpb_pack_weight_tf32": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "mpb_bn_bwd_apply": (_I, [_I, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _L, _I, _P, _P]),
    "mpb_bn_bwd_apply_pooled": (_I, [_I, _P, _P, _I, _P, _P, _L, _I, _P, _P]),
    "mpb_sa_first_layer": (_I, [_I, _P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _I, _I, _I, _I, _I, _P, _I, _I, _P, _P, _I, _P]),
    "mpb_sa_first_layer_bwd": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _I, _I, _I, _I, _I, _I,
                                   _P, _I, _P]),
    "mpb_adam_step_f32": (_I, [_I, _P, _P, _P, _P, _P, _F, _P, _D, _D, _D, _D, _F, _P, _P, _P]),
    "mpb_lap_f32": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "mpb_debug_mbar_state": (_I, [_P, _I]),
    "mpb_stage_batch": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mpb_loss_lengths_f32": (_I, [_P, _I, _I, _P, _I, _I, _I, _F, _P, _P, _P]),
    "mpb_mask_cost_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "mpb_asymm_v6_loss_value_f32": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _I, _I, _I, _I, _I, _P, _P, _P]),
    "mpb_asymm_v6_loss_bwd_f32": (_I, [_P] * 16 + [_F, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
}

_lib = None


def library_path():
    return _build.LIB


def load():
    """Load (building first if the .so is absent and nvcc exists).  Raises on any failure."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if _build.is_stale():   # missing, or built from other sources than the ones in csrc/ now (hash, not mtime)
        try:
            _build.build()
        except Exception as e:  # no silent degradation: surface why the native path is unavailable
            raise ImportError(
                "libmaskplanner_b200.so is missing or stale and could not be built (%s). maskplanner_b200 has no "
                "CPU or pure-torch fallback; run `python -m maskplanner_b200.build`." % (e,)) from e
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class MpbError(RuntimeError):
    pass


KERNEL_LAUNCHES = 0  # kernels of libmaskplanner_b200.so enqueued so far (bench.py reports the per-step delta)


def check(rc, what, launches=1):
    """Raise on a failed C-ABI call; otherwise account for the kernels it enqueued."""
    global KERNEL_LAUNCHES
    if rc != 0:
        msg = load().mpb_last_error_string()
        raise MpbError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
    KERNEL_LAUNCHES += launches


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("maskplanner_b200 runs on CUDA tensors only (sm_100a kernels, no CPU fallback); "
                               "got a %s tensor" % t.device)
