"""ctypes binding of librsb.so (include/rsb.h).  No torch types cross the boundary:
tensors are passed as raw device pointers + sizes, the stream as a void*.

The product path has NO fallback: if the shared library is missing or a call
fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "librsb.so")

# enums of include/rsb.h
KIND_VANILLA, KIND_QR_MULT, KIND_QR_ADD, KIND_QR_CAT, KIND_PEP, KIND_MASK, KIND_OPTEMBED = range(7)
PEP_GLOBAL, PEP_DIMENSION, PEP_FEATURE, PEP_FEATURE_DIM = range(4)
APPLY_DENSE, APPLY_SPARSE_ADAM, APPLY_SPARSE_SGD, APPLY_SHARD_ATOMIC = range(4)

PEP_TYPES = {"global": PEP_GLOBAL, "dimension": PEP_DIMENSION, "feature": PEP_FEATURE,
             "feature_dim": PEP_FEATURE_DIM}
QR_KINDS = {"mult": KIND_QR_MULT, "add": KIND_QR_ADD, "cat": KIND_QR_CAT}

_p, _i32, _i64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_float

class PlanesFormat(C.Structure):
    """rsb_planes_format of include/rsb.h."""

    _fields_ = [("format", C.c_int32), ("max_scale_exp", C.c_int32), ("amax", C.c_void_p)]


PLANES_BF16X3, PLANES_FP16X2 = 0, 1


class PlanesOperand(C.Structure):
    """rsb_planes_operand of include/rsb.h: an fp32 matrix held as three bf16 planes or two scaled fp16 planes."""

    _fields_ = [("planes", C.c_void_p), ("rows", C.c_int64), ("cols", C.c_int64), ("ld", C.c_int64),
                ("plane_stride", C.c_int64), ("mn_major", C.c_int32), ("batch_row_step", C.c_int64),
                ("batch_col_step", C.c_int64), ("fmt", PlanesFormat)]


class GemmEpilogue(C.Structure):
    """rsb_gemm_epilogue of include/rsb.h."""

    _fields_ = [("mode", C.c_int32), ("out_planes", C.c_void_p), ("out_ld", C.c_int64), ("out_plane_stride", C.c_int64),
                ("ones_col", C.c_int32), ("mask", C.c_void_p), ("p", C.c_float), ("d_amax", C.c_void_p),
                ("bn_partials", C.c_void_p)]


EPI_LINEAR, EPI_RELU_DROPOUT_PLANES, EPI_MASK_PLANES, EPI_MASK_F32 = range(4)

# name -> (restype, argtypes); must list every symbol include/rsb.h declares
PROTOTYPES = {
    "rsb_version": (C.c_char_p, []),
    "rsb_error_string": (C.c_char_p, [C.c_int]),
    "rsb_launch_count": (_i64, []),
    "rsb_row_width_supported": (C.c_int, [_i32]),
    "rsb_lookup_fwd": (C.c_int, [_i32, _p, _i32, _p, _i64, _i32, _i32, _p, _i64, _i64, _p, _i64, _p, _i32, _p,
                                 _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "rsb_lookup_bwd_rows": (C.c_int, [_i32, _p, _i64, _i32, _i32, _p, _i64, _p, _i64, _p, _i32, _p, _p, _p, _p, _p,
                                      _p, _p, _p, _p]),
    "rsb_fc_grad": (C.c_int, [_p, _p, _i64, _i32, _p, _p]),
    "rsb_fc_grad_sorted": (C.c_int, [_p, _p, _i64, _p, _i32, _p, _p, _i64, _p]),
    "rsb_qr_bwd_fused_workspace_bytes": (_i64, [_i64, _i32]),
    "rsb_qr_bwd_fused": (C.c_int, [_i32, _p, _i64, _i32, _i32, _p, _i64, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p,
                                   _i64, _p]),
    "rsb_sort_workspace_bytes": (_i64, [_i64]),
    "rsb_sort_rows": (C.c_int, [_p, _i64, _i64, _i64, _i64, _p, _p, _p, _i64, _p]),
    "rsb_segment_workspace_bytes": (_i64, [_i64, _i32]),
    "rsb_segment_reduce_apply": (C.c_int, [_i32, _p, _p, _i64, _p, _i32, _p, _p, _p, _f, _f, _f, _f, _i64, _p,
                                           _i64, _p]),
    "rsb_small_table_workspace_bytes": (_i64, [_i64, _i32]),
    "rsb_small_table_grad": (C.c_int, [_p, _i64, _i64, _i64, _p, _i32, _i64, _p, _p, _i64, _p]),
    "rsb_pep_threshold_table": (C.c_int, [_p, _p, _i32, _i64, _i32, _p, _p, _p]),
    "rsb_pep_dense_bwd": (C.c_int, [_p, _p, _i32, _i64, _i32, _p, _p, _p, _p]),
    "rsb_optembed_eval_weight": (C.c_int, [_p, _p, _p, _i32, _i64, _i32, _p, _p, _p]),
    "rsb_mask_table": (C.c_int, [_p, _p, _i64, _p, _p]),
    "rsb_sigmoid": (C.c_int, [_p, _i64, _p, _p]),
    "rsb_csr_lookup_fwd": (C.c_int, [_p, _i32, _p, _i64, _i32, _i32, _p, _p, _i32, _p, _i32, _i64, _p, _p, _p, _p,
                                     _p, _p]),
    "rsb_dhe_encode": (C.c_int, [_p, _i32, _i64, _i64, _p, _p, _p, _i32, _i64, _i32, _p, _p]),
    "rsb_split_planes": (C.c_int, [_p, _i64, _i64, _i64, _i32, _i32, _p, _i64, _i64, C.POINTER(PlanesFormat), _p]),
    "rsb_absmax": (C.c_int, [_p, _i64, _i64, _i64, _p, _p]),
    "rsb_rank1_absmax": (C.c_int, [_p, _i64, _p, _i64, _f, _p, _p]),
    "rsb_relu_dropout_planes": (C.c_int, [_p, _i64, _i32, _i64, _f, C.c_uint64, C.c_uint64, _p, _i32, _p, _i64, _i64, _p, _p]),
    "rsb_dropout_keep_mask": (C.c_int, [_p, _i64, _f, C.c_uint64, C.c_uint64, _p, _p]),
    "rsb_rank1_mask_planes": (C.c_int, [_p, _p, _p, _i64, _i32, _f, _p, _i64, _i64, _p]),
    "rsb_gemm_planes_workspace_bytes": (_i64, [_i64, _i64, _i64, _i64, _i32]),
    "rsb_gemm_planes": (C.c_int, [C.POINTER(PlanesOperand), C.POINTER(PlanesOperand), _i64, _i64, _i64, _i64, _i32, _p,
                                  _p, _i64, _i64, _p, _f, _f, C.POINTER(GemmEpilogue), _p, _i64, _p]),
    "rsb_relu_dropout_fwd": (C.c_int, [_p, _i64, _f, C.c_uint64, C.c_uint64, _p, _p, _p, _p]),
    "rsb_relu_dropout_bwd": (C.c_int, [_p, _p, _i64, _i32, _f, _p, _p, _p, _i64, _p]),
    "rsb_colsum_workspace_bytes": (_i64, [_i64, _i32]),
    "rsb_colsum": (C.c_int, [_p, _i64, _i32, _i64, _p, _p, _i64, _p]),
    "rsb_relu_dropout_dot_fwd": (C.c_int, [_p, _i64, _i32, _f, C.c_uint64, C.c_uint64, _p, _p, _p, _p, _p, _p, _p, _p]),
    "rsb_bn_workspace_bytes": (_i64, [_i64, _i32]),
    "rsb_gemm_bn_partials_bytes": (_i64, [_i64, _i64, _i64, _i32]),
    "rsb_bn_finalize_partials": (C.c_int, [_p, _i64, _i32, _p, _p, _f, _f, _p, _p, _p, _p, _f, _p, _p, _i64, _p]),
    "rsb_bn_train_fwd_stats": (C.c_int, [_p, _i64, _i32, _i64, _p, _p, _f, _f, _p, _p, _p, _p, _f, _p, _p, _i64, _p]),
    "rsb_bn_relu_dropout_planes": (C.c_int, [_p, _i64, _i32, _i64, _p, _f, C.c_uint64, C.c_uint64, _p, _i32, _p, _i64, _i64,
                                             _p, C.POINTER(PlanesFormat), _p]),
    "rsb_bn_train_bwd_planes": (C.c_int, [_p, _p, _i64, _i32, _i64, _i64, _p, _p, _p, _p, _i64, _i64,
                                          C.POINTER(PlanesFormat), _p, _p, _i64, _p]),
    "rsb_relu_dropout_bwd_rank1": (C.c_int, [_p, _p, _p, _i64, _i32, _f, _p, _p, _p, _i64, _p]),
    "rsb_colsum_weighted": (C.c_int, [_p, _p, _i64, _i32, _i64, _p, _p, _i64, _p]),
    "rsb_dcn_gate_mix_fwd": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _p, _p, _p, _p]),
    "rsb_dcn_cross_out_fwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _i32, _p, _p]),
    "rsb_dcn_cross_out_bwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _i32, _p, _p, _p, _p]),
    "rsb_dcn_gate_mix_bwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _p, _p, _p, _p]),
    "rsb_shared_alloc": (C.c_int, [_i64, C.POINTER(C.c_void_p)]),
    "rsb_shared_free": (C.c_int, [_p]),
    "rsb_ipc_get_handle": (C.c_int, [_p, C.c_char_p]),
    "rsb_ipc_open_handle": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "rsb_ipc_close_handle": (C.c_int, [_p]),
    "rsb_lookup_fwd_sharded": (C.c_int, [_p, _i32, _p, _i64, _i32, _i32, _p, _p, _i32, _i64, _p, _p, _p, _p, _p, _p,
                                         _p, _p, _p, _p]),
    "rsb_records_unpack": (C.c_int, [_p, _i64, _i32, _p, _p, _p]),
    "rsb_adam_dense": (C.c_int, [_p, _i32, _p, _i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _i64, _p]),
    "rsb_lookup_fwd_sharded_kind": (C.c_int, [_i32, _p, _i32, _p, _i64, _i32, _i32, _p, _i32, _i64, _i64, _p, _i64, _p, _p,
                                              _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "rsb_lookup_bwd_rows_sharded": (C.c_int, [_i32, _p, _i64, _i32, _i32, _p, _i32, _i64, _p, _i64, _p, _p, _i32, _p, _p,
                                              _p, _p, _p, _p, _p, _p]),
    "rsb_segment_scatter_shards": (C.c_int, [_p, _p, _i64, _p, _i32, _p, _i32, _f, _p, _i32, _p, _p, _i64, _p]),
}
LOOKUP_AMAX_SLOTS = 1024      # RSB_LOOKUP_AMAX_SLOTS of include/rsb.h
ADAM_CHUNK = 4096             # RSB_ADAM_CHUNK
ADAM_MAX_TENSORS = 64         # RSB_ADAM_MAX_TENSORS


class AdamTensor(C.Structure):
    """rsb_adam_tensor of include/rsb.h."""

    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64)]

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load librsb.so (built in-tree by build.py).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"rsb: {LIB_PATH} not found. Build it with `python recsys-benchmark_b200/build.py` "
            "(or __graft_entry__.build()). There is no CPU / PyTorch fallback for the hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def version() -> str:
    return load().rsb_version().decode()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().rsb_error_string(int(rc)).decode()
        raise RuntimeError(f"rsb {what} failed (code {rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr(device=None) -> int:
    """cudaStream_t of the current stream of `device` as an integer.  Called once per kernel launch: the raw-stream query
    (one C call) instead of building a torch.cuda.Stream object each time (~5 us of the ~20 us a launch costs the host)."""
    if _raw_stream is not None and type(device) is torch.device and device.index is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("rsb: the hot path runs on CUDA tensors only (no CPU fallback); "
                               "move the module and its inputs to a B200 device")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"rsb: tensors on different devices ({dev} vs {t.device})")
    return dev
