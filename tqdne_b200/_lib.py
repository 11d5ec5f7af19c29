"""ctypes binding of libtqdne_b200.so (the C-ABI declared in include/tqdne_b200.h).

There is no fallback: if the shared library is missing or an entry point fails, a ``RuntimeError``
is raised.  The library itself is built by ``__graft_entry__.build()`` / ``make -C tqdne_b200/csrc``.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

TQ_BF16, TQ_F32, TQ_F64 = 0, 1, 2
ABI_VERSION = 20  # TQ_ABI_VERSION of include/tqdne_b200.h

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("TQDNE_B200_LIB", _HERE / "libtqdne_b200.so"))


class TqSrc(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
        ("sn", C.c_int64), ("sy", C.c_int64), ("sx", C.c_int64),
    ]


class TqSlice(C.Structure):
    _fields_ = [
        ("src", C.c_int16), ("dx", C.c_int16), ("dy", C.c_int16), ("rsv", C.c_int16),
        ("c0", C.c_int32), ("kb", C.c_int32),
    ]


class TqConvDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("cout", C.c_int32), ("cout_pad", C.c_int32), ("ktot", C.c_int32),
        ("num_srcs", C.c_int32), ("num_classes", C.c_int32), ("num_slices", C.c_int32),
        ("srcs", TqSrc * 4),
        ("slices", C.POINTER(TqSlice)),
        ("weights", C.c_void_p),
        ("bias", C.c_void_p),
        ("emb", C.c_void_p),
        ("emb_ld", C.c_int32),
        ("residual", C.c_void_p),
        ("out", C.c_void_p),
        ("out_dtype", C.c_int32),
        ("out_sn", C.c_int64), ("out_sy", C.c_int64), ("out_sx", C.c_int64),
        ("out_class_off", C.c_int64 * 4),
        ("block_n", C.c_int32),
        ("stats", C.c_void_p),
        ("cta_group", C.c_int32),
        ("stats_parts", C.c_int32),
    ]


class TqGnDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("N", C.c_int32), ("P", C.c_int32), ("C0", C.c_int32), ("C1", C.c_int32),
        ("x0", C.c_void_p), ("x1", C.c_void_p),
        ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("eps", C.c_float),
        ("silu", C.c_int32),
        ("y", C.c_void_p),
        ("ws", C.c_void_p),
        ("stats0", C.c_void_p), ("stats1", C.c_void_p),
        ("parts0", C.c_int32), ("parts1", C.c_int32),
        ("drop_seed", C.c_void_p), ("drop_p", C.c_float), ("drop_site", C.c_int32),
        ("film", C.c_void_p), ("film_ld", C.c_int32),
    ]


class TqAttnDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("N", C.c_int32), ("T", C.c_int32), ("heads", C.c_int32), ("d", C.c_int32),
        ("qkv", C.c_void_p), ("out", C.c_void_p),
        ("causal", C.c_int32),
        ("lse", C.c_void_p),
    ]


class TqLinearDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("K", C.c_int32), ("Nout", C.c_int32), ("x_rows", C.c_int32),
        ("x", C.c_void_p), ("W", C.c_void_p), ("b", C.c_void_p),
        ("act_in", C.c_int32),
        ("add", C.c_void_p), ("add_rows", C.c_int32),
        ("y", C.c_void_p),
        ("y_act", C.c_void_p), ("y_act_dtype", C.c_int32),
    ]


class TqGnBwdDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("N", C.c_int32), ("P", C.c_int32), ("C0", C.c_int32), ("C1", C.c_int32),
                ("x0", C.c_void_p), ("x1", C.c_void_p), ("dy", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("eps", C.c_float), ("silu", C.c_int32), ("stats0", C.c_void_p), ("stats1", C.c_void_p), ("parts0", C.c_int32), ("parts1", C.c_int32),
                ("ws", C.c_void_p), ("dx0", C.c_void_p), ("dx1", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
                ("dx_add0", C.c_void_p), ("dx_add1", C.c_void_p), ("dx_sum", C.c_void_p), ("dx_sum_ld", C.c_int32),
                ("drop_seed", C.c_void_p), ("drop_p", C.c_float), ("drop_site", C.c_int32),
                ("dbias0", C.c_void_p), ("dbias1", C.c_void_p)]


class TqRepackJob(C.Structure):
    _fields_ = [("master", C.c_void_p), ("fwd", C.c_void_p), ("bwd", C.c_void_p),
                ("Op", C.c_int32), ("k", C.c_int32), ("Ip", C.c_int32), ("ci_off", C.c_int32), ("Cs", C.c_int32),
                ("block0", C.c_int32), ("nblocks", C.c_int32), ("pad_", C.c_int32)]


# name -> (restype, argtypes): every symbol include/tqdne_b200.h declares
_VP, _I32, _I64, _F, _D = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
SIGNATURES = {
    "tq_abi_version": (C.c_int, []),
    "tq_last_error": (C.c_char_p, []),
    "tq_launch_count": (_I64, []),
    "tq_launch_count_reset": (None, []),
    "tq_plan_create": (_VP, []),
    "tq_plan_destroy": (None, [_VP]),
    "tq_plan_num_ops": (C.c_int, [_VP]),
    "tq_plan_run": (C.c_int, [_VP, _VP]),
    "tq_plan_run_range": (C.c_int, [_VP, C.c_int, C.c_int, _VP]),
    "tq_plan_enable_graph": (C.c_int, [_VP, C.c_int]),
    "tq_plan_op_name": (C.c_char_p, [_VP, C.c_int]),
    "tq_plan_add_memset": (C.c_int, [_VP, _VP, _I64, _I32]),
    "tq_plan_add_conv": (C.c_int, [_VP, C.POINTER(TqConvDesc)]),
    "tq_conv_stats_parts": (_I32, [C.POINTER(TqConvDesc)]),
    "tq_plan_add_groupnorm": (C.c_int, [_VP, C.POINTER(TqGnDesc)]),
    "tq_groupnorm_ws_floats": (_I64, [C.POINTER(TqGnDesc)]),
    "tq_plan_add_attention": (C.c_int, [_VP, C.POINTER(TqAttnDesc)]),
    "tq_attention_writes_lse": (_I32, [C.POINTER(TqAttnDesc)]),
    "tq_plan_add_linear": (C.c_int, [_VP, C.POINTER(TqLinearDesc)]),
    "tq_plan_add_fourier": (C.c_int, [_VP, _VP, _VP, _I32, _I32, _VP]),
    "tq_plan_add_resample2": (C.c_int, [_VP, _I32, _VP, _VP, _I32, _I32, _I32, _I32, _I32]),
    "tq_plan_add_spatial_mean": (C.c_int, [_VP, _VP, _I32, _I32, _I32, _I32, _VP]),
    "tq_gn_silu_backward": (C.c_int, [C.POINTER(TqGnBwdDesc), _VP]),
    "tq_attention_backward": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _VP]),
    "tq_sample_channel_sums": (C.c_int, [_VP, _VP, _I32, _I32, _I64, _I32, _VP]),
    "tq_conv1d_wgrad": (C.c_int, [_VP, _VP, _VP, _VP, _I32, _I64, _I32, _I32, _I32, _I32, _I32, _VP]),
    "tq_conv2d_wgrad": (C.c_int, [_VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _VP]),
    "tq_rows_op": (C.c_int, [_VP, _VP, _VP, _I32, _I64, _I64, _I32, _VP]),
    "tq_linear_backward": (C.c_int, [_VP, _VP, _VP, _I32, _VP, _VP, _VP, _I32, _I32, _I32, _VP]),
    "tq_edm_noise": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _I64, _I64, _I32, _I32, _F, _VP]),
    "tq_edm_loss": (C.c_int, [_VP, _I32, _VP, _VP, _VP, _VP, _VP, _I64, _I64, _I32, _I32, _F, _VP]),
    "tq_dropout_mask": (C.c_int, [_VP, _I64, C.c_uint64, _F, _VP]),
    "tq_repack_conv_weights": (C.c_int, [_VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _VP]),
    "tq_repack_batch_prepare": (_I64, [C.POINTER(TqRepackJob), _I32]),
    "tq_repack_batch_run": (C.c_int, [_VP, _I32, _I64, _VP]),
    "tq_dropout_apply": (C.c_int, [_VP, _VP, _I64, C.c_uint64, _F, _VP]),
    "tq_adam_ema_step": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _I64, _F, _F, _F, _F, _I64, _F, _F, _VP]),
    "tq_edm_precondition": (C.c_int, [_VP, _VP, _I32, _I64, _I32, _I32, _F, _VP, _F, _VP]),
    "tq_edm_euler": (C.c_int, [_VP, _VP, _I32, _VP, _VP, _VP, _I32, _I64, _I32, _I32, _F, _F, _F, _F, _F, _I32, _VP, _F, _VP]),
    "tq_edm_heun": (C.c_int, [_VP, _VP, _VP, _VP, _I32, _VP, _I32, _I64, _I32, _I32, _F, _F, _F, _F, _F, _I32, _VP, _F, _VP]),
    "tq_edm_add_noise": (C.c_int, [_VP, _VP, _D, _I64, _VP]),
    "tq_nchw_to_nhwc": (C.c_int, [_VP, _I32, _VP, _I32, _I32, _I32, _I64, _I32, _VP]),
    "tq_nhwc_to_nchw": (C.c_int, [_VP, _I32, _I32, _VP, _I32, _I32, _I32, _I64, _VP]),
    "tq_griffinlim_ws_bytes": (_I64, [_I32, _I32, _I32, _I32]),
    "tq_logspec_griffinlim": (C.c_int, [_VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _D, _D, _D, _I32, _VP, _VP]),
    "tq_mavg_envelope_inverse": (C.c_int, [_VP, _VP, _I32, _I32, _I64, _D, _D, _VP]),
    "tq_logspec_forward": (C.c_int, [_VP, _VP, _I32, _I32, _I32, _I32, _I64, _I32, _D, _D, _I32, _VP]),
    "tq_mavg_envelope_forward": (C.c_int, [_VP, _VP, _I32, _I32, _I64, _I32, _D, _D, _VP]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the shared library once; fail loudly if it is missing or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"tqdne_b200: native library {LIB_PATH} not found -- run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C tqdne_b200/csrc`. There is no CPU / PyTorch fallback."
        )
    handle = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError as e:  # pragma: no cover
            raise RuntimeError(f"tqdne_b200: {LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if handle.tq_abi_version() != ABI_VERSION:
        raise RuntimeError("tqdne_b200: ABI version mismatch between _lib.py and libtqdne_b200.so")
    _lib = handle
    return handle


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().tq_last_error().decode(errors="replace")
        raise RuntimeError(f"tqdne_b200 {what} failed: {msg}")


def launch_count() -> int:
    return int(lib().tq_launch_count())


def launch_count_reset() -> None:
    lib().tq_launch_count_reset()
