"""ctypes binding of libqmc_b200.so (C ABI: include/qmc_b200.h).

There is no fallback: if the shared object is missing or a symbol is absent, importing this
module raises.  Build with ``python quantized_spectrum_cartography_b200/build.py``.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libqmc_b200.so")

QMC_MAX_BOUNDS = 257
QMC_MAX_RANK = 32
QMC_LOG_DOMAIN = 1 << 0
QMC_EPI_REFERENCE = 1 << 1
QMC_FORWARD_ONLY = 1 << 2
QMC_SKIP_GS = 1 << 3
QMC_SKIP_GC = 1 << 4
QMC_EPI_LSQ = 1 << 5
QMC_EPI_LOGISTIC = 1 << 6
QMC_ALGO_AUTO, QMC_ALGO_FLAT, QMC_ALGO_TILED, QMC_ALGO_LANES = 0, 1, 2, 3


class Likelihood(C.Structure):
    """qmc_likelihood_t"""
    _fields_ = [("n_bounds", C.c_int32), ("flags", C.c_uint32), ("noise_std", C.c_float),
                ("offset", C.c_float), ("bounds", C.c_float * QMC_MAX_BOUNDS)]


class ObsView(C.Structure):
    """qmc_obs_view_t"""
    _fields_ = [("idx_dev", C.c_void_p), ("lvl_dev", C.c_void_p), ("row_off_dev", C.c_void_p),
                ("n_sub", C.c_int32), ("sub_pixels", C.c_int32),
                ("words_dev", C.c_void_p), ("stream_off_dev", C.c_void_p), ("nrows_dev", C.c_void_p),
                ("stream_stride", C.c_int64), ("n_runs", C.c_int32), ("word_bits", C.c_int32),
                ("lvl_bits", C.c_int32), ("has_cont", C.c_int32), ("map_modulo", C.c_int32)]


QMC_PEER_MAX_WORLD = 8


class PeerExchange(C.Structure):
    """qmc_peer_exchange_t"""
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("slot_floats", C.c_int32), ("reserved", C.c_int32),
                ("region", C.c_void_p * QMC_PEER_MAX_WORLD)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float

# every symbol include/qmc_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "qmc_abi_version": (_I, []),
    "qmc_last_error": (C.c_char_p, []),
    "qmc_launch_count": (_L, []),
    "qmc_noisy_signal": (_I, [_P, _P, _F, _F, _I, _L, _P, _P]),
    "qmc_quantize_levels": (_I, [_P, _L, C.POINTER(C.c_float), _I, _P, _P, _P]),
    "qmc_obs_scan_ws_elems": (_L, [_L]),
    "qmc_obs_count_scan": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "qmc_obs_fill": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "qmc_lanes_stream_words": (_L, [_L, _I, _I, _I, _I]),
    "qmc_obs_build_lanes": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _L, _I, _I, _L, _P, _P, _P, _I, _I, _I, _P]),
    "qmc_lanes_smem_bytes": (_L, [_I, _I, _I, _I, _I, _I]),
    "qmc_sumsq_per_map": (_I, [_P, _I, _L, _P, _P]),
    "qmc_adam_frob_project": (_I, [_P, _P, _P, _P, _I, _L, _P, _P, _F, _F, _F, _F, _F, _I, _I, _P, _P]),
    "qmc_counter_add": (_I, [_P, _I, _P]),
    "qmc_solver_s_step_fused": (_I, [_P, _L, _L, _L, _P, C.POINTER(ObsView), C.POINTER(Likelihood), _I, _I, _I, _I, _I,
                                     _P, _P, _P, _P, _P, _F, _F, _F, _F, _F, _I, _I, _P, _P]),
    "qmc_nll_fwd_bwd_gather": (_I, [_P, _L, _L, _L, _P, C.POINTER(ObsView), C.POINTER(Likelihood),
                                    _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "qmc_tiled_smem_bytes": (_L, [_I, _I, _I, _I]),
    "qmc_nll_fwd_bwd_gather_host": (_I, [_P, _P, _P, _P, C.POINTER(ObsView), C.POINTER(Likelihood),
                                         _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "qmc_dense_pack": (_I, [_P, _I, _P, _I, _I, _I, _P, _P]),
    "qmc_nll_fwd_bwd_dense": (_I, [_P, _P, _P, C.POINTER(Likelihood), _I, _I, _I, _P, _P, _L, _P, _P]),
    "qmc_dense_smem_bytes": (_L, [_I, _I]),
    "qmc_peer_region_bytes": (_L, [_I, _I]),
    "qmc_peer_alloc": (_I, [_L, C.POINTER(C.c_void_p), _P]),
    "qmc_peer_open": (_I, [_P, C.POINTER(C.c_void_p)]),
    "qmc_peer_close": (_I, [_P]),
    "qmc_peer_free": (_I, [_P]),
    "qmc_peer_status": (_I, [_P, C.POINTER(C.c_int), _P]),
    "qmc_nll_fwd_bwd_dense_exchange": (_I, [_P, _P, _P, C.POINTER(Likelihood), _I, _I, _I, _P, _P, _L, _P,
                                            C.POINTER(PeerExchange), _P]),
    "qmc_get_tensor": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "qmc_nmse_terms": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _P]),
    "qmc_bce_one_bit": (_I, [_P, _P, _L, _F, _F, _I, _P, _P, _P]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first "
            "(python quantized_spectrum_cartography_b200/build.py).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class QmcError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        raise QmcError(f"libqmc_b200 error {rc}: {lib.qmc_last_error().decode()}")


def launch_count() -> int:
    return int(lib.qmc_launch_count())
