"""CPU: the C-ABI library builds, loads, and exports every symbol include/qmc_b200.h declares.
No compute calls here (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "qmc_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"^QMC_API\s+[\w\s\*]+?\b(qmc_\w+)\s*\(", text, flags=re.M)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ("qmc_quantize_levels", "qmc_obs_count_scan", "qmc_obs_fill", "qmc_nll_fwd_bwd_gather",
                 "qmc_nll_fwd_bwd_gather_host", "qmc_nll_fwd_bwd_dense", "qmc_dense_pack", "qmc_abi_version",
                 "qmc_last_error", "qmc_obs_build_lanes", "qmc_lanes_smem_bytes", "qmc_adam_frob_project",
                 "qmc_sumsq_per_map", "qmc_counter_add"):
        assert must in syms, must


def test_library_loads_and_exports_every_declared_symbol():
    import __graft_entry__
    path = __graft_entry__._load_builder().build()
    lib = ctypes.CDLL(path)
    for sym in declared_symbols():
        assert hasattr(lib, sym), f"{sym} declared in include/qmc_b200.h but not exported by {path}"
    lib.qmc_abi_version.restype = ctypes.c_int
    assert lib.qmc_abi_version() == int(re.search(r"#define QMC_ABI_VERSION (\d+)", open(HEADER).read()).group(1))


def test_python_binding_covers_every_declared_symbol():
    from quantized_spectrum_cartography_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_python_flag_constants_mirror_the_header():
    """Every flag of qmc_likelihood_t has the same value in _lib.py as in include/qmc_b200.h."""
    from quantized_spectrum_cartography_b200 import _lib
    text = open(HEADER).read()
    flags = dict(re.findall(r"\b(QMC_(?:LOG_DOMAIN|EPI_\w+|FORWARD_ONLY|SKIP_\w+))\s*=\s*1u\s*<<\s*(\d+)", text))
    assert {"QMC_LOG_DOMAIN", "QMC_EPI_REFERENCE", "QMC_FORWARD_ONLY", "QMC_SKIP_GS", "QMC_SKIP_GC", "QMC_EPI_LSQ", "QMC_EPI_LOGISTIC"} <= set(flags)
    for name, shift in flags.items():
        assert getattr(_lib, name) == 1 << int(shift), name
    assert len(set(flags.values())) == len(flags)


def test_least_squares_likelihood_packing():
    """make_likelihood(least_squares=True): flag set, table untouched (no sentinels: the reference's
    mid-point function leaves them commented out, quantization_model_log.py:45-46), noise_std optional."""
    import torch
    from quantized_spectrum_cartography_b200 import _lib, make_likelihood
    bb = torch.tensor([0.0, 0.25, 0.5, 1.0])
    lik = make_likelihood(bb, None, least_squares=True)
    assert lik.flags & _lib.QMC_EPI_LSQ and not lik.flags & _lib.QMC_LOG_DOMAIN
    assert [lik.bounds[i] for i in range(4)] == [0.0, 0.25, 0.5, 1.0] and lik.noise_std > 0
    lik = make_likelihood(bb, 0.1, offset=1e-3, least_squares=True)
    assert lik.flags & _lib.QMC_EPI_LSQ and lik.flags & _lib.QMC_LOG_DOMAIN and lik.bounds[0] == 0.0
    plain = make_likelihood(bb, 0.1)
    assert not plain.flags & _lib.QMC_EPI_LSQ and plain.bounds[0] == -100000.0


def test_argument_validation_needs_no_gpu():
    """Invalid arguments are rejected before any CUDA call, with a message."""
    from quantized_spectrum_cartography_b200 import _lib
    rc = _lib.lib.qmc_quantize_levels(None, 10, None, 3, None, None, None)
    assert rc == 1 and b"bad arguments" in _lib.lib.qmc_last_error()
    with pytest.raises(_lib.QmcError):
        _lib.check(rc)
    # cfg1/cfg3 tile: S + gS tiles, C, 8 private gC copies, 8 x 32 scratch rows, 8 x (K+2) row offsets
    assert _lib.lib.qmc_tiled_smem_bytes(64, 4, 326, 8) == (2 * 326 * 8 * 4 + 9 * 64 * 4 + 8 * 32 * 4) * 4 + 8 * 66 * 4 + 16
    assert _lib.lib.qmc_tiled_smem_bytes(64, 4, 100000, 8) == 0


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "quantized_spectrum_cartography_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_plan_tiles_covers_the_map():
    from quantized_spectrum_cartography_b200.obs import plan_tiles
    for IJ, K, R in ((2601, 64, 4), (10201, 128, 8), (262144, 256, 16), (7, 3, 1), (2601, 64, 3)):
        n_sub, sub, tw = plan_tiles(IJ, K, R)
        assert n_sub % tw == 0 and n_sub * sub >= IJ
        assert (n_sub - tw) * sub < IJ or n_sub == tw      # no entirely empty trailing tile
    assert plan_tiles(2601, 64, 4) == (8, 326, 8)          # cfg1/cfg3: one CTA per map
    assert plan_tiles(2601, 64, 4, lanes=True, max_level=1) == (8, 326, 8)   # 16-bit stream words: small ring
    assert plan_tiles(2601, 64, 4, lanes=True) == (16, 163, 8)               # 32-bit words: two tiles per map


def test_lanes_tile_of_cfg3_leaves_room_for_two_ctas_per_sm():
    """Lane-stream kernel at cfg1/cfg3: stream ring + S and gS tiles + C + 8 private gC copies; two CTAs
    (plus 1 KB reserved each) must fit the 228 KB of an SM."""
    from quantized_spectrum_cartography_b200 import _lib
    b = _lib.lib.qmc_lanes_smem_bytes(64, 4, 326, 8, 5, 16)
    # S and gS tiles | C (K+1 rows) | 8 gC copies of K+1+32 rows | run tables (5 entries per lane) | rings (2 slots)
    assert b == (2 * 326 * 8 * 4 + 65 * 4 + 8 * (65 + 32) * 4 + 8 * 5 * 32 + 8 * 2 * 128) * 4 + 16
    assert 2 * (b + 1024 + 512) <= 228 * 1024
    assert _lib.lib.qmc_lanes_smem_bytes(64, 4, 100000, 8, 5, 16) == 0


def test_exchange_region_geometry_and_descriptor_mirror():
    """The fused exchange's host-side helpers need no GPU: region size = 256-byte header + scratch slot + two sets of
    `world` slots; the ctypes mirror of qmc_peer_exchange_t has the header's layout; bad geometries give 0."""
    import ctypes as C

    from quantized_spectrum_cartography_b200 import _lib
    rb = _lib.lib.qmc_peer_region_bytes
    assert rb(8, 4100) == 256 + (1 + 2 * 8) * 4100 * 4
    assert rb(1, 4) == 256 + 3 * 16
    assert rb(0, 4100) == 0 and rb(9, 4100) == 0 and rb(2, 4098) == 0 and rb(2, 0) == 0
    px = _lib.PeerExchange
    assert _lib.QMC_PEER_MAX_WORLD == 8
    assert (px.rank.offset, px.world.offset, px.slot_floats.offset, px.region.offset) == (0, 4, 8, 16)
    assert C.sizeof(px) == 16 + 8 * C.sizeof(C.c_void_p)
    hdr = open(os.path.join(ROOT, "include", "qmc_b200.h")).read()
    assert "#define QMC_PEER_MAX_WORLD 8" in hdr
    # the dense smem budget the kernel asks for at cfg4 (K = 256, R = 16) fits the 227 KB a CTA may have
    assert 0 < _lib.lib.qmc_dense_smem_bytes(256, 16) <= 227 * 1024 - 4096
    assert _lib.lib.qmc_dense_smem_bytes(288, 16) == 0 and _lib.lib.qmc_dense_smem_bytes(256, 17) == 0
