"""Compact observation sets: the (Y, Wx) information the likelihood actually uses.

The reference carries a dense int64 ``Y [K,1,I,J]`` and a dense fp32 0/1 mask ``Wx`` through
every evaluation and multiplies by the mask (qmc/qmc.ipynb c1:70-72,114-115,150).  The CUDA path
builds, once per instance, a list of observed entries -- a 4-byte linear index ``k*IJ + p`` and a
1-byte level each -- grouped by (map, pixel sub-tile, band).  See include/qmc_b200.h.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import ObsView, check, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def plan_tiles(IJ: int, K: int, R: int, tile_warps: int = 8, smem_budget: int = 110 * 1024, lanes: bool = False,
               max_level: "int | None" = None):
    """Choose the pixel sub-tile size for the shared-memory (tiled / lane-stream) kernels.

    Returns (n_sub, sub_pixels, tile_warps).  A tile is ``tile_warps`` sub-tiles; its S and gS
    slices ([pixels][R] fp32 each) plus C and gC must fit ``smem_budget`` bytes.  A map that fits
    in one tile gets exactly one (no cross-CTA reduction of gC at all).  ``max_level`` (lane-stream layout):
    the largest level of the data, which decides whether the compact 16-bit stream words -- and their smaller
    ring -- can be used (default: plan for 32-bit words)."""
    RP = 1
    while RP < R:
        RP *= 2

    def plan(fixed, cap_pixels=None):
        max_tile_pixels = max((smem_budget - fixed) // (2 * RP * 4), tile_warps)
        if cap_pixels is not None:
            max_tile_pixels = min(max_tile_pixels, cap_pixels)
        if IJ <= max_tile_pixels:
            return tile_warps, -(-IJ // tile_warps), tile_warps
        sub = max(max_tile_pixels // tile_warps, 1)
        tiles = -(-IJ // (sub * tile_warps))
        return tiles * tile_warps, -(-IJ // (tiles * tile_warps)), tile_warps      # tiles evened out

    if not lanes:
        wc = tile_warps if tile_warps * K * RP * 4 <= 32 * 1024 else 1     # private gC copies (qmc_gather.cu)
        return plan((1 + wc) * K * RP * 4 + tile_warps * (K + 2) * 4 + tile_warps * 32 * RP * 4 + 16)
    # the lanes kernel always keeps one private gC copy per warp (K band rows + dummy + 32 continuation rows):
    # use fewer warps if K*R is large
    while tile_warps > 1 and (1 + tile_warps) * (K + 33) * RP * 4 > smem_budget // 2:
        tile_warps //= 2
    n_runs = lane_run_entries(K)

    def fixed(ring_bytes):     # C, gC copies, run tables, stream rings (lanes_layout, qmc_gather_common.cuh)
        return ((K + 1) + tile_warps * (K + 33)) * RP * 4 + tile_warps * (n_runs * 128 + ring_bytes) + 64

    if max_level is not None:
        lvl_bits = max(1, int(max_level).bit_length())
        if lvl_bits <= 8:
            n_sub, sub, tw = plan(fixed(2 * 512), cap_pixels=1 << (15 - lvl_bits))
            if sub * tw <= (1 << (15 - lvl_bits)):
                return n_sub, sub, tw
    return plan(fixed(4 * 512), cap_pixels=32768 - 32)


def lane_run_entries(K: int) -> int:
    """Run-table entries per lane of the lane-stream layout: a lane's quota of groups spans at most
    ceil(K/32) + 1 bands when the bands are of similar size; three more entries for uneven band sizes (the builder
    reports a lane that needs more and the caller retries with a longer table)."""
    return (K + 31) // 32 + 4


@dataclass
class ObsSet:
    """Device-resident compact observations of B maps of shape [K, IJ]."""
    idx: torch.Tensor        # int32 [nobs]
    lvl: torch.Tensor        # uint8 [nobs]
    row_off: torch.Tensor    # int64 [B*n_sub*K + 1]
    B: int
    K: int
    IJ: int
    n_sub: int
    sub_pixels: int
    tile_warps: int
    nobs: int
    max_level: int
    words: "torch.Tensor | None" = None        # lane-stream layout: uint32 words (stored as int32)
    stream_off: "torch.Tensor | None" = None   # lane-stream layout: int64 [B*n_sub + 1] word offsets
    nrows: "torch.Tensor | None" = None        # lane-stream layout: int32 [B*n_sub] steps per stream
    stream_stride: int = 0                     # lane-stream layout: > 0 when every stream has the same capacity
    n_runs: int = 0                            # lane-stream layout: run-table entries per lane
    word_bits: int = 32                        # lane-stream layout: 16 or 32
    lvl_bits: int = 0                          # lane-stream layout, 16-bit words: width of the level field
    has_cont: int = 0                          # lane-stream layout: bands split over lanes (continuation rows)

    @property
    def lanes(self) -> bool:
        return self.words is not None

    def view(self, map_modulo: int = 0) -> ObsView:
        return ObsView(self.idx.data_ptr(), self.lvl.data_ptr(), self.row_off.data_ptr(),
                       self.n_sub, self.sub_pixels,
                       self.words.data_ptr() if self.lanes else None,
                       self.stream_off.data_ptr() if self.lanes else None,
                       self.nrows.data_ptr() if self.lanes else None,
                       self.stream_stride if self.lanes else 0,
                       self.n_runs, self.word_bits, self.lvl_bits, self.has_cont, map_modulo)

    def padding_fraction(self) -> float:
        """Lane-stream layout: fraction of the walked slots that are padding."""
        if not self.lanes:
            return 0.0
        slots = int(self.nrows.sum().item()) * 32
        return 1.0 - self.nobs / max(slots, 1)

    @property
    def device(self):
        return self.idx.device

    def decode_stream(self, s: int):
        """Lane-stream layout, for tests and debugging: stream ``s`` as numpy arrays ``[step][lane]`` of
        (level, band, tile-local pixel, real, gC row) decoded from the run table and the words
        (include/qmc_b200.h).  ``real`` is False for padding."""
        import numpy as np
        if not self.lanes:
            raise ValueError("not a lane-stream observation set")
        base = s * self.stream_stride if self.stream_stride > 0 else int(self.stream_off[s].item())
        steps = int(self.nrows[s].item())
        ngroups = steps // 4
        nslots = -(-ngroups // 2) if self.word_bits == 16 else ngroups
        raw = self.words[base: base + self.n_runs * 32 + nslots * 128].cpu().numpy().view(np.uint32).astype(np.int64)
        table = raw[: self.n_runs * 32].reshape(self.n_runs, 32)
        band = np.full((ngroups, 32), self.K, dtype=np.int64)
        row = np.full((ngroups, 32), self.K, dtype=np.int64)
        for lane in range(32):
            g = 0
            for e in table[:, lane]:
                n = int(e >> 18)
                band[g: g + n, lane] = (e >> 9) & 0x1FF
                row[g: g + n, lane] = e & 0x1FF
                g += n
                if g >= ngroups:
                    break
            assert g >= ngroups, "run table shorter than the stream"
        body = raw[self.n_runs * 32:]
        if self.word_bits == 16:
            h = body.astype(np.uint32).view(np.uint16).astype(np.int64).reshape(nslots, 32, 2, 4)   # [slot][lane][group][step]
            w = h.transpose(0, 2, 3, 1).reshape(nslots * 8, 32)[:steps]
            lb = self.lvl_bits
            lv, pad, pix = w >> (16 - lb), (w >> (15 - lb)) & 1, w & ((1 << (15 - lb)) - 1)
            real = pad == 0
        else:
            w = body.reshape(nslots, 32, 4).transpose(0, 2, 1).reshape(nslots * 4, 32)[:steps]
            lv, pix = (((w >> 24) & 0x7F) << 1) | (w >> 31), w & 0x7FFF
            real = lv != 0xFF
        return lv, np.repeat(band, 4, axis=0), pix, real, np.repeat(row, 4, axis=0)

    def counts_per_map(self) -> torch.Tensor:
        ro = self.row_off[:: self.n_sub * self.K]
        return ro[1:] - ro[:-1]

    def algorithmic_bytes(self, R: int) -> int:
        """nobs*(4+1) + 2*4*R*(IJ+K) + 4 per map."""
        return self.nobs * 5 + self.B * (2 * 4 * R * (self.IJ + self.K) + 4)


def bank_mod_for_rank(R: int) -> int:
    """Residue modulus that makes the tiled kernel's gathers of [pixel][R_padded] fp32 rows
    conflict-free: rows of 4*R_padded bytes, 128 bytes served per shared-memory wavefront."""
    RP = 1
    while RP < R:
        RP *= 2
    return max(128 // (4 * RP), 1)


def build_obs(Y: torch.Tensor, Wx: torch.Tensor | None, K: int, IJ: int, B: int = 1, *,
              n_sub: int = 1, sub_pixels: int | None = None, tile_warps: int = 0, bank_mod: int = 0,
              lanes: bool = False) -> ObsSet:
    """(Y, Wx) dense ``[B][K][IJ]`` (any shape with that many elements; the reference's is
    ``[K,1,I,J]``) -> ObsSet on Y's CUDA device.  Y: int64 (reference dtype) or uint8."""
    if not Y.is_cuda:
        raise ValueError("build_obs needs CUDA tensors (no CPU path)")
    dev = Y.device
    if Y.dtype not in (torch.int64, torch.uint8):
        raise TypeError(f"Y must be int64 or uint8, got {Y.dtype}")
    Yc = Y.contiguous().reshape(-1)
    if Yc.numel() != B * K * IJ:
        raise ValueError(f"Y has {Yc.numel()} elements, expected B*K*IJ = {B * K * IJ}")
    if Y.dtype == torch.int64 and Yc.numel():
        lo, hi = int(Yc.min().item()), int(Yc.max().item())
        if lo < 0 or hi > 255:        # the compact format stores one byte per level: do not wrap silently
            raise ValueError(f"Y holds levels in [{lo}, {hi}]; the compact observation format supports 0..255")
    Wc = None
    if Wx is not None:
        Wc = Wx.to(device=dev, dtype=torch.float32).contiguous().reshape(-1)
        if Wc.numel() != B * K * IJ:
            raise ValueError("Wx and Y differ in size")
    if sub_pixels is None:
        sub_pixels = -(-IJ // n_sub)
    n_rows = B * n_sub * K
    with torch.cuda.device(dev):
        row_off = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
        ws = torch.empty(int(lib.qmc_obs_scan_ws_elems(n_rows)), dtype=torch.int64, device=dev)
        wptr = Wc.data_ptr() if Wc is not None else None
        check(lib.qmc_obs_count_scan(wptr, B, K, IJ, n_sub, sub_pixels, row_off.data_ptr(), ws.data_ptr(), _stream()))
        nobs = int(row_off[-1].item())
        idx = torch.empty(max(nobs, 1), dtype=torch.int32, device=dev)
        lvl = torch.empty(max(nobs, 1), dtype=torch.uint8, device=dev)
        check(lib.qmc_obs_fill(Yc.data_ptr(), int(Y.dtype == torch.int64), wptr, B, K, IJ, n_sub, sub_pixels,
                               bank_mod, row_off.data_ptr(), idx.data_ptr(), lvl.data_ptr(), _stream()))
        max_level = int(lvl[:nobs].max().item()) if nobs else 0
    obs = ObsSet(idx[:nobs] if nobs else idx[:0], lvl[:nobs] if nobs else lvl[:0], row_off, B, K, IJ,
                 n_sub, sub_pixels, tile_warps, nobs, max_level)
    return lane_streams(obs) if lanes else obs


def lane_word_format(max_level: int, tile_pixels: int) -> "tuple[int, int]":
    """(word_bits, lvl_bits) of the lane-stream words: 16-bit words when the level field, the padding flag and
    the tile-local pixel fit, 32-bit words otherwise (include/qmc_b200.h)."""
    lvl_bits = max(1, int(max_level).bit_length())
    if lvl_bits <= 8 and tile_pixels <= (1 << (15 - lvl_bits)):
        return 16, lvl_bits
    return 32, 0


def lane_streams(obs: ObsSet) -> ObsSet:
    """Re-cut a row-ordered observation set (built with ``bank_mod=0``) into the lane-stream layout
    (see qmc_obs_build_lanes): every lane of a stream's warp walks a list of (band, groups) runs, all lanes
    of a stream get the same number of groups, and the 32 entries of a step hit 32 different pixels."""
    if obs.K > 256:
        raise ValueError("the lane-stream layout supports at most 256 bands")
    if obs.max_level > 254:
        raise ValueError("the lane-stream layout supports levels 0..254")
    if obs.tile_warps <= 0:
        raise ValueError("the lane-stream layout needs a tiled observation set (tile_warps > 0)")
    dev = obs.device
    n_streams = obs.B * obs.n_sub
    per_stream = obs.row_off[:: obs.K][1:] - obs.row_off[:: obs.K][:-1]
    e_max = int(per_stream.max().item()) if n_streams else 0
    word_bits, lvl_bits = lane_word_format(obs.max_level, obs.tile_warps * obs.sub_pixels)
    n_runs = lane_run_entries(obs.K)
    extra = int(os.environ.get("QMC_LANES_EXTRA", "0"))
    split = int(os.environ.get("QMC_LANES_SPLIT", "1"))
    with torch.cuda.device(dev):
        nrows = torch.empty(n_streams, dtype=torch.int32, device=dev)
        overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        for attempt in range(12):
            # one capacity for all streams (that of the largest): a stream's address then needs no table look-up,
            # and a CTA can prefetch the data of the CTA that will follow it on its SM
            stride = int(lib.qmc_lanes_stream_words(e_max, obs.K, n_runs, word_bits, extra))
            words = torch.empty(max(n_streams * stride, 1), dtype=torch.int32, device=dev)
            check(lib.qmc_obs_build_lanes(obs.idx.data_ptr(), obs.lvl.data_ptr(), obs.row_off.data_ptr(), obs.B, obs.K,
                                          obs.IJ, obs.n_sub, obs.sub_pixels, obs.tile_warps, e_max, extra, split, stride,
                                          words.data_ptr(), nrows.data_ptr(), overflow.data_ptr(), n_runs, word_bits,
                                          lvl_bits, _stream()))
            ov = int(overflow.item())
            if not ov:
                break
            # dense sampling of small sub-tiles (every band wants the same few pixels in the same step) needs room
            # to spread out; many tiny bands per quota need a longer run table
            if os.environ.get("QMC_DEBUG_LANES"):
                print(f"lane_streams: attempt {attempt} overflow bits {ov:#x} (extra {extra}, n_runs {n_runs})")
            if ov & 13:
                extra = max(1, 2 * extra)
            if ov & 2:
                n_runs = min(64, n_runs + 4)
        else:
            raise RuntimeError("lane-stream layout: could not lay the observations out (pathological band/pixel structure)")
    stream_off = torch.arange(n_streams + 1, dtype=torch.int64, device=dev) * stride
    return ObsSet(obs.idx, obs.lvl, obs.row_off, obs.B, obs.K, obs.IJ, obs.n_sub, obs.sub_pixels, obs.tile_warps,
                  obs.nobs, obs.max_level, words, stream_off, nrows, stride, n_runs, word_bits, lvl_bits, int(split != 0))
