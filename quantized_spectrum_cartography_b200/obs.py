"""Compact observation sets: the (Y, Wx) information the likelihood actually uses.

The reference carries a dense int64 ``Y [K,1,I,J]`` and a dense fp32 0/1 mask ``Wx`` through
every evaluation and multiplies by the mask (qmc/qmc.ipynb c1:70-72,114-115,150).  The CUDA path
builds, once per instance, a list of observed entries -- a 4-byte linear index ``k*IJ + p`` and a
1-byte level each -- grouped by (map, pixel sub-tile, band).  See include/qmc_b200.h.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import ObsView, check, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def plan_tiles(IJ: int, K: int, R: int, tile_warps: int = 8, smem_budget: int = 110 * 1024, lanes: bool = False):
    """Choose the pixel sub-tile size for the shared-memory (tiled) kernel.

    Returns (n_sub, sub_pixels, tile_warps).  A tile is ``tile_warps`` sub-tiles; its S and gS
    slices ([pixels][R] fp32 each) plus C and gC must fit ``smem_budget`` bytes.  A map that fits
    in one tile gets exactly one (no cross-CTA reduction of gC at all)."""
    RP = 1
    while RP < R:
        RP *= 2
    if lanes:
        # the lanes kernel always keeps one private gC copy per warp: use fewer warps if K*R is large
        while tile_warps > 1 and (1 + tile_warps) * (K + 1) * RP * 4 > smem_budget // 2:
            tile_warps //= 2
        wc = tile_warps
        fixed = (1 + wc) * (K + 1) * RP * 4 + tile_warps * 4 * 512 + 16      # C, gC copies, stream rings
    else:
        wc = tile_warps if tile_warps * K * RP * 4 <= 32 * 1024 else 1     # private gC copies (qmc_gather.cu)
        fixed = (1 + wc) * K * RP * 4 + tile_warps * (K + 2) * 4 + tile_warps * 32 * RP * 4 + 16
    max_tile_pixels = max((smem_budget - fixed) // (2 * RP * 4), tile_warps)
    if IJ <= max_tile_pixels:
        sub = -(-IJ // tile_warps)
        return tile_warps, sub, tile_warps
    sub = max(max_tile_pixels // tile_warps, 1)
    tiles = -(-IJ // (sub * tile_warps))
    # even the tiles out
    sub = -(-IJ // (tiles * tile_warps))
    return tiles * tile_warps, sub, tile_warps


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

    @property
    def lanes(self) -> bool:
        return self.words is not None

    def view(self) -> ObsView:
        return ObsView(self.idx.data_ptr(), self.lvl.data_ptr(), self.row_off.data_ptr(),
                       self.n_sub, self.sub_pixels,
                       self.words.data_ptr() if self.lanes else None,
                       self.stream_off.data_ptr() if self.lanes else None,
                       self.nrows.data_ptr() if self.lanes else None,
                       self.stream_stride if self.lanes else 0)

    def padding_fraction(self) -> float:
        """Lane-stream layout: fraction of the walked slots that are padding."""
        if not self.lanes:
            return 0.0
        slots = int(self.nrows.sum().item()) * 32
        return 1.0 - self.nobs / max(slots, 1)

    @property
    def device(self):
        return self.idx.device

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


def lane_streams(obs: ObsSet) -> ObsSet:
    """Re-cut a row-ordered observation set (built with ``bank_mod=0``) into the lane-stream layout
    (see qmc_obs_build_lanes): every lane of a stream's warp walks one band at a time and the 32 entries
    of a step hit 32 different pixels.  The row-ordered arrays are consumed (permuted in place)."""
    if obs.K > 256:
        raise ValueError("the lane-stream layout supports at most 256 bands")
    if obs.max_level > 254:
        raise ValueError("the lane-stream layout supports levels 0..254")
    if obs.tile_warps <= 0:
        raise ValueError("the lane-stream layout needs a tiled observation set (tile_warps > 0)")
    dev = obs.device
    n_streams = obs.B * obs.n_sub
    per_stream = obs.row_off[:: obs.K][1:] - obs.row_off[:: obs.K][:-1]
    G = (obs.K + 31) // 32
    with torch.cuda.device(dev):
        nrows = torch.empty(n_streams, dtype=torch.int32, device=dev)
        overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        # capacity: 1.3x the even share of the busiest lane; dense sampling of small sub-tiles (every band
        # wants the same few pixels in the same step) can need more -> retry with twice the room
        for attempt in range(4):
            slack = 13 << attempt
            rows_cap = (slack * per_stream * G) // (10 * obs.K) + 4 * G + 8
            rows_cap = torch.clamp(((rows_cap + 3) // 4) * 4, min=16)     # the kernel loads the first four groups blindly
            # one capacity for all streams (the largest): a stream's address then needs no table look-up, and
            # a CTA can prefetch the data of the CTA that will follow it on its SM
            stride = int(rows_cap.max().item()) * 32 if n_streams else 512
            stream_off = torch.arange(n_streams + 1, dtype=torch.int64, device=dev) * stride
            words = torch.empty(max(n_streams * stride, 1), dtype=torch.int32, device=dev)
            check(lib.qmc_obs_build_lanes(obs.idx.data_ptr(), obs.lvl.data_ptr(), obs.row_off.data_ptr(), obs.B, obs.K,
                                          obs.IJ, obs.n_sub, obs.sub_pixels, obs.tile_warps, stream_off.data_ptr(),
                                          words.data_ptr(), nrows.data_ptr(), overflow.data_ptr(), _stream()))
            if not int(overflow.item()):
                break
        else:
            raise RuntimeError("lane-stream layout: a stream exceeded 8x its expected length (pathological band/pixel structure)")
    return ObsSet(obs.idx, obs.lvl, obs.row_off, obs.B, obs.K, obs.IJ, obs.n_sub, obs.sub_pixels, obs.tile_warps,
                  obs.nobs, obs.max_level, words, stream_off, nrows, stride)
