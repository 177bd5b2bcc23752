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


def plan_tiles(IJ: int, K: int, R: int, tile_warps: int = 8, smem_budget: int = 100 * 1024):
    """Choose the pixel sub-tile size for the shared-memory (tiled) kernel.

    Returns (n_sub, sub_pixels, tile_warps).  A tile is ``tile_warps`` sub-tiles; its S and gS
    slices ([pixels][R] fp32 each) plus C and gC must fit ``smem_budget`` bytes.  A map that fits
    in one tile gets exactly one (no cross-CTA reduction of gC at all)."""
    RP = 1
    while RP < R:
        RP *= 2
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

    def view(self) -> ObsView:
        return ObsView(self.idx.data_ptr(), self.lvl.data_ptr(), self.row_off.data_ptr(),
                       self.n_sub, self.sub_pixels)

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
              n_sub: int = 1, sub_pixels: int | None = None, tile_warps: int = 0, bank_mod: int = 0) -> ObsSet:
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
    return ObsSet(idx[:nobs] if nobs else idx[:0], lvl[:nobs] if nobs else lvl[:0], row_off, B, K, IJ,
                  n_sub, sub_pixels, tile_warps, nobs, max_level)
