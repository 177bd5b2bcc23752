"""Dense-sampling path (BASELINE config 4): one large instance with a large fraction of its entries
observed.  ``X = S*C^T`` and both gradient contractions run on the tcgen05 tensor cores with the
quantized likelihood as the epilogue (csrc/qmc_dense.cu); the observation format is one byte per
dense entry (level, or 255 = not observed), pixel-major."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import Likelihood, check, lib
from .fused import _with_flags


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


@dataclass
class DenseObs:
    code: torch.Tensor   # uint8 [IJ, K]
    IJ: int
    K: int
    nobs: int
    max_level: int

    def algorithmic_bytes(self, R: int) -> int:
        """One code byte per dense entry + factors read once + gradients written once + the NLL."""
        return self.IJ * self.K + 2 * 4 * R * (self.IJ + self.K) + 8


def dense_supported(K: int, R: int) -> bool:
    return int(lib.qmc_dense_smem_bytes(K, R)) > 0


def pack_dense(Y: torch.Tensor, Wx: Optional[torch.Tensor], K: int) -> DenseObs:
    """(Y, Wx) in the reference's band-major layout (``[K,1,I,J]`` or ``[K, IJ]``) -> DenseObs."""
    if not Y.is_cuda:
        raise ValueError("pack_dense needs CUDA tensors (no CPU path)")
    if Y.dtype not in (torch.int64, torch.uint8):
        raise TypeError(f"Y must be int64 or uint8, got {Y.dtype}")
    Yc = Y.contiguous().reshape(K, -1)
    IJ = Yc.shape[1]
    Wc = None if Wx is None else Wx.to(device=Y.device, dtype=torch.float32).contiguous().reshape(K, -1)
    max_level = int(Yc.max().item())
    if max_level >= 255:
        raise ValueError("the dense format reserves code 255 for 'not observed'")
    code = torch.empty(IJ, K, dtype=torch.uint8, device=Y.device)
    with torch.cuda.device(Y.device):
        check(lib.qmc_dense_pack(Yc.data_ptr(), int(Y.dtype == torch.int64), None if Wc is None else Wc.data_ptr(),
                                 1, K, IJ, code.data_ptr(), _stream()))
    nobs = int((code != 255).sum().item())
    return DenseObs(code, IJ, K, nobs, max_level)


class PeerRegions:
    """Exchange regions of the fused factor-gradient exchange (``qmc_nll_fwd_bwd_dense_exchange``): this rank's own
    region (``qmc_peer_alloc``) and the peers' regions opened through CUDA IPC.  ``share`` is the host mechanism that
    hands the 64-byte handles (and, in a second round, every rank's verdict) around: a callable ``obj -> list of every
    rank's obj`` (default: ``torch.distributed.all_gather_object``).  If any rank cannot allocate or map a region, ALL
    ranks raise ``QmcError`` -- nobody is left waiting."""

    def __init__(self, rank: int, world: int, slot_floats: int, device, share=None):
        if not 1 <= world <= _lib.QMC_PEER_MAX_WORLD:
            raise ValueError(f"world {world}: the fused exchange supports 1..{_lib.QMC_PEER_MAX_WORLD} ranks")
        self.rank, self.world, self.device = rank, world, torch.device(device)
        self.slot_floats = (slot_floats + 3) // 4 * 4
        nbytes = int(lib.qmc_peer_region_bytes(world, self.slot_floats))
        own = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        self._opened: list[int] = []
        self._own = None
        with torch.cuda.device(self.device):
            # A rank that fails must not leave its peers waiting in the hand-shake: every rank always takes part in both
            # rounds (handles, then "did every mapping work"), and all ranks raise together.
            err = None
            try:
                check(lib.qmc_peer_alloc(nbytes, C.byref(own), handle))
                self._own = own.value
            except Exception as e:          # noqa: BLE001 -- reported to every rank below
                err = repr(e)
            if share is None:
                import torch.distributed as dist

                def share(h):
                    got = [None] * world
                    dist.all_gather_object(got, h)
                    return got
            mine = None if err else bytes(handle)
            handles = share(mine) if world > 1 else [mine]
            self.px = _lib.PeerExchange()
            self.px.rank, self.px.world, self.px.slot_floats = rank, world, self.slot_floats
            if err is None and all(h is not None for h in handles):
                try:
                    for q in range(world):
                        if q == rank:
                            self.px.region[q] = self._own
                            continue
                        ptr = C.c_void_p()
                        check(lib.qmc_peer_open((C.c_ubyte * 64).from_buffer_copy(handles[q]), C.byref(ptr)))
                        self._opened.append(ptr.value)
                        self.px.region[q] = ptr.value
                except Exception as e:      # noqa: BLE001
                    err = repr(e)
            elif err is None:
                err = "a peer could not allocate its exchange region"
            verdicts = share(err) if world > 1 else [err]
            bad = [f"rank {q}: {v}" for q, v in enumerate(verdicts) if v is not None]
            if bad:
                self.close()
                raise _lib.QmcError("fused exchange unavailable (" + "; ".join(bad) + ")")

    def status(self) -> int:
        """0 = every exchange so far completed; 1 = a peer never arrived (the kernel gave up after ~2 s).  Synchronises
        the current stream."""
        st = C.c_int(0)
        with torch.cuda.device(self.device):
            check(lib.qmc_peer_status(self._own, C.byref(st), _stream()))
        return st.value

    def close(self):
        """Every rank must have finished its last exchange (e.g. after a barrier) before any rank closes."""
        with torch.cuda.device(self.device):
            for p in self._opened:
                check(lib.qmc_peer_close(p))
            self._opened = []
            if self._own:
                check(lib.qmc_peer_free(self._own))
                self._own = None


def nll_fwd_bwd_dense(S2: torch.Tensor, C2: torch.Tensor, obs: DenseObs, lik: Likelihood, want_grad: bool = True, out=None,
                      peers: Optional[PeerRegions] = None):
    """``S2 [R, IJ]``, ``C2 [R, K]`` fp32 CUDA.  Returns (nll fp64 0-dim, gS [R, IJ], gC [R, K]).
    ``out=(nll fp64 [1], gS, gC)`` writes into existing buffers; gS may then be a column slice of a wider
    row-major buffer (row stride >= IJ, unit column stride), gC must be contiguous.
    ``peers``: this rank evaluates its pixel block of a sharded instance and the kernel itself exchanges the partial
    gC and NLL with the peer ranks over NVLink: nll and gC come back summed over all ranks (same bits everywhere)."""
    if not (S2.is_cuda and C2.is_cuda):
        raise ValueError("nll_fwd_bwd_dense needs CUDA tensors: there is no CPU path")
    R, IJ = S2.shape
    K = C2.shape[1]
    if (IJ, K) != (obs.IJ, obs.K) or C2.shape[0] != R:
        raise ValueError("factor shapes do not match the observation set")
    if obs.max_level + 2 > lik.n_bounds:
        raise ValueError(f"Y contains level {obs.max_level} but the table has only {lik.n_bounds - 1} levels")
    S2 = S2.contiguous()
    C2 = C2.contiguous()
    lik = _with_flags(lik, not want_grad)
    with torch.cuda.device(S2.device):
        if out is not None:
            nll, gS, gC = out
            if nll.dtype != torch.float64 or nll.numel() < 1:
                raise ValueError("out[0] must hold one float64")
            if want_grad and (gS.shape != (R, IJ) or gS.stride(1) != 1 or gS.stride(0) < IJ or gS.dtype != torch.float32
                              or not gC.is_contiguous() or gC.shape != (R, K)):
                raise ValueError("out buffers: gS [R, IJ] fp32 with unit column stride, gC [R, K] contiguous")
        else:
            nll = torch.empty(1, dtype=torch.float64, device=S2.device)
            gS = torch.empty_like(S2) if want_grad else None
            gC = torch.empty_like(C2) if want_grad else None
        if peers is not None:
            if not want_grad:
                raise ValueError("the fused exchange combines gradients: want_grad=False is not supported with it")
            check(lib.qmc_nll_fwd_bwd_dense_exchange(S2.data_ptr(), C2.data_ptr(), obs.code.data_ptr(), C.byref(lik), IJ, K, R,
                                                     nll.data_ptr(), gS.data_ptr(), gS.stride(0), gC.data_ptr(),
                                                     C.byref(peers.px), _stream()))
        else:
            check(lib.qmc_nll_fwd_bwd_dense(S2.data_ptr(), C2.data_ptr(), obs.code.data_ptr(), C.byref(lik), IJ, K, R,
                                            nll.data_ptr(), gS.data_ptr() if want_grad else None,
                                            gS.stride(0) if want_grad else 0,
                                            gC.data_ptr() if want_grad else None, _stream()))
    return nll.reshape(-1)[0], gS, gC
