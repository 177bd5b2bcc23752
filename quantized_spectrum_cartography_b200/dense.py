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


def nll_fwd_bwd_dense(S2: torch.Tensor, C2: torch.Tensor, obs: DenseObs, lik: Likelihood, want_grad: bool = True, out=None):
    """``S2 [R, IJ]``, ``C2 [R, K]`` fp32 CUDA.  Returns (nll fp64 0-dim, gS [R, IJ], gC [R, K]).
    ``out=(nll fp64 [1], gS, gC)`` writes into existing buffers; gS may then be a column slice of a wider
    row-major buffer (row stride >= IJ, unit column stride), gC must be contiguous."""
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
        check(lib.qmc_nll_fwd_bwd_dense(S2.data_ptr(), C2.data_ptr(), obs.code.data_ptr(), C.byref(lik), IJ, K, R,
                                        nll.data_ptr(), gS.data_ptr() if want_grad else None,
                                        gS.stride(0) if want_grad else 0,
                                        gC.data_ptr() if want_grad else None, _stream()))
    return nll.reshape(-1)[0], gS, gC
