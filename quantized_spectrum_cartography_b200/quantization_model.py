"""Linear-domain call surface: same names and positional signatures as the reference's
``qmc/quantization_model.py`` so that ``from quantization_model import *`` callers
(qmc/qmc.ipynb c1:9-10, backup/notebooks/onebit_lowrank.ipynb c1:6) can switch imports.

Everything computes on the GPU.  The hot path is :func:`qmc_nll` (fused CUDA kernel); the
compositional pieces the notebooks chain by hand (``get_tensor``, ``prob_probit``, ``F_probit`` ...)
stay available as thin GPU versions with autograd so those notebooks still run, but composing them
re-materialises the dense tensor exactly like the reference does -- use ``qmc_nll`` in a loop.

CPU tensors are accepted (the reference works on CPU tensors): they are moved to the current CUDA
device and the result is returned on the input's device.  Without a CUDA device every function
raises; there is no CPU implementation.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import check, lib
from .fused import REF_SENTINEL, make_likelihood, make_obs, qmc_nll, qmc_nll_batched  # noqa: F401

__all__ = ["quantize", "prob_probit", "F_sigmoid", "dither_sigmoid", "F_probit", "dither_probit", "outer",
           "get_tensor", "NMSE", "NegLikelihood", "DeterministicCost", "qmc_nll", "make_obs"]

REF_SQRT2 = 1.414213  # quantization_model.py:61


def _cuda_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("quantized_spectrum_cartography_b200 needs a CUDA device (no CPU implementation)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_cuda else t.to(_cuda_device())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def assign_levels(noisy: torch.Tensor, bin_boundaries) -> torch.Tensor:
    """Level index (int64) of an already-noisy tensor: the loop of quantization_model.py:14-20 as
    one CUDA pass, bit-exact (left-open/right-closed cells, boundary 0 ignored, top level unbounded,
    NaN -> 0)."""
    bb = torch.as_tensor(bin_boundaries, dtype=torch.float32).detach().cpu().reshape(-1)
    n = bb.numel()
    src = _to_dev(noisy).to(torch.float32).contiguous()
    out = torch.empty(src.shape, dtype=torch.int64, device=src.device)
    arr = (C.c_float * n)(*bb.tolist())
    with torch.cuda.device(src.device):
        check(lib.qmc_quantize_levels(src.data_ptr(), src.numel(), arr, n, None, out.data_ptr(), _stream()))
    return out.to(noisy.device)


def _noisy(X: torch.Tensor, noise_std, offset=None) -> torch.Tensor:
    """X + randn*std (or log(X+offset) + randn*std), drawn where the reference draws it.

    For a CPU ``X`` the noise comes from torch's global CPU generator with the same call
    (``torch.randn(X.shape)``, quantization_model.py:13) and is added on the host with the same two
    roundings, so a seeded run sees bit-identical noisy values on both back ends.  For a CUDA ``X``
    the noise comes from the CUDA generator and the addition runs in qmc_noisy_signal."""
    std = float(noise_std.item()) if isinstance(noise_std, torch.Tensor) else float(noise_std)
    if not X.is_cuda:
        base = X if offset is None else torch.log(X + offset)
        return base + torch.randn(X.shape) * noise_std
    noise = torch.randn(X.shape, device=X.device, dtype=torch.float32)
    Xc = X.to(torch.float32).contiguous()
    out = torch.empty_like(Xc)
    with torch.cuda.device(X.device):
        check(lib.qmc_noisy_signal(Xc.data_ptr(), noise.data_ptr(), std, 0.0 if offset is None else float(offset),
                                   int(offset is not None), Xc.numel(), out.data_ptr(), _stream()))
    return out


def quantize(X, noise_std, bin_boundaries):
    """Y = Q(X + E), E ~ N(0, noise_std): bin index per entry (quantization_model.py:8-20)."""
    return assign_levels(_noisy(X, noise_std), bin_boundaries)


def F_sigmoid(y):
    """1/(1+exp(-y))  (quantization_model.py:43-47)."""
    yd = _to_dev(y)
    return (1 / (1 + torch.exp(-yd))).to(y.device)


def dither_sigmoid(y):
    """Bernoulli sample with parameter F_sigmoid(y)  (quantization_model.py:49-55)."""
    return torch.bernoulli(F_sigmoid(y))


def F_probit(y, std):
    """0.5*(1 + erf(y/(std*1.414213)))  (quantization_model.py:57-61), truncated sqrt(2) included."""
    yd = _to_dev(y)
    return ((1 / 2) * (1 + torch.erf(yd / (std * REF_SQRT2)))).to(y.device)


def dither_probit(y, std):
    """Bernoulli sample with parameter F_probit(y, std)  (quantization_model.py:63-68)."""
    return torch.bernoulli(F_probit(y, std))


def _effective_boundaries(bin_boundaries, sentinels: bool, device) -> torch.Tensor:
    bb = torch.as_tensor(bin_boundaries, dtype=torch.float32).detach().clone().to(device)
    if sentinels:
        bb[0] = -REF_SENTINEL
        bb[-1] = REF_SENTINEL
    return bb


def _prob_probit(Y, X_hat, bin_boundaries, noise_std, sentinels: bool):
    Xd = _to_dev(X_hat)
    Yd = Y.to(Xd.device)
    bb = _effective_boundaries(bin_boundaries, sentinels, Xd.device)
    lower, upper = bb[Yd], bb[Yd + 1]
    a = noise_std * REF_SQRT2
    P = (1 / 2) * (1 + torch.erf((upper - Xd) / a)) - (1 / 2) * (1 + torch.erf((lower - Xd) / a))
    return P.to(X_hat.device)


def prob_probit(Y, X_hat, bin_boundaries, noise_std):
    """P(Y | X_hat) = F(U - X_hat) - F(W - X_hat) with the outer boundaries replaced by -/+1e5
    (quantization_model.py:22-39).  Compositional GPU version (dense, same fp32 formula as the
    reference, so the same tail underflow); the solver path is :func:`qmc_nll`."""
    return _prob_probit(Y, X_hat, bin_boundaries, noise_std, sentinels=True)


class _GetTensor(torch.autograd.Function):
    @staticmethod
    def forward(ctx, S3, C3):
        B, R, IJ = S3.shape
        K = C3.shape[2]
        X = torch.empty(B, K, IJ, dtype=torch.float32, device=S3.device)
        with torch.cuda.device(S3.device):
            check(lib.qmc_get_tensor(S3.data_ptr(), C3.data_ptr(), B, IJ, K, R, X.data_ptr(), _stream()))
        ctx.save_for_backward(S3, C3)
        return X

    @staticmethod
    def backward(ctx, gX):
        S3, C3 = ctx.saved_tensors
        gS = torch.bmm(C3, gX) if ctx.needs_input_grad[0] else None                    # [B,R,K]@[B,K,IJ]
        gC = torch.bmm(S3, gX.transpose(1, 2)) if ctx.needs_input_grad[1] else None    # [B,R,IJ]@[B,IJ,K]
        return gS, gC


def get_tensor(S: torch.Tensor, C_: torch.Tensor):
    """sum_r S[r] o C[r] -> [K, I, J]  (quantization_model.py:79-86)."""
    R, K = C_.shape
    I, J = S.shape[-2:]
    S3 = _to_dev(S).to(torch.float32).reshape(1, R, I * J).contiguous()
    C3 = _to_dev(C_).to(torch.float32).reshape(1, R, K).contiguous()
    return _GetTensor.apply(S3, C3).reshape(K, I, J).to(S.device)


def outer(mat: torch.Tensor, vec: torch.Tensor):
    """prod[i] = mat * vec[i]  (quantization_model.py:70-77)."""
    return get_tensor(mat.reshape(1, 1, *mat.shape), vec.reshape(1, -1))


def NMSE(T: torch.Tensor, T_target: torch.Tensor):
    """||T - T*||_F / ||T*||_F -- not squared  (quantization_model.py:88-92)."""
    Td, Tt = _to_dev(T), _to_dev(T_target)
    return (torch.norm(Td - Tt, "fro") / torch.norm(Tt, "fro")).to(T.device)


def nmse_factors(S, C_, T_target, offset=None):
    """NMSE (or NMSE_LOG when ``offset`` is given) of S*C^T against a dense target without
    materialising the reconstruction (one fused pass; solver loops call this every iteration,
    qmc.ipynb c1:160,215)."""
    R, K = C_.shape[-2:]
    if C_.dim() == 3:                        # batched: S [B,R,IJ], C [B,R,K], T_target [B,K,IJ]
        S3 = _to_dev(S).to(torch.float32).reshape(C_.shape[0], R, -1).contiguous()
    else:                                    # reference shapes: S [R,1,I,J] or [R,I,J], C [R,K]
        S3 = _to_dev(S).to(torch.float32).reshape(1, R, -1).contiguous()
    B, _, IJ = S3.shape
    C3 = _to_dev(C_).to(torch.float32).reshape(B, R, K).contiguous()
    Xr = _to_dev(T_target).to(torch.float32).reshape(B, K, IJ).contiguous()
    out = torch.empty(B, 2, dtype=torch.float64, device=S3.device)
    with torch.cuda.device(S3.device):
        check(lib.qmc_nmse_terms(S3.data_ptr(), C3.data_ptr(), Xr.data_ptr(), B, IJ, K, R, int(offset is not None),
                                 0.0 if offset is None else float(offset), out.data_ptr(), _stream()))
    r = torch.sqrt(out[:, 0] / out[:, 1]).to(torch.float32)
    return r[0] if B == 1 else r


class _BceOneBit(torch.autograd.Function):
    """One launch: the BCE value and d loss / d T_sample (qmc_bce_one_bit)."""

    @staticmethod
    def forward(ctx, Ts, Tt, mean, std, probit):
        x = Ts.detach().to(torch.float32).contiguous()
        t = Tt.detach().to(torch.float32).contiguous()
        if x.numel() != t.numel():
            raise ValueError("T_sample and T_target differ in size")
        loss = torch.empty(1, dtype=torch.float64, device=x.device)
        gx = torch.empty_like(x) if Ts.requires_grad else None
        with torch.cuda.device(x.device):
            check(lib.qmc_bce_one_bit(x.data_ptr(), t.data_ptr(), x.numel(), float(mean), 0.0 if std is None else float(std),
                                      int(bool(probit)), loss.data_ptr(), None if gx is None else gx.data_ptr(), _stream()))
        if gx is not None:
            ctx.save_for_backward(gx)
        ctx.shape = Ts.shape
        return loss[0].to(torch.float32)

    @staticmethod
    def backward(ctx, grad_out):
        (gx,) = ctx.saved_tensors
        return gx.reshape(ctx.shape) * grad_out, None, None, None, None


class NegLikelihood(nn.Module):
    """One-bit BCE form (quantization_model.py:97-113): BCELoss(F_probit(T-mean, std) or
    F_sigmoid(T-mean), target), mean reduction, no mask, log clamped at -100 -- one fused CUDA launch for the
    value and the gradient with respect to T_sample."""

    def __init__(self, mean, std=None, probit=True):
        super().__init__()
        if probit:
            assert std is not None
        self.mean, self.std, self.probit = mean, std, probit

    def forward(self, T_sample, T_target):
        Ts, Tt = _to_dev(T_sample), _to_dev(T_target)
        mean = float(self.mean.item()) if isinstance(self.mean, torch.Tensor) else float(self.mean)
        std = None if self.std is None else (float(self.std.item()) if isinstance(self.std, torch.Tensor) else float(self.std))
        return _BceOneBit.apply(Ts, Tt, mean, std, self.probit).to(T_sample.device)


class DeterministicCost(nn.Module):
    """-lambda * sum((T-mean)*T_target) + ||T-mean||_F  (quantization_model.py:115-129)."""

    def __init__(self, mean=0):
        super().__init__()
        self.lambda_reg = 0.001
        self.mean = mean

    def forward(self, S, C_, T_target):
        T_hat = get_tensor(_to_dev(S), _to_dev(C_)) - self.mean
        out = -self.lambda_reg * ((T_hat * _to_dev(T_target)).sum()) + torch.norm(T_hat, "fro")
        return out.to(S.device)
