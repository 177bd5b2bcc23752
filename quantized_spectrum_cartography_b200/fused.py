"""Fused masked low-rank reconstruction + quantized NLL + factor gradients (the hot path).

``qmc_nll`` replaces the four lines every solver step of the reference repeats
(qmc/qmc.ipynb c1:145-151, again at :175-179, :189-193, :204-209)::

    T_hat = get_tensor(S, C).unsqueeze(dim=1)
    T_hat = torch.log(T_hat + offset)                      # log-domain variant only
    nll   = -torch.sum(Wx * torch.log(prob_probit(Y, T_hat, bin_boundaries, std_probit)))
    ...; cost.backward()                                   # S.grad, C.grad

with one CUDA launch that never materialises ``T_hat`` and visits only observed entries.  The
result is a 0-dim tensor whose ``backward()`` fills ``S.grad`` / ``C.grad`` (``S`` may be a
non-leaf, e.g. a generator output, as in c1:201-211).
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional

import torch

from . import _lib
from ._lib import Likelihood, check, lib
from .obs import ObsSet, bank_mod_for_rank, build_obs, plan_tiles

REF_SENTINEL = 100000.0  # quantization_model.py:32-33


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _as_float(v) -> float:
    return float(v.item()) if isinstance(v, torch.Tensor) else float(v)


def make_likelihood(bin_boundaries, noise_std, *, offset=None, log_domain: Optional[bool] = None,
                    sentinels: Optional[bool] = None, reference_epilogue: bool = False,
                    forward_only: bool = False, least_squares: bool = False, model: str = "probit") -> Likelihood:
    """Pack the model the way the reference's two files define it.

    ``offset is None`` -> linear-domain file: no log link, outer boundaries replaced by -/+1e5
    (quantization_model.py:31-33).  ``offset`` given -> log-domain file: x = log(t + offset),
    boundaries used as they are (quantization_model_log.py:32-34).  ``log_domain`` / ``sentinels``
    override either default.  The caller's table is never modified (the reference clones it).

    ``model="logistic"`` replaces the Gaussian CDF F_probit by the logistic CDF F_sigmoid
    (quantization_model.py:43-47) in ``P = F(U - x) - F(W - x)``; ``noise_std`` is then the logistic scale
    (1 = F_sigmoid as the reference writes it).

    ``least_squares=True`` selects the masked least-squares baseline on the bin mid-points instead of
    the likelihood (qmc_dowjons.ipynb c1:112; quantization_model_log.py:43-51): ``noise_std`` is not
    used and the table is taken as it is (the reference's mid-point function never substitutes
    sentinels) unless ``sentinels=True`` is passed explicitly."""
    if model not in ("probit", "logistic"):
        raise ValueError(f"model must be 'probit' or 'logistic', got {model!r}")
    if model == "logistic" and (least_squares or reference_epilogue):
        raise ValueError("model='logistic' cannot be combined with least_squares / reference_epilogue")
    bb = torch.as_tensor(bin_boundaries, dtype=torch.float32).detach().cpu().reshape(-1).clone()
    n = bb.numel()
    if not 2 <= n <= _lib.QMC_MAX_BOUNDS:
        raise ValueError(f"need 2..{_lib.QMC_MAX_BOUNDS} boundaries, got {n}")
    if log_domain is None:
        log_domain = offset is not None
    if sentinels is None:
        sentinels = not log_domain and not least_squares
    if sentinels:
        bb[0], bb[-1] = -REF_SENTINEL, REF_SENTINEL
    lik = Likelihood()
    lik.n_bounds = n
    lik.flags = ((_lib.QMC_LOG_DOMAIN if log_domain else 0) | (_lib.QMC_EPI_REFERENCE if reference_epilogue else 0)
                 | (_lib.QMC_FORWARD_ONLY if forward_only else 0) | (_lib.QMC_EPI_LSQ if least_squares else 0)
                 | (_lib.QMC_EPI_LOGISTIC if model == "logistic" else 0))
    lik.noise_std = 1.0 if least_squares and noise_std is None else _as_float(noise_std)
    lik.offset = 0.0 if offset is None else _as_float(offset)
    for i, v in enumerate(bb.tolist()):
        lik.bounds[i] = v
    return lik


def _with_flags(lik: Likelihood, forward_only: bool) -> Likelihood:
    out = Likelihood.from_buffer_copy(lik)
    out.flags = (lik.flags & ~_lib.QMC_FORWARD_ONLY) | (_lib.QMC_FORWARD_ONLY if forward_only else 0)
    return out


def _pixel_major(S3: torch.Tensor) -> bool:
    B, R, IJ = S3.shape
    return S3.stride(1) == 1 and S3.stride(2) == R and (B == 1 or S3.stride(0) == R * IJ)


def nll_candidates(S3: torch.Tensor, C3: torch.Tensor, obs: ObsSet, lik: Likelihood) -> torch.Tensor:
    """Forward-only NLL of D candidate factors per map in ONE launch: ``S3 [D*B, R, IJ]`` (candidate-major: rows
    d*B .. d*B+B-1 are candidate d of maps 0..B-1), ``C3 [B, R, K]`` shared by a map's candidates, ``obs`` the
    lane-stream observation set of the B maps (read D times, never copied).  Returns ``[D, B]`` fp64.  The batched
    random-restart latent search (qmc.ipynb c1:168-197)."""
    if not obs.lanes:
        raise ValueError("nll_candidates needs a lane-stream observation set")
    DB, R, IJ = S3.shape
    B, K = obs.B, C3.shape[2]
    if DB % B or C3.shape[:2] != (B, R) or (obs.K, obs.IJ) != (K, IJ):
        raise ValueError("shapes: S3 [D*B, R, IJ], C3 [B, R, K] for an observation set of B maps")
    if obs.max_level + 2 > lik.n_bounds:
        raise ValueError(f"Y contains level {obs.max_level} but the table has only {lik.n_bounds - 1} levels")
    if not (S3.is_contiguous() or _pixel_major(S3)):
        S3 = S3.contiguous()
    C3 = C3.contiguous()
    lik = _with_flags(lik, True)
    with torch.cuda.device(S3.device):
        nll = torch.empty(DB, dtype=torch.float64, device=S3.device)
        view = obs.view(map_modulo=B)
        check(lib.qmc_nll_fwd_bwd_gather(S3.data_ptr(), S3.stride(0), S3.stride(1), S3.stride(2), C3.data_ptr(), C.byref(view),
                                         C.byref(lik), DB, IJ, K, R, _lib.QMC_ALGO_LANES, obs.tile_warps, nll.data_ptr(),
                                         None, None, _stream()))
    return nll.reshape(DB // B, B)


def nll_fwd_bwd(S3: torch.Tensor, C3: torch.Tensor, obs: ObsSet, lik: Likelihood, *, algo: int = _lib.QMC_ALGO_AUTO,
                want_grad: bool = True, out=None, skip_gs: bool = False, skip_gc: bool = False):
    """Raw call: S3 [B,R,IJ] fp32 CUDA (emitter-major contiguous, or pixel-major storage viewed
    as [B,R,IJ]), C3 [B,R,K].  Returns (nll fp64 [B], gS like S3 or None, gC [B,R,K] or None).
    ``out=(nll, gS, gC)`` reuses buffers of those shapes, dtypes and strides.  ``skip_gs`` / ``skip_gc``
    tell the lane-stream kernel that one gradient is not needed (its buffer is then left untouched);
    the other kernels compute both regardless."""
    if not (S3.is_cuda and C3.is_cuda):
        raise ValueError("nll_fwd_bwd needs CUDA tensors: there is no CPU path")
    if S3.dtype != torch.float32 or C3.dtype != torch.float32:
        raise TypeError("S and C must be float32")
    B, R, IJ = S3.shape
    K = C3.shape[2]
    if C3.shape[:2] != (B, R):
        raise ValueError(f"C is {tuple(C3.shape)}, expected ({B}, {R}, K)")
    if (obs.B, obs.K, obs.IJ) != (B, K, IJ):
        raise ValueError(f"observation set is for (B,K,IJ)=({obs.B},{obs.K},{obs.IJ}), factors give ({B},{K},{IJ})")
    if obs.max_level + 2 > lik.n_bounds:
        raise ValueError(f"Y contains level {obs.max_level} but the table has only {lik.n_bounds - 1} levels")
    if obs.device != S3.device:
        raise ValueError("observations and factors live on different devices")
    if not (S3.is_contiguous() or _pixel_major(S3)):
        S3 = S3.contiguous()
    C3 = C3.contiguous()
    lik = _with_flags(lik, not want_grad)
    if want_grad and obs.lanes:
        lik.flags |= (_lib.QMC_SKIP_GS if skip_gs else 0) | (_lib.QMC_SKIP_GC if skip_gc else 0)
    with torch.cuda.device(S3.device):
        if out is not None:
            nll, gS, gC = out
            if want_grad and (gS.stride() != S3.stride() or not gC.is_contiguous()):
                raise ValueError("out buffers must have the strides of S3 (gS) and be contiguous (gC)")
        else:
            # a gradient the lane-stream kernel is told to skip is neither computed nor allocated
            need_gs = want_grad and not (skip_gs and obs.lanes)
            need_gc = want_grad and not (skip_gc and obs.lanes)
            nll = torch.empty(B, dtype=torch.float64, device=S3.device)
            gS = torch.empty_strided(S3.shape, S3.stride(), dtype=torch.float32, device=S3.device) if need_gs else None
            gC = torch.empty_like(C3) if need_gc else None
        view = obs.view()
        check(lib.qmc_nll_fwd_bwd_gather(
            S3.data_ptr(), S3.stride(0), S3.stride(1), S3.stride(2), C3.data_ptr(), C.byref(view), C.byref(lik),
            B, IJ, K, R, algo, obs.tile_warps, nll.data_ptr(),
            gS.data_ptr() if gS is not None else None, gC.data_ptr() if gC is not None else None, _stream()))
    return nll, gS, gC


class _QmcNll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, S3, C3, obs, lik, algo):
        want = S3.requires_grad or C3.requires_grad
        # the alternating solvers differentiate with respect to one factor at a time (qmc.ipynb c1:145,204:
        # the other one is detached): the lane-stream kernel then skips the gradient nobody asked for
        nll, gS, gC = nll_fwd_bwd(S3.detach(), C3.detach(), obs, lik, algo=algo, want_grad=want,
                                  skip_gs=want and not S3.requires_grad, skip_gc=want and not C3.requires_grad)
        ctx.has_grad = want
        if want:
            ctx.save_for_backward(*(x if x is not None else nll.new_empty(0) for x in (gS, gC)))
        return nll

    @staticmethod
    def backward(ctx, grad_nll):
        if not ctx.has_grad:
            return None, None, None, None, None
        gS, gC = ctx.saved_tensors
        w = grad_nll.to(torch.float32).reshape(-1, 1, 1)
        return (gS * w if ctx.needs_input_grad[0] else None, gC * w if ctx.needs_input_grad[1] else None,
                None, None, None)


def qmc_nll_batched(S3, C3, obs: ObsSet, lik: Likelihood, algo: int = _lib.QMC_ALGO_AUTO) -> torch.Tensor:
    """Per-map NLL ``[B]`` (fp64) of B independent maps; differentiable w.r.t. S3 and C3."""
    return _QmcNll.apply(S3, C3, obs, lik, algo)


# ---- the reference-shaped single-instance call ---------------------------------------------------
_OBS_CACHE: "dict[tuple, tuple]" = {}
_OBS_CACHE_MAX = 8


def _cached_obs(Y: torch.Tensor, Wx: Optional[torch.Tensor], K: int, IJ: int, device, R: int) -> ObsSet:
    """Observation sets are built once per (Y, Wx) pair and reused across solver iterations, the
    way the reference reuses Y and Wx themselves (qmc.ipynb c1:114-115 outside the loop)."""
    key = (id(Y), Y._version, None if Wx is None else id(Wx), None if Wx is None else Wx._version,
           K, IJ, str(device))
    hit = _OBS_CACHE.get(key)
    if hit is not None and hit[0]() is Y and (Wx is None or hit[1]() is Wx):
        return hit[2]
    n_sub, sub, tw = 1, IJ, 0
    obs = build_obs(Y.to(device), None if Wx is None else Wx.to(device), K, IJ, 1, n_sub=n_sub, sub_pixels=sub,
                    tile_warps=tw)
    if len(_OBS_CACHE) >= _OBS_CACHE_MAX:
        _OBS_CACHE.pop(next(iter(_OBS_CACHE)))
    _OBS_CACHE[key] = (weakref.ref(Y), weakref.ref(Wx) if Wx is not None else None, obs)
    return obs


def qmc_nll(S, C_, Y, Wx, bin_boundaries, noise_std, offset=None, log_domain=None, sentinels=None,
            reference_epilogue=False, obs: Optional[ObsSet] = None, device=None, model: str = "probit") -> torch.Tensor:
    """Drop-in for the reference idiom (module docstring).  Shapes and dtypes are the reference's:
    ``S [R,1,I,J]`` (or ``[R,I,J]``), ``C [R,K]``, ``Y`` int64 ``[K,1,I,J]``, ``Wx`` 0/1 float of
    the same shape, ``bin_boundaries`` 1-D (not modified), ``noise_std``/``offset`` floats or
    0-dim tensors.  CPU inputs are moved to the current CUDA device; the 0-dim fp32 result lives on
    S's device.  Pass ``obs`` (from :func:`make_obs`) to skip the (Y, Wx) cache lookup."""
    if not torch.cuda.is_available():
        raise RuntimeError("qmc_nll needs a CUDA device: this package has no CPU implementation")
    dev = torch.device(device) if device is not None else (S.device if S.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    R, K = C_.shape
    S3 = S.reshape(1, R, -1)
    IJ = S3.shape[2]
    lik = make_likelihood(bin_boundaries, noise_std, offset=offset, log_domain=log_domain, sentinels=sentinels,
                          reference_epilogue=reference_epilogue, model=model)
    if obs is None:
        obs = _cached_obs(Y, Wx, K, IJ, dev, R)
    S3d = S3.to(device=dev, dtype=torch.float32)
    C3d = C_.reshape(1, R, K).to(device=dev, dtype=torch.float32)
    nll = _QmcNll.apply(S3d, C3d, obs, lik, _lib.QMC_ALGO_AUTO)
    return nll[0].to(dtype=torch.float32, device=S.device)


def qmc_lsq(S, C_, Y, Wx, bin_boundaries, offset=None, log_domain=None, obs: Optional[ObsSet] = None,
            device=None) -> torch.Tensor:
    """The masked least-squares baseline as one fused call: drop-in for

        Obs   = get_quantized_obs_from_ordinal(Y, bin_boundaries, std)
        T_hat = get_tensor(S, C).unsqueeze(1); [T_hat = torch.log(T_hat + offset)]
        cost  = torch.norm(Wx * (T_hat - Obs)) ** 2

    (qmc_dowjons.ipynb c1:84,108-112).  Same shapes, dtypes, caching and autograd behaviour as
    :func:`qmc_nll`; the observed-entry kernels run with the least-squares epilogue."""
    if not torch.cuda.is_available():
        raise RuntimeError("qmc_lsq needs a CUDA device: this package has no CPU implementation")
    dev = torch.device(device) if device is not None else (S.device if S.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    R, K = C_.shape
    S3 = S.reshape(1, R, -1)
    IJ = S3.shape[2]
    lik = make_likelihood(bin_boundaries, None, offset=offset, log_domain=log_domain, least_squares=True)
    if obs is None:
        obs = _cached_obs(Y, Wx, K, IJ, dev, R)
    S3d = S3.to(device=dev, dtype=torch.float32)
    C3d = C_.reshape(1, R, K).to(device=dev, dtype=torch.float32)
    cost = _QmcNll.apply(S3d, C3d, obs, lik, _lib.QMC_ALGO_AUTO)
    return cost[0].to(dtype=torch.float32, device=S.device)


def make_obs(Y, Wx, K: Optional[int] = None, device=None, *, B: int = 1, R: Optional[int] = None,
             tiled: Optional[bool] = None, tile_warps: int = 8, lanes: Optional[bool] = None) -> ObsSet:
    """Build the compact observation set of one map (reference shapes ``[K,1,I,J]``) or of a batch
    ``[B,K,...]``.  ``tiled`` (default: batches of >= 64 maps) lays the entries out for the
    shared-memory kernels; it needs the rank R to size the tiles.  ``lanes`` (default: when
    32 <= K <= 256 and levels <= 254) additionally re-cuts every warp's stream into per-lane band walks."""
    dev = torch.device(device) if device is not None else (Y.device if Y.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    if K is None:
        K = Y.shape[0] if B == 1 else Y.shape[1]
    IJ = Y.numel() // (B * K)
    lanes_arg = lanes
    if tiled is None:
        tiled = B >= 64 and R is not None
    if tiled:
        if R is None:
            raise ValueError("tiled layout needs R")
        max_level = int(Y.max().item())                         # decides the width of the stream words
        if lanes is None:
            # every lane of the warp owns at least one band; level 255 is the lane streams' padding code: a
            # 256-level table (qmc/utils.py:24) goes through the tiled layout instead
            lanes = 32 <= K <= 256 and max_level <= 254
        n_sub, sub, tw = plan_tiles(IJ, K, R, tile_warps, lanes=lanes, max_level=max_level if lanes else None)
        bm = 0 if lanes else bank_mod_for_rank(R)
    else:
        n_sub, sub, tw, bm, lanes = 1, IJ, 0, 0, False
    auto_lanes = lanes_arg is None
    try:
        return build_obs(Y.to(dev), None if Wx is None else Wx.to(dev), K, IJ, B, n_sub=n_sub, sub_pixels=sub,
                         tile_warps=tw, bank_mod=bm, lanes=bool(lanes))
    except _lib.QmcError as e:
        # a stream too large for the lane-stream builder's shared memory (huge sub-tiles of a densely sampled
        # instance): the default selection falls back to the tiled layout, an explicit lanes=True fails loudly
        if not (lanes and auto_lanes and "builder's shared memory" in str(e)):
            raise
        n_sub, sub, tw = plan_tiles(IJ, K, R, tile_warps, lanes=False)
        return build_obs(Y.to(dev), None if Wx is None else Wx.to(dev), K, IJ, B, n_sub=n_sub, sub_pixels=sub,
                         tile_warps=tw, bank_mod=bank_mod_for_rank(R), lanes=False)
