"""CPU oracle for the QMC quantized-likelihood hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing under ``quantized_spectrum_cartography_b200/``
imports it, and the product has no CPU fallback.

It restates, in CPU PyTorch (fp32, same op sequence as the reference so that results are
comparable bit-for-bit where the ops are IEEE) and in NumPy float64 (a second, independent
statement of the same formulas used to calibrate tolerances), the algorithm of

    /root/reference/qmc/quantization_model.py       (linear domain, +-1e5 sentinels)
    /root/reference/qmc/quantization_model_log.py   (log domain, no sentinels)
    /root/reference/qmc/qmc.ipynb  cell 1           (the 4-line masked-NLL idiom and its caller)

Each function cites the reference ``file:line`` it follows.

Parity pin: the reference ships no tests for this path, so the oracle
is pinned against outputs of the reference itself, executed in the build container by
``tests/golden/make_golden.py`` (which imports ``/root/reference/qmc`` verbatim) and committed
as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every function here against
those vectors.
"""
from __future__ import annotations

import math

import numpy as np
import torch

# The reference writes sqrt(2) as this truncated literal (quantization_model.py:61).
REF_SQRT2 = 1.414213
# Sentinel magnitude the linear-domain file substitutes for the outer boundaries
# (quantization_model.py:32-33).
REF_SENTINEL = 100000.0


# ----------------------------------------------------------------------------------------
# a8  quantizer
# ----------------------------------------------------------------------------------------
def assign_levels(noisy: torch.Tensor, bin_boundaries: torch.Tensor) -> torch.Tensor:
    """Level index of every entry of an already-noisy tensor.

    Follows quantization_model.py:14-20 (identical in quantization_model_log.py:15-21):
    n boundaries give n-1 levels 0..n-2; boundary 0 is ignored (anything <= bb[1], NaN
    included, is level 0); cells are left-open/right-closed; the last boundary is replaced by
    +inf so the top level is unbounded above.
    """
    bb = bin_boundaries.clone()
    bb[-1] = float("inf")
    levels = torch.zeros(noisy.shape)
    for i in range(1, bb.numel() - 1):
        inside = (bb[i] < noisy) & (noisy <= bb[i + 1])
        levels[inside] = i
    return levels.long()


def assign_levels_closed_form(noisy: np.ndarray, bin_boundaries: np.ndarray) -> np.ndarray:
    """Closed form of :func:`assign_levels`.

    Replays the reference's overwrite loop cell by cell, so it agrees with it for any table
    (sorted or not): the last matching cell wins.  For a sorted table this is
    ``#{i in [1, n-2] : noisy > bb[i]}``.
    """
    bb = np.asarray(bin_boundaries, dtype=np.float32).copy()
    x = np.asarray(noisy, dtype=np.float32)
    n = bb.shape[0]
    out = np.zeros(x.shape, dtype=np.int64)
    upper = bb.copy()
    upper[-1] = np.inf
    for i in range(1, n - 1):
        with np.errstate(invalid="ignore"):
            hit = (x > bb[i]) & (x <= upper[i + 1])
        out[hit] = i
    return out


def noisy_signal(X: torch.Tensor, noise: torch.Tensor, noise_std, offset=None) -> torch.Tensor:
    """``X + noise*std`` (quantization_model.py:13) or ``log(X+offset) + noise*std``
    (quantization_model_log.py:14), with the reference's operation order (multiply, then add;
    no fused multiply-add on the CPU)."""
    base = X if offset is None else torch.log(X + offset)
    return base + noise * noise_std


def quantize(X, noise_std, bin_boundaries, offset=None, noise=None):
    """Y = Q(X + E).  ``noise`` (standard normal, same shape) may be supplied so that two
    back ends can be compared on identical noise; otherwise it is drawn exactly where the
    reference draws it (``torch.randn(X.shape)``, quantization_model.py:13)."""
    if noise is None:
        noise = torch.randn(X.shape)
    return assign_levels(noisy_signal(X, noise, noise_std, offset), bin_boundaries)


# ----------------------------------------------------------------------------------------
# a1/a2  low-rank tensor assembly  X = sum_r S_r o c_r
# ----------------------------------------------------------------------------------------
def outer_band_loop(mat: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    """Band-by-band outer product into a fresh fp32 buffer (quantization_model.py:70-77).

    Kept in the reference's shape on purpose -- one slice write per band -- because this loop
    (and the K in-place-copy autograd nodes it creates) is where the reference's time goes
    (SURVEY.md section 6); the CPU baseline must pay the same cost."""
    bands = vec.shape[0]
    cube = torch.zeros((bands,) + tuple(mat.shape), dtype=torch.float32)
    for b in range(bands):
        cube[b] = mat * vec[b]
    return cube


def get_tensor(S: torch.Tensor, C: torch.Tensor) -> torch.Tensor:
    """sum_r outer(S[r,0], C[r])  -> [K, I, J]   (quantization_model.py:79-86)."""
    total = 0
    for r in range(C.shape[0]):
        total = total + outer_band_loop(S[r, 0], C[r])
    return total


def get_tensor_vectorised(S: torch.Tensor, C: torch.Tensor) -> torch.Tensor:
    """Same values as :func:`get_tensor` (same fp32 multiply/add order over r), without the
    Python loop over bands.  This is the 'fair CPU' statement used for the second baseline
    row."""
    R = C.shape[0]
    S3 = S.reshape(R, *S.shape[-2:])
    total = S3[0].unsqueeze(0) * C[0].reshape(-1, 1, 1)
    for r in range(1, R):
        total = total + S3[r].unsqueeze(0) * C[r].reshape(-1, 1, 1)
    return total


# ----------------------------------------------------------------------------------------
# a4/a5  probit CDF and bin probability
# ----------------------------------------------------------------------------------------
def F_probit(y: torch.Tensor, std) -> torch.Tensor:
    """0.5*(1 + erf(y/(std*1.414213)))   (quantization_model.py:57-61)."""
    return (1 / 2) * (1 + torch.erf(y / (std * REF_SQRT2)))


def F_sigmoid(y: torch.Tensor) -> torch.Tensor:
    """1/(1+exp(-y))   (quantization_model.py:43-47)."""
    return 1 / (1 + torch.exp(-y))


def effective_boundaries(bin_boundaries: torch.Tensor, sentinels: bool) -> torch.Tensor:
    """Boundary table as the likelihood sees it: the linear-domain file overwrites the two
    outer boundaries with -/+1e5 (quantization_model.py:31-33); the log-domain file has those
    two lines commented out (quantization_model_log.py:32-34)."""
    bb = bin_boundaries.clone()
    if sentinels:
        bb[0] = -REF_SENTINEL
        bb[-1] = REF_SENTINEL
    return bb


def prob_probit(Y, X_hat, bin_boundaries, noise_std, sentinels=True):
    """P(Y | X_hat) = F(U - X_hat) - F(W - X_hat), W = bb[Y], U = bb[Y+1]
    (quantization_model.py:22-39 with ``sentinels=True``; quantization_model_log.py:23-41
    with ``sentinels=False``)."""
    bb = effective_boundaries(bin_boundaries, sentinels)
    lower = bb[Y]
    upper = bb[Y + 1]
    return F_probit(upper - X_hat, noise_std) - F_probit(lower - X_hat, noise_std)


def get_quantized_obs_from_ordinal(Y, bin_boundaries, noise_std=None):
    """Bin mid-points (W+U)/2   (quantization_model_log.py:43-51)."""
    bb = bin_boundaries.clone()
    return (bb[Y] + bb[Y + 1]) / 2.0


# ----------------------------------------------------------------------------------------
# a3/a6/a7  the masked negative log-likelihood idiom and its gradients
# ----------------------------------------------------------------------------------------
def masked_nll(S, C, Y, Wx, bin_boundaries, noise_std, offset=None, sentinels=True,
               vectorised=False):
    """The 4-line idiom of qmc.ipynb c1:145-150:

        T_hat = get_tensor(S, C).unsqueeze(1); [T_hat = log(T_hat + offset)]
        nll   = -sum(Wx * log(prob_probit(Y, T_hat, bb, std)))

    ``offset is None`` is the linear-domain model (no log link); a float applies the log
    link of c1:149.  Mask is applied by multiplication exactly as the reference does, so an
    unobserved entry with P == 0 yields NaN here too."""
    T_hat = (get_tensor_vectorised if vectorised else get_tensor)(S, C).unsqueeze(1)
    if offset is not None:
        T_hat = torch.log(T_hat + offset)
    P = prob_probit(Y, T_hat, bin_boundaries, noise_std, sentinels=sentinels)
    return -torch.sum(Wx * torch.log(P))


def nll_and_grads(S, C, Y, Wx, bin_boundaries, noise_std, offset=None, sentinels=True,
                  vectorised=False):
    """fp32 NLL and its autograd gradients w.r.t. S and C, i.e. ``cost.backward()`` of
    qmc.ipynb c1:153 without the regularisers."""
    S = S.detach().clone().requires_grad_(True)
    C = C.detach().clone().requires_grad_(True)
    nll = masked_nll(S, C, Y, Wx, bin_boundaries, noise_std, offset, sentinels, vectorised)
    nll.backward()
    return nll.detach(), S.grad.detach(), C.grad.detach()



# ----------------------------------------------------------------------------------------
# SURVEY 8(f)(4)  masked least-squares baseline on the de-quantised mid-points
# ----------------------------------------------------------------------------------------
def masked_lsq(S, C, Y, Wx, bin_boundaries, offset=None, vectorised=False):
    """The baseline cost of qmc_dowjons.ipynb c1:84,108-112 without the regularisers:

        Obs   = get_quantized_obs_from_ordinal(Y, bb, std)          # bin mid-points
        T_hat = get_tensor(S, C).unsqueeze(1); [T_hat = log(T_hat + offset)]
        cost  = torch.norm(Wx * (T_hat - Obs)) ** 2

    The boundary table is used as it is (quantization_model_log.py:44-46 leaves the sentinel
    lines commented out)."""
    Obs = get_quantized_obs_from_ordinal(Y, bin_boundaries)
    T_hat = (get_tensor_vectorised if vectorised else get_tensor)(S, C).unsqueeze(1)
    if offset is not None:
        T_hat = torch.log(T_hat + offset)
    return torch.norm(Wx * (T_hat - Obs)) ** 2


def lsq_and_grads(S, C, Y, Wx, bin_boundaries, offset=None, vectorised=False):
    """fp32 least-squares cost and its autograd gradients w.r.t. S and C (``cost.backward()`` of
    qmc_dowjons.ipynb c1:114,132 without the regularisers)."""
    S = S.detach().clone().requires_grad_(True)
    C = C.detach().clone().requires_grad_(True)
    cost = masked_lsq(S, C, Y, Wx, bin_boundaries, offset, vectorised)
    cost.backward()
    return cost.detach(), S.grad.detach(), C.grad.detach()


def lsq_and_grads_fp64(S, C, Y, Wx, bin_boundaries, offset=None):
    """float64 statement of the same cost with analytic gradients (observed entries selected, not
    multiplied): d cost/d x = 2 (x - mid).  Returns (cost, gS [R,1,I,J], gC [R,K])."""
    S64 = np.asarray(S.detach().cpu().numpy(), dtype=np.float64)
    C64 = np.asarray(C.detach().cpu().numpy(), dtype=np.float64)
    R, K = C64.shape
    Smat = S64.reshape(R, -1)
    Yk = np.asarray(Y.detach().cpu().numpy()).reshape(K, -1)
    obs = np.asarray(Wx.detach().cpu().numpy()).reshape(K, -1) != 0
    bb = np.asarray(bin_boundaries.detach().cpu().numpy(), dtype=np.float64)
    T = C64.T @ Smat
    X = np.log(T + float(offset)) if offset is not None else T
    d = np.where(obs, X - 0.5 * (bb[Yk] + bb[Yk + 1]), 0.0)
    gT = 2.0 * d / (T + float(offset)) if offset is not None else 2.0 * d
    return float(np.sum(d * d)), (C64 @ gT).reshape(S64.shape), Smat @ gT.T


# ----------------------------------------------------------------------------------------
# logistic noise model: prob_probit with F_sigmoid in place of F_probit
# ----------------------------------------------------------------------------------------
def prob_sigmoid(Y, X_hat, bin_boundaries, scale=1.0, sentinels=True):
    """``P = F_sigmoid((U - X)/s) - F_sigmoid((W - X)/s)``: the body of prob_probit
    (quantization_model.py:31-38) with the reference's logistic CDF (F_sigmoid, :43-47) in place of
    F_probit.  The reference ships F_sigmoid (and uses it in NegLikelihood(probit=False), :107-110) but no
    multi-level logistic function, so this composition is pinned through F_sigmoid's golden vectors only."""
    bb = effective_boundaries(bin_boundaries, sentinels)
    return F_sigmoid((bb[Y + 1] - X_hat) / scale) - F_sigmoid((bb[Y] - X_hat) / scale)


def logistic_nll_and_grads_fp64(S, C, Y, Wx, bin_boundaries, scale=1.0, offset=None, sentinels=True):
    """float64 masked NLL of the logistic model and analytic gradients, observed entries only, written
    without cancellation: log P = log(-expm1(-(zu-zl))) - softplus(-zu) - softplus(zl)."""
    S64 = np.asarray(S.detach().cpu().numpy(), dtype=np.float64)
    C64 = np.asarray(C.detach().cpu().numpy(), dtype=np.float64)
    R, K = C64.shape
    Smat = S64.reshape(R, -1)
    Yk = np.asarray(Y.detach().cpu().numpy()).reshape(K, -1)
    obs = np.asarray(Wx.detach().cpu().numpy()).reshape(K, -1) != 0
    bb = np.asarray(bin_boundaries.detach().cpu().numpy(), dtype=np.float64).copy()
    if sentinels:
        bb[0], bb[-1] = -REF_SENTINEL, REF_SENTINEL
    s = float(np.float32(scale))
    T = C64.T @ Smat
    X = np.log(T + float(offset)) if offset is not None else T
    zu, zl = (bb[Yk + 1] - X) / s, (bb[Yk] - X) / s

    def softplus(t):
        return np.maximum(t, 0.0) + np.log1p(np.exp(-np.abs(t)))

    def sigmoid(t):
        e = np.exp(-np.abs(t))
        return np.where(t >= 0, 1.0 / (1.0 + e), e / (1.0 + e))

    with np.errstate(all="ignore"):
        logP = np.log(-np.expm1(-(zu - zl))) - softplus(-zu) - softplus(zl)
        gX = np.where(obs, (sigmoid(-zu) - sigmoid(zl)) / s, 0.0)
    nll = -float(np.sum(logP[obs]))
    gT = gX / (T + float(offset)) if offset is not None else gX
    return nll, (C64 @ gT).reshape(S64.shape), Smat @ gT.T


def stable_logP_fp64(zl: np.ndarray, zu: np.ndarray):
    """float64 log P and (exp(-zu^2) - exp(-zl^2)) / P for P = 0.5*(erf(zu) - erf(zl)), zl < zu,
    evaluated without cancellation or underflow: both bounds in the right tail -> scaled
    complementary error functions relative to the nearer bound; left tail -> mirror image;
    bounds straddling zero -> the plain erf difference (no cancellation there)."""
    from scipy.special import erf, erfcx
    zl = np.asarray(zl, dtype=np.float64)
    zu = np.asarray(zu, dtype=np.float64)
    logP = np.empty_like(zl)
    ratio = np.empty_like(zl)
    right = zl >= 0
    left = (zu <= 0) & ~right
    mid = ~(right | left)
    with np.errstate(all="ignore"):
        for sel, near, far, sign in ((right, zl, zu, -1.0), (left, -zu, -zl, +1.0)):
            n, f = near[sel], far[sel]
            D = np.exp(n * n - f * f)                      # <= 1
            tail_far = np.where(D > 0, D * erfcx(f), 0.0)  # 0 * inf guard for infinite bounds
            core = 0.5 * (erfcx(n) - tail_far)
            logP[sel] = -n * n + np.log(core)
            ratio[sel] = sign * (1.0 - D) / core
        P = 0.5 * (erf(zu[mid]) - erf(zl[mid]))
        logP[mid] = np.log(P)
        ratio[mid] = (np.exp(-zu[mid] ** 2) - np.exp(-zl[mid] ** 2)) / P
    return logP, ratio


def nll_and_grads_fp64(S, C, Y, Wx, bin_boundaries, noise_std, offset=None, sentinels=True):
    """Independent float64 statement of the same formulas, analytic
    gradients, observed entries only (the mask *selects* instead of multiplying, so
    unobserved P == 0 entries do not poison the sum), tail-stable.  Used to calibrate
    tolerances: where the fp32 reference is accurate the two agree to ~1e-7, and it stays
    finite where the reference returns inf/NaN.

    Returns (nll, gS [R,1,I,J], gC [R,K], Pmin over observed entries)."""
    S64 = np.asarray(S.detach().cpu().numpy(), dtype=np.float64)
    C64 = np.asarray(C.detach().cpu().numpy(), dtype=np.float64)
    R, K = C64.shape
    Smat = S64.reshape(R, -1)                      # [R, IJ]
    Yk = np.asarray(Y.detach().cpu().numpy()).reshape(K, -1)
    Wk = np.asarray(Wx.detach().cpu().numpy(), dtype=np.float64).reshape(K, -1)
    bb = np.asarray(bin_boundaries.detach().cpu().numpy(), dtype=np.float64).copy()
    if sentinels:
        bb[0], bb[-1] = -REF_SENTINEL, REF_SENTINEL
    # the reference forms std*1.414213 in Python double, then divides an fp32 tensor by it
    a = float(np.float32(float(noise_std) * REF_SQRT2))
    T = C64.T @ Smat                               # [K, IJ]
    X = np.log(T + float(offset)) if offset is not None else T
    obs = Wk != 0
    zu = (bb[Yk + 1] - X)[obs] / a
    zl = (bb[Yk] - X)[obs] / a
    logP, ratio = stable_logP_fp64(zl, zu)
    nll = -np.sum(logP)
    gX = np.zeros_like(X)
    gX[obs] = ratio / (a * math.sqrt(math.pi))     # d nll / d x
    gT = gX / (T + float(offset)) if offset is not None else gX
    gS = (C64 @ gT).reshape(S64.shape)             # [R,K]@[K,IJ]
    gC = Smat @ gT.T                               # [R,IJ]@[IJ,K]
    pmin = float(np.exp(logP.min())) if obs.any() else float("nan")
    return float(nll), gS, gC, pmin


# ----------------------------------------------------------------------------------------
# a9  one-bit BCE formulation, a10 metrics
# ----------------------------------------------------------------------------------------
def neg_likelihood_bce(T_sample, T_target, mean, std=None, probit=True):
    """NegLikelihood.forward (quantization_model.py:97-113): BCELoss (mean reduction, log
    clamped at -100, no mask) of F_probit(T-mean, std) or F_sigmoid(T-mean)."""
    p = F_probit(T_sample - mean, std) if probit else F_sigmoid(T_sample - mean)
    return torch.nn.functional.binary_cross_entropy(p, T_target)


def deterministic_cost(S, C, T_target, mean=0.0, lambda_reg=0.001):
    """DeterministicCost.forward (quantization_model.py:115-129)."""
    T_hat = get_tensor(S, C) - mean
    return -lambda_reg * (T_hat * T_target).sum() + torch.norm(T_hat, "fro")


def NMSE(T, T_target):
    """||T - T*||_F / ||T*||_F  -- not squared (quantization_model.py:88-92)."""
    return torch.norm(T - T_target, "fro") / torch.norm(T_target, "fro")


def NMSE_LOG(T, T_target, offset):
    """NMSE after log(. + offset)   (quantization_model_log.py:104-111)."""
    a = torch.log(T + offset)
    b = torch.log(T_target + offset)
    return torch.norm(a - b, "fro") / torch.norm(b, "fro")


# ----------------------------------------------------------------------------------------
# compact observation format (what the CUDA path consumes), stated on the CPU
# ----------------------------------------------------------------------------------------
def observed_entries(Y, Wx):
    """(linear index k*IJ + p, level) of every observed entry in the reference's [K,1,I,J]
    layout, in increasing index order.  This is the information content of (Y, Wx) that the
    likelihood actually uses (qmc.ipynb c1:150: only Wx != 0 entries contribute)."""
    w = np.asarray(Wx.detach().cpu().numpy()).reshape(-1)
    y = np.asarray(Y.detach().cpu().numpy()).reshape(-1)
    idx = np.flatnonzero(w != 0)
    return idx.astype(np.int64), y[idx].astype(np.int64)
