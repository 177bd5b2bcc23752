"""Log-domain call surface, mirroring the reference's ``qmc/quantization_model_log.py`` (the file
qmc/qmc.ipynb actually imports, c1:10): the quantizer sees ``log(X + offset)``, the likelihood uses
the boundary table as it is (no -/+1e5 sentinels, quantization_model_log.py:32-34), plus
``NMSE_LOG`` and ``get_quantized_obs_from_ordinal``.  Everything else is shared with the linear
module.  GPU only."""
from __future__ import annotations

import torch

from .constants import LOG_OFFSET_7_ADJUSTED as LOG_OFFSET
from .fused import make_obs, qmc_lsq as _qmc_lsq, qmc_nll as _qmc_nll
from .quantization_model import (DeterministicCost, F_probit, F_sigmoid, NMSE, NegLikelihood, _noisy, _prob_probit,  # noqa: F401
                                 _to_dev, assign_levels, dither_probit, dither_sigmoid, get_tensor, nmse_factors, outer)

__all__ = ["quantize", "prob_probit", "get_quantized_obs_from_ordinal", "F_sigmoid", "dither_sigmoid", "F_probit",
           "dither_probit", "outer", "get_tensor", "NMSE", "NMSE_LOG", "NegLikelihood", "DeterministicCost",
           "qmc_nll", "qmc_lsq", "make_obs"]


def quantize(X, noise_std, bin_boundaries, offset=LOG_OFFSET):
    """Y = Q(log(X + offset) + E)  (quantization_model_log.py:9-21)."""
    return assign_levels(_noisy(X, noise_std, offset=offset), bin_boundaries)


def prob_probit(Y, X_hat, bin_boundaries, noise_std):
    """P(Y | X_hat) with the table used as given (quantization_model_log.py:23-41)."""
    return _prob_probit(Y, X_hat, bin_boundaries, noise_std, sentinels=False)


def get_quantized_obs_from_ordinal(Y, bin_boundaries, noise_std=None):
    """Bin mid-points (W+U)/2  (quantization_model_log.py:43-51)."""
    Yd = _to_dev(Y)
    bb = torch.as_tensor(bin_boundaries, dtype=torch.float32).detach().clone().to(Yd.device)
    return ((bb[Yd] + bb[Yd + 1]) / 2.0).to(Y.device)


def NMSE_LOG(T: torch.Tensor, T_target: torch.Tensor, offset: float):
    """NMSE after log(. + offset)  (quantization_model_log.py:104-111)."""
    a = torch.log(_to_dev(T) + offset)
    b = torch.log(_to_dev(T_target) + offset)
    return (torch.norm(a - b, "fro") / torch.norm(b, "fro")).to(T.device)


def qmc_nll(S, C, Y, Wx, bin_boundaries, noise_std, offset=LOG_OFFSET, **kw):
    """Fused log-domain NLL: ``-sum(Wx*log(prob_probit(Y, log(get_tensor(S,C)+offset), bb, std)))``
    (qmc.ipynb c1:145-150) in one launch; see :func:`..fused.qmc_nll`."""
    return _qmc_nll(S, C, Y, Wx, bin_boundaries, noise_std, offset=offset, **kw)


def qmc_lsq(S, C, Y, Wx, bin_boundaries, offset=LOG_OFFSET, **kw):
    """Log-domain masked least squares on the bin mid-points: ``torch.norm(Wx*(log(get_tensor(S,C)+offset) -
    get_quantized_obs_from_ordinal(Y, bb, std)))**2`` (qmc_dowjons.ipynb c1:84,108-112) in one launch."""
    return _qmc_lsq(S, C, Y, Wx, bin_boundaries, offset=offset, **kw)
