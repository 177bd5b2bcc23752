"""B200-native quantized-matrix-completion likelihood path.

Importing the package loads libqmc_b200.so and fails loudly if it has not been built: there is no
CPU implementation and no fallback."""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)
from .fused import make_likelihood, make_obs, nll_candidates, nll_fwd_bwd, qmc_lsq, qmc_nll, qmc_nll_batched  # noqa: F401
from .obs import ObsSet, bank_mod_for_rank, build_obs, lane_streams, plan_tiles  # noqa: F401

__version__ = "0.1.0"
