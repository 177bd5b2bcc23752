"""Synthetic radio maps with the statistics of the reference's MATLAB generator.

A torch restatement of qmc/generate_map.m + qmc/Shadowing_data.m + qmc/ColumnNormalization.m
(there is no MATLAB/Octave in the image, and the benchmark needs thousands of maps on the GPU):

* R emitters at uniform random positions on the grid (generate_map.m:109);
* path loss ``min(1, (d/d0)^-alpha)``, d0 = 2, alpha = 2 + 0.5 U(0,1) (:90-91, :113);
* log-normal shadowing ``10^(z/10)``, z Gaussian with ``E z(x) z(x') = sigma^2 p^|x-x'|``,
  p = exp(-1/Xc), sigma = 4 dB, Xc = 90 by default (generate_test_data.m:10-11;
  Shadowing_data.m:14-22 draws it through the Cholesky factor of the IJ x IJ correlation);
* every SLF scaled to unit Frobenius norm (:118);
* spectra: three Gaussian bumps per emitter, amplitudes 0.5 + 1.5 U, widths 2 + 2 U (:54-71, basis
  'g'), columns scaled to unit 2-norm (:88; ColumnNormalization.m);
* ``T_true = sum_r S_r o c_r`` (:128-131).

Shadowing: the dense Cholesky is used up to ``chol_max_pixels`` (the factor is shared by all
maps and emitters because the grid is the same, so a batch costs one factorisation and one GEMM);
beyond that the IJ x IJ matrix is infeasible (512 x 512 -> 2.7e11 entries) and the field is drawn by
FFT on a zero-padded torus with the same exponential correlation function, which reproduces the
covariance up to the (small) negative-eigenvalue clipping of the circulant embedding.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass
class SyntheticMaps:
    S_true: torch.Tensor   # [B, R, I*J]
    C_true: torch.Tensor   # [B, R, K]
    I: int
    J: int

    def tensor(self) -> torch.Tensor:
        """T_true [B, K, I*J] (band-major unfolding, the reference's layout after c1:75-76)."""
        return torch.einsum("brp,brk->bkp", self.S_true, self.C_true)


def _shadow_cholesky(I, J, p, device):
    ys, xs = torch.meshgrid(torch.arange(I, device=device, dtype=torch.float64),
                            torch.arange(J, device=device, dtype=torch.float64), indexing="ij")
    pts = torch.stack([xs.reshape(-1), ys.reshape(-1)], 1)
    dist = torch.cdist(pts, pts)
    corr = torch.pow(torch.tensor(p, dtype=torch.float64, device=device), dist)
    return torch.linalg.cholesky(corr).to(torch.float32)


def _shadow_fft(n_fields, I, J, p, gen, device):
    PI, PJ = 2 * I, 2 * J
    dy = torch.minimum(torch.arange(PI, device=device), PI - torch.arange(PI, device=device)).double()
    dx = torch.minimum(torch.arange(PJ, device=device), PJ - torch.arange(PJ, device=device)).double()
    corr = torch.pow(torch.tensor(p, dtype=torch.float64, device=device), torch.sqrt(dy[:, None] ** 2 + dx[None, :] ** 2))
    lam = torch.fft.fft2(corr).real.clamp_min(0).sqrt().float()
    out = torch.empty(n_fields, I * J, device=device)
    chunk = max(1, (1 << 26) // (PI * PJ))
    for a in range(0, n_fields, chunk):
        b = min(a + chunk, n_fields)
        w = torch.randn(b - a, PI, PJ, device=device, generator=gen)
        z = torch.fft.ifft2(torch.fft.fft2(w) * lam).real
        out[a:b] = z[:, :I, :J].reshape(b - a, -1)
    return out


def generate_maps(B: int, I: int, J: int, K: int, R: int, *, shadow_sigma: float = 4.0, Xc: float = 90.0,
                  seed: int = 0, device="cuda", chol_max_pixels: int = 12000) -> SyntheticMaps:
    device = torch.device(device)
    gen = torch.Generator(device=device).manual_seed(seed)
    n = B * R
    IJ = I * J
    # ---- spatial loss fields ------------------------------------------------------------------
    ys, xs = torch.meshgrid(torch.arange(I, device=device, dtype=torch.float32),
                            torch.arange(J, device=device, dtype=torch.float32), indexing="ij")
    loc = torch.rand(n, 2, device=device, generator=gen) * torch.tensor([J - 1.0, I - 1.0], device=device)
    d = torch.sqrt((xs.reshape(1, -1) - loc[:, :1]) ** 2 + (ys.reshape(1, -1) - loc[:, 1:]) ** 2)
    alpha = 2.0 + 0.5 * torch.rand(n, 1, device=device, generator=gen)
    loss = torch.clamp(torch.pow(d / 2.0, -alpha), max=1.0)   # d = 0 -> inf -> clamped to 1
    p = math.exp(-1.0 / Xc)
    if shadow_sigma == 0:
        shadow = torch.zeros(n, IJ, device=device)
    elif IJ <= chol_max_pixels:
        L = _shadow_cholesky(I, J, p, device)
        shadow = shadow_sigma * (torch.randn(n, IJ, device=device, generator=gen) @ L.T)
    else:
        shadow = shadow_sigma * _shadow_fft(n, I, J, p, gen, device)
    S = loss * torch.pow(10.0, shadow / 10.0)
    S = S / torch.linalg.norm(S, dim=1, keepdim=True)
    # ---- power spectra --------------------------------------------------------------------------
    k = torch.arange(1, K + 1, device=device, dtype=torch.float32).reshape(1, 1, K)
    peaks = 3
    centre = 1 + (K - 2) * torch.rand(n, peaks, 1, device=device, generator=gen)
    amp = 0.5 + 1.5 * torch.rand(n, peaks, 1, device=device, generator=gen)
    width = 2.0 + 2.0 * torch.rand(n, peaks, 1, device=device, generator=gen)
    Cm = (amp * torch.exp(-(k - centre) ** 2 / (2 * width ** 2))).sum(1)
    Cm = Cm / torch.linalg.norm(Cm, dim=1, keepdim=True)
    return SyntheticMaps(S.reshape(B, R, IJ).contiguous(), Cm.reshape(B, R, K).contiguous(), I, J)


def bernoulli_mask(shape, f: float, seed: int, device="cuda") -> torch.Tensor:
    """Per-entry Bernoulli(f) sampling mask as qmc.ipynb c1:70-72 (float 0/1)."""
    gen = torch.Generator(device=torch.device(device)).manual_seed(seed)
    return torch.bernoulli(torch.full(shape, f, device=device), generator=gen)


def equal_mass_boundaries(samples: torch.Tensor, levels: int) -> torch.Tensor:
    """Boundaries that put (about) the same number of samples in every bin, first = min and
    last = max of the samples: the goal of qmc/utils.py:57-74 (`_find_boundaries`), via quantiles."""
    x = samples.reshape(-1).float()
    if x.numel() > 4_000_000:
        x = x[:: x.numel() // 4_000_000]          # deterministic subsample (every rank must get the same table)
    qs = torch.linspace(0, 1, levels + 1, device=x.device)
    return torch.quantile(x, qs).cpu()
