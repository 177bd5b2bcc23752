"""Deep-prior variant of the QMC-MLE solver: S = generator(Z), optimise C and the latents Z.

The reference's ``qmc/dip.py`` is an empty file; the deep-prior step is the Z-update inside
``qmc/qmc.ipynb`` cell 1 (:199-212), with the random-restart latent search of :167-196.  This module
restates both for B maps at once.  The generator runs in stock PyTorch (BASELINE config 5); the
quantized likelihood is the fused CUDA op, which accepts the generator output as a non-leaf and
hands gS back to autograd.

    C-step:  cost = nll(generator(Z).detach(), C) + lam_c ||C||_F + lam_s ||Z||_F   Adam(0.005) on C
             C[C < 0] = 0
    Z-step:  cost = nll(generator(Z), C) + lam_c ||C||_F + lam_s ||Z||_F            Adam(0.01) on Z

``Generator256`` below has the architecture (and state-dict key names) of the reference's
``deep_prior/networks/gan.py:83-126`` so that a reference checkpoint (``g_model_state_dict``) loads;
the trained weights are not shipped with the reference (.MISSING_LARGE_BLOBS), so tests and
benchmarks use a seeded random initialisation in eval() mode.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.nn as nn


class _Unflatten(nn.Module):
    def __init__(self, shape):
        super().__init__()
        self.shape = tuple(shape)

    def forward(self, x):
        return x.reshape(x.shape[0], *self.shape)


class Generator256(nn.Module):
    """z [N, 256] -> SLF [N, 1, 51, 51] in (0, 1).

    (256,1,1) -ConvT k3-> (128,3,3) -ConvT k4 s2 p1-> (64,6,6) -ConvT k4 s2 p1-> (32,12,12)
    -ConvT k4 s2-> (16,26,26) -ConvT k4 s2-> (2,54,54) -Conv k4-> (1,51,51) -Sigmoid; BatchNorm + ReLU
    after every transposed convolution."""

    # (in, out, kernel, stride, padding) of the transposed convolutions
    STACK = ((256, 128, 3, 1, 0), (128, 64, 4, 2, 1), (64, 32, 4, 2, 1), (32, 16, 4, 2, 0), (16, 2, 4, 2, 0))

    def __init__(self):
        super().__init__()
        layers = [_Unflatten((256, 1, 1))]
        for cin, cout, k, s, p in self.STACK:
            layers += [nn.ConvTranspose2d(cin, cout, k, s, p), nn.BatchNorm2d(cout), nn.ReLU(True)]
        layers += [nn.Conv2d(2, 1, 4, 1, 0), nn.Sigmoid()]
        self.main = nn.Sequential(*layers)

    def forward(self, z):
        return self.main(z)


@dataclass
class DipConfig:
    iters: int = 500
    lr_c: float = 0.005         # c1:126
    lr_z: float = 0.01          # c1:127
    lam_c: float = 100.0
    lam_s: float = 100.0        # multiplies ||Z||_F (c1:150, c1:208)
    z_dim: int = 256
    search_at: int = 1          # iteration at which the random-restart search runs (c1:167); -1 = never
    search_draws: int = 200     # c1:170
    search_refine: int = 200    # c1:186
    refine_scale: float = 0.2   # c1:187


def _frob(x):
    return torch.linalg.vector_norm(x.reshape(x.shape[0], -1), dim=1)


def latent_search(generator, Z, C, nll_fn, cfg: DipConfig, gen: Optional[torch.Generator] = None, candidates_fn=None,
                  max_images: int = 1 << 17):
    """Random-restart latent search (c1:167-196), vectorised over the B maps and over the draws: a phase of
    ``search_draws`` fresh latents, then a phase of ``search_refine`` perturbations (scale ``refine_scale``) of the
    incumbent of that phase; per map the candidate with the lowest NLL replaces the incumbent if it beats it (the
    earliest draw wins ties).  Forward-only evaluations.

    ``candidates_fn(S [D*B, R, IJ], C [B, R, K]) -> [D, B]`` scores all D draws of a phase with ONE batched generator
    forward and ONE likelihood launch (``fused.nll_candidates`` on a lane-stream observation set: the observation
    streams are shared by reference, not copied); phases whose D*B*R images exceed ``max_images`` are cut into the
    fewest equal chunks.  Without it the draws are scored one by one through ``nll_fn`` -- the same candidates in the
    same order, hence the same result (tests/test_gpu_parity.py).

    (The notebook's refinement loop evaluates the *previous* generator output instead of the perturbed latent --
    c1:187-188 reuses ``temp_out`` -- so it can never improve; here the perturbed latents are evaluated, which is
    what the comment above that loop says it does.  Refinement draws perturb the incumbent the phase started
    with, which is what makes them independent of each other and batchable.)"""
    B, R, zd = Z.shape
    with torch.no_grad():
        best = nll_fn(generator(Z.reshape(B * R, zd)).reshape(B, R, -1), C).to(torch.float64)
        for phase, n in (("draw", cfg.search_draws), ("refine", cfg.search_refine)):
            if n <= 0:
                continue
            noise = torch.randn((n,) + tuple(Z.shape), device=Z.device, generator=gen)      # [D, B, R, zd]
            cand = noise if phase == "draw" else Z.unsqueeze(0) + cfg.refine_scale * noise
            if candidates_fn is None:
                vals = torch.stack([nll_fn(generator(cand[d].reshape(B * R, zd)).reshape(B, R, -1), C).to(torch.float64)
                                    for d in range(n)])
            else:
                chunks = max(1, -(-(n * B * R) // max_images))
                per = -(-n // chunks)
                parts = []
                for d0 in range(0, n, per):
                    dc = min(per, n - d0)
                    parts.append(candidates_fn(generator(cand[d0: d0 + dc].reshape(dc * B * R, zd)).reshape(dc * B, R, -1), C))
                vals = torch.cat(parts)
            vmin, dmin = vals.min(dim=0)                                                       # first minimum per map
            better = vmin < best
            pick = cand[dmin, torch.arange(B, device=Z.device)]                                 # [B, R, zd]
            Z[better] = pick[better]
            best = torch.where(better, vmin, best)
    return Z, best


def solve_deep_prior(generator: nn.Module, Z0: torch.Tensor, C0: torch.Tensor, nll_fn: Callable,
                     cfg: DipConfig = DipConfig(), nmse_fn: Optional[Callable] = None, track_every: int = 0,
                     candidates_fn: Optional[Callable] = None):
    """``Z0 [B, R, z_dim]``, ``C0 [B, R, K]``; ``nll_fn(S [B,R,IJ], C) -> [B]``; ``candidates_fn``: the batched
    scorer of the latent search (:func:`latent_search`)."""
    B, R, zd = Z0.shape
    Z = Z0.detach().clone().requires_grad_(True)
    Cf = C0.detach().clone().requires_grad_(True)
    opt_c = torch.optim.Adam([Cf], lr=cfg.lr_c)
    opt_z = torch.optim.Adam([Z], lr=cfg.lr_z)
    generator.eval()

    def slf(z):
        return generator(z.reshape(B * R, zd)).reshape(B, R, -1)

    trace = []
    if Z.is_cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        S = slf(Z)
    for it in range(cfg.iters):
        opt_c.zero_grad(set_to_none=True)
        cost = nll_fn(S.detach(), Cf).to(torch.float32) + cfg.lam_c * _frob(Cf) + cfg.lam_s * _frob(Z.detach())
        cost.sum().backward()
        opt_c.step()
        with torch.no_grad():
            Cf.clamp_(min=0)
        if it == cfg.search_at:
            with torch.no_grad():
                latent_search(generator, Z.data, Cf.detach(), nll_fn, cfg, candidates_fn=candidates_fn)
        opt_z.zero_grad(set_to_none=True)
        S = slf(Z)
        cost = nll_fn(S, Cf.detach()).to(torch.float32) + cfg.lam_c * _frob(Cf.detach()) + cfg.lam_s * _frob(Z)
        cost.sum().backward()
        opt_z.step()
        if track_every and (it % track_every == 0 or it == cfg.iters - 1):
            with torch.no_grad():
                trace.append((cost.detach().clone(), nmse_fn(S.detach(), Cf.detach()) if nmse_fn else None))
    if Z.is_cuda:
        torch.cuda.synchronize()
    return dict(Z=Z.detach(), C=Cf.detach(), S=S.detach(), trace=trace, seconds=time.perf_counter() - t0,
                iterations=cfg.iters)
