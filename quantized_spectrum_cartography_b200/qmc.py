"""QMC maximum-likelihood solver: alternating Adam on the factors with the fused likelihood op.

The reference's ``qmc/qmc.py`` is a 29-line setup stub; its actual solver lives in the notebook
``qmc/qmc.ipynb`` cell 1 (:136-229), and the pure low-rank variant (S optimised directly, no
generator) in ``backup/notebooks/onebit_lowrank.ipynb`` cell 1.  This module restates that loop for
B independent maps at once:

    repeat:
        C-step:  cost = nll(S.detach(), C) + lam_c * ||C||_F     -> Adam(lr_c) on C     (c1:140-154)
                 C[C < 0] = 0                                                           (c1:156-157)
        S-step:  cost = nll(S, C.detach()) + lam_s * ||S||_F     -> Adam(lr_s) on S     (c1:199-212,
                 S[S < 0] = 0                                       with S in place of generator(Z))
        track cost and NMSE(get_tensor(S, C), T_true)                                   (c1:214-217)

``nll`` is one fused CUDA launch per evaluation (:func:`..fused.qmc_nll_batched`); the norms are
non-squared Frobenius norms per map exactly as in the notebook (their sub-gradient at 0 is 0).  The
likelihood backend is injectable only so that the parity test can drive the same loop with the
checker; the default and only product backend is the CUDA kernel.

CLI:  python -m quantized_spectrum_cartography_b200.qmc --config cfg1|cfg2 [--maps B] [--iters N]
"""
from __future__ import annotations

import argparse
import time
from dataclasses import dataclass, field
from typing import Callable, Optional

import torch


@dataclass
class SolverConfig:
    iters: int = 500            # maxIter, c1:41
    lr_c: float = 0.005         # c1:126
    lr_s: float = 0.001
    lam_c: float = 100.0        # c1:51
    lam_s: float = 100.0        # c1:52
    c_inner: int = 1            # cinnerIter, c1:62
    s_inner: int = 1            # sinnerIter, c1:63
    project_c: bool = True      # c1:156-157
    project_s: bool = True
    track_every: int = 1        # the notebook evaluates NMSE every iteration (a host sync each time)
    cuda_graph: bool = False    # capture one alternating iteration (2 evaluations, 2 Adam steps, the
                                # projections) in a CUDA graph and replay it: removes the ~20 host-side
                                # launches per iteration that bound small batches


@dataclass
class SolverResult:
    S: torch.Tensor             # [B, R, IJ]
    C: torch.Tensor             # [B, R, K]
    cost: list = field(default_factory=list)   # per tracked iteration: [B] tensors (last S-step cost)
    nmse: list = field(default_factory=list)
    seconds: float = 0.0
    iterations: int = 0


def _frob(x: torch.Tensor) -> torch.Tensor:
    """Per-map non-squared Frobenius norm (torch.norm(., 'fro') of the notebook, one per map)."""
    return torch.linalg.vector_norm(x.reshape(x.shape[0], -1), dim=1)


def solve_lowrank(S0: torch.Tensor, C0: torch.Tensor, nll_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
                  cfg: SolverConfig = SolverConfig(), nmse_fn: Optional[Callable] = None) -> SolverResult:
    """``S0 [B,R,IJ]``, ``C0 [B,R,K]`` initial factors; ``nll_fn(S, C) -> [B]`` differentiable NLL
    (sum over maps is what gets back-propagated: maps are independent).  ``nmse_fn(S, C) -> [B]``."""
    S = S0.detach().clone().requires_grad_(True)
    Cf = C0.detach().clone().requires_grad_(True)
    capturable = bool(cfg.cuda_graph)
    opt_c = torch.optim.Adam([Cf], lr=cfg.lr_c, capturable=capturable)
    opt_s = torch.optim.Adam([S], lr=cfg.lr_s, capturable=capturable)
    res = SolverResult(S, Cf)
    if S.is_cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    def c_step():
        opt_c.zero_grad(set_to_none=False)
        cost = nll_fn(S.detach(), Cf).to(torch.float32) + cfg.lam_c * _frob(Cf)
        cost.sum().backward()
        opt_c.step()
        return cost

    def s_step():
        opt_s.zero_grad(set_to_none=False)
        cost = nll_fn(S, Cf.detach()).to(torch.float32) + cfg.lam_s * _frob(S)
        cost.sum().backward()
        opt_s.step()
        return cost

    def iteration():
        for _ in range(cfg.c_inner):
            c_step()
        if cfg.project_c:
            with torch.no_grad():
                Cf.clamp_(min=0)
        cost = None
        for _ in range(cfg.s_inner):
            cost = s_step()
        if cfg.project_s:
            with torch.no_grad():
                S.clamp_(min=0)
        return cost

    graph = None
    if cfg.cuda_graph:
        if not S.is_cuda:
            raise ValueError("cuda_graph needs CUDA tensors")
        # warm up on a side stream (allocator, Adam state, kernel attributes), restore, then capture
        S_keep, C_keep = S.detach().clone(), Cf.detach().clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                iteration()
        torch.cuda.current_stream().wait_stream(side)
        with torch.no_grad():
            S.copy_(S_keep)
            Cf.copy_(C_keep)
        for opt in (opt_c, opt_s):              # reset the moments the warm-up touched
            for st in opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_cost = iteration()
        torch.cuda.synchronize()
        t0 = time.perf_counter()

    for it in range(cfg.iters):
        if graph is not None:
            graph.replay()
            cost = static_cost
        else:
            cost = iteration()
        if cfg.track_every and (it % cfg.track_every == 0 or it == cfg.iters - 1):
            res.cost.append(cost.detach().clone())
            if nmse_fn is not None:
                with torch.no_grad():
                    res.nmse.append(nmse_fn(S, Cf))
    if S.is_cuda:
        torch.cuda.synchronize()
    res.seconds = time.perf_counter() - t0
    res.iterations = cfg.iters
    res.S, res.C = S.detach(), Cf.detach()
    return res


def solve_lowrank_fused(S0: torch.Tensor, C0: torch.Tensor, obs, lik, cfg: SolverConfig = SolverConfig(),
                        nmse_fn: Optional[Callable] = None, betas=(0.9, 0.999), eps: float = 1e-8,
                        fuse_s_step: bool = True) -> SolverResult:
    """The iteration of :func:`solve_lowrank` without autograd and without torch.optim: every step is
    one fused likelihood evaluation (``fused.nll_fwd_bwd``) followed by one fused update
    (``qmc_adam_frob_project``: regulariser gradient, Adam, projection, next squared norm) -- four
    kernels per alternating iteration instead of ~60.  Same arithmetic as the torch path to fp32
    rounding (``tests/test_gpu_parity.py::test_fused_solver_matches_torch_solver``).

    ``S0`` may be emitter-major contiguous or pixel-major storage viewed as ``[B,R,IJ]``; the result
    comes back in the same layout (internally S lives pixel-major when R is a multiple of 4).  With a
    lane-stream observation set (one tile per map) the S-step is a single launch
    (``qmc_solver_s_step_fused``: the update is applied from the gS tile in shared memory);
    ``fuse_s_step=False`` keeps evaluation and update apart.  ``cfg.cuda_graph`` captures one iteration (the Adam step number then lives on the device)."""
    import ctypes as C
    from . import _lib
    from ._lib import check, lib
    from .fused import nll_fwd_bwd
    if not S0.is_cuda:
        raise ValueError("solve_lowrank_fused needs CUDA tensors: there is no CPU path")
    dev = S0.device
    B, R, IJ = S0.shape
    if R % 4 == 0:
        # pixel-major storage [B][IJ][R] viewed as [B,R,IJ]: the kernels stage and write whole tiles with
        # TMA bulk copies; the caller gets S back in the layout it passed
        S = torch.empty(B, IJ, R, dtype=torch.float32, device=dev).transpose(1, 2)
    else:
        S = torch.empty_strided(S0.shape, S0.stride(), dtype=torch.float32, device=dev)
    S.copy_(S0.detach())
    Cf = C0.detach().to(torch.float32).contiguous().clone()
    K = Cf.shape[2]
    nS, nC = R * IJ, R * K
    dense_s = S.is_contiguous() or (S.stride(1) == 1 and S.stride(2) == R and S.stride(0) == nS)
    if not dense_s:
        raise ValueError("S must be a dense [B,R,IJ] tensor (emitter-major or pixel-major storage)")
    mS, vS = torch.zeros_like(S), torch.zeros_like(S)
    mC, vC = torch.zeros_like(Cf), torch.zeros_like(Cf)
    if mS.stride() != S.stride():
        raise RuntimeError("moment buffers did not keep the layout of S")
    gS = torch.empty_strided(S.shape, S.stride(), dtype=torch.float32, device=dev)
    gC = torch.empty_like(Cf)
    nll = torch.empty(B, dtype=torch.float64, device=dev)
    ssS, ssS_next = (torch.empty(B, dtype=torch.float64, device=dev) for _ in range(2))   # ||S_b||_F^2
    ssC, ssC_next = (torch.empty(B, dtype=torch.float64, device=dev) for _ in range(2))
    view = obs.view()
    fused_s = bool(fuse_s_step and obs.lanes and R in (4, 8, 16, 32) and obs.n_sub == obs.tile_warps and not lik.flags & _lib.QMC_FORWARD_ONLY
                   and S.stride(1) == 1 and S.stride(2) == R)
    ctr = torch.zeros(2, dtype=torch.int32, device=dev)          # Adam steps taken on C, on S
    ctr_c, ctr_s = ctr.data_ptr(), ctr.data_ptr() + 4
    res = SolverResult(S, Cf)

    def stream():
        return torch.cuda.current_stream().cuda_stream

    def update(p, g, m, v, n, ss, ss_next, lr, lam, project, ctr_ptr):
        check(lib.qmc_adam_frob_project(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), B, n,
                                        ss.data_ptr(), ss_next.data_ptr(), lr, betas[0], betas[1], eps, lam,
                                        int(project), 1, ctr_ptr, stream()))
        check(lib.qmc_counter_add(ctr_ptr, 1, stream()))
        ss.copy_(ss_next)

    def iteration(want_cost: bool = True):
        for j in range(cfg.c_inner):
            nll_fwd_bwd(S, Cf, obs, lik, out=(nll, gS, gC), skip_gs=True)
            update(Cf, gC, mC, vC, nC, ssC, ssC_next, cfg.lr_c, cfg.lam_c, cfg.project_c and j == cfg.c_inner - 1, ctr_c)
        cost = None
        for j in range(cfg.s_inner):
            project = cfg.project_s and j == cfg.s_inner - 1
            if fused_s:
                # evaluation + update of S in one launch: the gS tile never leaves shared memory
                check(lib.qmc_solver_s_step_fused(
                    S.data_ptr(), S.stride(0), S.stride(1), S.stride(2), Cf.data_ptr(), C.byref(view), C.byref(lik), B, IJ,
                    K, R, obs.tile_warps, nll.data_ptr(), mS.data_ptr(), vS.data_ptr(), ssS.data_ptr(), ssS_next.data_ptr(),
                    cfg.lr_s, betas[0], betas[1], eps, cfg.lam_s, int(project), 1, ctr_s, stream()))
                if want_cost and j == cfg.s_inner - 1:
                    cost = nll.to(torch.float32) + cfg.lam_s * ssS.sqrt().to(torch.float32)
                check(lib.qmc_counter_add(ctr_s, 1, stream()))
                ssS.copy_(ssS_next)
                continue
            nll_fwd_bwd(S, Cf, obs, lik, out=(nll, gS, gC), skip_gc=True)
            if want_cost and j == cfg.s_inner - 1:   # the tracked quantity of the notebook (c1:209,214)
                cost = nll.to(torch.float32) + cfg.lam_s * ssS.sqrt().to(torch.float32)
            update(S, gS, mS, vS, nS, ssS, ssS_next, cfg.lr_s, cfg.lam_s, cfg.project_s and j == cfg.s_inner - 1, ctr_s)
        return cost

    with torch.cuda.device(dev):
        check(lib.qmc_sumsq_per_map(S.data_ptr(), B, nS, ssS.data_ptr(), stream()))
        check(lib.qmc_sumsq_per_map(Cf.data_ptr(), B, nC, ssC.data_ptr(), stream()))
        graph = None
        if cfg.cuda_graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):      # warm-up launch outside the capture (kernel attributes)
                nll_fwd_bwd(S, Cf, obs, lik, out=(nll, gS, gC))
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_cost = iteration(bool(cfg.track_every))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for it in range(cfg.iters):
            tracked = bool(cfg.track_every) and (it % cfg.track_every == 0 or it == cfg.iters - 1)
            if graph is not None:
                graph.replay()
                cost = static_cost
            else:
                cost = iteration(tracked)
            if cfg.track_every and (it % cfg.track_every == 0 or it == cfg.iters - 1):
                res.cost.append(cost.detach().clone())
                if nmse_fn is not None:
                    res.nmse.append(nmse_fn(S, Cf))
        torch.cuda.synchronize()
        res.seconds = time.perf_counter() - t0
    res.iterations = cfg.iters
    S_out = torch.empty_strided(S0.shape, S0.stride(), dtype=torch.float32, device=dev)
    S_out.copy_(S)
    res.S, res.C = S_out, Cf
    return res


def cuda_nll_fn(obs, lik):
    """The product backend: fused CUDA likelihood of a batch of maps."""
    from .fused import qmc_nll_batched
    return lambda S, C: qmc_nll_batched(S, C, obs, lik)


def cuda_nmse_fn(T_true: torch.Tensor, offset=None):
    """NMSE of S*C^T against ``T_true [B,K,IJ]`` without materialising the reconstruction."""
    from .quantization_model import nmse_factors

    def fn(S, C):
        return nmse_factors(S, C, T_true, offset=offset).reshape(-1)
    return fn


# ---- synthetic configurations of BASELINE.json ---------------------------------------------------
CONFIGS = {
    # one-bit, linear domain (quantization_model.py semantics: +-1e5 sentinels)
    "cfg1": dict(I=51, J=51, K=64, R=4, f=0.10, levels=2, log_domain=False),
    # 3-bit / 8 levels (9 boundaries), log domain (quantization_model_log.py semantics)
    "cfg2": dict(I=101, J=101, K=128, R=8, f=0.20, levels=8, log_domain=True),
    # one large, densely sampled instance (the tcgen05 path; sharded by pixel blocks over several GPUs)
    "cfg4": dict(I=512, J=512, K=256, R=16, f=0.50, levels=8, log_domain=True),
}


def synth_problem(name: str, B: int, device, seed: int = 0):
    """Synthetic instance(s) of a BASELINE config: maps, observations, likelihood, start point."""
    from . import synth
    from .fused import make_likelihood, make_obs
    from .quantization_model import assign_levels
    c = CONFIGS[name]
    I, J, K, R = c["I"], c["J"], c["K"], c["R"]
    maps = synth.generate_maps(B, I, J, K, R, seed=seed, device=device)
    T = maps.tensor()
    gen = torch.Generator(device=device).manual_seed(seed + 1)
    if c["log_domain"]:
        offset = float(T.median()) * 0.1 + 1e-12
        X = torch.log(T + offset)
        bb = synth.equal_mass_boundaries(X, c["levels"])
        sigma = float((bb[1:] - bb[:-1]).min()) * 2.0
    else:
        offset = None
        X = T
        if c["levels"] == 2:
            thr = float(T.median())
            bb = torch.tensor([0.0, thr, 1.0])
            sigma = thr
        else:
            bb = synth.equal_mass_boundaries(X, c["levels"])
            sigma = float((bb[1:] - bb[:-1]).min()) * 2.0
    noisy = X + sigma * torch.randn(X.shape, device=device, generator=gen)
    Y = assign_levels(noisy, bb)
    Wx = torch.bernoulli(torch.full(T.shape, c["f"], device=device), generator=gen)
    lik = make_likelihood(bb, sigma, offset=offset)
    obs = make_obs(Y, Wx, K, device, B=B, R=R)
    return dict(maps=maps, T=T, Y=Y, Wx=Wx, bb=bb, sigma=sigma, offset=offset, lik=lik, obs=obs, cfg=c)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg1", choices=sorted(CONFIGS))
    ap.add_argument("--maps", type=int, default=1)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--graph", action="store_true", help="replay one captured iteration (CUDA graph)")
    ap.add_argument("--torch-optim", action="store_true", help="autograd + torch.optim.Adam instead of the fused update kernel")
    args = ap.parse_args(argv)
    dev = torch.device("cuda", torch.cuda.current_device())
    pb = synth_problem(args.config, args.maps, dev, args.seed)
    maps = pb["maps"]
    cfg = SolverConfig(iters=args.iters, lam_c=1.0, lam_s=1.0, track_every=max(1, args.iters // 10), cuda_graph=args.graph)
    if args.torch_optim:
        res = solve_lowrank(0.7 * maps.S_true, 0.9 * maps.C_true, cuda_nll_fn(pb["obs"], pb["lik"]), cfg,
                            cuda_nmse_fn(pb["T"]))
    else:
        res = solve_lowrank_fused(0.7 * maps.S_true, 0.9 * maps.C_true, pb["obs"], pb["lik"], cfg, cuda_nmse_fn(pb["T"]))
    for i, (c, n) in enumerate(zip(res.cost, res.nmse)):
        print(f"track {i}: cost[0]={c[0].item():.4f} nmse[0]={n.reshape(-1)[0].item():.5f}")
    print(f"{res.iterations} iterations x {args.maps} maps in {res.seconds:.3f} s -> "
          f"{res.iterations / res.seconds:.1f} solver iterations/s, "
          f"{res.iterations * 2 * pb['obs'].nobs / res.seconds:.3e} observed entries/s")


if __name__ == "__main__":
    main()
