#!/usr/bin/env python
"""cfg5 of BASELINE.json: deep-prior variant, 256 independent 51x51x64 one-bit maps (R = 4 emitters each
-> generator batch 1024, SURVEY 8 ambiguity note), generator forward/backward in stock PyTorch, the
lane-stream likelihood kernel as the loss.  Prints one JSON line: deep-prior iterations/s and where the
time of one iteration goes.  Random-init Generator256 (the trained weights are not shipped)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import quantized_spectrum_cartography_b200 as q  # noqa: E402
from quantized_spectrum_cartography_b200 import dip, qmc  # noqa: E402
from quantized_spectrum_cartography_b200.quantization_model import assign_levels  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    B, R, K, I, J = 256, 4, 64, 51, 51
    torch.manual_seed(0)
    gen = dip.Generator256().eval().to(dev)
    for p in gen.parameters():
        p.requires_grad_(False)
    g = torch.Generator(device=dev).manual_seed(1)
    with torch.no_grad():
        S_true = gen(torch.randn(B * R, 256, device=dev, generator=g)).reshape(B, R, -1)
    C_true = torch.rand(B, R, K, device=dev, generator=g) * 0.2 + 0.05
    T = torch.einsum("brp,brk->bkp", S_true, C_true)
    thr = T.median().item()
    bb = torch.tensor([0.0, thr, 10.0])
    sigma = 0.5 * thr
    Y = assign_levels(T + sigma * torch.randn(T.shape, device=dev, generator=g), bb)
    Wx = torch.bernoulli(torch.full(T.shape, 0.1, device=dev), generator=g)
    lik = q.make_likelihood(bb, sigma)
    obs = q.make_obs(Y, Wx, K, dev, B=B, R=R)
    nll_fn = qmc.cuda_nll_fn(obs, lik)
    Z0 = torch.randn(B, R, 256, device=dev, generator=g)
    C0 = 0.9 * C_true
    cfg = dip.DipConfig(iters=5, lam_c=1.0, lam_s=0.1, search_at=-1)
    dip.solve_deep_prior(gen, Z0, C0, nll_fn, cfg)                       # warm-up (cuDNN autotune, allocator)
    cfg = dip.DipConfig(iters=30, lam_c=1.0, lam_s=0.1, search_at=-1)
    res = dip.solve_deep_prior(gen, Z0, C0, nll_fn, cfg)
    it_s = res["iterations"] / res["seconds"]

    def timed(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    S = gen(Z0.reshape(B * R, 256)).reshape(B, R, -1)
    t_nll = timed(lambda: q.nll_fwd_bwd(S, C0, obs, lik))
    Zg = Z0.clone().requires_grad_(True)

    def gen_fwd_bwd():
        out = gen(Zg.reshape(B * R, 256))
        out.backward(torch.ones_like(out))
    t_gen = timed(gen_fwd_bwd)
    print(json.dumps({"case": "cfg5 deep prior", "maps": B, "emitters": R, "generator_batch": B * R, "shape": "51x51x64",
                      "observed_entries": obs.nobs, "obs_layout": "lanes" if obs.lanes else "rows",
                      "deep_prior_iterations_per_s": it_s, "map_iterations_per_s": it_s * B,
                      "ms_per_iteration": 1e3 / it_s, "likelihood_eval_ms": t_nll, "generator_fwd_bwd_ms": t_gen,
                      "note": "one iteration = C-step (1 evaluation) + Z-step (generator forward, 1 evaluation, generator backward)"}))


if __name__ == "__main__":
    main()
