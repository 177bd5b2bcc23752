"""GPU: the BASELINE.json configurations at FULL size, checked through size-independent properties
(the oracle cannot run 4096 maps or a 512x512x256 instance in seconds):

* additivity -- NLL and both gradients are sums over observed entries, so evaluating two disjoint
  masks separately and adding must equal evaluating their union;
* batch consistency -- a map evaluated inside the 4096-map launch equals the same map evaluated alone
  with the other kernel, and permuting the maps permutes the outputs;
* path agreement -- gather-flat, gather-tiled, gather-lanes and dense-tcgen05 agree on the same instance;
* reproducibility -- the lanes and tiled kernels use no atomics at one tile per map: bitwise equal reruns;
* a sampled subset of maps against the float64 oracle.
"""
import numpy as np
import pytest
import torch

from oracle import qmc_oracle as oc

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.fixture(scope="module")
def q():
    import quantized_spectrum_cartography_b200 as pkg
    return pkg


def _one_bit_problem(q, B, I, J, K, R, f, seed):
    from quantized_spectrum_cartography_b200 import synth
    from quantized_spectrum_cartography_b200.quantization_model import assign_levels
    dev = torch.device("cuda", 0)
    maps = synth.generate_maps(B, I, J, K, R, seed=seed, device=dev)
    T = maps.tensor()
    thr = float(T.reshape(-1)[:: max(1, T.numel() // 1_000_000)].median())
    bb = torch.tensor([0.0, thr, 1.0])
    gen = torch.Generator(device=dev).manual_seed(seed + 1)
    Y = assign_levels(T + thr * torch.randn(T.shape, device=dev, generator=gen), bb).to(torch.uint8)
    Wx = torch.bernoulli(torch.full(T.shape, f, device=dev), generator=gen)
    return maps, Y, Wx, bb, thr


def test_cfg3_full_batch_lanes_kernel(q):
    """cfg3 through the headline path (lane-stream layout, gather_lanes_kernel, S pixel-major): agreement
    with the tiled kernel on all 4096 maps, additivity over a mask split, map permutation, bitwise
    reproducibility over repeated launches (no atomics: any shared-memory race would show up here), both S
    storage orders, the single-gradient modes, and sampled maps against the float64 oracle."""
    from quantized_spectrum_cartography_b200 import _lib
    B, I, J, K, R = 4096, 51, 51, 64, 4
    IJ = I * J
    maps, Y, Wx, bb, sigma = _one_bit_problem(q, B, I, J, K, R, 0.10, seed=0)
    lik = q.make_likelihood(bb, sigma)
    S, C = (0.8 * maps.S_true).contiguous(), maps.C_true.contiguous()
    S_pm = S.transpose(1, 2).contiguous().transpose(1, 2)
    build = lambda y, w: q.make_obs(y, w, K, y.device, B=y.shape[0], R=R, tiled=True, lanes=True)
    obs = build(Y, Wx)
    assert obs.lanes and obs.n_sub == obs.tile_warps == 8 and obs.padding_fraction() < 0.12
    nll, gS, gC = q.nll_fwd_bwd(S_pm, C, obs, lik)
    assert torch.isfinite(nll).all() and torch.isfinite(gS).all() and torch.isfinite(gC).all()
    # the previous headline kernel on the same batch
    n_sub, sub, tw = q.plan_tiles(IJ, K, R)
    obs_t = q.build_obs(Y, Wx, K, IJ, B, n_sub=n_sub, sub_pixels=sub, tile_warps=tw, bank_mod=q.bank_mod_for_rank(R))
    t = q.nll_fwd_bwd(S, C, obs_t, lik, algo=_lib.QMC_ALGO_TILED)
    assert torch.allclose(nll, t[0], rtol=1e-6)
    assert rel(gS, t[1]) < 1e-5 and rel(gC, t[2]) < 1e-5
    worst = ((gS - t[1]).flatten(1).norm(dim=1) / t[1].flatten(1).norm(dim=1)).max().item()
    assert worst < 1e-4, worst                                   # every single map, not only the batch norm
    # bitwise reproducible, launch after launch
    for _ in range(5):
        again = q.nll_fwd_bwd(S_pm, C, obs, lik)
        assert torch.equal(again[0], nll) and torch.equal(again[1], gS) and torch.equal(again[2], gC)
    # emitter-major S (the reference's layout) takes the transposing staging path: same arithmetic
    e = q.nll_fwd_bwd(S, C, obs, lik)
    assert torch.equal(e[0], nll) and torch.equal(e[1], gS.contiguous()) and torch.equal(e[2], gC)
    # one gradient at a time (the solver's C-step and S-step)
    c_only = q.nll_fwd_bwd(S_pm, C, obs, lik, skip_gs=True)
    s_only = q.nll_fwd_bwd(S_pm, C, obs, lik, skip_gc=True)
    assert torch.equal(c_only[2], gC) and torch.equal(s_only[1], gS) and torch.equal(c_only[0], nll) and torch.equal(s_only[0], nll)
    # additivity over a split of the mask (different streams, different padding, same sums)
    half = (torch.rand(Wx.shape, device=Wx.device, generator=torch.Generator(device=Wx.device).manual_seed(5)) < 0.5).float()
    a = q.nll_fwd_bwd(S_pm, C, build(Y, Wx * half), lik)
    b = q.nll_fwd_bwd(S_pm, C, build(Y, Wx * (1 - half)), lik)
    assert torch.allclose(a[0] + b[0], nll, rtol=5e-7)
    assert rel(a[1] + b[1], gS) < 1e-5 and rel(a[2] + b[2], gC) < 1e-5
    # permuting the maps permutes the outputs, bitwise
    perm = torch.randperm(B, device=S.device, generator=torch.Generator(device=S.device).manual_seed(6))
    p = q.nll_fwd_bwd(S_pm[perm].transpose(1, 2).contiguous().transpose(1, 2), C[perm].contiguous(),
                      build(Y[perm].contiguous(), Wx[perm].contiguous()), lik)
    assert torch.equal(p[0], nll[perm]) and torch.equal(p[1], gS[perm]) and torch.equal(p[2], gC[perm])
    for m in (0, 2048, B - 1):
        want = oc.nll_and_grads_fp64(S[m].cpu().reshape(R, 1, I, J), C[m].cpu(), Y[m].cpu().long().reshape(K, 1, I, J),
                                     Wx[m].cpu().reshape(K, 1, I, J), bb, sigma)
        assert abs(nll[m].item() / want[0] - 1) < 1e-5
        assert np.linalg.norm(gS[m].cpu().numpy() - want[1].reshape(R, -1)) / np.linalg.norm(want[1]) < 1e-4
        assert np.linalg.norm(gC[m].cpu().numpy() - want[2]) / np.linalg.norm(want[2]) < 1e-4


def test_cfg3_full_batch_properties(q):
    """cfg3: 4096 maps 51x51x64, R=4, 10 %, one-bit -- the benchmark workload itself."""
    from quantized_spectrum_cartography_b200 import _lib
    B, I, J, K, R = 4096, 51, 51, 64, 4
    IJ = I * J
    maps, Y, Wx, bb, sigma = _one_bit_problem(q, B, I, J, K, R, 0.10, seed=0)
    lik = q.make_likelihood(bb, sigma)
    S, C = (0.8 * maps.S_true).contiguous(), maps.C_true.contiguous()
    n_sub, sub, tw = q.plan_tiles(IJ, K, R)
    build = lambda y, w: q.build_obs(y, w, K, IJ, y.shape[0], n_sub=n_sub, sub_pixels=sub, tile_warps=tw,
                                     bank_mod=q.bank_mod_for_rank(R))
    obs = build(Y, Wx)
    assert abs(obs.nobs / (B * K * IJ) - 0.10) < 1e-3
    nll, gS, gC = q.nll_fwd_bwd(S, C, obs, lik, algo=_lib.QMC_ALGO_TILED)
    assert torch.isfinite(nll).all() and torch.isfinite(gS).all() and torch.isfinite(gC).all()
    # additivity over a split of the mask
    half = (torch.rand(Wx.shape, device=Wx.device, generator=torch.Generator(device=Wx.device).manual_seed(5)) < 0.5).float()
    a = q.nll_fwd_bwd(S, C, build(Y, Wx * half), lik, algo=_lib.QMC_ALGO_TILED)
    b = q.nll_fwd_bwd(S, C, build(Y, Wx * (1 - half)), lik, algo=_lib.QMC_ALGO_TILED)
    assert torch.allclose(a[0] + b[0], nll, rtol=5e-7)      # per-lane fp32 partial sums regroup
    assert rel(a[1] + b[1], gS) < 1e-5 and rel(a[2] + b[2], gC) < 1e-5
    # permuting the maps permutes the outputs (bitwise: the tiled kernel has no atomics here)
    perm = torch.randperm(B, device=S.device, generator=torch.Generator(device=S.device).manual_seed(6))
    p = q.nll_fwd_bwd(S[perm].contiguous(), C[perm].contiguous(), build(Y[perm].contiguous(), Wx[perm].contiguous()),
                      lik, algo=_lib.QMC_ALGO_TILED)
    assert torch.equal(p[0], nll[perm]) and torch.equal(p[1], gS[perm]) and torch.equal(p[2], gC[perm])
    # maps from inside the batch, evaluated alone with the other kernel and against the oracle
    for m in (0, 1777, B - 1):
        solo = q.nll_fwd_bwd(S[m:m + 1].contiguous(), C[m:m + 1].contiguous(),
                             q.build_obs(Y[m:m + 1].contiguous(), Wx[m:m + 1].contiguous(), K, IJ, 1), lik, algo=_lib.QMC_ALGO_FLAT)
        assert abs(solo[0][0].item() / nll[m].item() - 1) < 1e-6
        assert rel(solo[1][0], gS[m]) < 1e-5 and rel(solo[2][0], gC[m]) < 1e-5
        want = oc.nll_and_grads_fp64(S[m].cpu().reshape(R, 1, I, J), C[m].cpu(), Y[m].cpu().long().reshape(K, 1, I, J),
                                     Wx[m].cpu().reshape(K, 1, I, J), bb, sigma)
        assert abs(nll[m].item() / want[0] - 1) < 1e-5
        assert np.linalg.norm(gS[m].cpu().numpy() - want[1].reshape(R, -1)) / np.linalg.norm(want[1]) < 1e-4
        assert np.linalg.norm(gC[m].cpu().numpy() - want[2]) / np.linalg.norm(want[2]) < 1e-4


def test_cfg4_full_instance_paths_agree(q):
    """cfg4: 512x512x256, R=16, 50 % -- gather-flat vs dense-tcgen05 vs gather-tiled, and additivity."""
    from quantized_spectrum_cartography_b200 import _lib, dense, synth
    from quantized_spectrum_cartography_b200.quantization_model import assign_levels
    dev = torch.device("cuda", 0)
    I, J, K, R = 512, 512, 256, 16
    IJ = I * J
    maps = synth.generate_maps(1, I, J, K, R, seed=3, device=dev)
    T = maps.tensor()[0]
    off = float(T.median()) * 0.1
    X = torch.log(T + off)
    bb = synth.equal_mass_boundaries(X, 8)
    sigma = float((bb[1:] - bb[:-1]).min()) * 2.0
    gen = torch.Generator(device=dev).manual_seed(4)
    Y = assign_levels(X + sigma * torch.randn(X.shape, device=dev, generator=gen), bb).to(torch.uint8)
    Wx = torch.bernoulli(torch.full(T.shape, 0.5, device=dev), generator=gen)
    lik = q.make_likelihood(bb, sigma, offset=off)
    S, C = (0.8 * maps.S_true).contiguous(), maps.C_true.contiguous()
    flat = q.nll_fwd_bwd(S, C, q.build_obs(Y, Wx, K, IJ, 1), lik, algo=_lib.QMC_ALGO_FLAT)
    d = dense.nll_fwd_bwd_dense(S[0], C[0], dense.pack_dense(Y, Wx, K), lik)
    assert abs(d[0].item() / flat[0][0].item() - 1) < 1e-6
    assert rel(d[1], flat[1][0]) < 1e-4 and rel(d[2], flat[2][0]) < 1e-4
    n_sub, sub, tw = q.plan_tiles(IJ, K, R)
    tiled = q.nll_fwd_bwd(S, C, q.build_obs(Y, Wx, K, IJ, 1, n_sub=n_sub, sub_pixels=sub, tile_warps=tw,
                                            bank_mod=q.bank_mod_for_rank(R)), lik, algo=_lib.QMC_ALGO_TILED)
    assert abs(tiled[0][0].item() / flat[0][0].item() - 1) < 1e-6
    assert rel(tiled[1][0], flat[1][0]) < 1e-4 and rel(tiled[2][0], flat[2][0]) < 1e-4
    # additivity through the dense path: pixel halves of the mask
    left = torch.zeros_like(Wx)
    left[:, : IJ // 2] = 1
    a = dense.nll_fwd_bwd_dense(S[0], C[0], dense.pack_dense(Y, Wx * left, K), lik)
    b = dense.nll_fwd_bwd_dense(S[0], C[0], dense.pack_dense(Y, Wx * (1 - left), K), lik)
    assert abs((a[0] + b[0]).item() / d[0].item() - 1) < 1e-7
    assert rel(a[1] + b[1], d[1]) < 1e-5 and rel(a[2] + b[2], d[2]) < 1e-4


def test_cfg2_instance_vs_oracle_subsample(q):
    """cfg2: 101x101x128, R=8, 20 %, 8 levels, log domain: the whole instance against the float64
    oracle (1.3 M dense entries: the oracle still runs in about a second)."""
    from quantized_spectrum_cartography_b200 import qmc
    pb = qmc.synth_problem("cfg2", 1, torch.device("cuda", 0), seed=2)
    maps, lik, obs = pb["maps"], pb["lik"], pb["obs"]
    I, J, K, R = 101, 101, 128, 8
    S, C = (0.8 * maps.S_true).contiguous(), maps.C_true.contiguous()
    nll, gS, gC = q.nll_fwd_bwd(S, C, obs, lik)
    want = oc.nll_and_grads_fp64(S[0].cpu().reshape(R, 1, I, J), C[0].cpu(), pb["Y"][0].cpu().reshape(K, 1, I, J),
                                 pb["Wx"][0].cpu().reshape(K, 1, I, J), pb["bb"], pb["sigma"], offset=pb["offset"],
                                 sentinels=False)
    assert abs(nll[0].item() / want[0] - 1) < 1e-5
    assert np.linalg.norm(gS[0].cpu().numpy() - want[1].reshape(R, -1)) / np.linalg.norm(want[1]) < 1e-4
    assert np.linalg.norm(gC[0].cpu().numpy() - want[2]) / np.linalg.norm(want[2]) < 1e-4


@pytest.mark.parametrize("epilogue", ["lsq", "logistic"])
def test_cfg3_full_batch_other_epilogues(q, epilogue):
    """cfg3 at full size with the least-squares and the logistic epilogue: lane-stream kernel against the
    tiled kernel on all 4096 maps, additivity over a mask split, and sampled maps against the float64 oracle."""
    from quantized_spectrum_cartography_b200 import _lib
    B, I, J, K, R = 4096, 51, 51, 64, 4
    IJ = I * J
    maps, Y, Wx, bb, sigma = _one_bit_problem(q, B, I, J, K, R, 0.10, seed=1)
    if epilogue == "lsq":
        lik = q.make_likelihood(bb, None, least_squares=True)
    else:
        lik = q.make_likelihood(bb, 0.5 * sigma, model="logistic")
    S, C = (0.8 * maps.S_true).contiguous(), maps.C_true.contiguous()
    S_pm = S.transpose(1, 2).contiguous().transpose(1, 2)
    build = lambda y, w: q.make_obs(y, w, K, y.device, B=y.shape[0], R=R, tiled=True, lanes=True)
    obs = build(Y, Wx)
    assert obs.lanes
    nll, gS, gC = q.nll_fwd_bwd(S_pm, C, obs, lik)
    assert torch.isfinite(nll).all() and torch.isfinite(gS).all() and torch.isfinite(gC).all()
    n_sub, sub, tw = q.plan_tiles(IJ, K, R)
    obs_t = q.build_obs(Y, Wx, K, IJ, B, n_sub=n_sub, sub_pixels=sub, tile_warps=tw, bank_mod=q.bank_mod_for_rank(R))
    t = q.nll_fwd_bwd(S, C, obs_t, lik, algo=_lib.QMC_ALGO_TILED)
    assert ((nll - t[0]).abs() / t[0].abs()).max().item() < 1e-6
    assert ((gS - t[1]).flatten(1).norm(dim=1) / t[1].flatten(1).norm(dim=1)).max().item() < 1e-5
    assert ((gC - t[2]).flatten(1).norm(dim=1) / t[2].flatten(1).norm(dim=1)).max().item() < 1e-5
    # additivity over a split of the mask
    half = (torch.rand(Wx.shape, device=Wx.device, generator=torch.Generator(device=Wx.device).manual_seed(5)) < 0.5).float()
    a = q.nll_fwd_bwd(S_pm, C, build(Y, Wx * half), lik)
    b = q.nll_fwd_bwd(S_pm, C, build(Y, Wx * (1 - half)), lik)
    assert ((a[0] + b[0] - nll).abs() / nll.abs()).max().item() < 1e-6   # fp32 partial sums per lane
    assert rel(a[1] + b[1], gS) < 1e-5 and rel(a[2] + b[2], gC) < 1e-5
    # sampled maps against the float64 oracle
    for m in (0, 1777, 4095):
        args = (S[m].cpu().reshape(R, 1, I, J), C[m].cpu(), Y[m].cpu().long().reshape(K, 1, I, J), Wx[m].cpu().reshape(K, 1, I, J), bb)
        want = oc.lsq_and_grads_fp64(*args) if epilogue == "lsq" else oc.logistic_nll_and_grads_fp64(*args, 0.5 * sigma, None, True)
        assert nll[m].item() == pytest.approx(want[0], rel=1e-5)
        assert rel(gS[m].cpu(), torch.from_numpy(want[1].reshape(R, -1))) < 1e-4
        assert rel(gC[m].cpu(), torch.from_numpy(want[2])) < 1e-4
