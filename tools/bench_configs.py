#!/usr/bin/env python
"""Secondary measurements for DESIGN.md: the single-instance configurations of BASELINE.json
(cfg1, cfg2: flat gather kernel; cfg4: dense tcgen05 kernel vs the gather kernels on the same
instance).  Not the contract benchmark (that is bench.py, cfg3).  Prints one JSON line per case."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import quantized_spectrum_cartography_b200 as q  # noqa: E402
from quantized_spectrum_cartography_b200 import _lib, dense, synth  # noqa: E402
from quantized_spectrum_cartography_b200.quantization_model import assign_levels  # noqa: E402


def timeit(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()                    # > L2: evict the inputs between iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def problem(I, J, K, R, f, levels, log_domain, dev, seed=0):
    maps = synth.generate_maps(1, I, J, K, R, seed=seed, device=dev)
    T = maps.tensor()
    gen = torch.Generator(device=dev).manual_seed(seed + 1)
    if log_domain:
        offset = float(T.median()) * 0.1 + 1e-12
        X = torch.log(T + offset)
    else:
        offset, X = None, T
    if levels == 2 and not log_domain:
        thr = float(T.median())
        bb, sigma = torch.tensor([0.0, thr, 1.0]), thr
    else:
        bb = synth.equal_mass_boundaries(X, levels)
        sigma = float((bb[1:] - bb[:-1]).min()) * 2.0
    Y = assign_levels(X + sigma * torch.randn(X.shape, device=dev, generator=gen), bb)
    Wx = torch.bernoulli(torch.full(T.shape, f, device=dev), generator=gen)
    return maps, Y, Wx, q.make_likelihood(bb, sigma, offset=offset)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="cfg1,cfg2,cfg4")
    ap.add_argument("--lsq-maps", type=int, default=4096)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak = 6455.9
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = {"cfg1": (51, 51, 64, 4, 0.10, 2, False), "cfg2": (101, 101, 128, 8, 0.20, 8, True),
              "cfg4": (512, 512, 256, 16, 0.50, 8, True)}
    for name in args.cases.split(","):
        if name not in shapes:
            continue
        I, J, K, R, f, levels, logd = shapes[name]
        IJ = I * J
        maps, Y, Wx, lik = problem(I, J, K, R, f, levels, logd, dev)
        S = (0.8 * maps.S_true).contiguous()
        C = maps.C_true.contiguous()
        out = {"case": name, "shape": f"{I}x{J}x{K}", "R": R, "sampling": f, "levels": levels, "log_domain": logd}
        obs_f = q.build_obs(Y[0], Wx[0], K, IJ, 1)
        out["nobs"] = obs_f.nobs
        t = timeit(lambda: q.nll_fwd_bwd(S, C, obs_f, lik, algo=_lib.QMC_ALGO_FLAT), flush=flush)
        out["flat_ms"] = t
        out["flat_entries_per_s"] = obs_f.nobs / (t * 1e-3)
        out["flat_hbm_frac"] = obs_f.algorithmic_bytes(R) / (t * 1e-3) / 1e9 / peak
        n_sub, sub, tw = q.plan_tiles(IJ, K, R)
        if n_sub // tw >= 16:
            obs_t = q.build_obs(Y[0], Wx[0], K, IJ, 1, n_sub=n_sub, sub_pixels=sub, tile_warps=tw, bank_mod=q.bank_mod_for_rank(R))
            t = timeit(lambda: q.nll_fwd_bwd(S, C, obs_t, lik, algo=_lib.QMC_ALGO_TILED), flush=flush)
            out["tiled_ms"] = t
            out["tiled_entries_per_s"] = obs_t.nobs / (t * 1e-3)
            out["tiled_hbm_frac"] = obs_t.algorithmic_bytes(R) / (t * 1e-3) / 1e9 / peak
            out["tiled_geometry"] = {"tiles": n_sub // tw, "tile_warps": tw, "sub_pixels": sub}
        if n_sub // tw >= 16 and 32 <= K <= 256:
            n_sub_l, sub_l, tw_l = q.plan_tiles(IJ, K, R, lanes=True, max_level=levels - 1)
            try:
                obs_l = q.build_obs(Y[0], Wx[0], K, IJ, 1, n_sub=n_sub_l, sub_pixels=sub_l, tile_warps=tw_l, lanes=True)
            except _lib.QmcError as e:              # streams too large for the builder's shared memory
                obs_l = None
                out["lanes_skipped"] = str(e)[:120]
            if obs_l is not None:
                S_pm = S.transpose(1, 2).contiguous().transpose(1, 2)
                t = timeit(lambda: q.nll_fwd_bwd(S_pm, C, obs_l, lik), flush=flush)
                out["lanes_ms"] = t
                out["lanes_entries_per_s"] = obs_l.nobs / (t * 1e-3)
                out["lanes_hbm_frac"] = obs_f.algorithmic_bytes(R) / (t * 1e-3) / 1e9 / peak
                out["lanes_geometry"] = {"tiles": n_sub_l // tw_l, "tile_warps": tw_l, "sub_pixels": sub_l,
                                         "padding": round(obs_l.padding_fraction(), 4)}
        if dense.dense_supported(K, R):
            dobs = dense.pack_dense(Y[0], Wx[0], K)
            t = timeit(lambda: dense.nll_fwd_bwd_dense(S[0], C[0], dobs, lik), flush=flush)
            out["dense_ms"] = t
            out["dense_entries_per_s"] = dobs.nobs / (t * 1e-3)
            out["dense_hbm_frac"] = dobs.algorithmic_bytes(R) / (t * 1e-3) / 1e9 / peak
            out["dense_mma_tflops"] = 3 * 2.0 * IJ * K * R / (t * 1e-3) / 1e12
            a = q.nll_fwd_bwd(S, C, obs_f, lik, algo=_lib.QMC_ALGO_FLAT)
            b = dense.nll_fwd_bwd_dense(S[0], C[0], dobs, lik)
            out["dense_vs_flat_nll_rel"] = abs(b[0].item() / a[0][0].item() - 1)
            out["dense_vs_flat_gS_rel"] = float((b[1] - a[1][0]).norm() / a[1][0].norm())
            out["dense_vs_flat_gC_rel"] = float((b[2] - a[2][0]).norm() / a[2][0].norm())
        print(json.dumps(out))
    if "cfg2" in args.cases.split(","):
        # batched multi-bit maps: 256 x cfg2 through the lane-stream kernel (general epilogue, log domain)
        from quantized_spectrum_cartography_b200 import qmc
        B = 256
        pb = qmc.synth_problem("cfg2", B, dev, seed=0)
        obs, lik = pb["obs"], pb["lik"]
        S = (0.8 * pb["maps"].S_true).transpose(1, 2).contiguous().transpose(1, 2)
        C = pb["maps"].C_true.contiguous()
        t = timeit(lambda: q.nll_fwd_bwd(S, C, obs, lik))
        R = C.shape[1]
        alg = obs.nobs * 5 + B * (2 * 4 * R * (S.shape[2] + C.shape[2]) + 4)
        print(json.dumps({"case": "cfg2 x 256 batched", "layout": "lanes" if obs.lanes else "rows", "nobs": obs.nobs, "ms": t,
                          "entries_per_s": obs.nobs / (t * 1e-3), "hbm_frac": alg / (t * 1e-3) / 1e9 / peak,
                          "tile_warps": obs.tile_warps, "tiles_per_map": obs.n_sub // max(obs.tile_warps, 1),
                          "padding": round(obs.padding_fraction(), 4)}))
    if "lsq" in args.cases.split(","):
        # the masked least-squares baseline on the cfg3 batch (same observation set, same
        # kernel, epilogue without the SFU chain) next to the likelihood on the same inputs
        from quantized_spectrum_cartography_b200 import qmc
        B = args.lsq_maps
        pb = qmc.synth_problem("cfg1", B, dev, seed=0)          # cfg3 = B x cfg1
        obs = pb["obs"]
        S = (0.8 * pb["maps"].S_true).transpose(1, 2).contiguous().transpose(1, 2)
        C = pb["maps"].C_true.contiguous()
        R = C.shape[1]
        alg = obs.nobs * 5 + B * (2 * 4 * R * (S.shape[2] + C.shape[2]) + 4)
        lik_lsq = q.make_likelihood(pb["bb"], None, least_squares=True)
        lik_logit = q.make_likelihood(pb["bb"], 0.5 * pb["sigma"], model="logistic")
        for name, lk in (("likelihood (one-bit probit)", pb["lik"]), ("least squares on bin mid-points", lik_lsq),
                         ("likelihood (logistic, stable log-difference)", lik_logit)):
            t = timeit(lambda: q.nll_fwd_bwd(S, C, obs, lk))
            print(json.dumps({"case": f"cfg3 x {B}", "epilogue": name, "layout": "lanes" if obs.lanes else "rows",
                              "nobs": obs.nobs, "ms": t, "entries_per_s": obs.nobs / (t * 1e-3),
                              "hbm_frac": alg / (t * 1e-3) / 1e9 / peak}))


if __name__ == "__main__":
    main()
