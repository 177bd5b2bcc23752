#!/usr/bin/env python
"""Dense tcgen05 kernel at cfg4's shape for pixel blocks of different sizes (one GPU): the share of one of N GPUs of
the sharded instance (262144 / N pixels) and exact multiples of 148 full tiles -- to separate the per-tile time from
the fixed cost of a launch, and to compare balanced short tiles (default) with full 128-pixel tiles
(QMC_DENSE_TILE_PIX=128).  One JSON line per case."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from quantized_spectrum_cartography_b200 import dense, qmc


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    c = qmc.CONFIGS["cfg4"]
    K, R = c["K"], c["R"]
    pb = qmc.synth_problem("cfg4", 1, dev, seed=0)
    Y, Wx, lik = pb["Y"][0].reshape(K, -1), pb["Wx"][0].reshape(K, -1), pb["lik"]
    S = (0.8 * pb["maps"].S_true[0]).reshape(R, -1).contiguous()
    Cm = pb["maps"].C_true[0].contiguous()
    IJ = S.shape[1]
    sizes = [IJ // 8, IJ // 4, IJ // 2, IJ] + [148 * 128 * w for w in (1, 2, 3, 4)]
    for n in sizes:
        Yl, Wl, Sl = Y[:, :n].contiguous(), Wx[:, :n].contiguous(), S[:, :n].contiguous()
        obs = dense.pack_dense(Yl, Wl, K)
        out = None
        res = {}
        for tp in ("auto", "128"):
            if tp == "auto":
                os.environ.pop("QMC_DENSE_TILE_PIX", None)
            else:
                os.environ["QMC_DENSE_TILE_PIX"] = tp
            for _ in range(3):
                r = dense.nll_fwd_bwd_dense(Sl, Cm, obs, lik)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                r = dense.nll_fwd_bwd_dense(Sl, Cm, obs, lik)
            for _ in range(5):
                g.replay()
            torch.cuda.synchronize()
            best = 1e9
            for _rep in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(20):
                    g.replay()
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b) / 20)
            res[tp] = (best, [x.double().clone() for x in r])
            del g
        ea = [float(((x - y).norm() / y.norm()).item()) for x, y in zip(res["auto"][1], res["128"][1])]
        print(json.dumps({"pixels": n, "full_tiles": -(-n // 128), "rounds_of_148": round(-(-n // 128) / 148, 2),
                          "observed": obs.nobs, "ms_auto": round(res["auto"][0], 4), "ms_tile128": round(res["128"][0], 4),
                          "auto_vs_128_rel_err": ea}), flush=True)
    os.environ.pop("QMC_DENSE_TILE_PIX", None)


if __name__ == "__main__":
    main()
