#!/usr/bin/env python
"""Strong-scaling study on one GPU: the per-GPU share of 4096 cfg3 maps at 8 / 4 / 2 GPUs (512 / 1024 / 2048 maps),
with the map's eight sub-tiles grouped into 1, 2 or 4 CTAs (8-, 4- or 2-warp CTAs).  One evaluation per step,
replayed from a CUDA graph, device-timed.  Prints one JSON line per case."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from quantized_spectrum_cartography_b200 import parallel


def main():
    maps = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["512", "1024", "4096"])]
    splits = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "2", "4"])]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ref = {}
    for n in maps:
        for split in splits:
            wl = bench.build_workload(n, dev, seed=100, ctas_per_map=split)
            obs = wl["obs"]
            bm = parallel.BatchedMaps(0, n, n, obs, wl["lik"])
            S = wl["S"].transpose(1, 2).contiguous().transpose(1, 2)
            Cf = wl["C"]
            out = (torch.empty(n, dtype=torch.float64, device=dev),
                   torch.empty_strided(S.shape, S.stride(), dtype=torch.float32, device=dev), torch.empty_like(Cf))
            for _ in range(3):
                bm.evaluate(S, Cf, out=out)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                bm.evaluate(S, Cf, out=out)
            for _ in range(10):
                g.replay()
            torch.cuda.synchronize()
            best = 1e9
            for _rep in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(50):
                    g.replay()
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b) / 50)
            res = [t.double().clone() for t in out]
            if n not in ref:
                ref[n] = res
            errs = [float(((x - y).norm() / y.norm()).item()) for x, y in zip(res, ref[n])]
            print(json.dumps({"maps": n, "ctas_per_map": split, "tile_warps": obs.tile_warps, "n_sub": obs.n_sub,
                              "us": round(best * 1e3, 2), "entries_per_s": obs.nobs / (best * 1e-3),
                              "padding": round(obs.padding_fraction(), 4), "err_vs_split1": errs}), flush=True)
            del wl, obs, bm, g
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
