#!/usr/bin/env python
"""Where the dense tcgen05 kernel's time goes at cfg4 (one GPU): the full kernel against runs with parts switched off
through the QMC_DENSE_DEBUG measurement hook (1 = no MMA3, 2 = no MMA2, 4 = no likelihood evaluation, 8 = no G stores;
results are then wrong, only the time means something).  One JSON line."""
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
    obs = dense.pack_dense(Y, Wx, K)
    res = {}
    for dbg in [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3, 4, 8, 12, 15, 7]:
        os.environ["QMC_DENSE_DEBUG"] = str(dbg)
        for _ in range(3):
            dense.nll_fwd_bwd_dense(S, Cm, obs, lik)
        torch.cuda.synchronize()
        best = 1e9
        for _rep in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                dense.nll_fwd_bwd_dense(S, Cm, obs, lik)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / 10)
        res[str(dbg)] = round(best, 4)
    os.environ.pop("QMC_DENSE_DEBUG", None)
    print(json.dumps({"ms_by_debug_mask": res}))


if __name__ == "__main__":
    main()
