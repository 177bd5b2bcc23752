#!/usr/bin/env python
"""Time the experimental variants of gather_lanes_kernel (QMC_LANES_VARIANT) on the cfg3 batch.
Development tool: variants other than 0 may rely on builder guarantees the current streams do not give,
so only their duration is meaningful."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from quantized_spectrum_cartography_b200 import _lib
from quantized_spectrum_cartography_b200._lib import check, lib


def main():
    variants = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,1,2,3".split(","))]
    layout = sys.argv[2] if len(sys.argv) > 2 else "pixel_major"
    dev = torch.device("cuda", 0)
    wl = bench.build_workload(4096, dev, seed=0)
    obs, lik, R, K, IJ = wl["obs"], wl["lik"], wl["R"], wl["K"], wl["IJ"]
    B = 4096
    S = wl["S"]
    if layout == "pixel_major":
        S = S.transpose(1, 2).contiguous().transpose(1, 2)
    Cf = wl["C"]
    nll = torch.empty(B, dtype=torch.float64, device=dev)
    gS = torch.empty_strided(S.shape, S.stride(), dtype=torch.float32, device=dev)
    gC = torch.empty_like(Cf)
    view = obs.view()
    st = torch.cuda.current_stream()
    ref = None
    out = {}
    for v in variants:
        os.environ["QMC_LANES_VARIANT"] = str(v)

        def step():
            check(lib.qmc_nll_fwd_bwd_gather(S.data_ptr(), S.stride(0), S.stride(1), S.stride(2), Cf.data_ptr(),
                                             C.byref(view), C.byref(lik), B, IJ, K, R, _lib.QMC_ALGO_AUTO,
                                             obs.tile_warps, nll.data_ptr(), gS.data_ptr(), gC.data_ptr(), st.cuda_stream))
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(20):
            step()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res = (nll.clone(), gS.clone(), gC.clone())
        if ref is None:
            ref = res
        errs = [float(((a - b).norm() / b.norm()).item()) for a, b in zip(res, ref)]
        out[v] = {"ms": ms, "err_vs_first": errs}
        print(f"variant {v} layout {layout}: {ms * 1e3:.1f} us  err vs first {errs}", flush=True)
    print(json.dumps({"layout": layout, "padding": obs.padding_fraction(), "results": out}))


if __name__ == "__main__":
    main()
