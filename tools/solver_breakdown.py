"""Time the pieces of one fused solver iteration at cfg3 (CUDA events, 30 repetitions each)."""
import sys
import torch
sys.path.insert(0, ".")
import bench
import quantized_spectrum_cartography_b200 as q
from quantized_spectrum_cartography_b200 import qmc
from quantized_spectrum_cartography_b200._lib import check, lib

dev = torch.device("cuda")
wl = bench.build_workload(4096, dev, seed=0)
obs, lik, S, Cf = wl["obs"], wl["lik"], wl["S"], wl["C"].contiguous()
B, R, IJ = S.shape
K = Cf.shape[2]
nll = torch.empty(B, dtype=torch.float64, device=dev)
gS = torch.empty_strided(S.shape, S.stride(), dtype=torch.float32, device=dev)
gC = torch.empty_like(Cf)
m, v = torch.zeros_like(S), torch.zeros_like(S)
ss, ss2 = torch.ones(B, dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.float64, device=dev)
p = S.clone()
st = lambda: torch.cuda.current_stream().cuda_stream


def timeit(name, fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / n * 1000:8.1f} us")


print("S strides", S.stride())
timeit("eval both gradients", lambda: q.nll_fwd_bwd(S, Cf, obs, lik, out=(nll, gS, gC)))
timeit("eval skip gS (C-step)", lambda: q.nll_fwd_bwd(S, Cf, obs, lik, out=(nll, gS, gC), skip_gs=True))
timeit("eval skip gC (S-step)", lambda: q.nll_fwd_bwd(S, Cf, obs, lik, out=(nll, gS, gC), skip_gc=True))
timeit("eval forward only", lambda: q.nll_fwd_bwd(S, Cf, obs, lik, want_grad=False))
timeit("S update (adam_frob)", lambda: check(lib.qmc_adam_frob_project(p.data_ptr(), gS.data_ptr(), m.data_ptr(), v.data_ptr(), B, R * IJ,
                                                                    ss.data_ptr(), ss2.data_ptr(), 1e-3, 0.9, 0.999, 1e-8, 1.0, 1, 5, None, st())))
for graph in (False, True):
    cfg = qmc.SolverConfig(iters=40, lam_c=1.0, lam_s=1.0, track_every=0, cuda_graph=graph)
    res = qmc.solve_lowrank_fused(S, Cf, obs, lik, cfg)
    print(f"solver graph={graph}: {res.seconds / 40 * 1e6:.1f} us per iteration")
