#!/usr/bin/env python
"""Headline benchmark of the QMC quantized-likelihood hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

Workload (config.workload = "cfg3"): batched recovery of 4096 independent 51x51x64 one-bit maps,
rank R = 4, 10 % per-entry sampling -- BASELINE.json configs[2], the configuration the metric's
"% HBM roofline" is meaningful on (a single cfg1/cfg2 instance moves 0.2-2 MB and is pure launch
latency).  One *step* = one evaluation of the whole batch: fused NLL forward
+ gradients w.r.t. S and C.  With N GPUs every rank owns its own 4096 maps (weak scaling, no
collective on the data path); `value` = observed entries of all ranks / max-over-ranks time.

Printed JSON (one line, rank 0): the base contract's keys plus
  roofline     -- algorithmic bytes per launch / mean launch time, against the measured HBM peak
  cpu_baseline -- the oracle's torch-CPU port of the reference algorithm, timed on this box's cores
  e2e          -- same metric through the host-buffer C-ABI call (H2D of S,C and D2H of nll,gS,gC
                  inside the timed region)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG3 = dict(workload="cfg3", maps=4096, I=51, J=51, K=64, R=4, sampling=0.10, levels=2)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--maps", type=int, default=CFG3["maps"], help="maps per GPU (default: the cfg3 batch)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-solver", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the strong-scaling, cfg4 (dense / sharded) and emitter-major records")
    ap.add_argument("--tile-warps", type=int, default=8)
    ap.add_argument("--smem-budget-kb", type=int, default=110)
    ap.add_argument("--obs-layout", default="lanes", choices=["lanes", "rows"],
                    help="lanes: per-lane band walks, conflict-free steps (default); rows: band rows per sub-tile")
    ap.add_argument("--bank-mod", type=int, default=-1, help="-1: conflict-free order for the rank; 0: pixel order")
    ap.add_argument("--layout", default="pixel_major", choices=["pixel_major", "emitter_major"],
                    help="device storage of S for the kernel-only number")
    return ap.parse_args()


# -------------------------------------------------------------------------------------------------
# clocks
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
# workload
# -------------------------------------------------------------------------------------------------
def build_workload(n_maps: int, device, seed: int, tile_warps: int = 8, smem_budget_kb: int = 110, bank_mod: int = -1,
                   lanes: bool = True, ctas_per_map: int = 1):
    """cfg3 on one GPU: synthetic maps, one-bit observations, tiled compact observation set.
    Seeds: `seed` data, `seed+1` noise and mask.  ``ctas_per_map`` > 1 keeps the sub-tiles (one per
    warp, hence the streams' lengths) and groups them into that many smaller CTAs per map: a finer unit for the
    block scheduler when a GPU holds only a wave or two of maps."""
    import torch

    import quantized_spectrum_cartography_b200 as q
    from quantized_spectrum_cartography_b200 import _lib, synth
    from quantized_spectrum_cartography_b200._lib import check, lib

    c = CFG3
    I, J, K, R = c["I"], c["J"], c["K"], c["R"]
    IJ = I * J
    maps = synth.generate_maps(n_maps, I, J, K, R, seed=seed, device=device)
    T = maps.tensor()                                                    # [B, K, IJ]
    thr = T.reshape(-1)[:: max(1, T.numel() // 2_000_000)].median().item()
    sigma = thr                                                          # keeps the fp32 reference finite
    bb = torch.tensor([0.0, thr, 1.0])
    gen = torch.Generator(device=device).manual_seed(seed + 1)
    noise = torch.randn(T.shape, device=device, generator=gen)
    noisy = torch.empty_like(T)
    check(lib.qmc_noisy_signal(T.data_ptr(), noise.data_ptr(), sigma, 0.0, 0, T.numel(), noisy.data_ptr(),
                               torch.cuda.current_stream().cuda_stream))
    del noise
    Y = torch.empty(T.shape, dtype=torch.uint8, device=device)
    import ctypes as C
    arr = (C.c_float * 3)(*bb.tolist())
    check(lib.qmc_quantize_levels(noisy.data_ptr(), noisy.numel(), arr, 3, Y.data_ptr(), None,
                                  torch.cuda.current_stream().cuda_stream))
    del noisy
    Wx = torch.bernoulli(torch.full(T.shape, c["sampling"], device=device), generator=gen)
    n_sub, sub, tw = q.plan_tiles(IJ, K, R, tile_warps, smem_budget_kb * 1024, lanes=lanes, max_level=c["levels"] - 1)
    if ctas_per_map > 1 and n_sub == tw and tw % ctas_per_map == 0:
        tw //= ctas_per_map
    obs = q.build_obs(Y, Wx, K, IJ, n_maps, n_sub=n_sub, sub_pixels=sub, tile_warps=tw,
                      bank_mod=0 if lanes else (bank_mod if bank_mod >= 0 else q.bank_mod_for_rank(R)), lanes=lanes)
    lik = q.make_likelihood(bb, sigma)
    S_eval = (0.8 * maps.S_true).contiguous()                            # evaluation point
    C_eval = maps.C_true.contiguous()
    host_sample = dict(Y=Y[:256].cpu(), Wx=Wx[:256].cpu(), S=S_eval[:256].cpu(), C=C_eval[:256].cpu(),
                       bb=bb, sigma=sigma)
    del T, Y, Wx
    torch.cuda.empty_cache()
    return dict(obs=obs, lik=lik, S=S_eval, C=C_eval, IJ=IJ, K=K, R=R, I=I, J=J, host_sample=host_sample,
                thr=thr, sigma=sigma)


def config_keys(n_maps_per_gpu: int):
    """The `config` keys both arms print (the b200 arm adds its layout details)."""
    c = CFG3
    IJ, K, R = c["I"] * c["J"], c["K"], c["R"]
    gb = n_maps_per_gpu * (c["sampling"] * K * IJ * 5 + 2 * 4 * R * (IJ + K) + 4) / 1e9    # SURVEY 8(d), expected value
    return {"workload": c["workload"], "maps_per_gpu": n_maps_per_gpu, "shape": "51x51x64", "rank": c["R"],
            "sampling": c["sampling"], "levels": c["levels"], "evaluation_point": "0.8*S_true, C_true",
            "l2": "inputs ~%.2f GB/step per GPU > 126 MB L2 (no flush needed)" % gb}


def load_reference():
    """The reference's own implementation of the path: oracle/_ref/quantization_model.py, copied verbatim from
    /root/reference/qmc by oracle/make_ref.py in the build container (kind "reference"); the oracle's port of it
    when that copy is missing (kind "port").  Test infrastructure: only the CPU legs of this file use it."""
    import importlib.util
    path = os.path.join(ROOT, "oracle", "_ref", "quantization_model.py")
    if os.path.exists(path):
        spec = importlib.util.spec_from_file_location("reference_quantization_model", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod, "reference"
    from oracle import qmc_oracle as oc
    return oc, "port"


def reference_evaluation(mod, kind, S, Cm, Y, Wx, bb, sigma, vectorised=False):
    """One evaluation the way the reference does it (qmc/qmc.ipynb c1:145-153 with the functions of
    qmc/quantization_model.py:22-39,57-61,70-86): NLL forward + autograd backward to S and C.  Returns the NLL."""
    import torch
    if kind == "port" or vectorised:
        from oracle import qmc_oracle as oc
        return oc.nll_and_grads(S, Cm, Y, Wx, bb, sigma, vectorised=vectorised)[0]
    S = S.clone().requires_grad_(True)
    Cm = Cm.clone().requires_grad_(True)
    T_hat = mod.get_tensor(S, Cm).unsqueeze(dim=1)
    nll = -torch.sum(Wx * torch.log(mod.prob_probit(Y, T_hat, bb, sigma)))
    nll.backward()
    return nll.detach()


def cpu_reference_sample(hs, I, J, K, R, budget_s: float, vectorised: bool = False, max_maps: int = 256):
    """Time the reference's CPU implementation of the path (get_tensor -> prob_probit -> -sum(Wx*log P) -> backward)
    map by map until the budget is spent.  Returns (obs/s, maps, seconds, kind)."""
    import torch

    mod, kind = load_reference()
    torch.set_num_threads(os.cpu_count() or 1)
    n_obs, t_used, done = 0, 0.0, 0
    n_avail = min(max_maps, hs["S"].shape[0])
    while t_used < budget_s:                       # cycle over the sample until the budget is spent
        b = done % n_avail
        S = hs["S"][b].reshape(R, 1, I, J)
        Cm = hs["C"][b]
        Y = hs["Y"][b].reshape(K, 1, I, J).long()
        Wx = hs["Wx"][b].reshape(K, 1, I, J)
        t0 = time.perf_counter()
        nll = reference_evaluation(mod, kind, S, Cm, Y, Wx, hs["bb"], hs["sigma"], vectorised=vectorised)
        t_used += time.perf_counter() - t0
        assert torch.isfinite(nll), "reference NLL is not finite on the benchmark workload"
        n_obs += int(Wx.sum().item())
        done += 1
    return n_obs / t_used, done, t_used, kind


def cpu_workload_slice(n_maps: int, seed: int = 0):
    """The first maps of the cfg3 workload synthesised on the host cores alone: same generator code
    (quantized_spectrum_cartography_b200/synth.py, loaded by path: it needs torch only, not the CUDA library), same
    quantizer semantics through the oracle.  Nothing of the product is imported."""
    import importlib.util

    import torch

    from oracle import qmc_oracle as oc
    spec = importlib.util.spec_from_file_location("qmc_synth_cpu", os.path.join(ROOT, "quantized_spectrum_cartography_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    sys.modules["qmc_synth_cpu"] = synth          # dataclasses look the defining module up
    spec.loader.exec_module(synth)
    c = CFG3
    I, J, K, R = c["I"], c["J"], c["K"], c["R"]
    maps = synth.generate_maps(n_maps, I, J, K, R, seed=seed, device="cpu")
    T = maps.tensor()
    thr = T.reshape(-1)[:: max(1, T.numel() // 2_000_000)].median().item()
    sigma = thr
    bb = torch.tensor([0.0, thr, 1.0])
    gen = torch.Generator().manual_seed(seed + 1)
    noisy = oc.noisy_signal(T, torch.randn(T.shape, generator=gen), sigma)
    Y = oc.assign_levels(noisy, bb)
    Wx = torch.bernoulli(torch.full(T.shape, c["sampling"]), generator=gen)
    return dict(Y=Y, Wx=Wx, S=(0.8 * maps.S_true).contiguous(), C=maps.C_true.contiguous(), bb=bb, sigma=sigma)


# -------------------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref, the verbatim modules;
    the oracle port if they are missing) on this box's host cores, on a slice of the same workload synthesised on
    the CPU.  Does not import the product and needs no GPU.  Rank 0 only."""
    if rank != 0:
        return
    import torch

    c = CFG3
    I, J, K, R = c["I"], c["J"], c["K"], c["R"]
    maps_per_step = 8
    n_slice = min(256, maps_per_step * (args.steps + args.warmup))
    hs = cpu_workload_slice(n_slice, seed=0)
    torch.set_num_threads(os.cpu_count() or 1)
    mod, kind = load_reference()

    def step(i):
        tot = 0
        for b in range(maps_per_step):
            m = (i * maps_per_step + b) % hs["S"].shape[0]
            Wx = hs["Wx"][m].reshape(K, 1, I, J)
            nll = reference_evaluation(mod, kind, hs["S"][m].reshape(R, 1, I, J), hs["C"][m], hs["Y"][m].reshape(K, 1, I, J), Wx,
                                       hs["bb"], hs["sigma"])
            assert torch.isfinite(nll)
            tot += int(Wx.sum().item())
        return tot

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    total = 0
    for i in range(args.steps):
        total += step(args.warmup + i)
    dt = time.perf_counter() - t0
    value = total / dt
    cores = torch.get_num_threads()
    sample = f"{maps_per_step} of {c['maps']} maps per step, {args.steps} steps, maps synthesised on the host"
    print(json.dumps({
        "impl": "reference", "metric": "QMC observed-entries/s (fused NLL fwd + gS + gC)", "value": value,
        "unit": "observed-entries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_keys(args.maps),
        "cpu_baseline": {"value": value, "unit": "observed-entries/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "observed-entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def measured_hbm_peak():
    """(GB/s, where it comes from): the driver-measured copy bandwidth, else the profiling recipe's fallback."""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _timed(fn, n, stream, barrier):
    """Mean device time of fn() in ms over n calls, max over ranks taken by the caller."""
    import torch
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(n):
        fn()
    b.record(stream)
    barrier()
    return a.elapsed_time(b) / n


def strong_scaling_record(args, rank, world, dev, barrier, reduce_max, reduce_sum):
    """BASELINE configs[2] as written: 4096 maps in TOTAL, partitioned over the ranks by parallel.partition_maps /
    BatchedMaps, no collective on the data path.  One evaluation per step, replayed from a CUDA graph (at 8 GPUs a
    step is ~30 us of kernel: the Python call would be as long)."""
    import torch

    import quantized_spectrum_cartography_b200 as q
    from quantized_spectrum_cartography_b200 import parallel
    total = CFG3["maps"]
    lo, hi = parallel.partition_maps(total, world, rank)
    n = hi - lo
    wl = build_workload(n, dev, seed=100 + rank)
    bm = parallel.BatchedMaps(lo, hi, total, wl["obs"], wl["lik"])
    S = wl["S"].transpose(1, 2).contiguous().transpose(1, 2)
    Cf = wl["C"]
    out = (torch.empty(n, dtype=torch.float64, device=dev), torch.empty_strided(S.shape, S.stride(), dtype=torch.float32, device=dev),
           torch.empty_like(Cf))
    stream = torch.cuda.current_stream()
    for _ in range(3):
        bm.evaluate(S, Cf, out=out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        bm.evaluate(S, Cf, out=out)
    for _ in range(args.warmup):
        g.replay()
    cap = torch.cuda.current_stream()
    ms = reduce_max(_timed(g.replay, args.steps, cap, barrier))
    nobs = reduce_sum(float(wl["obs"].nobs))
    ctas = n * (wl["obs"].n_sub // wl["obs"].tile_warps)
    slots = 2 * torch.cuda.get_device_properties(dev).multi_processor_count
    waves = ctas / slots
    return {"maps_total": total, "maps_per_gpu": n, "ms_per_step": ms, "value": nobs / (ms * 1e-3), "unit": "observed-entries/s",
            "ctas_per_gpu": ctas, "resident_ctas_per_gpu": slots, "waves": round(waves, 3),
            "wave_quantisation_ceiling": round(waves / -(-ctas // slots), 3), "collective": "none",
            "api": "parallel.BatchedMaps.evaluate (one CUDA graph replay per step)"}


def cfg4_record(args, rank, world, dev, barrier, reduce_max):
    """BASELINE configs[3]: one 512x512x256 instance, R = 16, 50 % sampling, 8 levels, log domain -- the dense
    tcgen05 kernel on this rank's pixel block (parallel.ShardedInstance; the whole instance at one GPU) and,
    at N > 1, the NCCL all-reduce of the factor gradients: the contract form ("flat": [gS|gC|nll], 16.8 MB) and the
    pixel-block form ([gC|nll] only).  Per mode: time of one evaluation (local kernel + collective, one CUDA graph),
    of the local part alone and of the collective alone."""
    import torch
    import torch.distributed as dist

    import quantized_spectrum_cartography_b200 as q
    from quantized_spectrum_cartography_b200 import dense, parallel, qmc
    c = qmc.CONFIGS["cfg4"]
    I, J, K, R = c["I"], c["J"], c["K"], c["R"]
    IJ = I * J
    pb = qmc.synth_problem("cfg4", 1, dev, seed=0)          # same seed on every rank: replicated inputs
    Y, Wx, lik = pb["Y"][0], pb["Wx"][0], pb["lik"]
    nobs = int(pb["obs"].nobs)
    S = (0.8 * pb["maps"].S_true[0]).contiguous()
    Cm = pb["maps"].C_true[0].contiguous()
    ref = None
    if world > 1:                                           # single-GPU answer for the error columns
        d1 = dense.pack_dense(Y, Wx, K)
        ref = [x.clone() for x in dense.nll_fwd_bwd_dense(S, Cm, d1, lik)]
        del d1
    del pb
    torch.cuda.empty_cache()
    stream = torch.cuda.current_stream()
    rec = {"shape": f"{I}x{J}x{K}", "rank": R, "sampling": c["f"], "levels": c["levels"], "log_domain": bool(c["log_domain"]),
           "observed_entries": nobs, "kernel": "dense_kernel (tcgen05 kind::tf32, 3xTF32)", "modes": {}}
    n_it = max(args.steps, 10)
    for tag in (("flat", "pixel_block", "pixel_block_peer") if world > 1 else ("pixel_block",)):
        mode, exchange = ("pixel_block", "peer") if tag == "pixel_block_peer" else (tag, "nccl")
        inst = parallel.ShardedInstance.from_dense(Y, Wx, K, R, lik, mode=mode, align=128, dense=True, exchange=exchange)
        Sl = S[:, inst.lo:inst.hi].contiguous()
        use_graph = True
        try:
            inst.evaluate(Sl, Cm, gather_gS=False, cuda_graph=True)
        except Exception as e:                               # NCCL capture unavailable: time the eager sequence
            if exchange == "peer":
                # no CUDA IPC between the ranks' processes on this box (every rank gets the same verdict and raises
                # together, dense.PeerRegions): the NCCL forms above stand
                rec["modes"][tag] = {"unavailable": str(e)[:300]}
                del inst
                continue
            use_graph = False
            inst._graph = None
            torch.cuda.synchronize()
        if use_graph:
            Sl, Cm_in = inst._S_in, inst._C_in                 # the graph's own input buffers: no staging copy
        else:
            Cm_in = Cm
        ev = lambda: inst.evaluate(Sl, Cm_in, gather_gS=False, cuda_graph=use_graph)
        for _ in range(3):
            ev()
        ms = reduce_max(_timed(lambda: (inst._graph.replay() if use_graph else inst._step(Sl, Cm_in)), n_it, stream, barrier))
        nll, gS, gC = ev()
        m = {"ms_per_eval": ms, "entries_per_s": nobs / (ms * 1e-3), "exchange_bytes": inst.exchange_bytes() if world > 1 else 0,
             "cuda_graph": use_graph, "pixels_per_gpu": inst.hi - inst.lo}
        off = R * IJ if mode == "flat" else R * (inst.hi - inst.lo)
        if exchange == "peer":
            m["exchange"] = ("fused into dense_kernel: the last CTA stores [gC | nll] into every peer's region over NVLink "
                             "(CUDA IPC peer mappings), flags, waits, sums in rank order; no collective call")
            m["exchange_status"] = inst.exchange_status()
        else:
            # the local part alone (kernel + packing of the NLL words), as its own CUDA graph
            local = lambda: inst._local_into(inst._buf, Sl, Cm_in, off)
            for _ in range(3):
                local()
            torch.cuda.synchronize()
            gl = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gl):
                local()
            for _ in range(3):
                gl.replay()
            m["local_kernel_ms"] = reduce_max(_timed(gl.replay, n_it, torch.cuda.current_stream(), barrier))
            del gl
        if world > 1 and exchange == "peer":
            nll, gS, gC = inst.evaluate(Sl, Cm_in, gather_gS=False, cuda_graph=use_graph)
            gS_ref = ref[1][:, inst.lo:inst.hi]
            m["err_vs_one_gpu"] = {"nll": abs(nll.item() / ref[0].item() - 1), "gS": float((gS - gS_ref).norm() / gS_ref.norm()),
                                   "gC": float((gC - ref[2]).norm() / ref[2].norm())}
            inst.close()
        elif world > 1:
            coll = (lambda: dist.all_reduce(inst._buf)) if mode == "flat" else (lambda: dist.all_reduce(inst._buf[off:]))
            for _ in range(3):
                coll()
            m["all_reduce_ms"] = reduce_max(_timed(coll, n_it, stream, barrier))
            nll, gS, gC = inst.evaluate(Sl, Cm_in, gather_gS=False, cuda_graph=use_graph)   # buffers were clobbered by the timing loops
            gS_ref = ref[1] if mode == "flat" else ref[1][:, inst.lo:inst.hi]
            m["err_vs_one_gpu"] = {"nll": abs(nll.item() / ref[0].item() - 1), "gS": float((gS - gS_ref).norm() / gS_ref.norm()),
                                   "gC": float((gC - ref[2]).norm() / ref[2].norm())}
        rec["modes"][tag] = m
        del inst
    d = dense.DenseObs(torch.empty(0), IJ, K, nobs, 0)
    best = min(v["ms_per_eval"] for v in rec["modes"].values() if "ms_per_eval" in v)
    rec["algorithmic_bytes"] = d.algorithmic_bytes(R)
    rec["mma_tflops"] = 3 * 2.0 * IJ * K * R / (best * 1e-3) / 1e12
    rec["hbm_frac"] = rec["algorithmic_bytes"] / (best * 1e-3) / 1e9 / measured_hbm_peak()[0] if world == 1 else None
    cpath = os.path.join(ROOT, "profiles", "dense_counters.json")     # ncu counters of the same kernel (not live)
    if os.path.exists(cpath):
        with open(cpath) as f:
            rec["ncu"] = json.load(f)
    return rec


def bind_to_gpu_numa_node(dev):
    """Pin this process (and with it the pages of the pinned host buffers it allocates next: first touch) to the NUMA
    node the GPU hangs off, so that the end-to-end copies do not cross the socket interconnect.  Best effort."""
    import torch
    try:
        p = torch.cuda.get_device_properties(dev)
        name = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{name}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        pass
    return None


def run_b200(args, rank: int, world: int, local_rank: int):
    import ctypes as C

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import quantized_spectrum_cartography_b200 as q
    from quantized_spectrum_cartography_b200 import _lib
    from quantized_spectrum_cartography_b200._lib import check, lib

    wl = build_workload(args.maps, dev, seed=2 * rank, tile_warps=args.tile_warps, smem_budget_kb=args.smem_budget_kb,
                        bank_mod=args.bank_mod, lanes=args.obs_layout == "lanes")
    obs, lik, R, K, IJ = wl["obs"], wl["lik"], wl["R"], wl["K"], wl["IJ"]
    B = args.maps
    S = wl["S"]
    if args.layout == "pixel_major":
        S = S.transpose(1, 2).contiguous().transpose(1, 2)     # [B,R,IJ] view of [B,IJ,R] storage
    Cf = wl["C"]
    nll = torch.empty(B, dtype=torch.float64, device=dev)
    gS = torch.empty_strided(S.shape, S.stride(), dtype=torch.float32, device=dev)
    gC = torch.empty_like(Cf)
    view = obs.view()
    stream = torch.cuda.current_stream()

    def step():
        check(lib.qmc_nll_fwd_bwd_gather(S.data_ptr(), S.stride(0), S.stride(1), S.stride(2), Cf.data_ptr(),
                                         C.byref(view), C.byref(lik), B, IJ, K, R, _lib.QMC_ALGO_AUTO,
                                         obs.tile_warps, nll.data_ptr(), gS.data_ptr(), gC.data_ptr(),
                                         stream.cuda_stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    # clocks: nvidia-smi needs a few hundred ms to start and samples every 100 ms, the timed region is a few ms.
    # The sampler therefore starts here and the same step runs back to back for ~0.7 s (an extended warm-up, so that
    # the samples are taken under exactly the timed load) straight into the timed region.
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 0.7:
        for _ in range(50):
            step()
        torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches0 = _lib.launch_count()
    barrier()
    ev[0].record(stream)
    for i in range(args.steps):
        step()
        ev[i + 1].record(stream)
    barrier()
    launches = _lib.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks, units summed over ranks
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    n = torch.tensor([float(obs.nobs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    total_ms = t.item()
    nobs_all = n.item()
    value = nobs_all * args.steps / (total_ms * 1e-3)

    # ---- the same step with S in the reference's own layout (emitter-major [R][IJ]) ---------------------------
    em_ms = None
    if not args.no_extra and args.layout == "pixel_major":
        S_em = wl["S"]
        gS_em = torch.empty_like(S_em)

        def step_em():
            check(lib.qmc_nll_fwd_bwd_gather(S_em.data_ptr(), S_em.stride(0), S_em.stride(1), S_em.stride(2), Cf.data_ptr(),
                                             C.byref(view), C.byref(lik), B, IJ, K, R, _lib.QMC_ALGO_AUTO,
                                             obs.tile_warps, nll.data_ptr(), gS_em.data_ptr(), gC.data_ptr(),
                                             stream.cuda_stream))
        for _ in range(args.warmup):
            step_em()
        em = torch.tensor([_timed(step_em, args.steps, stream, barrier)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(em, op=dist.ReduceOp.MAX)
        em_ms = em.item()
        del gS_em

    # ---- end to end through the host-buffer C-ABI call -----------------------------------------
    e2e = None
    if not args.no_e2e:
        Sh = wl["S"].cpu().pin_memory()
        Ch = wl["C"].cpu().pin_memory()
        gSh = torch.empty_like(Sh).pin_memory()
        gCh = torch.empty_like(Ch).pin_memory()
        nllh = torch.empty(B, dtype=torch.float64).pin_memory()
        Sd, gSd = torch.empty_like(wl["S"]), torch.empty_like(wl["S"])
        Cd, gCd = torch.empty_like(Cf), torch.empty_like(Cf)

        def e2e_step():
            check(lib.qmc_nll_fwd_bwd_gather_host(
                Sh.data_ptr(), Ch.data_ptr(), Sd.data_ptr(), Cd.data_ptr(), C.byref(view), C.byref(lik), B, IJ, K, R,
                _lib.QMC_ALGO_AUTO, obs.tile_warps, nll.data_ptr(), gSd.data_ptr(), gCd.data_ptr(), nllh.data_ptr(),
                gSh.data_ptr(), gCh.data_ptr(), stream.cuda_stream))

        for _ in range(max(3, min(args.warmup, 5))):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()                     # synchronises the stream itself: results are on the host
        barrier()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = Sh.numel() * 4 + Ch.numel() * 4
        d2h = gSh.numel() * 4 + gCh.numel() * 4 + nllh.numel() * 8
        e2e = {"value": nobs_all * args.steps / te.item(), "unit": "observed-entries/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * te.item() / args.steps,
               "h2d_GBps_per_rank": h2d * args.steps / dt / 1e9, "d2h_GBps_per_rank": d2h * args.steps / dt / 1e9,
               "host_binding": numa,
               "api": "qmc_nll_fwd_bwd_gather_host (C ABI, pinned host S/C in, nll/gS/gC out; observation set resident)"}
        assert torch.isfinite(nllh).all()

    # ---- solver iterations/s (second half of the BASELINE metric) ---------------------------------
    solver = None
    if not args.no_solver:
        from quantized_spectrum_cartography_b200 import qmc
        n_it = 40
        cfg = qmc.SolverConfig(iters=n_it, lam_c=1.0, lam_s=1.0, track_every=0, cuda_graph=True)
        res = qmc.solve_lowrank_fused(wl["S"], wl["C"], obs, lik, cfg)
        ts = torch.tensor([res.seconds], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        solver = {"iterations_per_s": n_it / ts.item(), "maps_per_gpu": B, "map_iterations_per_s": n_it * B * world / ts.item(),
                  "iteration": "C-step + S-step: 2 fused evaluations + 2 fused updates (Frobenius-regulariser gradient, "
                               "Adam, projection, next norm); one CUDA graph replayed",
                  "observed_entries_per_s": 2 * nobs_all * n_it / ts.item()}

    # ---- the two multi-GPU configurations BASELINE.json names, and the dense path ------------------------------
    def reduce_max(x):
        tt = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt.item()

    def reduce_sum(x):
        tt = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        return tt.item()

    strong = cfg4 = None
    if not args.no_extra:
        del gS, gC, nll
        torch.cuda.empty_cache()
        strong = strong_scaling_record(args, rank, world, dev, barrier, reduce_max, reduce_sum) if world > 1 else None
        cfg4 = cfg4_record(args, rank, world, dev, barrier, reduce_max)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline -----------------------------------------------------------------------------------
    peak, peak_src = measured_hbm_peak()
    alg_bytes = obs.algorithmic_bytes(R)
    mean_launch_ms = statistics.mean(per_launch_ms)
    achieved = alg_bytes / (mean_launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "gather_lanes_kernel<4,ONEBIT>" if obs.lanes else "gather_tiled_kernel<4,ONEBIT>",
                "algorithmic_bytes_per_launch": alg_bytes, "mean_launch_ms": mean_launch_ms,
                "frac_of_nominal_8TBs": achieved / 8000.0}

    roofline_em = None
    if em_ms is not None:
        roofline_em = {"bound": "hbm", "achieved": alg_bytes / (em_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                       "frac": alg_bytes / (em_ms * 1e-3) / 1e9 / peak, "mean_launch_ms": em_ms,
                       "S_layout": "emitter_major (the reference's [R,1,I,J]; slices transposed through registers: coalesced row loads, 16-byte shared-memory stores)"}

    cpu = None
    if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only
        v, nmaps, secs, kind = cpu_reference_sample(wl["host_sample"], wl["I"], wl["J"], K, R, args.cpu_seconds)
        vf, nmaps_f, secs_f, _ = cpu_reference_sample(wl["host_sample"], wl["I"], wl["J"], K, R,
                                                      min(4.0, args.cpu_seconds), vectorised=True)
        import torch as _t
        cpu = {"value": v, "unit": "observed-entries/s", "cores": _t.get_num_threads(), "kind": kind,
               "sample": f"{nmaps} map evaluations (fwd+bwd) cycling over the first 256 of {B} maps, {secs:.1f} s of CPU work",
               "value_vectorised_fair_cpu": vf}

    c = CFG3
    out = {
        "metric": "QMC observed-entries/s (fused NLL fwd + gS + gC)", "value": value, "unit": "observed-entries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_keys(B),
        "layout": {"observed_entries_per_gpu": obs.nobs, "S_layout": args.layout, "tile_warps": obs.tile_warps,
                   "tiles_per_map": obs.n_sub // obs.tile_warps, "obs_layout": "lanes" if obs.lanes else "rows",
                   "obs_word_bits": obs.word_bits if obs.lanes else None, "obs_padding": round(obs.padding_fraction(), 4),
                   "threshold": wl["thr"], "sigma": wl["sigma"]},
        "roofline": roofline, "roofline_emitter_major": roofline_em, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "solver": solver, "strong": strong, ("dense_cfg4" if world == 1 else "sharded_cfg4"): cfg4,
        "evaluations_per_s": args.steps / (total_ms * 1e-3) * world,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # not launched by torchrun: do it ourselves
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
