#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU, NCCL): one large instance sharded by pixel
blocks, factor gradients combined by an NCCL all-reduce, against the single-GPU evaluation of the
whole instance.  Also times one sharded evaluation of the cfg4 shape.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_sharded_nccl.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import quantized_spectrum_cartography_b200 as q  # noqa: E402
from quantized_spectrum_cartography_b200 import _lib, dense, parallel, synth  # noqa: E402
from quantized_spectrum_cartography_b200.quantization_model import assign_levels  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}
    for name, (I, J, K, R, f) in {"small": (96, 100, 64, 6, 0.5), "cfg4": (512, 512, 256, 16, 0.5)}.items():
        IJ = I * J
        maps = synth.generate_maps(1, I, J, K, R, seed=0, device=dev)      # same seed: replicated inputs
        T = maps.tensor()[0]
        gen = torch.Generator(device=dev).manual_seed(1)
        off = float(T.median()) * 0.1
        X = torch.log(T + off)
        bb = synth.equal_mass_boundaries(X, 8)
        sigma = float((bb[1:] - bb[:-1]).min()) * 2.0
        Y = assign_levels(X + sigma * torch.randn(X.shape, device=dev, generator=gen), bb)
        Wx = torch.bernoulli(torch.full(T.shape, f, device=dev), generator=gen)
        lik = q.make_likelihood(bb, sigma, offset=off)
        S = (0.8 * maps.S_true[0]).contiguous()
        C = maps.C_true[0].contiguous()
        # single-GPU truth (every rank computes it; cheap)
        obs_all = q.build_obs(Y, Wx, K, IJ, 1)
        ref = q.nll_fwd_bwd(S.unsqueeze(0), C.unsqueeze(0), obs_all, lik, algo=_lib.QMC_ALGO_FLAT)
        for use_dense in (False, True):
            for mode in ("flat", "pixel_block"):
                inst = parallel.ShardedInstance.from_dense(Y, Wx, K, R, lik, mode=mode, align=128, dense=use_dense)
                nll, gS, gC = inst.evaluate(S, C)
                e = dict(nll=abs(nll.item() / ref[0][0].item() - 1),
                         gS=float((gS - ref[1][0]).norm() / ref[1][0].norm()),
                         gC=float((gC - ref[2][0]).norm() / ref[2][0].norm()))
                assert e["nll"] < 1e-5 and e["gS"] < 1e-4 and e["gC"] < 1e-4, (name, use_dense, mode, e)
                # timing: barrier, K evaluations, max over ranks
                for _ in range(3):
                    inst.evaluate(S, C, gather_gS=(mode == "flat"))
                dist.barrier(); torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                n_it = 10
                for _ in range(n_it):
                    inst.evaluate(S, C, gather_gS=(mode == "flat"))
                b.record()
                dist.barrier(); torch.cuda.synchronize()
                t = torch.tensor([a.elapsed_time(b) / n_it], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                out[f"{name}_{'dense' if use_dense else 'gather'}_{mode}"] = dict(
                    err=e, ms_per_eval=t.item(), entries_per_s=obs_all.nobs / (t.item() * 1e-3),
                    exchange_bytes=4 * inst.flat_size() if mode == "flat" else 8 * (R * K + 1))
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
