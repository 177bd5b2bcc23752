#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU, NCCL): one large instance sharded by pixel
blocks, factor gradients combined by an NCCL all-reduce (contract form and pixel-block form, eager and as one
CUDA graph, tcgen05 dense kernel and observed-entry kernel), against the single-GPU evaluation of the whole
instance; and a batch of independent maps partitioned over the ranks.  Prints one JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_sharded_nccl.py [--cfg4]
"""
import datetime
import json
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def run():
    import quantized_spectrum_cartography_b200 as q
    from quantized_spectrum_cartography_b200 import _lib, dense, parallel, synth
    from quantized_spectrum_cartography_b200.quantization_model import assign_levels
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    out = {"world": world}
    shapes = {"small": (96, 100, 64, 6, 0.5)}
    if "--cfg4" in sys.argv:
        shapes["cfg4"] = (512, 512, 256, 16, 0.5)
    for name, (I, J, K, R, f) in shapes.items():
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
        obs_all = q.build_obs(Y, Wx, K, IJ, 1)
        ref = q.nll_fwd_bwd(S.unsqueeze(0), C.unsqueeze(0), obs_all, lik, algo=_lib.QMC_ALGO_FLAT)   # single-GPU truth
        for use_dense, mode, exchange in ((True, "flat", "nccl"), (True, "pixel_block", "nccl"), (True, "pixel_block", "peer"),
                                          (False, "flat", "nccl"), (False, "pixel_block", "nccl")):
            if True:
                for graph in (False, True):
                    inst = parallel.ShardedInstance.from_dense(Y, Wx, K, R, lik, mode=mode, align=128, dense=use_dense,
                                                               exchange=exchange)
                    assert isinstance(inst.obs, dense.DenseObs) == use_dense
                    for _ in range(2):                                      # the second call replays / reuses buffers
                        nll, gS, gC = inst.evaluate(S, C, cuda_graph=graph)
                    e = dict(nll=abs(nll.item() / ref[0][0].item() - 1),
                             gS=float((gS - ref[1][0]).norm() / ref[1][0].norm()),
                             gC=float((gC - ref[2][0]).norm() / ref[2][0].norm()))
                    assert e["nll"] < 1e-5 and e["gS"] < 1e-4 and e["gC"] < 1e-4, (name, use_dense, mode, graph, e)
                    # local pixel block in, local gS block out (what a solver with a sharded S would use)
                    nll2, gS_loc, _ = inst.evaluate(S[:, inst.lo:inst.hi], C, gather_gS=False, cuda_graph=graph)
                    want = ref[1][0][:, inst.lo:inst.hi]
                    got = gS_loc if mode == "pixel_block" else gS_loc[:, inst.lo:inst.hi]
                    assert float((got - want).norm() / want.norm()) < 1e-4
                    assert abs(nll2.item() / ref[0][0].item() - 1) < 1e-5
                    tag = f"{name}_{'dense' if use_dense else 'gather'}_{mode}_{'graph' if graph else 'eager'}"
                    if exchange == "peer":
                        # the kernel's own exchange over NVLink peer memory: no collective call; every rank must hold
                        # the same bits (slots are added in rank order), across repeated evaluations (epoch parity)
                        for _ in range(5):
                            nll3, _, gC3 = inst.evaluate(S[:, inst.lo:inst.hi], C, gather_gS=False, cuda_graph=graph)
                        assert inst.exchange_status() == 0
                        mine = torch.cat([gC3.reshape(-1).double(), nll3.reshape(1)])
                        both = [torch.empty_like(mine) for _ in range(world)]
                        dist.all_gather(both, mine)
                        assert all(torch.equal(b, both[0]) for b in both), "ranks disagree bitwise"
                        assert float((gC3 - ref[2][0]).norm() / ref[2][0].norm()) < 1e-4
                        inst.close()
                        tag += "_peer"
                    out[tag] = e
    # batched maps: contiguous chunks, no collective on the data path
    B = 6
    lo, hi = parallel.partition_maps(B, world, rank)
    g = torch.Generator().manual_seed(3)
    Sb = torch.rand(B, 4, 35 * 31, generator=g) * 0.1 + 0.01
    Cb = torch.rand(B, 4, 64, generator=g) * 0.2 + 0.02
    Tb = torch.einsum("brp,brk->bkp", Sb, Cb)
    thr = Tb.median().item()
    bb2 = torch.tensor([0.0, thr, 1.0])
    Yb = (Tb + 0.5 * thr * torch.randn(Tb.shape, generator=g) > thr).to(torch.uint8)
    Wb = torch.bernoulli(torch.full(Tb.shape, 0.2), generator=g)
    lik2 = q.make_likelihood(bb2, 0.5 * thr)
    bm = parallel.BatchedMaps.from_dense(
        Yb[lo:hi].to(dev), Wb[lo:hi].to(dev), 64, 4, lik2, B, tile_warps=4,
        build=lambda y, w: q.make_obs(y, w, 64, y.device, B=y.shape[0], R=4, tiled=True, tile_warps=4, lanes=True))
    nll, gS, gC = bm.evaluate(Sb[lo:hi].to(dev), Cb[lo:hi].to(dev))
    full = q.nll_fwd_bwd(Sb.to(dev), Cb.to(dev), q.build_obs(Yb.to(dev), Wb.to(dev), 64, 35 * 31, B), lik2, algo=_lib.QMC_ALGO_FLAT)
    e_nll = float(((nll - full[0][lo:hi]).abs() / full[0][lo:hi].abs()).max())
    e_gs = float((gS - full[1][lo:hi]).norm() / full[1][lo:hi].norm())
    assert e_nll < 1e-6 and e_gs < 2e-5, (e_nll, e_gs)
    all_nll = bm.gather_nll(nll)
    if rank == 0:
        assert float(((all_nll - full[0]).abs() / full[0].abs()).max()) < 1e-6
        out["partitioned_maps"] = {"nll": e_nll, "gS": e_gs}
        out["launches"] = _lib.launch_count()
        print(json.dumps(out))


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    try:
        run()
    except BaseException:
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)     # fail fast: a rank that raises must not leave its peers waiting in a collective
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
