"""Two GPUs, NCCL: the sharded single instance (BASELINE configs[3]) with the CUDA evaluators -- the
tcgen05 dense kernel writing straight into the all-reduce buffer, and the observed-entry kernel -- against the
single-GPU evaluation of the whole instance; and the batched maps partitioned over the ranks.  Skipped with fewer
than two devices (the driver's 1-GPU box); run by `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import quantized_spectrum_cartography_b200 as q
        from quantized_spectrum_cartography_b200 import _lib, dense, parallel, synth
        from quantized_spectrum_cartography_b200.quantization_model import assign_levels
        I, J, K, R = 96, 100, 64, 6
        IJ = I * J
        maps = synth.generate_maps(1, I, J, K, R, seed=0, device=dev)            # same seed: replicated inputs
        T = maps.tensor()[0]
        gen = torch.Generator(device=dev).manual_seed(1)
        off = float(T.median()) * 0.1
        X = torch.log(T + off)
        bb = synth.equal_mass_boundaries(X, 8)
        sigma = float((bb[1:] - bb[:-1]).min()) * 2.0
        Y = assign_levels(X + sigma * torch.randn(X.shape, device=dev, generator=gen), bb)
        Wx = torch.bernoulli(torch.full(T.shape, 0.5, device=dev), generator=gen)
        lik = q.make_likelihood(bb, sigma, offset=off)
        S = (0.8 * maps.S_true[0]).contiguous()
        C = maps.C_true[0].contiguous()
        ref = q.nll_fwd_bwd(S.unsqueeze(0), C.unsqueeze(0), q.build_obs(Y, Wx, K, IJ, 1), lik, algo=_lib.QMC_ALGO_FLAT)
        before = _lib.launch_count()
        for use_dense in (True, False):
            for mode in ("flat", "pixel_block"):
                for graph in (False, True):
                    inst = parallel.ShardedInstance.from_dense(Y, Wx, K, R, lik, mode=mode, align=128, dense=use_dense)
                    assert isinstance(inst.obs, dense.DenseObs) == use_dense
                    for _ in range(2):                                               # the second call replays / reuses buffers
                        nll, gS, gC = inst.evaluate(S, C, cuda_graph=graph)
                    assert abs(nll.item() / ref[0][0].item() - 1) < 1e-5
                    assert float((gS - ref[1][0]).norm() / ref[1][0].norm()) < 1e-4
                    assert float((gC - ref[2][0]).norm() / ref[2][0].norm()) < 1e-4
                    # local pixel block in, local gS block out (what a solver with a sharded S would use)
                    nll2, gS_loc, gC2 = inst.evaluate(S[:, inst.lo:inst.hi], C, gather_gS=False)
                    want = ref[1][0][:, inst.lo:inst.hi]
                    got = gS_loc if mode == "pixel_block" else gS_loc[:, inst.lo:inst.hi]
                    assert float((got - want).norm() / want.norm()) < 1e-4
                    assert abs(nll2.item() / ref[0][0].item() - 1) < 1e-5
        assert _lib.launch_count() > before
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
        bm = parallel.BatchedMaps.from_dense(Yb[lo:hi].cuda(), Wb[lo:hi].cuda(), 64, 4, lik2, B, tile_warps=4,
                                             build=lambda y, w: q.make_obs(y, w, 64, y.device, B=y.shape[0], R=4, tiled=True,
                                                                           tile_warps=4, lanes=True))
        nll, gS, gC = bm.evaluate(Sb[lo:hi].cuda(), Cb[lo:hi].cuda())
        full = q.nll_fwd_bwd(Sb.cuda(), Cb.cuda(), q.build_obs(Yb.cuda(), Wb.cuda(), 64, 35 * 31, B), lik2, algo=_lib.QMC_ALGO_FLAT)
        np.testing.assert_allclose(nll.cpu().numpy(), full[0][lo:hi].cpu().numpy(), rtol=1e-6)
        assert float((gS - full[1][lo:hi]).norm() / full[1][lo:hi].norm()) < 2e-5
        all_nll = bm.gather_nll(nll)
        if rank == 0:
            np.testing.assert_allclose(all_nll.cpu().numpy(), full[0].cpu().numpy(), rtol=1e-6)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_instance_and_partitioned_maps_on_two_gpus(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
