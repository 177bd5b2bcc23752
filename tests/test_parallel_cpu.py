"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU paths -- partitioning,
packing, the collective, unpacking -- with the oracle injected as the local evaluator (the product
default is the CUDA kernel; nothing here is a product fallback)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quantized_spectrum_cartography_b200 import parallel as par


def test_partitions_cover_without_overlap():
    for n, w in ((4096, 8), (4096, 3), (7, 8), (1, 2), (262144, 8)):
        spans = [par.partition_maps(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    for IJ, w, al in ((10201, 4, 1), (262144, 8, 128), (2601, 2, 32), (100, 8, 64)):
        spans = [par.partition_pixels(IJ, w, r, al) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == IJ
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(l % al == 0 for l, h in spans if h > l)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_eval(S3, C3, obs, lik, want_grad=True):
    """Local evaluator for the tests: the float64 oracle, map by map."""
    from oracle import qmc_oracle as oc
    Y, Wx, bb, sigma = obs
    B, R, IJ = S3.shape
    K = C3.shape[2]
    nll = torch.empty(B, dtype=torch.float64)
    gS = torch.empty(B, R, IJ)
    gC = torch.empty(B, R, K)
    for b in range(B):
        n, gs, gc, _ = oc.nll_and_grads_fp64(S3[b].reshape(R, 1, 1, IJ), C3[b], Y[b].reshape(K, 1, 1, IJ),
                                             Wx[b].reshape(K, 1, 1, IJ), bb, sigma)
        nll[b] = n
        gS[b] = torch.from_numpy(gs.reshape(R, IJ)).float()
        gC[b] = torch.from_numpy(gc).float()
    return nll, gS, gC


def _instance(B, IJ, K, R, seed):
    from oracle import qmc_oracle as oc
    g = torch.Generator().manual_seed(seed)
    S = torch.rand(B, R, IJ, generator=g) * 0.1 + 0.01
    C = torch.rand(B, R, K, generator=g) * 0.2 + 0.02
    T = torch.einsum("brp,brk->bkp", S, C)
    bb = torch.linspace(T.min().item(), T.max().item(), 5)
    sigma = 0.5 * (bb[1] - bb[0]).item()
    Y = oc.assign_levels(T + sigma * torch.randn(T.shape, generator=g), bb)
    Wx = torch.bernoulli(torch.full(T.shape, 0.4), generator=g)
    return S, C, Y, Wx, bb, sigma


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- one instance, entries sharded by pixel block -------------------------------------
        S, C, Y, Wx, bb, sigma = _instance(1, 37, 6, 3, seed=5)
        want = _oracle_eval(S, C, (Y, Wx, bb, sigma), None)
        for mode in ("flat", "pixel_block"):
            inst = par.ShardedInstance.from_dense(
                Y[0], Wx[0], 6, 3, None, mode=mode, align=4,
                build=lambda Yl, Wl: (Yl.unsqueeze(0), Wl.unsqueeze(0), bb, sigma), local_eval=_oracle_eval)
            assert (inst.lo, inst.hi) == par.partition_pixels(37, world, rank, 4)
            nll, gS, gC = inst.evaluate(S[0], C[0])
            assert nll.item() == pytest.approx(want[0][0].item(), rel=1e-6)
            np.testing.assert_allclose(gS.numpy(), want[1][0].numpy(), rtol=1e-5, atol=1e-4 * want[1].abs().max().item())
            np.testing.assert_allclose(gC.numpy(), want[2][0].numpy(), rtol=1e-4, atol=1e-4 * want[2].abs().max().item())
            if mode == "pixel_block":
                # local S block in, local gS block out, only gC/nll exchanged
                nll2, gS_loc, gC2 = inst.evaluate(S[0][:, inst.lo:inst.hi], C[0], gather_gS=False)
                np.testing.assert_allclose(gS_loc.numpy(), want[1][0][:, inst.lo:inst.hi].numpy(), rtol=1e-5,
                                           atol=1e-4 * want[1].abs().max().item())
                assert nll2.item() == pytest.approx(want[0][0].item(), rel=1e-6)
        # ---- batched maps: contiguous chunks, no collective on the data path ----------------------
        B = 5
        S, C, Y, Wx, bb, sigma = _instance(B, 23, 4, 2, seed=9)
        want = _oracle_eval(S, C, (Y, Wx, bb, sigma), None)
        lo, hi = par.partition_maps(B, world, rank)
        bm = par.BatchedMaps.from_dense(Y[lo:hi], Wx[lo:hi], 4, 2, None, B,
                                        build=lambda Yl, Wl: (Yl, Wl, bb, sigma), local_eval=_oracle_eval)
        nll, gS, gC = bm.evaluate(S[lo:hi], C[lo:hi])
        np.testing.assert_allclose(nll.numpy(), want[0][lo:hi].numpy(), rtol=1e-12)
        np.testing.assert_allclose(gS.numpy(), want[1][lo:hi].numpy(), rtol=1e-6)
        all_nll = bm.gather_nll(nll)
        if rank == 0:
            np.testing.assert_allclose(all_nll.numpy(), want[0].numpy(), rtol=1e-12)
        else:
            assert all_nll is None
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_and_batched_paths_world_size_2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_fused_exchange_is_refused_where_it_cannot_run():
    """exchange="peer" is part of the dense tcgen05 kernel: it carries [gC | nll] only (pixel_block mode) and needs the
    CUDA evaluator on a dense observation set; anything else is refused when the instance is built, not at run time."""
    Y = torch.zeros(8, 12, dtype=torch.int64)
    Wx = torch.ones(8, 12)
    lik = object()
    with pytest.raises(ValueError, match="carries"):
        par.ShardedInstance.from_dense(Y, Wx, 8, 2, lik, mode="flat", exchange="peer")
    with pytest.raises(ValueError, match="exchange"):
        par.ShardedInstance.from_dense(Y, Wx, 8, 2, lik, mode="pixel_block", exchange="nvshmem")
    with pytest.raises(ValueError, match="dense tcgen05"):   # an injected evaluator / a non-dense observation set
        par.ShardedInstance.from_dense(Y, Wx, 8, 2, lik, mode="pixel_block", exchange="peer", build=lambda y, w: (y, w),
                                       local_eval=lambda *a, **k: None)
    assert par.ShardedInstance.from_dense(Y, Wx, 8, 2, lik, mode="pixel_block", build=lambda y, w: (y, w),
                                          local_eval=lambda *a, **k: None).exchange_status() == 0
