"""Multi-GPU scheduling of the likelihood path: one process per GPU, ``torch.distributed``.

Two cases, matching how the path shards:

* **Batched independent maps** (cfg3, cfg5): every map has its own S, C, Y, Wx and its own NLL, so
  the batch is cut into contiguous chunks, one per rank, and *no collective* touches the data path
  (:func:`partition_maps`, :class:`BatchedMaps`).  Only per-map scalars are gathered, on request.

* **One oversized instance** (cfg4): NLL and both factor gradients are sums over observed entries,
  so the entries are sharded and one exchange step per evaluation combines the partial sums
  (:class:`ShardedInstance`).  Entries are sharded by contiguous *pixel blocks*: rank g owns pixels
  ``[lo_g, hi_g)`` for all K bands, hence its rows of gS are complete locally and only ``gC`` and
  the scalar NLL are partial.  Two exchange forms:

  - ``"flat"``  -- the contract form BASELINE.json names: one all-reduce (sum, fp32) of the flat
    ``[gS | gC | nll]`` buffer, 4*(R*IJ + R*K + 1) bytes (16.8 MB at cfg4);
  - ``"pixel_block"`` -- all-reduce only ``[gC | nll]`` (16 KB at cfg4) and, if the caller wants the
    full gS everywhere, all-gather the disjoint gS slices.

The local evaluation is injectable (``local_eval``) so that the host-side logic -- partitioning,
packing, the collective, unpacking -- runs under ``gloo`` on CPU in the tests with the checker as
the local evaluator; the product default is the CUDA kernel and nothing else.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.distributed as dist


def partition_maps(n_maps: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous chunk ``[lo, hi)`` of rank ``rank``: sizes differ by at most one, earlier ranks
    take the remainder."""
    base, rem = divmod(n_maps, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def partition_pixels(IJ: int, world: int, rank: int, align: int = 1) -> tuple[int, int]:
    """Contiguous pixel block of rank ``rank`` (block starts are multiples of ``align``)."""
    blocks = -(-IJ // align)
    lo_b, hi_b = partition_maps(blocks, world, rank)
    return min(lo_b * align, IJ), min(hi_b * align, IJ)


def _world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


# -------------------------------------------------------------------------------------------------
# batched independent maps: no collective
# -------------------------------------------------------------------------------------------------
def _cuda_local_eval(S3, C3, obs, lik, want_grad=True):
    """Product evaluator: the gather kernels for an ObsSet, the tcgen05 kernel for a DenseObs."""
    from .dense import DenseObs, nll_fwd_bwd_dense
    from .fused import nll_fwd_bwd
    if isinstance(obs, DenseObs):
        nll, gS, gC = nll_fwd_bwd_dense(S3[0], C3[0], obs, lik, want_grad=want_grad)
        return nll.reshape(1), None if gS is None else gS.unsqueeze(0), None if gC is None else gC.unsqueeze(0)
    return nll_fwd_bwd(S3, C3, obs, lik, want_grad=want_grad)


@dataclass
class BatchedMaps:
    """This rank's share of a batch of independent maps."""
    lo: int
    hi: int
    n_maps: int
    obs: object
    lik: object
    local_eval: Callable = _cuda_local_eval

    @classmethod
    def from_dense(cls, Y, Wx, K: int, R: int, lik, n_maps: int, *, device=None, tile_warps: int = 8,
                   build: Optional[Callable] = None, local_eval: Callable = _cuda_local_eval) -> "BatchedMaps":
        """``Y``/``Wx``: this rank's maps only, ``[hi-lo, K, IJ]`` (each rank loads or synthesises
        its own chunk; nothing is broadcast)."""
        rank, world = _world()
        lo, hi = partition_maps(n_maps, world, rank)
        if Y.shape[0] != hi - lo:
            raise ValueError(f"rank {rank} owns maps [{lo},{hi}) but was given {Y.shape[0]} maps")
        if build is None:
            from .fused import make_obs
            obs = make_obs(Y, Wx, K, device, B=hi - lo, R=R, tiled=True, tile_warps=tile_warps)
        else:
            obs = build(Y, Wx)
        return cls(lo, hi, n_maps, obs, lik, local_eval)

    def evaluate(self, S3, C3, want_grad: bool = True):
        """NLL ``[hi-lo]`` (fp64), gS, gC of this rank's maps.  No communication."""
        return self.local_eval(S3, C3, self.obs, self.lik, want_grad)

    def gather_nll(self, nll_local: torch.Tensor) -> Optional[torch.Tensor]:
        """Per-map NLL of the whole batch on rank 0 (None elsewhere): the only thing that ever
        leaves a device in the batched case."""
        rank, world = _world()
        if world == 1:
            return nll_local
        sizes = [partition_maps(self.n_maps, world, r) for r in range(world)]
        width = max(h - l for l, h in sizes)
        pad = torch.zeros(width, dtype=nll_local.dtype, device=nll_local.device)
        pad[: nll_local.numel()] = nll_local
        out = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, out, dst=0)
        if rank != 0:
            return None
        return torch.cat([o[: h - l] for o, (l, h) in zip(out, sizes)])


# -------------------------------------------------------------------------------------------------
# one oversized instance: entries sharded by pixel block, one exchange step per evaluation
# -------------------------------------------------------------------------------------------------
@dataclass
class ShardedInstance:
    IJ: int
    K: int
    R: int
    lo: int
    hi: int
    obs: object
    lik: object
    mode: str = "flat"
    local_eval: Callable = _cuda_local_eval
    align: int = 1

    @classmethod
    def from_dense(cls, Y, Wx, K: int, R: int, lik, *, mode: str = "flat", device=None, align: int = 1,
                   dense: Optional[bool] = None, build: Optional[Callable] = None,
                   local_eval: Callable = _cuda_local_eval) -> "ShardedInstance":
        """``Y``/``Wx``: the full instance ``[K, IJ]`` (reference layout ``[K,1,I,J]`` accepted); each
        rank keeps only its pixel block.  ``dense``: use the tcgen05 dense kernel for the local block
        (default: when the geometry is supported and at least a quarter of the entries are observed)."""
        if mode not in ("flat", "pixel_block"):
            raise ValueError(mode)
        rank, world = _world()
        Yk = Y.reshape(K, -1)
        IJ = Yk.shape[1]
        lo, hi = partition_pixels(IJ, world, rank, align)
        Yl = Yk[:, lo:hi].contiguous()
        Wl = None if Wx is None else Wx.reshape(K, -1)[:, lo:hi].contiguous()
        if build is None:
            from .dense import dense_supported, pack_dense
            from .fused import make_obs
            if dense is None:
                dense = dense_supported(K, R) and Wl is not None and float(Wl.float().mean()) >= 0.25
            obs = pack_dense(Yl, Wl, K) if dense else make_obs(Yl, Wl, K, device, B=1, R=R, tiled=False)
        else:
            obs = build(Yl, Wl)
        return cls(IJ, K, R, lo, hi, obs, lik, mode, local_eval, align)

    def flat_size(self) -> int:
        return self.R * self.IJ + self.R * self.K + 1

    def evaluate(self, S, C, gather_gS: bool = True):
        """``S [R, IJ]`` (full, replicated) or ``[R, hi-lo]`` (this rank's block), ``C [R, K]``
        replicated.  Returns (nll 0-dim fp64, gS, gC [R, K]) where gS is ``[R, IJ]`` (complete on
        every rank) in "flat" mode or with ``gather_gS``; otherwise this rank's ``[R, hi-lo]``."""
        rank, world = _world()
        R, K, IJ = self.R, self.K, self.IJ
        S2 = S.reshape(R, -1)
        Sl = S2[:, self.lo:self.hi] if S2.shape[1] == IJ else S2
        if Sl.shape[1] != self.hi - self.lo:
            raise ValueError("S has neither the full nor the local pixel extent")
        nll, gSl, gC = self.local_eval(Sl.contiguous().unsqueeze(0), C.reshape(1, R, K).contiguous(), self.obs,
                                       self.lik, True)
        gSl, gC, nll = gSl.reshape(R, -1), gC.reshape(R, K), nll.reshape(())
        if self.mode == "flat":
            # the contract form: one fp32 all-reduce of [gS | gC | nll]
            flat = torch.zeros(self.flat_size(), dtype=torch.float32, device=gC.device)
            flat[: R * IJ].view(R, IJ)[:, self.lo:self.hi] = gSl
            flat[R * IJ: R * IJ + R * K] = gC.reshape(-1)
            flat[-1] = nll.to(torch.float32)
            if world > 1:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            return (flat[-1].to(torch.float64), flat[: R * IJ].view(R, IJ), flat[R * IJ: R * IJ + R * K].view(R, K))
        # pixel-block form: only gC and nll are partial sums (nll kept in fp64)
        small = torch.empty(R * K + 1, dtype=torch.float64, device=gC.device)
        small[: R * K] = gC.reshape(-1).to(torch.float64)
        small[-1] = nll
        if world > 1:
            dist.all_reduce(small, op=dist.ReduceOp.SUM)
        gC_all = small[: R * K].to(torch.float32).view(R, K)
        if not gather_gS or world == 1:
            return small[-1], gSl, gC_all       # (with one rank the local block is the whole map)
        blocks = [partition_pixels(IJ, world, r, self.align) for r in range(world)]
        width = max(h - l for l, h in blocks)
        pad = torch.zeros(R, width, dtype=torch.float32, device=gC.device)
        pad[:, : gSl.shape[1]] = gSl
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        gS_all = torch.cat([p[:, : h - l] for p, (l, h) in zip(parts, blocks)], dim=1)
        return small[-1], gS_all, gC_all
