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
    ``[gS | gC | nll]`` buffer, 4*(R*IJ + R*K + 2) bytes (16.8 MB at cfg4); the tcgen05 kernel writes its pixel
    block of gS, its gC and its NLL straight into that persistent buffer;
  - ``"pixel_block"`` -- all-reduce only ``[gC | nll]`` (16 KB at cfg4) and, if the caller wants the
    full gS everywhere, all-gather the disjoint gS slices.  With ``exchange="peer"`` (dense tcgen05 kernel, one
    node, at most 8 ranks) there is no collective call at all: the kernel's last CTA pushes the rank's partial
    ``[gC | nll]`` into every peer's exchange region with plain stores over NVLink and sums the peers' slots
    (``qmc_nll_fwd_bwd_dense_exchange``, csrc/qmc_dense.cu).

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
def _cuda_local_eval(S3, C3, obs, lik, want_grad=True, out=None):
    """Product evaluator: the gather kernels for an ObsSet, the tcgen05 kernel for a DenseObs."""
    from .dense import DenseObs, nll_fwd_bwd_dense
    from .fused import nll_fwd_bwd
    if isinstance(obs, DenseObs):
        nll, gS, gC = nll_fwd_bwd_dense(S3[0], C3[0], obs, lik, want_grad=want_grad)
        return nll.reshape(1), None if gS is None else gS.unsqueeze(0), None if gC is None else gC.unsqueeze(0)
    return nll_fwd_bwd(S3, C3, obs, lik, want_grad=want_grad, out=out)


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

    def evaluate(self, S3, C3, want_grad: bool = True, out=None):
        """NLL ``[hi-lo]`` (fp64), gS, gC of this rank's maps.  No communication.  ``out=(nll, gS, gC)`` reuses
        buffers (CUDA evaluator only)."""
        if out is not None:
            return self.local_eval(S3, C3, self.obs, self.lik, want_grad, out=out)
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
    exchange: str = "nccl"                    # "nccl" (torch.distributed all-reduce) | "peer" (fused into the dense kernel)
    _peers: object = None
    _buf: Optional[torch.Tensor] = None      # persistent exchange buffer (fp32): [gS | gC | nll_hi, nll_lo]
    _nll64: Optional[torch.Tensor] = None
    _graph: object = None
    _S_in: Optional[torch.Tensor] = None     # staging copies of the inputs the captured graph reads
    _C_in: Optional[torch.Tensor] = None

    @classmethod
    def from_dense(cls, Y, Wx, K: int, R: int, lik, *, mode: str = "flat", device=None, align: int = 1,
                   dense: Optional[bool] = None, build: Optional[Callable] = None,
                   local_eval: Callable = _cuda_local_eval, exchange: str = "nccl") -> "ShardedInstance":
        """``Y``/``Wx``: the full instance ``[K, IJ]`` (reference layout ``[K,1,I,J]`` accepted); each
        rank keeps only its pixel block.  ``dense``: use the tcgen05 dense kernel for the local block
        (default: when the geometry is supported and at least a quarter of the entries are observed)."""
        if mode not in ("flat", "pixel_block"):
            raise ValueError(mode)
        if exchange not in ("nccl", "peer") or (exchange == "peer" and mode != "pixel_block"):
            raise ValueError(f"exchange={exchange!r} with mode={mode!r}: the fused exchange carries [gC | nll] only")
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
        if exchange == "peer":
            from .dense import DenseObs
            if local_eval is not _cuda_local_eval or not isinstance(obs, DenseObs):
                raise ValueError('exchange="peer" is part of the dense tcgen05 kernel: it needs a supported geometry '
                                 "and the CUDA evaluator")
        return cls(IJ, K, R, lo, hi, obs, lik, mode, local_eval, align, exchange)

    def flat_size(self) -> int:
        """Elements of the contract-form exchange buffer [gS | gC | nll] (the NLL travels as two fp32 words,
        high and low part of the fp64 partial sum, so that the fp32 all-reduce loses nothing of it)."""
        return self.R * self.IJ + self.R * self.K + 2

    def exchange_bytes(self) -> int:
        return 4 * (self.flat_size() if self.mode == "flat" else self.R * self.K + 2)

    def exchange_status(self) -> int:
        """Fused exchange only: 0 = fine, 1 = a peer never arrived in some evaluation (results are partial sums)."""
        return 0 if self._peers is None else self._peers.status()

    def close(self):
        """Release the exchange regions (all ranks, after their last evaluation)."""
        if self._peers is not None:
            torch.cuda.synchronize()
            _, world = _world()
            if world > 1:
                dist.barrier()
            self._peers.close()
            self._peers = None
            self._graph = None

    # ---- one evaluation: local kernel into the persistent buffer, one collective ---------------------------
    def _local_into(self, buf, Sl, C, off_gs):
        """Run the local evaluation with its outputs inside ``buf`` where the kernel can write there directly
        (tcgen05 dense kernel: gS rows with the buffer's row stride, gC in place); otherwise copy."""
        R, K, IJ, n = self.R, self.K, self.IJ, self.hi - self.lo
        gC_view = buf[off_gs: off_gs + R * K].view(R, K)
        tail = buf[off_gs + R * K: off_gs + R * K + 2]
        direct = False
        if self.local_eval is _cuda_local_eval:
            from .dense import DenseObs, nll_fwd_bwd_dense
            if isinstance(self.obs, DenseObs):
                gS_view = (buf[: R * IJ].view(R, IJ)[:, self.lo:self.hi] if self.mode == "flat" else
                           buf[: R * n].view(R, n))
                nll_fwd_bwd_dense(Sl, C, self.obs, self.lik, out=(self._nll64, gS_view, gC_view), peers=self._peers)
                direct = True
        if not direct:
            nll, gSl, gC = self.local_eval(Sl.unsqueeze(0), C.reshape(1, R, K), self.obs, self.lik, True)
            if self.mode == "flat":
                buf[: R * IJ].view(R, IJ)[:, self.lo:self.hi].copy_(gSl.reshape(R, -1))
            else:
                buf[: R * n].view(R, n).copy_(gSl.reshape(R, -1))
            gC_view.copy_(gC.reshape(R, K))
            self._nll64.copy_(nll.reshape(1).to(torch.float64))
        if self._peers is not None or _world()[1] == 1:
            return                           # the global sums are in place already: nothing travels through `tail`
        hi32 = self._nll64.to(torch.float32)
        tail[0:1].copy_(hi32)
        tail[1:2].copy_((self._nll64 - hi32.to(torch.float64)).to(torch.float32))

    def _step(self, Sl, C):
        rank, world = _world()
        R, K, IJ, n = self.R, self.K, self.IJ, self.hi - self.lo
        buf = self._buf
        if self.mode == "flat":
            buf.zero_()                      # the other ranks' pixel blocks must not carry the previous sum
            self._local_into(buf, Sl, C, R * IJ)
            if world > 1:
                dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        else:
            self._local_into(buf, Sl, C, R * n)
            if world > 1 and self._peers is None:
                dist.all_reduce(buf[R * n:], op=dist.ReduceOp.SUM)

    def evaluate(self, S, C, gather_gS: bool = True, cuda_graph: bool = False):
        """``S [R, IJ]`` (full, replicated) or ``[R, hi-lo]`` (this rank's block), ``C [R, K]``
        replicated.  Returns (nll 0-dim fp64, gS, gC [R, K]) where gS is ``[R, IJ]`` (complete on
        every rank) in "flat" mode or with ``gather_gS``; otherwise this rank's ``[R, hi-lo]``.  The returned
        tensors are views of a persistent buffer that the next call overwrites.  ``cuda_graph``: capture the
        local kernel and the collective once and replay them (the inputs are copied into staging buffers; pass
        ``inst._S_in`` / ``inst._C_in`` themselves to skip the copy).  All ranks must make the same sequence of
        calls with the same ``cuda_graph`` flags."""
        rank, world = _world()
        R, K, IJ, n = self.R, self.K, self.IJ, self.hi - self.lo
        S2 = S.reshape(R, -1)
        Sl = S2[:, self.lo:self.hi] if S2.shape[1] == IJ else S2
        if Sl.shape[1] != n:
            raise ValueError("S has neither the full nor the local pixel extent")
        Sl = Sl.contiguous()
        C = C.reshape(R, K).contiguous()
        dev = C.device
        size = self.flat_size() if self.mode == "flat" else R * n + R * K + 2
        if self._buf is None or self._buf.device != dev or self._buf.numel() != size:
            self._buf = torch.zeros(size, dtype=torch.float32, device=dev)
            self._nll64 = torch.zeros(1, dtype=torch.float64, device=dev)
            self._graph = None
        if self.exchange == "peer" and self._peers is None:
            from .dense import PeerRegions
            self._peers = PeerRegions(rank, world, R * K + 2, dev)   # collective: the ranks swap IPC handles
        if cuda_graph and dev.type == "cuda":
            # The graph is captured ONCE per instance, on staging buffers the instance owns: whether a rank
            # re-captures must not depend on where its allocator happened to put the caller's tensors (a rank that
            # re-captures issues one collective more than a rank that replays: deadlock).
            if self._graph is None:
                self._S_in, self._C_in = Sl.clone(), C.clone()
                self._step(self._S_in, self._C_in)   # warm-up outside capture (lazy initialisation of kernels and NCCL)
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step(self._S_in, self._C_in)
                self._graph = g
            if Sl.data_ptr() != self._S_in.data_ptr():
                self._S_in.copy_(Sl)
            if C.data_ptr() != self._C_in.data_ptr():
                self._C_in.copy_(C)
            self._graph.replay()
        else:
            self._step(Sl, C)
        buf = self._buf
        off = R * IJ if self.mode == "flat" else R * n
        gC_all = buf[off: off + R * K].view(R, K)
        if self._peers is not None or world == 1:
            nll = self._nll64[0].clone()
        else:
            nll = buf[off + R * K].to(torch.float64) + buf[off + R * K + 1].to(torch.float64)
        if self.mode == "flat":
            return nll, buf[: R * IJ].view(R, IJ), gC_all
        gSl = buf[: R * n].view(R, n)
        if not gather_gS or world == 1:
            return nll, gSl, gC_all         # (with one rank the local block is the whole map)
        blocks = [partition_pixels(IJ, world, r, self.align) for r in range(world)]
        width = max(h - l for l, h in blocks)
        pad = torch.zeros(R, width, dtype=torch.float32, device=dev)
        pad[:, :n] = gSl
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        gS_all = torch.cat([p[:, : h - l] for p, (l, h) in zip(parts, blocks)], dim=1)
        return nll, gS_all, gC_all
