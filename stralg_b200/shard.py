"""Read sharding for batched exact search over a REPLICATED index (BASELINE configs[3]).

The reference maps one read at a time (tools/readmappers/bwt_readmapper/bwt_readmapper.c:128-161);
reads are independent, so the batched path splits them into contiguous shards, one per rank
(one process per GPU), every rank searches its shard against its own copy of the index, and the
fixed-width (L, R) pairs end up on rank 0: either stored there directly by the search kernels
through NVLink peer memory, or moved by ONE gather collective.  No collective sits on the data
path of the search itself.

Nothing here computes: ``search_fn`` is the engine's device search (``SuffixArrayIndex.search_device``)
in the product; the world_size-2 ``gloo`` tests on CPU plug the oracle in instead to check the
partition / gather logic.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of ``total`` reads for ``rank``: sizes differ by at most one,
    the first ``total % world`` ranks take the extra read."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, extra = divmod(int(total), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_shard(total: int, world: int) -> int:
    return -(-int(total) // world)


class ShardedSearch:
    """Searches this rank's shard and delivers all (L, R) pairs on ``dst``.

    ``dist`` is ``torch.distributed`` (initialised, NCCL on GPUs / gloo in the CPU tests) or None
    for a single process.  Buffers are allocated once; ``step()`` is what a bench step times.

    Two transports for the one exchange of the path:

    * ``"p2p"`` (GPUs of one NVLink / NVSwitch node): the result array lives in symmetric memory;
      every rank's search kernel stores its (L, R) pairs STRAIGHT into ``dst``'s HBM through the
      peer mapping, so the transfer rides along with the kernel and nothing is left to gather --
      one device-side barrier tells ``dst`` that all peers are done.
    * ``"gather"``: the shard's pairs land in local memory and ONE ``gather`` collective moves
      them (the CPU / gloo tests, and the fallback when peer mapping is unavailable).  The shard
      can be cut in ``chunks`` pieces whose gathers are issued asynchronously.
    """

    def __init__(self, total_reads: int, read_len: int, device, dist=None, dst: int = 0, chunks: int = 1,
                 transport: str = "auto"):
        import torch
        self.torch = torch
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.dst = dst
        self.total = int(total_reads)
        self.m = int(read_len)
        self.lo, self.hi = shard_bounds(self.total, self.world, self.rank)
        self.count = self.hi - self.lo
        self.cap = max_shard(self.total, self.world)  # gather needs equal-size pieces
        self.transport = "local" if dist is None else "gather"
        self.hdl = None
        if dist is not None and transport in ("auto", "p2p") and str(device).startswith("cuda") and self.total:
            self._try_p2p(device, strict=(transport == "p2p"))
        if self.transport == "p2p":
            return
        self.chunks = max(1, min(int(chunks), self.cap)) if self.cap else 1
        # piece k covers shard-local reads [k * piece, (k + 1) * piece), the same split on every rank
        self.piece = -(-self.cap // self.chunks) if self.cap else 0
        self.LR = [torch.zeros((2, self.piece), dtype=torch.int32, device=device) for _ in range(self.chunks)]
        self.gathered = None
        if dist is not None and self.rank == dst:
            self.gathered = [[torch.empty((2, self.piece), dtype=torch.int32, device=device)
                              for _ in range(self.world)] for _ in range(self.chunks)]

    def _try_p2p(self, device, strict: bool) -> None:
        """Symmetric allocation of the (2, total) result array + peer view of ``dst``'s copy.
        All ranks must succeed, else everyone uses the gather transport."""
        torch, dist = self.torch, self.dist
        ok = 1
        try:
            import torch.distributed._symmetric_memory as symm_mem
            self._sym = symm_mem.empty(2 * self.total, dtype=torch.int32, device=device)
            self._sym.zero_()
            self.hdl = symm_mem.rendezvous(self._sym, dist.group.WORLD)
            self._dst_view = self.hdl.get_buffer(self.dst, (2, self.total), torch.int32)
        except Exception as ex:  # no peer access / API missing
            if strict:
                raise
            self._p2p_error = str(ex)[:200]
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            self.transport = "p2p"
            self.chunks = 1
        else:
            self.hdl = None

    def piece_bounds(self, k: int, count: int) -> Tuple[int, int]:
        lo = min(k * self.piece, count)
        return lo, min(lo + self.piece, count)

    def step(self, search_fn: Callable, reads) -> None:
        """search_fn(reads, read_len, count, L_out, R_out) fills the first ``count`` slots;
        ``reads`` holds this rank's shard (count * read_len codes)."""
        if self.transport == "p2p":
            if self.count:
                search_fn(reads, self.m, self.count, self._dst_view[0, self.lo:self.hi],
                          self._dst_view[1, self.lo:self.hi])
            self.hdl.barrier()  # device-side: dst's stream proceeds once every peer's kernel has finished
            return
        works = []
        for k in range(self.chunks):
            lo, hi = self.piece_bounds(k, self.count)
            if hi > lo:
                search_fn(reads[lo * self.m: hi * self.m], self.m, hi - lo, self.LR[k][0], self.LR[k][1])
            if self.dist is not None:
                works.append(self.dist.gather(self.LR[k], self.gathered[k] if self.gathered else None,
                                              dst=self.dst, async_op=True))
        for w in works:
            w.wait()

    def local_result(self):
        """(L, R) of this rank's shard (after a step; with p2p read back from ``dst``'s array)."""
        torch = self.torch
        if self.transport == "p2p":
            return self._dst_view[0, self.lo:self.hi].clone(), self._dst_view[1, self.lo:self.hi].clone()
        Ls, Rs = [], []
        for k in range(self.chunks):
            lo, hi = self.piece_bounds(k, self.count)
            Ls.append(self.LR[k][0, :hi - lo])
            Rs.append(self.LR[k][1, :hi - lo])
        return torch.cat(Ls), torch.cat(Rs)

    def result(self) -> Optional[Tuple["object", "object"]]:
        """(L, R) of all reads in input order on ``dst`` (int32 bit patterns of the uint32 values);
        None on the other ranks."""
        torch = self.torch
        if self.dist is None:
            return self.local_result()
        if self.rank != self.dst:
            return None
        if self.transport == "p2p":
            v = self._sym.view(2, self.total)
            return v[0], v[1]
        Ls, Rs = [], []
        for g in range(self.world):
            glo, ghi = shard_bounds(self.total, self.world, g)
            for k in range(self.chunks):
                lo, hi = self.piece_bounds(k, ghi - glo)
                Ls.append(self.gathered[k][g][0, :hi - lo])
                Rs.append(self.gathered[k][g][1, :hi - lo])
        return torch.cat(Ls), torch.cat(Rs)
