"""Id-sharded index over the GPUs of one box (north_star subsystem 4). One process per GPU.

Rows are partitioned by id: global id g lives on rank ``g % world`` as local id ``g // world`` (the
reference's ids are dense insertion counters, hnsw.zig:77, so round-robin keeps shards balanced
under streaming insert). Every rank holds an independent per-shard index -- built by the reference's
own insert on that shard's rows in global-id order -- and searches the FULL query batch on it; the
per-shard top-k are then exchanged once and merged by (distance, global id):

  * ``exchange="p2p"``  (default on CUDA): ONE kernel launch per step. The search kernel's epilogue
    stores each shard's top-k straight into every peer's gather buffer over NVLink (CUDA IPC
    mappings) as 128-byte self-validating records (payload + epoch in one warp-wide store, atomic
    over NVLink: no flag store, no fence); one wave later the same kernel merges each query whose
    records have all arrived -- no collective call, no second launch
    (zvdb_search_batch_exchange);
  * ``exchange="p2pb"``: the same one-launch step with the results travelling as per-rank blocks plus
    per-query release flags instead of self-validating 128-byte records (the first fused form, A/B);
  * ``exchange="p2p3"``: round 1's form of the same exchange as three launches (search with peer
    stores, a flag kernel, a merge kernel that waits on the flags), kept for A/B;
  * ``exchange="nccl"``: search into a packed block, ONE ``all_gather_into_tensor`` of the blocks,
    then the merge kernel (the formulation north_star states; also the baseline the fused path is
    measured against).

The oracle for this path is G independent reference indexes plus the same merge; it is NOT one
reference index over all rows (different graphs).

The compute hooks (``backend``) default to the CUDA library. Tests inject a CPU backend to exercise
this host logic (sharding, id mapping, block layout, gather order) under ``gloo`` without a GPU;
the product has no CPU backend of its own.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L
from .hnsw import HNSW


def block_bytes(nq: int, k: int) -> int:
    """Size of one packed per-shard result block (matches zvdb_shard_block_bytes)."""
    return (nq * k * 12 + nq * 4 + 255) // 256 * 256


def unpack_block(block: np.ndarray, nq: int, k: int):
    """(ids u64[nq,k], dist f32[nq,k], counts u32[nq]) views of a packed block (uint8 array)."""
    b = np.ascontiguousarray(block, np.uint8)
    ids = b[: nq * k * 8].view(np.uint64).reshape(nq, k)
    dist = b[nq * k * 8: nq * k * 12].view(np.float32).reshape(nq, k)
    counts = b[nq * k * 12: nq * k * 12 + nq * 4].view(np.uint32)
    return ids, dist, counts


def pack_block(ids, dist, counts) -> np.ndarray:
    nq, k = ids.shape
    b = np.zeros(block_bytes(nq, k), np.uint8)
    b[: nq * k * 8] = np.ascontiguousarray(ids, np.uint64).view(np.uint8).reshape(-1)
    b[nq * k * 8: nq * k * 12] = np.ascontiguousarray(dist, np.float32).view(np.uint8).reshape(-1)
    b[nq * k * 12: nq * k * 12 + nq * 4] = np.ascontiguousarray(counts, np.uint32).view(np.uint8).reshape(-1)
    return b


def shard_rows(n_total: int, rank: int, world: int) -> np.ndarray:
    """Global ids owned by `rank`, ascending: rank, rank + world, ..."""
    return np.arange(rank, n_total, world, dtype=np.int64)


def per_shard_ef(ef: int, k: int, world: int) -> int:
    """Pop budget of one shard when a total budget of `ef` pops is split over `world` shards."""
    return max(k, -(-ef // world))


class CudaBackend:
    """Shard-local compute through libzvdb_b200.so on this rank's GPU."""

    def __init__(self, index: HNSW, rank: int, world: int):
        import torch
        self.torch = torch
        self.index = index
        self.rank, self.world = rank, world
        self.device = torch.device("cuda", index.device)
        self.exchange = C.c_void_p()

    # -- nccl formulation -----------------------------------------------------------------------
    def search_packed(self, d_queries, nq: int, k: int, ef: int):
        torch = self.torch
        block = torch.empty(block_bytes(nq, k), dtype=torch.uint8, device=self.device)
        L.check(L.lib().zvdb_search_batch_packed_device(self.index._h, d_queries.data_ptr(), nq, k, ef, block.data_ptr(),
                                                        self.world, self.rank, torch.cuda.current_stream().cuda_stream))
        return block

    def merge_packed(self, gathered, nq: int, k: int):
        torch = self.torch
        ids = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        dist = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        counts = torch.empty(nq, dtype=torch.int32, device=self.device)
        L.check(L.lib().zvdb_merge_topk_packed_device(gathered.data_ptr(), self.world, nq, k, dist.data_ptr(), ids.data_ptr(),
                                                      counts.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return ids, dist, counts

    def to_device(self, queries: np.ndarray):
        return self.torch.from_numpy(np.ascontiguousarray(queries, np.float32)).to(self.device)

    def to_host(self, ids, dist, counts):
        return (ids.cpu().numpy().view(np.uint64), dist.cpu().numpy(), counts.cpu().numpy().view(np.uint32))

    # -- fused formulation ------------------------------------------------------------------------
    def open_exchange(self, nq_max: int, k_max: int, group, dim_max: int = 0) -> None:
        """dim_max > 0 also reserves the query slices of the host step (search_exchange_host)."""
        import torch.distributed as dist
        torch = self.torch
        L.check(L.lib().zvdb_exchange_create_host(C.byref(self.exchange), self.index.device, self.world, self.rank, nq_max, k_max, dim_max))
        mine = (C.c_uint8 * 64)()
        L.check(L.lib().zvdb_exchange_ipc_handle(self.exchange, mine))
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, bytes(mine), group=group)
        else:
            handles[0] = bytes(mine)
        blob = (C.c_uint8 * (64 * self.world)).from_buffer_copy(b"".join(handles))
        L.check(L.lib().zvdb_exchange_open_peers(self.exchange, blob))
        self._cap = (nq_max, k_max, dim_max)
        if self.world > 1:
            dist.barrier(group=group)

    def search_exchange(self, d_queries, nq: int, k: int, ef: int, out=None):
        torch = self.torch
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.int64, device=self.device),
                   torch.empty((nq, k), dtype=torch.float32, device=self.device),
                   torch.empty(nq, dtype=torch.int32, device=self.device))
        ids, dist, counts = out
        L.check(L.lib().zvdb_search_batch_exchange(self.index._h, self.exchange, d_queries.data_ptr(), nq, k, ef,
                                                   ids.data_ptr(), dist.data_ptr(), counts.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream))
        return ids, dist, counts

    def search_exchange_host(self, q_ptr: int, nq: int, dim: int, k: int, ef: int, ids_ptr: int, dist_ptr: int, counts_ptr: int):
        """zvdb_search_batch_exchange_host on raw HOST addresses: this rank copies in its slice of the batch and writes
        its slice of the merged results (gather to owner); asynchronous on the current stream."""
        L.check(L.lib().zvdb_search_batch_exchange_host(self.index._h, self.exchange, q_ptr, nq, dim, k, ef, ids_ptr, dist_ptr,
                                                        counts_ptr, self.torch.cuda.current_stream().cuda_stream))

    def close(self) -> None:
        if self.exchange:
            L.lib().zvdb_exchange_destroy(self.exchange)
            self.exchange = C.c_void_p()


class ShardedHNSW:
    """`HNSW` over `world` id-shards; every method is collective (call it on all ranks)."""

    def __init__(self, m: int = 16, ef_construction: int = 200, *, metric: int = L.METRIC_L2, rank: Optional[int] = None,
                 world: Optional[int] = None, device: Optional[int] = None, group=None, exchange: str = "p2p",
                 backend=None):
        import torch.distributed as dist
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        if exchange not in ("p2p", "p2pb", "p2p3", "nccl"):
            raise ValueError("exchange must be 'p2p', 'p2pb', 'p2p3' or 'nccl'")
        self.exchange = exchange
        self.m = m
        self.n_total = 0
        if backend is None:
            self.index = HNSW(m, ef_construction, metric=metric, device=self.rank if device is None else device)
            if exchange == "p2p3":
                self.index.set_kernel_variant(0x1000)          # bit 12: the sharded step as three launches
            if exchange == "p2pb":
                self.index.set_kernel_variant(0x2000)          # bit 13: fused step through result blocks + release flags
            self.backend = CudaBackend(self.index, self.rank, self.world)
        else:
            self.index = None
            self.backend = backend
        self._exchange_open = False

    # -- ids --------------------------------------------------------------------------------------
    def owner(self, gid: int) -> int:
        return gid % self.world

    def local_id(self, gid: int) -> int:
        return gid // self.world

    def global_id(self, local: int) -> int:
        return local * self.world + self.rank

    def count(self) -> int:
        return self.n_total

    # -- insert -------------------------------------------------------------------------------------
    def insert_batch(self, points) -> None:
        """Append rows with global ids n_total, n_total+1, ...; this rank keeps the ones it owns."""
        p = np.ascontiguousarray(points, np.float32)
        first = self.n_total
        start = (self.rank - first) % self.world          # first row of `points` owned by this rank
        mine = p[start::self.world]
        if len(mine):
            (self.index.insert_batch if self.index is not None else self.backend.insert_batch)(mine)
        self.n_total += len(p)

    # -- search -------------------------------------------------------------------------------------
    def search_batch_device(self, d_queries, nq: int, k: int, ef: int = 0, ef_per_shard: Optional[int] = None):
        """Merged top-k of the whole index for nq queries resident on this rank's device.
        Returns device tensors (ids int64 [nq,k] holding global ids, dist [nq,k], counts [nq])."""
        import torch.distributed as dist
        ef = ef or k
        e = ef_per_shard if ef_per_shard is not None else per_shard_ef(ef, k, self.world)
        if self.exchange in ("p2p", "p2pb", "p2p3") and hasattr(self.backend, "search_exchange"):
            self.ensure_exchange(nq, k)
            return self.backend.search_exchange(d_queries, nq, k, e)
        block = self.backend.search_packed(d_queries, nq, k, e)
        if self.world == 1:
            gathered = block
        else:
            gathered = block.new_empty(self.world * block.numel())
            dist.all_gather_into_tensor(gathered, block, group=self.group)     # the ONE collective of the path
        return self.backend.merge_packed(gathered, nq, k)

    def ensure_exchange(self, nq: int, k: int, dim: int = 0) -> None:
        """(Re)open the peer-mapped exchange so that it holds nq x k results per rank and, for the host step, query slices
        of `dim` floats. Collective: every rank must call it with the same arguments."""
        cap = getattr(self.backend, "_cap", (0, 0, 0))
        if not self._exchange_open or cap[0] < nq or cap[1] < k or cap[2] < dim:
            self.backend.close()
            self.backend.open_exchange(max(nq, cap[0] if self._exchange_open else 0, 1), max(k, cap[1] if self._exchange_open else 0), self.group,
                                       dim_max=max(dim, cap[2] if self._exchange_open else 0))
            self._exchange_open = True

    def search_batch(self, queries, k: int, ef: int = 0, ef_per_shard: Optional[int] = None):
        """Host arrays in, host arrays out: (ids u64 [nq,k] global, dist f32 [nq,k], counts u32 [nq])."""
        q = np.ascontiguousarray(queries, np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        out = self.search_batch_device(self.backend.to_device(q), q.shape[0], k, ef, ef_per_shard)
        return self.backend.to_host(*out)

    def search_batch_host_slice(self, queries, k: int, ef: int = 0, ef_per_shard: Optional[int] = None):
        """The host step (gather to owner): every rank passes the SAME host batch; rank r gets back
        (lo, hi, ids[lo:hi], dist[lo:hi], counts[lo:hi]) -- its slice of the merged top-k of the whole index. Every query
        and every result row crosses PCIe once over the whole job. CUDA backend only."""
        import torch
        q = np.ascontiguousarray(queries, np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        nq, dim = q.shape
        ef = ef or k
        e = ef_per_shard if ef_per_shard is not None else per_shard_ef(ef, k, self.world)
        self.ensure_exchange(nq, k, dim)
        ids = np.full((nq, k), 0xFFFFFFFFFFFFFFFF, np.uint64)
        dist = np.zeros((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        self.backend.search_exchange_host(q.ctypes.data, nq, dim, k, e, ids.ctypes.data, dist.ctypes.data, counts.ctypes.data)
        torch.cuda.current_stream().synchronize()
        per = -(-nq // self.world)
        lo, hi = min(nq, per * self.rank), min(nq, per * (self.rank + 1))
        return lo, hi, ids[lo:hi], dist[lo:hi], counts[lo:hi]

    def deinit(self) -> None:
        if hasattr(self.backend, "close"):
            self.backend.close()
        if self.index is not None:
            self.index.deinit()
