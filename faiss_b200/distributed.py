"""Database sharding across processes (one process per GPU) with IndexShards semantics.

Reference: faiss::IndexShards (faiss/IndexShards.cpp:172-264) splits the database contiguously,
sends every query to every shard and heap-merges the per-shard top-k on the host
(merge_knn_results, faiss/utils/Heap.cpp:166-238) after one D2H copy per shard.  Here each rank
owns one shard on its GPU; the per-shard [nq, k] results are exchanged with ONE all-gather
(NCCL over NVLink on GPUs, gloo on CPU tensors) and merged on every rank by the device merge
kernel (b200_topk_merge) -- or by the library's host merge when the tensors live on the CPU
(host-logic tests under gloo).  Ids are translated like successive_ids=True: global id =
local id + number of vectors in lower-ranked shards.
"""
import ctypes

import numpy as np

from . import _c_f, _c_i64, _ptr, check, lib, topk_merge


def shard_bounds(n, rank, world):
    """Contiguous split i0 = rank*n/world (faiss/IndexShards.cpp:172-190)."""
    return rank * n // world, (rank + 1) * n // world


def merge_host(all_D, all_I, k, metric):
    """[nshard, n, kin] numpy arrays -> merged [n, k] with the library's host merge."""
    all_D = np.ascontiguousarray(all_D, dtype=np.float32)
    all_I = np.ascontiguousarray(all_I, dtype=np.int64)
    ns, n, kin = all_D.shape
    assert kin == k
    D = np.empty((n, k), dtype=np.float32)
    I = np.empty((n, k), dtype=np.int64)
    check(
        lib.faiss_b200_merge_knn_results_host(
            ctypes.c_int64(n), ctypes.c_int64(k), int(ns), int(metric), _ptr(all_D, _c_f), _ptr(all_I, _c_i64),
            _ptr(D, _c_f), _ptr(I, _c_i64),
        )
    )
    return D, I


class ShardedSearcher:
    """One shard per rank.  `local_search(xq, k) -> (D, I)` runs the rank's sub-index (ids local to
    the shard); `ntotal_local` is the shard size.  `group` is a torch.distributed process group."""

    def __init__(self, local_search, ntotal_local, metric, res=None, group=None, device=0):
        import torch
        import torch.distributed as dist

        self.local_search = local_search
        self.metric = metric
        self.res = res
        self.group = group
        self.device = device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        # id translation table = exclusive prefix sum of shard sizes (successive_ids)
        sizes = [None] * self.world
        dist.all_gather_object(sizes, int(ntotal_local), group=group)
        self.sizes = sizes
        self.offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        self._torch = torch
        self._dist = dist
        self._off_dev = None

    @property
    def ntotal(self):
        return int(sum(self.sizes))

    def search(self, xq, k):
        """xq: identical on every rank (torch tensor, CUDA or CPU).  Returns merged (D, I) on every rank."""
        torch, dist = self._torch, self._dist
        D, I = self.local_search(xq, k)
        if not torch.is_tensor(D):
            D = torch.from_numpy(D)
            I = torch.from_numpy(I)
        nq = D.shape[0]
        # ONE all-gather: distances (fp32) and labels (int64) travel in one byte message per rank
        nD, nI = nq * k * 4, nq * k * 8
        send = torch.empty(nD + nI, dtype=torch.uint8, device=D.device)
        send[:nD] = D.contiguous().view(torch.uint8).view(-1)
        send[nD:] = I.contiguous().view(torch.uint8).view(-1)
        recv = torch.empty((self.world, nD + nI), dtype=torch.uint8, device=D.device)
        dist.all_gather_into_tensor(recv.view(-1), send, group=self.group)
        allD = recv[:, :nD].contiguous().view(torch.float32).view(self.world, nq, k)
        allI = recv[:, nD:].contiguous().view(torch.int64).view(self.world, nq, k)
        if D.is_cuda:
            if self._off_dev is None:
                self._off_dev = torch.from_numpy(self.offsets).to(D.device)
            # device merge kernel wants [nq, nshard, k]
            mD, mI = topk_merge(
                self.res,
                allD.permute(1, 0, 2).contiguous(),
                allI.permute(1, 0, 2).contiguous(),
                k,
                self.metric,
                id_offsets=self._off_dev,
                device=self.device,
            )
            return mD, mI
        aI = allI.numpy().copy()
        for s in range(self.world):
            m = aI[s] >= 0
            aI[s][m] += self.offsets[s]
        mD, mI = merge_host(allD.numpy(), aI, k, self.metric)
        return torch.from_numpy(mD), torch.from_numpy(mI)
