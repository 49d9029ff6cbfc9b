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

from . import _c_f, _c_i64, _ptr, check, lib, rand_perm, split_clusters, topk_merge


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


def sharded_kmeans(x_local, k, niter, local_assign, local_accumulate, seed=1234, group=None):
    """k-means with the training set sharded over the ranks (contiguous, rank order) and the centroid
    table replicated: faiss::Clustering::train (faiss/Clustering.cpp:60-380) with the per-iteration
    reduction of SURVEY 8(e) -- all-reduce(sum) of the per-centroid partial sums, counts and the
    objective.  Everything else follows the single-process algorithm on the concatenated set:
    initial centroids = rows rand_perm(n_total, seed + 1)[:k], empty clusters refilled by the
    reference's deterministic split_clusters on every rank identically.

    x_local:          this rank's rows, torch tensor [n_local, d] (CUDA under NCCL, CPU under gloo)
    local_assign:     (centroids, x_local) -> (sqdist [n_local], assign int64 [n_local])   -- Flat k=1 search
    local_accumulate: (x_local, assign, k) -> (sums [k, d], counts [k]) float32, same device as x_local
    Returns (centroids [k, d] tensor on x_local's device, objective per iteration as numpy float32).
    Subsampling to max_points_per_centroid is the caller's job (each rank passes the rows it wants used)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local, d = x_local.shape
    dev = x_local.device
    sizes = [None] * world
    dist.all_gather_object(sizes, int(n_local), group=group)
    n_total = int(sum(sizes))
    off = int(sum(sizes[:rank]))
    assert n_total >= k, "need at least k training points"
    # ---- initial centroids: the global rows perm[:k]; every row is owned by exactly one rank
    perm = rand_perm(n_total, seed + 1)[:k].astype(np.int64)
    mine = (perm >= off) & (perm < off + n_local)
    cent = torch.zeros((k, d), dtype=torch.float32, device=dev)
    if mine.any():
        rows = torch.from_numpy(perm[mine] - off).to(dev)
        cent[torch.from_numpy(np.nonzero(mine)[0]).to(dev)] = x_local[rows]
    dist.all_reduce(cent, op=dist.ReduceOp.SUM, group=group)
    objs = []
    for _ in range(niter):
        dis, assign = local_assign(cent, x_local)
        sums, counts = local_accumulate(x_local, assign, k)
        # one packed reduction per iteration: [k, d] sums | k counts | objective
        packed = torch.cat([sums.reshape(-1), counts.reshape(-1), dis.double().sum().float().reshape(1)])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        sums = packed[: k * d].reshape(k, d)
        counts = packed[k * d : k * d + k]
        objs.append(float(packed[-1]))
        nz = counts > 0
        new = torch.zeros_like(cent)
        new[nz] = sums[nz] * (1.0 / counts[nz]).unsqueeze(1)
        # empty clusters: the reference's split_clusters on the host (k x d floats), identical on every rank
        if bool((~nz).any()):
            h = counts.cpu().numpy().astype(np.float32)
            c = new.cpu().numpy()
            split_clusters(h, c, n_total)
            new = torch.from_numpy(c).to(dev)
        cent = new
    return cent, np.array(objs, dtype=np.float32)
