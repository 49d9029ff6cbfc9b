"""world_size-2 gloo test of the sharded search host logic (IndexShards semantics): contiguous
split, successive-id translation, one all-gather, merge == unsharded result."""
import os
import socket
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from faiss_b200.distributed import ShardedSearcher, shard_bounds
    from oracle import oracle_np as o

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rs = np.random.RandomState(0)
    N, d, nq, k = 1501, 16, 23, 7
    xb = np.floor(rs.rand(N, d) * 8).astype(np.float32)  # integer regime: exact, with ties
    xq = np.floor(rs.rand(nq, d) * 8).astype(np.float32)
    i0, i1 = shard_bounds(N, rank, world)

    def local_search(x, kk):
        D, I = o.knn_flat(x.numpy(), xb[i0:i1], kk, 1)  # stands in for the rank's GPU sub-index
        return D, I

    s = ShardedSearcher(local_search, i1 - i0, 1)
    D, I = s.search(torch.from_numpy(xq), k)
    rD, rI = o.knn_flat(xq, xb, k, 1)
    ok = bool(np.array_equal(I.numpy(), rI) and np.array_equal(D.numpy(), rD) and s.ntotal == N)

    # inner product (descending order, ties by id) through the same exchange
    def local_search_ip(x, kk):
        return o.knn_flat(x.numpy(), xb[i0:i1], kk, 0)

    sip = ShardedSearcher(local_search_ip, i1 - i0, 0)
    D, I = sip.search(torch.from_numpy(xq), k)
    rD, rI = o.knn_flat(xq, xb, k, 0)
    ok = ok and bool(np.array_equal(I.numpy(), rI) and np.array_equal(D.numpy(), rD))

    # a database smaller than k: shards pad with id -1 (Heap.cpp:200,224 skips them), the merged tail
    # is -1 / +inf-like exactly where the unsharded index has no result
    small = xb[:9]
    j0, j1 = shard_bounds(9, rank, world)

    def local_search_small(x, kk):
        return o.knn_flat(x.numpy(), small[j0:j1], kk, 1)

    ss = ShardedSearcher(local_search_small, j1 - j0, 1)
    D, I = ss.search(torch.from_numpy(xq), 12)
    rD, rI = o.knn_flat(xq, small, 12, 1)
    ok = ok and bool(np.array_equal(I.numpy(), rI) and np.array_equal(D.numpy()[:, :9], rD[:, :9]) and (I.numpy()[:, 9:] == -1).all())
    q.put((rank, ok))
    dist.destroy_process_group()


def test_sharded_search_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def _kmeans_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from faiss_b200.distributed import shard_bounds, sharded_kmeans
    from oracle import oracle_np as o

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rs = np.random.RandomState(5)
    N, d, k, niter = 3001, 8, 37, 6
    # integer-valued rows: partial sums are exact in fp32 whatever the reduction order, so the sharded
    # run must reproduce the single-process algorithm; a far-away blob forces empty clusters -> split_clusters
    x = np.floor(rs.rand(N, d) * 16).astype(np.float32)
    x[:40] += 500.0
    i0, i1 = shard_bounds(N, rank, world)

    def local_assign(cent, xl):
        D, I = o.knn_flat(xl.numpy(), cent.numpy(), 1, 1)
        return torch.from_numpy(D[:, 0].copy()), torch.from_numpy(I[:, 0].copy())

    def local_accumulate(xl, assign, kk):
        sums = np.zeros((kk, xl.shape[1]), dtype=np.float64)
        np.add.at(sums, assign.numpy(), xl.numpy().astype(np.float64))
        counts = np.bincount(assign.numpy(), minlength=kk).astype(np.float32)
        return torch.from_numpy(sums.astype(np.float32)), torch.from_numpy(counts)

    cent, objs = sharded_kmeans(torch.from_numpy(x[i0:i1]), k, niter, local_assign, local_accumulate, seed=1234)
    rc, robj = o.kmeans(x, k, niter=niter, seed=1234, max_points_per_centroid=1 << 20)
    ok = bool(np.allclose(cent.numpy(), rc, rtol=1e-6, atol=1e-6) and np.allclose(objs, robj, rtol=1e-6))
    q.put((rank, ok, float(np.abs(cent.numpy() - rc).max())))
    dist.destroy_process_group()


def test_sharded_kmeans_gloo_world2():
    """training set split over 2 ranks, all-reduce of sums | counts | objective per iteration ==
    the single-process reference algorithm on the concatenated set (incl. split_clusters)"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_kmeans_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
