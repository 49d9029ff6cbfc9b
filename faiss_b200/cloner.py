"""Cloner helpers at payload level: the host format either side of the hot path.

The reference clones a CPU index to GPUs with `index_cpu_to_gpu[_multiple]` (faiss/gpu/GpuCloner.cpp:
124-498): coarse centroids, PQ centroids and every `ArrayInvertedLists` list (codes bytes + int64 ids)
are copied verbatim; with `GpuMultipleClonerOptions.shard` the inverted lists are split over the GPUs
by `shard_type` (GpuCloner.cpp:287-322, `IndexIVF::copy_subset_to`, faiss/IndexIVF.cpp).  These
helpers work on that payload (numpy arrays), so any producer of the format -- the reference CPU index,
a file reader -- can feed them; nothing here depends on the reference's classes.

    payload = {"d", "nlist", "metric", "centroids" [nlist, d] f32,
               "pq" [M, 256, dsub] f32 (IVFPQ only), "codes": [nlist] uint8 arrays, "ids": [nlist] int64 arrays}
"""
import numpy as np

SHARD_BY_ID_MOD = 1      # id % nshard == i          (the reference's default)
SHARD_BY_ID_RANGE = 2    # i*ntotal/nshard <= id < (i+1)*ntotal/nshard
SHARD_BY_LIST_RANGE = 4  # whole lists  i*nlist/nshard <= l < (i+1)*nlist/nshard


def shard_ivf_lists(codes, ids, code_size, nshard, shard_type=SHARD_BY_ID_MOD, ntotal=None):
    """Split inverted lists over `nshard` sub-indexes with the reference's rules
    (ToGpuClonerMultiple::copy_ivf_shard, GpuCloner.cpp:287-322).  Entry order inside a list is kept
    (copy_subset_to appends in list order).  Returns [(codes_i, ids_i)] * nshard, each a list over all
    nlist lists (empty arrays where a shard holds nothing of a list)."""
    nlist = len(ids)
    assert len(codes) == nlist and nshard >= 1
    ids = [np.ascontiguousarray(a, dtype=np.int64).reshape(-1) for a in ids]
    codes = [np.ascontiguousarray(c, dtype=np.uint8).reshape(-1, code_size) for c in codes]
    for c, a in zip(codes, ids):
        assert c.shape[0] == a.size, "codes / ids length mismatch"
    if ntotal is None:
        ntotal = int(sum(a.size for a in ids))
    out = []
    for i in range(nshard):
        ci, ii = [], []
        if shard_type == SHARD_BY_ID_RANGE:
            i0, i1 = i * ntotal // nshard, (i + 1) * ntotal // nshard
        elif shard_type == SHARD_BY_LIST_RANGE:
            l0, l1 = i * nlist // nshard, (i + 1) * nlist // nshard
        elif shard_type != SHARD_BY_ID_MOD:
            raise ValueError("shard_type %d not implemented" % shard_type)  # as the reference
        for l in range(nlist):
            if shard_type == SHARD_BY_ID_MOD:
                keep = (ids[l] % nshard) == i
            elif shard_type == SHARD_BY_ID_RANGE:
                keep = (ids[l] >= i0) & (ids[l] < i1)
            else:
                keep = np.full(ids[l].shape, l0 <= l < l1)
            ci.append(np.ascontiguousarray(codes[l][keep]).reshape(-1))
            ii.append(np.ascontiguousarray(ids[l][keep]))
        out.append((ci, ii))
    return out


def gpu_ivf_from_payload(res, payload, device=0):
    """GpuIndexIVFFlat / GpuIndexIVFPQ holding exactly the payload (the role of copyFrom,
    faiss/gpu/GpuIndexIVFPQ.cu:105-217, GpuIndexIVFFlat.cu:89-150)."""
    import faiss_b200 as fb

    d, nlist, metric = int(payload["d"]), int(payload["nlist"]), int(payload.get("metric", fb.METRIC_L2))
    if "pq" in payload and payload["pq"] is not None:
        pq = np.ascontiguousarray(payload["pq"], dtype=np.float32)
        index = fb.GpuIndexIVFPQ(res, d, nlist, int(pq.shape[0]), 8, metric, device=device)
        index.setCoarseCentroids(payload["centroids"])
        index.setPQCentroids(pq)
    else:
        index = fb.GpuIndexIVFFlat(res, d, nlist, metric, device=device)
        index.setCoarseCentroids(payload["centroids"])
    # all list lengths are known up front: ONE arena relayout with exact capacities, then plain copies
    # (per-list growth would re-layout the whole arena once per list: quadratic in nlist)
    index.setListSizes(np.array([len(payload["ids"][l]) for l in range(nlist)], dtype=np.int64))
    for l in range(nlist):
        if len(payload["ids"][l]):
            index.setList(l, payload["codes"][l], payload["ids"][l])
    index.setIsTrained(True)
    return index


def gpu_ivf_shards_from_payload(resources, payload, shard_type=SHARD_BY_ID_MOD, devices=None, threaded=True):
    """index_cpu_to_gpu_multiple(..., shard=True): one sub-index per resources object, the same coarse
    quantiser (and PQ) everywhere, lists split by `shard_type`, wrapped in IndexShards with explicit ids
    (successive_ids=False, GpuCloner.cpp:417)."""
    import faiss_b200 as fb

    n = len(resources)
    devices = list(devices) if devices is not None else [0] * n
    code_size = int(payload["pq"].shape[0]) if payload.get("pq") is not None else 4 * int(payload["d"])
    parts = shard_ivf_lists(payload["codes"], payload["ids"], code_size, n, shard_type)
    shards = fb.IndexShards(int(payload["d"]), threaded=threaded, successive_ids=False)
    for r, dev, (ci, ii) in zip(resources, devices, parts):
        sub = dict(payload)
        sub["codes"], sub["ids"] = ci, ii
        shards.add_shard(gpu_ivf_from_payload(r, sub, device=dev))
    return shards
