"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference CPU algorithms on the hot path.

Each function cites the reference code it follows.  This module is the checker used by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg; nothing under faiss_b200/ imports it, and
it is never the thing measured or shipped.

Pinning: tests/test_oracle.py checks every function here against (a) the unmodified reference CPU
library compiled into oracle/_ref (when present) and (b) the golden fixtures in tests/golden/ that
were generated from that library by tests/golden/make_golden.py.
"""
import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


# ----------------------------------------------------------------------------- RNG
def _mt_raw(seed, n):
    """First n 32-bit outputs of std::mt19937((unsigned)seed)."""
    bg = np.random.MT19937()
    bg._legacy_seeding(int(seed) & 0xFFFFFFFF)
    return bg.random_raw(n).astype(np.uint32)


def float_rand(n, seed):
    """faiss::float_rand (faiss/utils/random.cpp:95-113): 1024 independently seeded blocks,
    x = mt() / float(mt.max())."""
    nblock = 1 if n < 1024 else 1024
    r0 = _mt_raw(seed, 2)
    a0 = int(r0[0] & 0x7FFFFFFF)
    b0 = int(r0[1] & 0x7FFFFFFF)
    out = np.empty(n, dtype=np.float32)
    for j in range(nblock):
        i0 = j * n // nblock
        i1 = (j + 1) * n // nblock
        raw = _mt_raw((a0 + j * b0) & 0xFFFFFFFF, i1 - i0)
        out[i0:i1] = raw.astype(np.float32) / np.float32(4294967295.0)
    return out


def rand_perm(n, seed):
    """faiss::rand_perm (faiss/utils/random.cpp:188-199): Fisher-Yates with mt() % (n - i)."""
    perm = np.arange(n, dtype=np.int64)
    if n <= 1:
        return perm
    raw = _mt_raw(seed, n - 1).astype(np.int64)
    for i in range(n - 1):
        i2 = i + int(raw[i] % (n - i))
        perm[i], perm[i2] = perm[i2], perm[i]
    return perm


# ----------------------------------------------------------------------------- ordering helpers
def _topk_sorted(keys, ids, k):
    """Per row: the k best by (key asc, id asc) -- the order the reference result handlers
    deliver (faiss/utils/ordered_key_value.h:40-75, faiss/impl/ResultHandler.h:275-282)."""
    n = keys.shape[0]
    outK = np.full((n, k), np.inf, dtype=keys.dtype)
    outI = np.full((n, k), -1, dtype=np.int64)
    for r in range(n):
        kk = keys[r]
        ii = ids[r] if ids.ndim == 2 else ids
        valid = ii >= 0
        kv, iv = kk[valid], ii[valid]
        if kv.size > 4 * k:
            # pre-select: everything <= the k-th smallest key (keeps all boundary ties)
            kth = np.partition(kv, k - 1)[k - 1]
            m = kv <= kth
            kv, iv = kv[m], iv[m]
        order = np.lexsort((iv, kv))[:k]
        outK[r, : order.size] = kv[order]
        outI[r, : order.size] = iv[order]
    return outK, outI


def _finish(D, I, metric):
    """Missing results: id -1 and +/-FLT_MAX (faiss/gpu/impl/Distance.cu:152-164, heap init)."""
    D = D.astype(np.float32)
    miss = I < 0
    if metric == METRIC_L2:
        D[miss] = np.finfo(np.float32).max
    else:
        D = -D
        D[miss] = -np.finfo(np.float32).max
    return D, I


# ----------------------------------------------------------------------------- Flat
def pairwise(xq, xb, metric=METRIC_L2, exact=True):
    """Distance matrix.  exact=True: direct form in float64 rounded to float32 (ground truth).
    exact=False: the norm expansion of exhaustive_L2sqr_blas (faiss/utils/distances.cpp:424-511):
    ip via sgemm, dis = ||x||^2 + ||y||^2 - 2 ip, clamped at 0, all in float32."""
    xq = np.asarray(xq, dtype=np.float32)
    xb = np.asarray(xb, dtype=np.float32)
    if metric == METRIC_INNER_PRODUCT:
        if exact:
            return (xq.astype(np.float64) @ xb.astype(np.float64).T).astype(np.float32)
        return xq @ xb.T
    if exact:
        q = xq.astype(np.float64)
        b = xb.astype(np.float64)
        d2 = (q * q).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * (q @ b.T)
        # the expansion is exact enough in float64 (53-bit) for float32 inputs
        return np.maximum(d2, 0).astype(np.float32)
    xn = (xq * xq).sum(1)
    yn = (xb * xb).sum(1)
    ip = xq @ xb.T
    dis = xn[:, None] + yn[None, :] - 2 * ip
    dis[dis < 0] = 0
    return dis.astype(np.float32)


def knn_flat(xq, xb, k, metric=METRIC_L2, exact=True, block=256):
    """IndexFlat::search -> knn_L2sqr / knn_inner_product (faiss/IndexFlat.cpp:29-60,
    faiss/utils/distances.cpp:768-890).  Result order: (distance asc [IP: desc], id asc)."""
    xq = np.asarray(xq, dtype=np.float32)
    nq = xq.shape[0]
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    ids = np.arange(xb.shape[0], dtype=np.int64)
    for q0 in range(0, nq, block):
        dis = pairwise(xq[q0 : q0 + block], xb, metric, exact)
        keys = dis if metric == METRIC_L2 else -dis
        kk, ii = _topk_sorted(keys, ids, k)
        D[q0 : q0 + block], I[q0 : q0 + block] = _finish(kk, ii, metric)
    return D, I


def merge_knn_results(all_D, all_I, metric=METRIC_L2):
    """faiss::merge_knn_results (faiss/utils/Heap.cpp:166-238): S-way merge of per-shard sorted
    lists [nshard, n, k] -> [n, k]; labels < 0 are skipped."""
    ns, n, k = all_D.shape
    keys = np.transpose(all_D, (1, 0, 2)).reshape(n, ns * k).astype(np.float32)
    ids = np.transpose(all_I, (1, 0, 2)).reshape(n, ns * k)
    if metric != METRIC_L2:
        keys = -keys
    kk, ii = _topk_sorted(keys, ids, k)
    return _finish(kk, ii, metric)


# ----------------------------------------------------------------------------- PQ / IVF
def pq_encode(x, pq_centroids):
    """ProductQuantizer::compute_code (faiss/impl/ProductQuantizer.cpp:282-309): per sub-vector,
    argmin of the direct L2 distance, first minimum wins."""
    M, ksub, dsub = pq_centroids.shape
    n = x.shape[0]
    codes = np.empty((n, M), dtype=np.uint8)
    for m in range(M):
        xs = x[:, m * dsub : (m + 1) * dsub].astype(np.float32)
        c = pq_centroids[m].astype(np.float32)
        d2 = ((xs[:, None, :] - c[None, :, :]) ** 2).sum(-1, dtype=np.float32)
        codes[:, m] = np.argmin(d2, axis=1)
    return codes


def ivf_assign(x, centroids, metric=METRIC_L2):
    D, I = knn_flat(x, centroids, 1, metric)
    return I[:, 0]


def ivfpq_lut(q, c_list, pq_centroids, metric=METRIC_L2):
    """Per-(query, list) distance table.  L2 by_residual: tab[m][c] = ||(q - c_list)_m - pq[m][c]||^2
    (IVFPQ_QueryTables.cpp:194-244, residual form; equal to the precomputed-table form
    coarse_dis + T2[list] - 2 q.pq of :126-192 up to rounding).  IP: tab[m][c] = q_m . pq[m][c]."""
    M, ksub, dsub = pq_centroids.shape
    if metric == METRIC_L2:
        r = (q - c_list).astype(np.float32).reshape(M, 1, dsub)
        return ((r - pq_centroids) ** 2).sum(-1, dtype=np.float32)
    qq = q.astype(np.float32).reshape(M, 1, dsub)
    return (qq * pq_centroids).sum(-1, dtype=np.float32)


def ivfpq_precomputed_table(centroids, pq_centroids):
    """IndexIVFPQ::precompute_table (faiss/IndexIVFPQ.cpp:376-458), use_precomputed_table = 1:
    T2[list][m][c] = ||y_{m,c}||^2 + 2 <centroid_list | m, y_{m,c}>."""
    M, ksub, dsub = pq_centroids.shape
    nlist = centroids.shape[0]
    r_norms = (pq_centroids.astype(np.float32) ** 2).sum(-1, dtype=np.float32)  # [M, ksub]
    c = centroids.astype(np.float32).reshape(nlist, M, 1, dsub)
    cross = (c * pq_centroids[None]).sum(-1, dtype=np.float32)  # [nlist, M, ksub]
    return r_norms[None] + np.float32(2.0) * cross


def ivfpq_lut_precomputed(q, c_list, t2_list, pq_centroids):
    """Precomputed-table form of the L2 table (IVFPQ_QueryTables.cpp:126-192): the distance is
    term1 + sum_m (T2[list][m][c] - 2 <q|m, y_{m,c}>) with term1 = ||q - c_list||^2.  Returns (term1, tab)."""
    M, ksub, dsub = pq_centroids.shape
    qq = q.astype(np.float32).reshape(M, 1, dsub)
    t3 = np.float32(-2.0) * (qq * pq_centroids).sum(-1, dtype=np.float32)
    r = (q - c_list).astype(np.float32)
    return np.float32((r * r).sum(dtype=np.float32)), (t2_list + t3).astype(np.float32)


def ivfpq_search(xq, k, nprobe, centroids, pq_centroids, lists_codes, lists_ids, metric=METRIC_L2, probes=None, precomputed=False):
    """IndexIVFPQ::search = coarse quantisation + search_preassigned with the table scanner
    (faiss/IndexIVF.cpp:305-760, faiss/impl/pq_code_distance/IVFPQScanner_impl.h:122-198).
    precomputed=True evaluates the L2 distance in the precomputed-table decomposition instead of
    the residual form (the CPU reference picks either, `use_precomputed_table`)."""
    M = pq_centroids.shape[0]
    t2 = ivfpq_precomputed_table(centroids, pq_centroids) if (precomputed and metric == METRIC_L2) else None
    nq = xq.shape[0]
    if probes is None:
        cD, probes = knn_flat(xq, centroids, nprobe, metric)
    else:
        cD = None
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    mrange = np.arange(M)
    for qi in range(nq):
        keys, ids = [], []
        for p in range(probes.shape[1]):
            l = probes[qi, p]
            if l < 0:
                continue
            codes = lists_codes[l].reshape(-1, M)
            if codes.shape[0] == 0:
                continue
            if t2 is not None:
                term1, tab = ivfpq_lut_precomputed(xq[qi], centroids[l], t2[l], pq_centroids)
                dis = tab[mrange[None, :], codes].sum(1, dtype=np.float32) + term1
            else:
                tab = ivfpq_lut(xq[qi], centroids[l], pq_centroids, metric)
                dis = tab[mrange[None, :], codes].sum(1, dtype=np.float32)
            if metric == METRIC_INNER_PRODUCT:
                dis = dis + np.float32(np.dot(xq[qi].astype(np.float32), centroids[l].astype(np.float32)))
                dis = -dis
            keys.append(dis)
            ids.append(lists_ids[l])
        if keys:
            kk, ii = _topk_sorted(np.concatenate(keys)[None, :], np.concatenate(ids)[None, :], k)
        else:
            kk = np.full((1, k), np.inf, dtype=np.float32)
            ii = np.full((1, k), -1, dtype=np.int64)
        D[qi], I[qi] = _finish(kk, ii, metric)
    return D, I


def ivfflat_search(xq, k, nprobe, centroids, lists_vecs, lists_ids, metric=METRIC_L2, probes=None):
    """IndexIVFFlat::search (faiss/IndexIVFFlat.cpp, scanner = exact fvec_L2sqr / inner product)."""
    nq, d = xq.shape
    if probes is None:
        _, probes = knn_flat(xq, centroids, nprobe, metric)
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    for qi in range(nq):
        keys, ids = [], []
        for p in range(probes.shape[1]):
            l = probes[qi, p]
            if l < 0 or lists_ids[l].size == 0:
                continue
            v = lists_vecs[l].reshape(-1, d)
            dis = pairwise(xq[qi : qi + 1], v, metric, exact=True)[0]
            keys.append(dis if metric == METRIC_L2 else -dis)
            ids.append(lists_ids[l])
        if keys:
            kk, ii = _topk_sorted(np.concatenate(keys)[None, :], np.concatenate(ids)[None, :], k)
        else:
            kk = np.full((1, k), np.inf, dtype=np.float32)
            ii = np.full((1, k), -1, dtype=np.int64)
        D[qi], I[qi] = _finish(kk, ii, metric)
    return D, I


def build_ivf_lists(assign, rows, ids, nlist):
    """ArrayInvertedLists after add_with_ids in batch order (faiss/invlists/InvertedLists.h):
    list l holds the rows with assign == l, in insertion order."""
    codes, lid = [], []
    for l in range(nlist):
        m = np.nonzero(assign == l)[0]
        codes.append(np.ascontiguousarray(rows[m]).reshape(-1).view(np.uint8))
        lid.append(ids[m].astype(np.int64))
    return codes, lid


# ----------------------------------------------------------------------------- k-means
def split_clusters(d, k, n, hassign, centroids):
    """faiss/impl/ClusteringHelpers.cpp:177-240 (EPS = 1/1024, RandomGenerator(1234))."""
    EPS = np.float32(1.0 / 1024.0)
    nsplit = 0
    stream = None
    pos = 0

    def rand_float():
        nonlocal stream, pos
        if stream is None or pos >= stream.size:
            base = 0 if stream is None else stream.size
            stream = _mt_raw(1234, max(1 << 16, base * 2))
        v = np.float32(stream[pos]) / np.float32(4294967295.0)
        pos += 1
        return v

    for ci in range(k):
        if hassign[ci] != 0:
            continue
        cj, tries, found = 0, 0, False
        while tries < 10 * k:
            p = np.float32((hassign[cj] - 1.0) / np.float32(n - k))
            if rand_float() < p:
                found = True
                break
            tries += 1
            cj = (cj + 1) % k
        if not found:
            cj = int(np.argmax(hassign))
        centroids[ci] = centroids[cj]
        even = np.arange(d) % 2 == 0
        centroids[ci, even] *= 1 + EPS
        centroids[cj, even] *= 1 - EPS
        centroids[ci, ~even] *= 1 - EPS
        centroids[cj, ~even] *= 1 + EPS
        hassign[ci] = hassign[cj] / 2
        hassign[cj] -= hassign[ci]
        nsplit += 1
    return nsplit


def renorm_l2(c):
    """fvec_renorm_L2 (faiss/utils/distances.cpp:238-251): rows with non-zero norm scaled by 1/sqrtf(||row||^2)."""
    c = np.asarray(c, dtype=np.float32)
    nr = (c.astype(np.float32) ** 2).sum(axis=1, dtype=np.float32)
    inv = np.ones_like(nr)
    inv[nr > 0] = (np.float32(1.0) / np.sqrt(nr[nr > 0], dtype=np.float32)).astype(np.float32)
    return (c * inv[:, None]).astype(np.float32)


def kmeans(x, k, niter=25, seed=1234, max_points_per_centroid=256, metric=METRIC_L2, spherical=False):
    """faiss::Clustering::train_encoded (faiss/Clustering.cpp:60-380) with an exact assignment index of
    the given metric: subsample by rand_perm(seed), init = x[rand_perm(seed+1)[:k]], post_process_centroids
    (spherical: fvec_renorm_L2, Clustering.cpp:35-45) after the init and after every iteration, Lloyd
    iterations with compute_centroids (ClusteringHelpers.cpp:101-172) and split_clusters, early stop when
    the objective did not change (early_stop_threshold = 0, Clustering.cpp:360-377)."""
    x = np.asarray(x, dtype=np.float32)
    n, d = x.shape
    if n > k * max_points_per_centroid:
        perm = rand_perm(n, seed)
        x = x[perm[: k * max_points_per_centroid]]
        n = x.shape[0]
    perm = rand_perm(n, seed + 1)
    cent = x[perm[:k]].copy()
    if spherical:
        cent = renorm_l2(cent)
    objs = []
    for it in range(niter):
        D, I = knn_flat(x, cent, 1, metric)
        objs.append(np.float32(D.sum(dtype=np.float64)))
        a = I[:, 0]
        hassign = np.bincount(a, minlength=k).astype(np.float32)
        sums = np.zeros((k, d), dtype=np.float64)
        np.add.at(sums, a, x.astype(np.float64))
        nz = hassign > 0
        new = cent.copy()
        new[nz] = (sums[nz] * (1.0 / hassign[nz])[:, None]).astype(np.float32)
        new[~nz] = 0
        cent = new
        split_clusters(d, k, n, hassign, cent)
        if spherical:
            cent = renorm_l2(cent)
        if it > 0 and objs[-2] != 0 and abs(float(objs[-2]) - float(objs[-1])) / abs(float(objs[-2])) <= 0.0:
            break
    return cent, np.array(objs, dtype=np.float32)


def pq_train(x, M, niter=25, seed=1234, max_points_per_centroid=256):
    """ProductQuantizer::train, Train_default (faiss/impl/ProductQuantizer.cpp:130-195): M independent
    k-means (256 centroids) on the column slices, each a fresh Clustering(dsub, 256, cp)."""
    x = np.asarray(x, dtype=np.float32)
    n, d = x.shape
    dsub = d // M
    out = np.empty((M, 256, dsub), dtype=np.float32)
    for m in range(M):
        out[m], _ = kmeans(np.ascontiguousarray(x[:, m * dsub : (m + 1) * dsub]), 256, niter=niter, seed=seed,
                           max_points_per_centroid=max_points_per_centroid)
    return out


# ----------------------------------------------------------------------------- comparison helpers
def recall_at(I, gt, r):
    """fraction of queries whose true nearest neighbour is in the first r results"""
    return float((I[:, :r] == gt[:, :1]).any(axis=1).mean())


def intersection_recall(I, gt):
    n = I.shape[0]
    tot = 0
    for a, b in zip(I, gt):
        tot += np.intersect1d(a[a >= 0], b[b >= 0]).size
    return tot / float(gt.size)


def compare_lists(refD, refI, D, I, eps=6e-3, pct_max_diff1=0.1, pct_max_diffN=0.015):
    """Semantics of compareLists (faiss/gpu/test/TestUtils.cpp:234-443): unique ids per query,
    -1 in the same places, relative distance error <= eps where both have the same id at the same
    rank-or-nearby, bounded fraction of rank differences."""
    n, k = refI.shape
    diff1 = diffN = 0
    max_rel = 0.0
    for q in range(n):
        a, b = refI[q], I[q]
        va = a[a >= 0]
        vb = b[b >= 0]
        assert np.unique(vb).size == vb.size, "duplicate ids in query %d" % q
        assert ((a < 0) == (b < 0)).all(), "-1 placement differs in query %d" % q
        pos = {int(v): i for i, v in enumerate(b)}
        for i, v in enumerate(a):
            if v < 0:
                continue
            j = pos.get(int(v))
            if j is None:
                diffN += 1
                continue
            if i != j:
                diff1 += 1
                if abs(i - j) > 1:
                    diffN += 1
            x, y = float(refD[q, i]), float(D[q, j])
            den = 0.5 * (abs(x) + abs(y))
            if den > 0:
                max_rel = max(max_rel, abs(x - y) / den)
    total = n * k
    assert max_rel <= eps, "relative distance error %g > %g" % (max_rel, eps)
    assert diff1 <= pct_max_diff1 * total, "rank differences %d > %g of %d" % (diff1, pct_max_diff1, total)
    assert diffN <= pct_max_diffN * total, "rank differences >1: %d > %g of %d" % (diffN, pct_max_diffN, total)
    return {"max_rel_err": max_rel, "diff1": diff1, "diffN": diffN}
