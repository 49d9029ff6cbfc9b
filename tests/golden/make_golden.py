"""Generates tests/golden/golden.npz from the UNMODIFIED reference CPU library (oracle/_ref, built
from /root/reference by oracle/Makefile).  Run in the authoring container:

    python tests/golden/make_golden.py

The fixtures are small seeded input/output pairs for every row of the hot path; the committed
file lets the oracle restatement and the CUDA path be checked where /root/reference is absent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402


def main():
    out = {}
    # ---- RNG
    out["float_rand_5000_s1234"] = ref.float_rand(5000, 1234)
    out["float_rand_300_s7"] = ref.float_rand(300, 7)
    out["rand_perm_1000_s42"] = ref.rand_perm(1000, 42)

    # ---- Flat, uniform floats (faiss::float_rand seeds 1234 / 1235 as in SURVEY 8(d))
    N, d, nq, k = 3000, 32, 24, 10
    xb = ref.float_rand(N * d, 1234).reshape(N, d)
    xq = ref.float_rand(nq * d, 1235).reshape(nq, d)
    for metric, name in ((1, "l2"), (0, "ip")):
        idx = ref.IndexFlat(d, metric)
        idx.add(xb)
        D, I = idx.search(xq, k)
        out["flat_%s_D" % name] = D
        out["flat_%s_I" % name] = I
    out["flat_shape"] = np.array([N, d, nq, k])

    # ---- Flat, integer-valued regime: exact arithmetic -> id order must match bit for bit
    N, d, nq = 4000, 64, 16
    xbi = np.floor(ref.float_rand(N * d, 11).reshape(N, d) * 16).astype(np.float32)
    xqi = np.floor(ref.float_rand(nq * d, 12).reshape(nq, d) * 16).astype(np.float32)
    for k in (10, 100):  # heap handler (k<100) and reservoir handler (k>=100)
        idx = ref.IndexFlat(d, 1)
        idx.add(xbi)
        D, I = idx.search(xqi, k)
        out["flatint_l2_k%d_D" % k] = D
        out["flatint_l2_k%d_I" % k] = I
    out["flatint_shape"] = np.array([N, d, nq])

    # ---- merge_knn_results (IndexShards host merge)
    rs = np.random.RandomState(5)
    allD = np.sort(rs.rand(3, 7, 5).astype(np.float32), axis=2)
    allI = rs.permutation(3 * 7 * 5).reshape(3, 7, 5).astype(np.int64)
    allI[1, :, 4] = -1
    D, I = ref.merge_knn_results(allD, allI, 1)
    out["merge_allD"], out["merge_allI"], out["merge_D"], out["merge_I"] = allD, allI, D, I

    # ---- IVFPQ: train on the CPU, export centroids / PQ / lists, search
    N, d, nlist, M, nq, k, nprobe = 6000, 32, 16, 8, 20, 10, 4
    xb = ref.float_rand(N * d, 21).reshape(N, d)
    xq = ref.float_rand(nq * d, 22).reshape(nq, d)
    for metric, name in ((1, "l2"), (0, "ip")):
        ivf = ref.IndexIVFPQ(d, nlist, M, 8, metric)
        ivf.set_cp(niter=5)
        ivf.set_pq_cp(niter=5)
        ivf.train(xb)
        ivf.add(xb)
        ivf.set_nprobe(nprobe)
        D, I = ivf.search(xq, k)
        out["ivfpq_%s_centroids" % name] = ivf.centroids()
        out["ivfpq_%s_pq" % name] = ivf.pq_centroids()
        out["ivfpq_%s_use_precomputed" % name] = np.array([ivf.use_precomputed_table])
        lens = []
        codes, ids = [], []
        for l in range(nlist):
            c, i = ivf.get_list(l)
            lens.append(i.size)
            codes.append(c)
            ids.append(i)
        out["ivfpq_%s_lens" % name] = np.array(lens)
        out["ivfpq_%s_codes" % name] = np.concatenate(codes)
        out["ivfpq_%s_ids" % name] = np.concatenate(ids)
        out["ivfpq_%s_D" % name] = D
        out["ivfpq_%s_I" % name] = I
    out["ivfpq_shape"] = np.array([N, d, nlist, M, nq, k, nprobe])

    # ---- IVFFlat
    ivf = ref.IndexIVFFlat(d, nlist, 1)
    ivf.set_cp(niter=5)
    ivf.train(xb)
    ivf.add(xb)
    ivf.set_nprobe(nprobe)
    D, I = ivf.search(xq, k)
    out["ivfflat_centroids"] = ivf.centroids()
    out["ivfflat_D"], out["ivfflat_I"] = D, I
    # list membership (ids per list) so the oracle restatement can be pinned on the reference's own lists
    out["ivfflat_lens"] = np.array([ivf.list_size(l) for l in range(nlist)])
    out["ivfflat_ids"] = np.concatenate([ivf.get_list(l)[1] for l in range(nlist)])

    # ---- ProductQuantizer::train (faiss/impl/ProductQuantizer.cpp:130-195): M independent k-means
    xp = ref.float_rand(4000 * 16, 41).reshape(4000, 16)
    out["pqtrain_centroids"] = ref.pq_train(xp, 4, 8, niter=6, seed=1234)
    out["pqtrain_shape"] = np.array([4000, 16, 4, 6, 1234])

    # ---- spherical k-means over an inner-product index (GpuIndexIVF's IP coarse training)
    xs = ref.float_rand(3000 * 8, 51).reshape(3000, 8) - 0.5
    cent, obj = ref.kmeans_spherical_ip(xs, 12, niter=6, seed=77)
    out["kmeans_sph_centroids"], out["kmeans_sph_obj"] = cent, obj

    # ---- k-means (Clustering with a CPU IndexFlatL2)
    x = ref.float_rand(5000 * 8, 31).reshape(5000, 8)
    cent, obj, nsplit = ref.kmeans(x, 20, niter=8, seed=123)
    out["kmeans_centroids"], out["kmeans_obj"], out["kmeans_nsplit"] = cent, obj, nsplit
    # a run that needs subsampling (n > k * max_points_per_centroid)
    cent, obj, nsplit = ref.kmeans(x, 4, niter=5, seed=99, max_points_per_centroid=256)
    out["kmeans_sub_centroids"], out["kmeans_sub_obj"] = cent, obj

    path = os.path.join(ROOT, "tests", "golden", "golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", ref.compile_options(), "omp", ref.omp_threads())


if __name__ == "__main__":
    main()
