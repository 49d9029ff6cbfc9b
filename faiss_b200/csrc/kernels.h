// faiss_b200 -- launcher declarations for every device kernel (L1).  Host code (index objects,
// C ABI) only sees these; all pointers are DEVICE pointers, all work is enqueued on `stream`.
#pragma once

#include <cuda_fp16.h>

#include "common.h"
#include "resources.h"

namespace fb200 {

// ---------------------------------------------------------------- flat_exact.cu
// ||x||^2 per row (role of runL2Norm, faiss/gpu/impl/L2Norm.cu:176)
void runL2Norms(const float* x, int64_t n, int d, float* norms, cudaStream_t stream);

// Exact fp32 brute-force k-NN (SIMT, direct-form sum (q-y)^2 / sum q*y accumulated in dimension
// order).  This is the always-correct path: small problems, odd dimensions, the fallback of the
// tensor-core path, and the canonical arithmetic the tensor-core re-rank reproduces.
// Role of runDistance<float> (faiss/gpu/impl/Distance.cu:121-405) with the k-select fused in.
//   Q [nq,d], Y [n,d] row-major; outD [nq,k], outI [nq,k] (int64, row index + idBase; -1 missing)
//   yHalf: Y holds __half rows (GpuIndexFlatConfig::useFloat16 storage), widened to fp32 on load
void runFlatExact(
        GpuResources* res,
        int device,
        const float* Q,
        int64_t nq,
        const void* Y,
        int64_t n,
        int d,
        int k,
        MetricType metric,
        int64_t idBase,
        float* outD,
        idx_t* outI,
        cudaStream_t stream,
        int yHalf = 0);

// k = 1 convenience (assignment): outI int64 [nq], outD optional
void runFlatArgmin(
        GpuResources* res,
        int device,
        const float* Q,
        int64_t nq,
        const float* Y,
        int64_t n,
        int d,
        MetricType metric,
        float* outD,
        idx_t* outI,
        cudaStream_t stream);

// Row-wise top-k over candidate lists (role of runBlockSelectPair / merge_knn_results,
// faiss/gpu/utils/BlockSelectFloat.cu:98, faiss/utils/Heap.cpp:166-238).
//   inD/inI: [rows, nlists, kin]; ids < 0 are skipped; idOffsets (optional, [nlists]) is added to
//   ids of list l (IndexShards successive_ids translation, faiss/IndexShards.cpp:214-220).
//   Keys are user-facing distances (L2: smaller better; IP: larger better).
void runMergeTopK(
        const float* inD,
        const idx_t* inI,
        int64_t rows,
        int nlists,
        int kin,
        const idx_t* idOffsets,
        int k,
        MetricType metric,
        float* outD,
        idx_t* outI,
        cudaStream_t stream);

// same merge over inputs laid out [nlists][rows][kin] (the layout an all-gather of per-shard results
// produces: no permute copy between the collective and the merge)
void runMergeTopKListMajor(
        const float* inD,
        const idx_t* inI,
        int64_t rows,
        int nlists,
        int kin,
        const idx_t* idOffsets,
        int k,
        MetricType metric,
        float* outD,
        idx_t* outI,
        cudaStream_t stream);

// residual x - c[assign] (NaN if assign = -1) ; role of runCalcResidual (VectorResidual.cu:26-176)
void runCalcResidual(
        const float* x,
        const void* centroids,
        const idx_t* assign,
        int64_t n,
        int d,
        float* out,
        cudaStream_t stream,
        int yHalf = 0);
// gather rows by id (reconstruct_batch) / by range
void runGatherRows(const void* src, const idx_t* ids, int64_t n, int d, float* out, cudaStream_t stream, int yHalf = 0);

// ---------------------------------------------------------------- flat_tc.cu  (tcgen05 path)
struct FlatTcPlan; // opaque: tensor maps + scratch sizing for one (index, nq, k) shape

// fp32 rows -> scaled fp16 rows (padded to dpad, multiple of 64) + score bias (bias = -||y||^2/2 for
// L2, 0 for IP) + per-256-row-tile maximum bias.  With perm != null (L2) the fp16 copy is stored in
// order of increasing norm: perm[stored position] = row id.  norms[] stays in row order.
void runFlatTcPrepareRows(
        GpuResources* res,
        int device,
        const void* Y, // fp32 rows, or __half rows when yHalf
        int64_t n,
        int d,
        int dpad,
        float scale,
        MetricType metric,
        __half* Y16,
        float* bias,
        int* perm,
        float* tileMaxBias,
        float* norms,
        cudaStream_t stream,
        int yHalf = 0);

// max |x| over a matrix (device scalar, float) -- used to pick the power-of-two fp16 scale
void runAbsMax(const void* x, int64_t count, float* out /*device, must be zeroed*/, cudaStream_t stream, int yHalf = 0);
// max row norm^2
void runMaxOf(const float* x, int64_t count, float* out /*device, zeroed; x >= 0*/, cudaStream_t stream);

bool flatTcSupported(int d, int k, int64_t n);

// Sharded search (one shard per NCCL rank): thresholds are pooled across the ranks after every round
// (one all-reduce of 2 floats per query), so a 1/S-size shard filters as tightly as the whole database would
// and keeps only its share of the global top-k; all ranks must call with the same queries and k.
class Communicator;
struct FlatTcShard {
    const Communicator* comm; // this rank
    int64_t maxTiles;         // max over ranks of ceil(n_r / 256): the common round schedule
};

// Full certified search: fp16 tcgen05 scoring + candidate emission + exact fp32 re-rank, with the
// exact SIMT kernel as fallback for queries whose certificate fails.  See flat_tc.cu.
void runFlatTcSearch(
        GpuResources* res,
        int device,
        const float* Q,
        int64_t nq,
        const void* Y,       // stored rows [n,d] (exact re-rank): fp32, or __half when yHalf
        const __half* Y16,   // fp16 scaled rows [n,dpad], stored order
        const float* bias,   // [n] stored order
        const int* perm,     // stored position -> row id (null: identity)
        const float* tileMaxBias, // [ceil(n/256)]
        float yScale,        // power of two applied to Y16
        float yMaxNorm,      // max ||y|| (unscaled)
        int64_t n,
        int d,
        int dpad,
        int k,
        MetricType metric,
        float* outD,
        idx_t* outI,
        cudaStream_t stream,
        const FlatTcShard* shard = nullptr,
        int yHalf = 0);

// number of queries the last runFlatTcSearch on this thread recomputed with the exact kernel
int& lastFlatTcFallbacks();

// debug / unit-test seam: raw fp16 tensor-core score tile  S[nq,n] = Q16 . Y16^T  (fp32 out)
void runFlatTcScoresDebug(
        const __half* Q16,
        int64_t nq,
        const __half* Y16,
        int64_t n,
        int dpad,
        float* S,
        cudaStream_t stream);

// ---------------------------------------------------------------- kmeans.cu
// centroid update (role of compute_centroids, faiss/impl/ClusteringHelpers.cpp:101-172):
// sums[c] += x_i for assign[i]=c, counts[c] += 1 ; then centroids = sums / counts
void runKmeansAccumulate(
        const float* x,
        const idx_t* assign,
        int64_t n,
        int d,
        int64_t k,
        float* sums /*[k,d] zeroed*/,
        float* counts /*[k] zeroed*/,
        cudaStream_t stream);
void runKmeansFinalize(
        const float* sums,
        const float* counts,
        int64_t k,
        int d,
        float* centroids /* in: previous, out: new (unchanged where count==0) */,
        cudaStream_t stream);

// post_process_centroids (faiss/Clustering.cpp:35-45): spherical renormalisation and/or rounding to integers
void runKmeansPostProcess(float* centroids, int64_t k, int d, bool spherical, bool intCentroids, cudaStream_t stream);

// ---------------------------------------------------------------- ivf.cu
// PQ encode: codes[i][m] = argmin_c || r_i[m*dsub:(m+1)*dsub] - pq[m][c] ||^2
//   (role of IVFPQ::appendVectors_ per-subquantizer k=1 search, faiss/gpu/impl/IVFPQ.cu:129-257;
//    arithmetic follows faiss/impl/ProductQuantizer.cpp compute_code: direct L2, first min wins)
void runPQEncode(
        const float* resid,
        int64_t n,
        int d,
        int M,
        int ksub,
        const float* pqCentroids /*[M][ksub][dsub]*/,
        uint8_t* codes /*[n][M]*/,
        cudaStream_t stream);

// histogram of list assignments + stable scatter positions (device-side append bookkeeping;
// replaces the host unordered_map pass of IVFBase::addVectorsToLists_, IVFBase.cu:693-905)
void runIvfCountAssign(const idx_t* assign, int64_t n, int64_t nlist, int* counts /*[nlist] += */, cudaStream_t stream);
// offsets[i] = position of vector i inside its list (listLen[assign[i]] before this batch + rank of
// i among batch vectors with the same list, in batch order)
void runIvfAppendOffsets(
        const idx_t* assign,
        int64_t n,
        int64_t nlist,
        const int* listLenBefore /*[nlist]*/,
        int* offsets /*[n]*/,
        int* scratch /*[nlist]*/,
        cudaStream_t stream);
// scatter rows (codeSize bytes each) + ids to list storage: dst = base + listStart[l]*codeSize
void runIvfScatter(
        const uint8_t* rows,
        const idx_t* ids,
        const idx_t* assign,
        const int* offsets,
        int64_t n,
        int codeSize,
        const int64_t* listStart /*[nlist] element offset of each list in the arena*/,
        uint8_t* arenaCodes,
        idx_t* arenaIds,
        cudaStream_t stream);

// IVF-Flat list scan (role of runIVFInterleavedScan, faiss/gpu/impl/IVFInterleaved.cu:179).
//   probes [nq,nprobe] list ids (-1 = skip), lists are row-major fp32 [len,d] inside the arena.
void runIvfFlatScan(
        GpuResources* res,
        int device,
        const float* Q,
        int64_t nq,
        int d,
        const idx_t* probes,
        int nprobe,
        const int64_t* listStart,
        const int* listLen,
        const float* arenaVecs,
        const idx_t* arenaIds,
        int64_t arenaElems,
        int k,
        MetricType metric,
        float* outD,
        idx_t* outI,
        cudaStream_t stream);

// IVF-PQ list scan (role of runPQScanMultiPassNoPrecomputed + pqCodeDistances,
// faiss/gpu/impl/PQScanMultiPassNoPrecomputed-inl.cuh:527, PQCodeDistances-inl.cuh:591): the
// per-(query,list) LUT is built in shared memory, codes are streamed with 128-bit loads, the
// running top-k stays on chip.  Distance form follows the reference CPU scanner
// (faiss/impl/pq_code_distance/IVFPQ_QueryTables.cpp:126-192): L2 by_residual:
//   dis = sum_m || (q - c_list)_m - pq[m][code_m] ||^2 ; IP: q.c_list + sum_m q_m . pq[m][code_m]
void runIvfPqScan(
        GpuResources* res,
        int device,
        const float* Q,
        int64_t nq,
        int d,
        const idx_t* probes,
        const float* coarseDis,
        int nprobe,
        const float* coarseCentroids /*[nlist,d]*/,
        const float* pqCentroids /*[M][ksub][dsub]*/,
        int M,
        const int64_t* listStart,
        const int* listLen,
        const uint8_t* arenaCodes,
        const idx_t* arenaIds,
        int k,
        MetricType metric,
        float* outD,
        idx_t* outI,
        cudaStream_t stream);

// ---- "rotated, interleaved-by-32" PQ code layout (B200-native storage for M % 16 == 0, M <= 32) ----
// List-relative vector v = 32*g + t is stored in group g; byte position j of the vector holds
// code[(j + t) % M] and lives at  g*32*M + (j/16)*512 + t*16 + (j%16).  A warp therefore loads a
// group with fully coalesced 128-bit loads (512 B per instruction), and at step j lane t needs the
// LUT entry of sub-quantiser (j + t) % M: with the LUT laid out [code][slot] (slot = sub-quantiser,
// duplicated up to 64 slots = 256 B per code) lane t reads bank (t + j) % 32 -- conflict-free by
// construction.  copyTo / getListVectorData undo the permutation, so the external format stays the
// CPU ArrayInvertedLists byte layout.
inline bool ivfPqInterleavedSupported(int M) {
    return (M == 16 || M == 32);
}
// flat [n][M] codes -> arena (append): position = listStart[assign[i]] + offsets[i]
void runIvfPqScatterInterleaved(
        const uint8_t* codesFlat,
        const idx_t* ids,
        const idx_t* assign,
        const int* offsets,
        int64_t n,
        int M,
        const int64_t* listStart,
        uint8_t* arenaCodes,
        idx_t* arenaIds,
        cudaStream_t stream);
// one list: flat [len][M] <-> interleaved bytes at `listCodes` (arena + listStart*M)
void runIvfPqListToInterleaved(const uint8_t* flat, int64_t len, int M, uint8_t* listCodes, cudaStream_t stream);
void runIvfPqListFromInterleaved(const uint8_t* listCodes, int64_t len, int M, uint8_t* flat, cudaStream_t stream);

// T2[l][c][m] = ||y_{m,c}||^2 + 2 <centroid_l restricted to sub-space m, y_{m,c}>   (pqT = [256][M][dsub])
void runIvfPqPrecomputeTerm2(
        const float* coarse, const float* pqT, int64_t nlist, int d, int M, float* term2, cudaStream_t stream);

// scan over the interleaved layout; pqCentroidsT is the [ksub][M][dsub] transpose of pqCentroids
void runIvfPqScanInterleaved(
        GpuResources* res,
        int device,
        const float* Q,
        int64_t nq,
        int d,
        const idx_t* probes,
        const float* coarseDis,
        int nprobe,
        const float* coarseCentroids,
        const float* pqCentroidsT,
        const float* term2, // precomputed [nlist][256][M] table (L2) or null
        int M,
        const int64_t* listStart,
        const int* listLen,
        const uint8_t* arenaCodes,
        const idx_t* arenaIds,
        int64_t arenaElems,
        int k,
        MetricType metric,
        float* outD,
        idx_t* outI,
        cudaStream_t stream);

} // namespace fb200
