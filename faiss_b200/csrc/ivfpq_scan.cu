// faiss_b200 -- IVF-PQ scan over the rotated, interleaved-by-32 code layout (see kernels.h).
//
// What it replaces: pqCodeDistances (LUT [nq,nprobe,M,256] written to HBM,
// faiss/gpu/impl/PQCodeDistances-inl.cuh:34-285) + pqScanNoPrecomputedMultiPass (per-thread 32-byte
// strided code loads, every distance written to HBM, PQScanMultiPassNoPrecomputed-inl.cuh:174-270,
// PQCodeLoad.cuh:439-454) + pass1/pass2SelectLists (IVFUtilsSelect1/2.cu).
//
// Bound: HBM (codes: M bytes per scanned vector), co-limited by shared-memory gathers (M lookups per
// vector).  Both limits are attacked by the layout:
//   * a warp reads a 32-vector group as M/16 fully coalesced 512-byte loads;
//   * per lookup the inner loop is PRMT (byte -> LUT row address | lane slot) + LDS + FADD, and the
//     LUT access is bank-conflict-free by construction (lane t reads bank (t + j) % 32).
// LUT and distances never touch HBM; the running top-k stays in shared memory (select.cuh).
#include <cfloat>
#include <cstdlib>

#include "kernels.h"
#include "select.cuh"

namespace fb200 {

void runMergeTopKKeyspace(
        const float*, const idx_t*, int64_t, int, int, int, MetricType, int64_t, float*, idx_t*, cudaStream_t);
int ivfScanChunks(int device, int64_t nq, int nprobe, int* probesPerCta);

namespace {

constexpr int kBuf = 64;
constexpr int kLutSlots = 64; // 256 B per code value

__device__ __forceinline__ int64_t interleaved_pos(int64_t v, int j, int M) {
    // byte position j of list-relative vector v
    const int64_t g = v >> 5;
    const int t = (int)(v & 31);
    return g * 32 * M + (j >> 4) * 512 + t * 16 + (j & 15);
}

__global__ void pq_scatter_interleaved_kernel(
        const uint8_t* __restrict__ flat,
        const idx_t* __restrict__ ids,
        const idx_t* __restrict__ assign,
        const int* __restrict__ offsets,
        int64_t n,
        int M,
        const int64_t* __restrict__ listStart,
        uint8_t* __restrict__ arenaCodes,
        idx_t* __restrict__ arenaIds) {
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n)
        return;
    const int off = offsets[i];
    if (off < 0)
        return;
    const int64_t ls = listStart[assign[i]];
    uint8_t* base = arenaCodes + ls * M;
    const int t = off & 31;
    for (int j = lane_id(); j < M; j += 32)
        base[interleaved_pos(off, j, M)] = flat[i * M + ((j + t) % M)];
    if (lane_id() == 0)
        arenaIds[ls + off] = ids[i];
}

__global__ void pq_list_to_interleaved_kernel(const uint8_t* __restrict__ flat, int64_t len, int M, uint8_t* __restrict__ dst) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= len * M)
        return;
    const int64_t v = e / M;
    const int j = (int)(e - v * M);
    dst[interleaved_pos(v, j, M)] = flat[v * M + ((j + (int)(v & 31)) % M)];
}

__global__ void pq_list_from_interleaved_kernel(const uint8_t* __restrict__ src, int64_t len, int M, uint8_t* __restrict__ flat) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= len * M)
        return;
    const int64_t v = e / M;
    const int j = (int)(e - v * M);
    flat[v * M + ((j + (int)(v & 31)) % M)] = src[interleaved_pos(v, j, M)];
}

// block-level merge of the per-warp lists into warp 0 + write-out.  List ids are arena positions;
// the user labels are looked up only for the k survivors.
template <typename IdT, int kWarps>
__device__ void merge_and_write(
        WarpTopK<IdT>& w,
        int warp,
        unsigned char* lists,
        size_t perWarp,
        int LIST,
        int k,
        const idx_t* __restrict__ arenaIds,
        float* __restrict__ outD,
        idx_t* __restrict__ outI) {
    w.finish();
    __syncthreads();
    if (warp == 0) {
        for (int ow = 1; ow < kWarps; ow++) {
            const float* ok = reinterpret_cast<const float*>(lists + perWarp * ow);
            const IdT* oi = reinterpret_cast<const IdT*>(lists + perWarp * ow + sizeof(float) * (LIST + kBuf));
            for (int e0 = 0; e0 < k; e0 += 32) {
                int e = e0 + lane_id();
                bool valid = e < k;
                float key = valid ? ok[e] : 0.f;
                IdT id = valid ? oi[e] : 0;
                valid = valid && id != IdLimits<IdT>::max();
                if (!__any_sync(kFullMask, valid && key <= w.thr))
                    break;
                w.add(valid, key, id);
            }
        }
        w.finish();
        for (int j = lane_id(); j < k; j += 32) {
            IdT id = w.q.ids[j];
            bool ok2 = id != IdLimits<IdT>::max();
            outD[j] = ok2 ? w.q.keys[j] : CUDART_INF_F;
            outI[j] = ok2 ? arenaIds[id] : -1;
        }
    }
}

// One CTA = one query x one chunk of its probes.  The per-warp top-k lists (and their thresholds)
// live across the probes of the chunk, so the number of candidates that pass the threshold grows with
// log(vectors scanned per CTA), not with the number of (query, probe) pairs; per probe only the LUT is
// rebuilt.  Keys: L2 -> sum of LUT entries; IP -> -(q.centroid) - sum (coarse term folded in per probe).
template <int M, bool IS_L2, bool PRECOMP, typename IdT, int kWarps, int LU>
__global__ void __launch_bounds__(kWarps * 32, 1024 / (kWarps * 32)) ivfpq_scan_interleaved_kernel(
        const float* __restrict__ Q,
        int d,
        const idx_t* __restrict__ probes,
        const float* __restrict__ coarseDis,
        int nprobe,
        int probesPerCta,
        const float* __restrict__ coarse,
        const float* __restrict__ pqT, // [256][M][dsub]
        const float* __restrict__ term2, // PRECOMP: [nlist][256][M]
        const int64_t* __restrict__ listStart,
        const int* __restrict__ listLen,
        const uint8_t* __restrict__ arenaCodes,
        const idx_t* __restrict__ arenaIds,
        int k,
        int LIST,
        float* __restrict__ partD,
        idx_t* __restrict__ partI) {
    static_assert(!PRECOMP || IS_L2, "precomputed tables are an L2 decomposition");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int q = blockIdx.y, chunk = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int dsub = d / M;
    float* lut = reinterpret_cast<float*>(smem_raw);                 // [256][kLutSlots]
    float* rs = lut + 256 * kLutSlots;                               // [d]
    unsigned char* lists = reinterpret_cast<unsigned char*>(rs) + round_up(sizeof(float) * d, 16);
    const size_t perWarp = SmemTopK<IdT>::bytes(LIST, kBuf);
    float* oD = partD + ((int64_t)q * gridDim.x + chunk) * k;
    idx_t* oI = partI + ((int64_t)q * gridDim.x + chunk) * k;

    WarpTopK<IdT> w;
    unsigned char* mine = lists + perWarp * warp;
    w.init(reinterpret_cast<float*>(mine), reinterpret_cast<IdT*>(mine + sizeof(float) * (LIST + kBuf)), LIST, kBuf, k);
    const unsigned char* lutB = reinterpret_cast<const unsigned char*>(lut);
    const unsigned lane4 = (unsigned)lane << 2;

    // kU groups (kU * 32 * M bytes) per warp iteration, all 128-bit loads issued before the first lookup;
    // groups are dealt to the warps round-robin.  (Measured alternatives that did NOT help on B200, N=100M:
    // a register double buffer for the next iteration's codes, 71 vs 66 ms; claiming groups from a shared
    // counter to balance the per-probe barrier, 59.2 vs 57.7 ms; prefetch.global.L2 of the next chunk, 60.1.)
    constexpr int kU = 4;
    constexpr int kStride = kWarps * kU;
    constexpr int kThreads = kWarps * 32;
    constexpr int kEntriesPerThread = 256 * M / kThreads;
    static_assert(256 * M % kThreads == 0, "LUT entries must divide evenly among the threads");

    // direct LUT entry e = c*M + m from the vector r[d] in shared memory:
    //   L2: ||r|m - y||^2 (r = query - list centroid);  IP / PRECOMP term 3: <r|m, y> (r = query)
    auto entry = [&](int e, bool l2form) {
        const int c = e / M, m = e - c * M;
        const float* cp = pqT + (size_t)e * dsub;
        const float* rp = rs + m * dsub;
        float acc = 0.f;
        if ((dsub & 3) == 0) {
            for (int j = 0; j < dsub; j += 4) {
                const float4 cv = __ldg(reinterpret_cast<const float4*>(cp + j));
                const float4 rv = *reinterpret_cast<const float4*>(rp + j);
                if (l2form) {
                    float d0 = rv.x - cv.x, d1 = rv.y - cv.y, d2 = rv.z - cv.z, d3 = rv.w - cv.w;
                    acc = fmaf(d0, d0, acc);
                    acc = fmaf(d1, d1, acc);
                    acc = fmaf(d2, d2, acc);
                    acc = fmaf(d3, d3, acc);
                } else {
                    acc = fmaf(rv.x, cv.x, acc);
                    acc = fmaf(rv.y, cv.y, acc);
                    acc = fmaf(rv.z, cv.z, acc);
                    acc = fmaf(rv.w, cv.w, acc);
                }
            }
        } else {
            for (int j = 0; j < dsub; j++) {
                if (l2form) {
                    float df = rp[j] - cp[j];
                    acc = fmaf(df, df, acc);
                } else {
                    acc = fmaf(rp[j], cp[j], acc);
                }
            }
        }
        return acc;
    };
    auto store = [&](int e, float val) {
        const int c = e / M, m = e - c * M;
#pragma unroll
        for (int s = 0; s < kLutSlots; s += M)
            lut[c * kLutSlots + s + m] = val; // entry (c, m) -> slots m, m+M, ... (< 64), conflict-free
    };

    // Query-only part, once per CTA.  IP: the whole LUT (-<x|m, y> does not depend on the list).
    // PRECOMP: term 3 = -2 <x|m, y> for this thread's entries, kept in registers across the probes.
    float t3[PRECOMP ? kEntriesPerThread : 1];
    if (!IS_L2 || PRECOMP) {
        for (int i = threadIdx.x; i < d; i += blockDim.x)
            rs[i] = Q[(int64_t)q * d + i];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kEntriesPerThread; i++) {
            const int e = threadIdx.x + i * kThreads;
            const float dot = entry(e, false);
            if (PRECOMP)
                t3[i] = -2.f * dot;
            else
                store(e, -dot);
        }
        __syncthreads();
    }

    const int pEnd = min(nprobe, (chunk + 1) * probesPerCta);
    for (int p = chunk * probesPerCta; p < pEnd; p++) {
        const idx_t l = probes[(int64_t)q * nprobe + p];
        if (l < 0)
            continue; // block-uniform
        const int len = listLen[l];
        const int64_t ls = listStart[l];
        const uint8_t* codes = arenaCodes + ls * (int64_t)M;
        const int ngroups = (len + 31) >> 5;
        if (ngroups == 0)
            continue;
        float term1 = 0.f;
        if (IS_L2) {
            __syncthreads(); // every warp is done with the previous probe's LUT
            if (PRECOMP) {
                // LUT = T2[list] + term 3 (one coalesced load and one add per entry); term 1 = ||x - c||^2 is
                // recomputed here by every warp identically, so the result does not depend on the caller's
                // coarse distances (search == search_preassigned bit for bit)
                const float* t2 = term2 + (size_t)l * 256 * M;
                float v2[kEntriesPerThread];
#pragma unroll
                for (int i = 0; i < kEntriesPerThread; i++)
                    v2[i] = __ldg(t2 + threadIdx.x + i * kThreads);
                float part = 0.f;
                for (int i = lane; i < d; i += 32) {
                    const float df = rs[i] - __ldg(coarse + l * d + i);
                    part = fmaf(df, df, part);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    part += __shfl_xor_sync(kFullMask, part, o);
                term1 = part;
#pragma unroll
                for (int i = 0; i < kEntriesPerThread; i++)
                    store(threadIdx.x + i * kThreads, v2[i] + t3[i]);
            } else {
                for (int i = threadIdx.x; i < d; i += blockDim.x)
                    rs[i] = Q[(int64_t)q * d + i] - coarse[l * d + i];
                __syncthreads();
#pragma unroll LU
                for (int e = threadIdx.x; e < 256 * M; e += kThreads)
                    store(e, entry(e, true));
            }
            __syncthreads();
        }

        const float add = IS_L2 ? term1 : -coarseDis[(int64_t)q * nprobe + p];
        for (int g0 = warp * kU; g0 < ngroups; g0 += kStride) {
            uint4 cur[kU][M / 16];
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const int g = min(g0 + u, ngroups - 1); // clamped: tail groups re-read the last one (masked below)
                const uint4* gp = reinterpret_cast<const uint4*>(codes + (int64_t)g * 32 * M) + lane;
#pragma unroll
                for (int h = 0; h < M / 16; h++)
                    cur[u][h] = __ldg(gp + h * 32);
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int h = 0; h < M / 16; h++) {
                    const unsigned wds[4] = {cur[u][h].x, cur[u][h].y, cur[u][h].z, cur[u][h].w};
#pragma unroll
                    for (int wi = 0; wi < 4; wi++) {
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            const int j = h * 16 + wi * 4 + b;
                            // R = (byte << 8) | (lane << 2): LUT row of this code value + this lane's slot
                            const unsigned R = __byte_perm(wds[wi], lane4, 0x6504 | (b << 4));
                            const float val = *reinterpret_cast<const float*>(lutB + R + j * 4);
                            if (j & 1)
                                a1 += val;
                            else
                                a0 += val;
                        }
                    }
                }
                const int v = (g0 + u) * 32 + lane;
                const float key = (IS_L2 && !PRECOMP) ? a0 + a1 : (a0 + a1) + add;
                w.add(g0 + u < ngroups && v < len, key, (IdT)(ls + v));
            }
        }
    }
    merge_and_write<IdT, kWarps>(w, warp, lists, perWarp, LIST, k, arenaIds, oD, oI);
}

// T2[l][e] (e = c*M + m) = ||y_e||^2 + 2 <centroid_l | m, y_e>
__global__ void ivfpq_term2_kernel(
        const float* __restrict__ coarse, const float* __restrict__ pqT, int64_t nlist, int d, int M, float* __restrict__ term2) {
    const int64_t l = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 256 * M)
        return;
    const int dsub = d / M;
    const int m = e % M;
    const float* y = pqT + (size_t)e * dsub;
    const float* c = coarse + l * d + m * dsub;
    float acc = 0.f;
    for (int j = 0; j < dsub; j++) {
        const float yj = y[j];
        acc = fmaf(yj, yj, acc);
        acc = fmaf(2.f * c[j], yj, acc);
    }
    term2[(size_t)l * 256 * M + e] = acc;
}

} // namespace

void runIvfPqPrecomputeTerm2(
        const float* coarse, const float* pqT, int64_t nlist, int d, int M, float* term2, cudaStream_t stream) {
    if (nlist == 0)
        return;
    FB_THROW_IF_NOT_MSG(nlist <= 65535 * 32, "nlist too large for the term-2 precompute grid");
    // grid.y is limited to 65535: fold the lists into (x = entries, y = lists) with y chunks
    for (int64_t l0 = 0; l0 < nlist; l0 += 65535) {
        const int64_t nl = std::min<int64_t>(65535, nlist - l0);
        dim3 grid((unsigned)ceil_div(256 * M, 256), (unsigned)nl);
        ivfpq_term2_kernel<<<grid, 256, 0, stream>>>(coarse + l0 * d, pqT, nl, d, M, term2 + (size_t)l0 * 256 * M);
        CUDA_CHECK_LAST();
    }
}

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
        cudaStream_t stream) {
    if (n == 0)
        return;
    int warps = 8;
    pq_scatter_interleaved_kernel<<<(unsigned)ceil_div(n, warps), warps * 32, 0, stream>>>(
            codesFlat, ids, assign, offsets, n, M, listStart, arenaCodes, arenaIds);
    CUDA_CHECK_LAST();
}

void runIvfPqListToInterleaved(const uint8_t* flat, int64_t len, int M, uint8_t* listCodes, cudaStream_t stream) {
    if (len == 0)
        return;
    pq_list_to_interleaved_kernel<<<(unsigned)ceil_div(len * M, 256), 256, 0, stream>>>(flat, len, M, listCodes);
    CUDA_CHECK_LAST();
}

void runIvfPqListFromInterleaved(const uint8_t* listCodes, int64_t len, int M, uint8_t* flat, cudaStream_t stream) {
    if (len == 0)
        return;
    pq_list_from_interleaved_kernel<<<(unsigned)ceil_div(len * M, 256), 256, 0, stream>>>(listCodes, len, M, flat);
    CUDA_CHECK_LAST();
}

template <int M, bool IS_L2, bool PRECOMP, typename IdT, int kWarps, int LU>
static void launchScanV(
        dim3 grid,
        size_t smem,
        cudaStream_t stream,
        const float* Q,
        int d,
        const idx_t* probes,
        const float* coarseDis,
        int nprobe,
        int probesPerCta,
        const float* coarse,
        const float* pqT,
        const float* term2,
        const int64_t* listStart,
        const int* listLen,
        const uint8_t* codes,
        const idx_t* ids,
        int k,
        int LIST,
        float* partD,
        idx_t* partI) {
    auto kern = ivfpq_scan_interleaved_kernel<M, IS_L2, PRECOMP, IdT, kWarps, LU>;
    CUDA_VERIFY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KernelTiming::begin("ivfpq_scan", stream);
    kern<<<grid, kWarps * 32, smem, stream>>>(
            Q, d, probes, coarseDis, nprobe, probesPerCta, coarse, pqT, term2, listStart, listLen, codes, ids, k, LIST,
            partD, partI);
    KernelTiming::end("ivfpq_scan", stream);
    CUDA_CHECK_LAST();
}

constexpr int kScanWarps = 16; // warps per CTA (8 -> 16: 66 -> 60 ms on the N=100M workload)

template <int M, bool IS_L2, bool PRECOMP, typename IdT, typename... Args>
static void launchScan(Args... args) {
    launchScanV<M, IS_L2, PRECOMP, IdT, kScanWarps, 8>(args...);
}

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
        const float* term2,
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
        cudaStream_t stream) {
    if (nq == 0)
        return;
    FB_THROW_IF_NOT(ivfPqInterleavedSupported(M));
    const int LIST = std::max(64, next_pow2(k));
    const bool wide = arenaElems >= (int64_t(1) << 31) - 1; // arena positions need 64-bit list ids
    const size_t listBytes = wide ? SmemTopK<long long>::bytes(LIST, kBuf) : SmemTopK<int>::bytes(LIST, kBuf);
    size_t smem = sizeof(float) * 256 * kLutSlots + round_up(sizeof(float) * d, 16) + listBytes * kScanWarps;
    FB_THROW_IF_NOT_MSG(smem <= 220 * 1024, "LUT + top-k lists do not fit shared memory");
    const bool l2 = metric == METRIC_L2;
    int probesPerCta = 1;
    const int chunks = ivfScanChunks(device, nq, nprobe, &probesPerCta);
    const int64_t maxQ = std::max<int64_t>(1, std::min<int64_t>(65535, (int64_t(1) << 30) / ((int64_t)chunks * k * 12)));
    for (int64_t q0 = 0; q0 < nq; q0 += maxQ) {
        int64_t nb = std::min(maxQ, nq - q0);
        auto partD = res->temp(device, sizeof(float) * nb * chunks * k);
        auto partI = res->temp(device, sizeof(idx_t) * nb * chunks * k);
        dim3 grid((unsigned)chunks, (unsigned)nb);
#define SCAN(M_, L2_, PRE_, ID_)                                                                                   \
    launchScan<M_, L2_, PRE_, ID_>(                                                                                \
            grid, smem, stream, Q + q0 * d, d, probes + q0 * nprobe, coarseDis + q0 * nprobe, nprobe, probesPerCta, \
            coarseCentroids, pqCentroidsT, term2, listStart, listLen, arenaCodes, arenaIds, k, LIST,               \
            partD.as<float>(), partI.as<idx_t>())
#define SCAN_ID(M_, L2_, PRE_)          \
    do {                                \
        if (wide)                       \
            SCAN(M_, L2_, PRE_, long long); \
        else                            \
            SCAN(M_, L2_, PRE_, int);   \
    } while (0)
        const bool pre = l2 && term2 != nullptr;
        if (M == 32) {
            if (pre)
                SCAN_ID(32, true, true);
            else if (l2)
                SCAN_ID(32, true, false);
            else
                SCAN_ID(32, false, false);
        } else {
            if (pre)
                SCAN_ID(16, true, true);
            else if (l2)
                SCAN_ID(16, true, false);
            else
                SCAN_ID(16, false, false);
        }
#undef SCAN_ID
#undef SCAN
        runMergeTopKKeyspace(
                partD.as<float>(), partI.as<idx_t>(), nb, chunks, k, k, metric, 0, outD + q0 * k, outI + q0 * k, stream);
    }
}

} // namespace fb200
