// faiss_b200 -- the tcgen05 Flat scoring + filter kernel (included by flat_tc.cu).
//
// Roles in one persistent CTA (320 threads, one CTA per SM):
//   warp 0   : TMA producer -- database tiles (128 rows x dpad fp16, 128B-swizzled K-major) through a
//              multi-stage mbarrier ring, plus a small ring with each tile's bias row and tile id
//   warp 1   : single-thread tcgen05.mma issuer.  The A operand (queries) lives in TENSOR MEMORY
//              (TS mode): a work unit covers TWO query tiles (256 queries), so every database tile
//              that lands in shared memory feeds two MMA groups, and shared memory only serves B.
//              (SS mode with one query tile needs ~96 KB of smem traffic per 512-cycle tile and
//              measured ~50% of the MMA floor; profiles/r01_*.)
//   warps 2-9: epilogue.  A thread owns one TMEM lane = one query row of each of the unit's two query
//              tiles, and 64 of a tile's 128 columns.  It streams accumulators with tcgen05.ld in
//              32-column chunks (software-pipelined against the filter), computes
//              score = acc * inv + bias with FFMA2, folds the chunk with FMNMX3 and compares one
//              maximum per 8 columns against the query's threshold.  Survivors (rare) are appended
//              with plain stores to a thread-private candidate segment.
//   TMEM map : [0, dpad) columns = the two A tiles (dpad/2 columns each, fp16 pairs per column),
//              then (512 - dpad)/128 accumulator stages of 128 columns.
#pragma once

#include <cuda_fp16.h>

#include "select.cuh"
#include "tc_ptx.cuh"

namespace fb200 {
namespace tc {

constexpr int kTileM = 128;       // queries per MMA tile (TMEM lanes)
constexpr int kPairM = 256;       // queries per work unit (two MMA tiles)
constexpr int kTileN = 128;       // database rows per tile (TMEM columns per accumulator stage)
constexpr int kKBlock = 64;       // fp16 elements per 128-byte swizzle row
constexpr int kKBlockBytes = kTileN * kKBlock * 2; // 16 KiB per (128 rows x 64 halfs)
constexpr int kThreads = 320;
constexpr int kEpiWarps = 8;
constexpr int kMaxYStages = 6;
constexpr int kMaxAccStages = 4;
constexpr int kBiasSlots = 8;
constexpr int kSegsPerUnit = 512; // 256 query rows x 2 column halves

struct TcParams {
    int numUnits;
    int slices;
    int qPairs;         // unit u = slice * qPairs + pair: neighbouring CTAs stream the SAME database tiles (L2 reuse)
    int tileBegin;      // permuted position range of this round
    int tileEnd;
    int tilesPerSlice;
    unsigned long long permA, permB, numTiles;
    int KB;             // dpad / 64
    int yStages;
    int accStages;      // (512 - dpad) / 128
    const __half* Q16;  // [qPairs*256][dpad] scaled fp16 queries, zero padded
    const float* invScalePtr; // device scalar: 1 / (qScale * yScale)
    const float* bias;  // [numTiles*128], -inf padded
    const float* thr;   // [nq]  pass if score > thr
    uint2* cand;        // [numUnits*512][cap] (score bits, row)
    int cap;
    int* candCount;     // [numUnits*512]
    float* dump;        // debug: raw accumulators [nq][dumpLd]
    long long dumpLd;
    int nq;
    int debugSkip;      // timing experiments only: 1 = skip the filter (TMEM loads still issued)
};

__device__ __forceinline__ int perm_tile(const TcParams& p, int pos) {
    return (int)(((unsigned long long)pos * p.permA + p.permB) % p.numTiles);
}

// D[tmem] (+)= A[tmem] * B[smem desc]^T   (TS mode)
__device__ __forceinline__ void mma_f16_ts(
        uint32_t tmem_d,
        uint32_t tmem_a,
        uint64_t desc_b,
        uint32_t idesc,
        uint32_t accumulate) {
    asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
            "}\n" ::"r"(tmem_d),
            "r"(tmem_a),
            "l"(desc_b),
            "r"(idesc),
            "r"(accumulate)
            : "memory");
}

// 32 lanes x 32 columns registers -> TMEM
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
            "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
            "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
            "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
            "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
            : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Filter 32 columns of one query row: score = acc * inv + bias; one maximum per 8 columns against the
// query's threshold; the rare survivors are appended to the thread-private candidate segment.
template <bool DUMP>
__device__ __forceinline__ void epi_filter32(
        const TcParams& p,
        const uint32_t (&r)[32],
        int q,
        long long colBase, // global row index of column 0 of this chunk
        float inv,
        float thr,
        uint32_t bp, // shared address of the 32 biases
        uint2* buf,
        int& cnt) {
    if (DUMP) {
        if (q < p.nq) {
            float* dst = p.dump + (long long)q * p.dumpLd + colBase;
#pragma unroll
            for (int j = 0; j < 32; j++)
                dst[j] = __uint_as_float(r[j]);
        }
        return;
    }
    float v[32];
    float mg[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const float4 b0 = ptx::lds128(bp + (2 * g) * 16);
        const float4 b1 = ptx::lds128(bp + (2 * g + 1) * 16);
        const int o = 8 * g;
        ptx::fma2(v[o + 0], v[o + 1], __uint_as_float(r[o + 0]), __uint_as_float(r[o + 1]), inv, b0.x, b0.y);
        ptx::fma2(v[o + 2], v[o + 3], __uint_as_float(r[o + 2]), __uint_as_float(r[o + 3]), inv, b0.z, b0.w);
        ptx::fma2(v[o + 4], v[o + 5], __uint_as_float(r[o + 4]), __uint_as_float(r[o + 5]), inv, b1.x, b1.y);
        ptx::fma2(v[o + 6], v[o + 7], __uint_as_float(r[o + 6]), __uint_as_float(r[o + 7]), inv, b1.z, b1.w);
        const float a = ptx::max3(v[o + 0], v[o + 1], v[o + 2]);
        const float c = ptx::max3(v[o + 3], v[o + 4], v[o + 5]);
        mg[g] = ptx::max3(a, c, fmaxf(v[o + 6], v[o + 7]));
    }
    if (ptx::max3(mg[0], mg[1], fmaxf(mg[2], mg[3])) > thr) {
        const unsigned rowBase = (unsigned)colBase;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            if (mg[g] > thr) {
#pragma unroll
                for (int j = 8 * g; j < 8 * g + 8; j++) {
                    if (v[j] > thr) {
                        if (cnt < p.cap)
                            buf[cnt] = make_uint2(__float_as_uint(v[j]), rowBase + j);
                        cnt++;
                    }
                }
            }
        }
    }
}

template <bool DUMP>
__global__ void __launch_bounds__(kThreads, 1) flat_tc_kernel(const __grid_constant__ CUtensorMap mapY, const TcParams p) {
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte aligned carve-up (SWIZZLE_128B atoms need it)
    unsigned char* smem = reinterpret_cast<unsigned char*>(
            (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const int stageBytes = p.KB * kKBlockBytes;
    unsigned char* sY = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sY + (size_t)p.yStages * stageBytes);
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* y_full = bars + 2;
    uint64_t* y_empty = y_full + kMaxYStages;
    uint64_t* t_full = y_empty + kMaxYStages;
    uint64_t* t_empty = t_full + kMaxAccStages;
    uint64_t* b_full = t_empty + kMaxAccStages;
    uint64_t* b_empty = b_full + kBiasSlots;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + kBiasSlots);
    int* tileS = reinterpret_cast<int*>(tmem_slot + 2);                // [kBiasSlots]
    float* biasS = reinterpret_cast<float*>(tileS + kBiasSlots + 2);   // [kBiasSlots][128], 16B aligned

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int colsA = p.KB * 32;            // TMEM columns per A tile
    const uint32_t accBase = 2 * colsA;     // first accumulator column

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&mapY);
        ptx::mbar_init(a_full, kEpiWarps);
        ptx::mbar_init(a_empty, 1);
        for (int i = 0; i < p.yStages; i++) {
            ptx::mbar_init(&y_full[i], 1);
            ptx::mbar_init(&y_empty[i], 1);
        }
        for (int i = 0; i < p.accStages; i++) {
            ptx::mbar_init(&t_full[i], 1);
            ptx::mbar_init(&t_empty[i], kEpiWarps);
        }
        for (int i = 0; i < kBiasSlots; i++) {
            ptx::mbar_init(&b_full[i], 1);
            ptx::mbar_init(&b_empty[i], kEpiWarps);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc<512>(tmem_slot);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int ys = 0, bs = 0;
            uint32_t yphase = 0, bphase = 0;
            for (int u = blockIdx.x; u < p.numUnits && p.debugSkip < 3; u += gridDim.x) {
                const int sl = u / p.qPairs;
                const int pb = p.tileBegin + sl * p.tilesPerSlice;
                const int pe = min(p.tileEnd, pb + p.tilesPerSlice);
                for (int pp = pb; pp < pe; pp++) {
                    const int t = perm_tile(p, pp);
                    ptx::mbar_wait(&y_empty[ys], yphase ^ 1);
                    ptx::mbar_arrive_expect_tx(&y_full[ys], (uint32_t)stageBytes);
                    ptx::tma_load_3d(sY + (size_t)ys * stageBytes, &mapY, &y_full[ys], 0, t * kTileN, 0);
                    if (++ys == p.yStages) {
                        ys = 0;
                        yphase ^= 1;
                    }
                    // per-tile bias + tile id for the epilogue
                    ptx::mbar_wait(&b_empty[bs], bphase ^ 1);
                    ptx::sts32(ptx::smem_u32(tileS + bs), t);
                    ptx::mbar_arrive_expect_tx(&b_full[bs], kTileN * 4);
                    ptx::bulk_load_1d(biasS + bs * kTileN, p.bias + (long long)t * kTileN, kTileN * 4, &b_full[bs]);
                    if (++bs == kBiasSlots) {
                        bs = 0;
                        bphase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_f16(kTileM, kTileN);
            int ys = 0, as = 0;
            uint32_t yphase = 0, aphase = 0;
            int it = 0;
            const uint32_t sYaddr = ptx::smem_u32(sY);
            for (int u = blockIdx.x; u < p.numUnits; u += gridDim.x, it++) {
                const int sl = u / p.qPairs;
                const int pb = p.tileBegin + sl * p.tilesPerSlice;
                const int pe = min(p.tileEnd, pb + p.tilesPerSlice);
                ptx::mbar_wait(a_full, it & 1); // both query tiles are in tensor memory
                ptx::tc_fence_after();
                for (int pp = pb; pp < pe; pp++) {
                    if (p.debugSkip < 3)
                        ptx::mbar_wait(&y_full[ys], yphase);
                    ptx::tc_fence_after();
                    const uint32_t yaddr = sYaddr + (uint32_t)ys * (uint32_t)stageBytes;
#pragma unroll 1
                    for (int h = 0; h < 2; h++) {
                        if (p.debugSkip < 4)
                            ptx::mbar_wait(&t_empty[as], aphase ^ 1);
                        ptx::tc_fence_after();
                        const uint32_t dcol = tmem_base + accBase + (uint32_t)as * kTileN;
                        const uint32_t acol = tmem_base + (uint32_t)(h * colsA);
                        for (int kb = 0; kb < p.KB; kb++) {
#pragma unroll
                            for (int k4 = 0; k4 < 4; k4++) {
                                uint64_t db = ptx::make_smem_desc_sw128(yaddr + kb * kKBlockBytes + k4 * 32);
                                mma_f16_ts(dcol, acol + kb * 32 + k4 * 8, db, idesc, (kb | k4) != 0 ? 1u : 0u);
                            }
                        }
                        ptx::mma_commit(&t_full[as]); // accumulator stage ready for the epilogue
                        if (++as == p.accStages) {
                            as = 0;
                            aphase ^= 1;
                        }
                    }
                    if (p.debugSkip < 3)
                        ptx::mma_commit(&y_empty[ys]); // smem stage reusable once these MMAs retire
                    if (++ys == p.yStages) {
                        ys = 0;
                        yphase ^= 1;
                    }
                }
                ptx::mma_commit(a_empty); // the A tiles may be overwritten
            }
        }
    } else {
        // ================================ epilogue ================================
        const int ew = warp - 2;
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        const int half = ew >> 2;      // which 64 columns of a tile; also which A tile this warp loads
        const int row = quarter * 32 + lane;
        const float inv = *p.invScalePtr;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t lane_acc = lane_base + accBase + (uint32_t)(half * 64);
        int as = 0, bs = 0;
        uint32_t aphase = 0, bphase = 0;
        int it = 0;
        for (int u = blockIdx.x; u < p.numUnits; u += gridDim.x, it++) {
            const int pair = u % p.qPairs;
            const int sl = u / p.qPairs;
            // ---- load this warp's share of the A operand into tensor memory
            {
                ptx::mbar_wait(a_empty, (it & 1) ^ 1); // previous unit's MMAs have retired
                ptx::tc_fence_after();
                const long long qa = (long long)pair * kPairM + half * kTileM + row;
                const uint4* src = reinterpret_cast<const uint4*>(p.Q16 + qa * (p.KB * kKBlock));
                for (int kb = 0; kb < p.KB; kb++) {
                    uint32_t w[32];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const uint4 v = __ldg(src + kb * 8 + j);
                        w[4 * j + 0] = v.x;
                        w[4 * j + 1] = v.y;
                        w[4 * j + 2] = v.z;
                        w[4 * j + 3] = v.w;
                    }
                    tmem_st_32x32b_x32(lane_base + (uint32_t)(half * colsA + kb * 32), w);
                }
                tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0)
                    ptx::mbar_arrive(a_full);
            }
            // ---- per-thread filter state for its two queries (one per query tile of the pair)
            const int q0 = pair * kPairM + row;
            const int q1 = q0 + kTileM;
            const float thr0 = (!DUMP && q0 < p.nq) ? p.thr[q0] : CUDART_INF_F;
            const float thr1 = (!DUMP && q1 < p.nq) ? p.thr[q1] : CUDART_INF_F;
            const long long seg0 = ((long long)u * kPairM + row) * 2 + half;
            const long long seg1 = ((long long)u * kPairM + kTileM + row) * 2 + half;
            uint2* buf0 = DUMP ? nullptr : p.cand + seg0 * p.cap;
            uint2* buf1 = DUMP ? nullptr : p.cand + seg1 * p.cap;
            int cnt0 = 0, cnt1 = 0;
            const int pb = p.tileBegin + sl * p.tilesPerSlice;
            const int pe = min(p.tileEnd, pb + p.tilesPerSlice);

            // Software pipeline at 32-column granularity: while chunk A (columns 0..31 of this warp's
            // half) is filtered, the TMEM load of chunk B is in flight, and vice versa across stages.
            uint32_t ra[32], rb[32];
            ptx::mbar_wait(&t_full[as], aphase);
            ptx::tc_fence_after();
            if (p.debugSkip != 2)
                ptx::tmem_ld_32x32b_x32(lane_acc + (uint32_t)(as * kTileN), ra);
            for (int pp = pb; pp < pe; pp++) {
                if (p.debugSkip < 3)
                    ptx::mbar_wait(&b_full[bs], bphase);
                const int t = p.debugSkip < 3 ? ptx::lds32(ptx::smem_u32(tileS + bs)) : 0;
                const long long colBase = (long long)t * kTileN + half * 64;
                const uint32_t bp = ptx::smem_u32(biasS + bs * kTileN + half * 64);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int q = h ? q1 : q0;
                    const float thr = h ? thr1 : thr0;
                    uint2* buf = h ? buf1 : buf0;
                    int& cnt = h ? cnt1 : cnt0;
                    ptx::tmem_ld_wait(); // chunk A landed
                    if (p.debugSkip != 2)
                        ptx::tmem_ld_32x32b_x32(lane_acc + (uint32_t)(as * kTileN + 32), rb);
                    if (!p.debugSkip)
                        epi_filter32<DUMP>(p, ra, q, colBase, inv, thr, bp, buf, cnt);
                    ptx::tmem_ld_wait(); // chunk B landed: hand the accumulator stage back to the MMA warp
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0)
                        ptx::mbar_arrive(&t_empty[as]);
                    if (++as == p.accStages) {
                        as = 0;
                        aphase ^= 1;
                    }
                    if (h == 0 || pp + 1 < pe) {
                        ptx::mbar_wait(&t_full[as], aphase);
                        ptx::tc_fence_after();
                        if (p.debugSkip != 2)
                            ptx::tmem_ld_32x32b_x32(lane_acc + (uint32_t)(as * kTileN), ra);
                    }
                    if (!p.debugSkip)
                        epi_filter32<DUMP>(p, rb, q, colBase + 32, inv, thr, bp + 128, buf, cnt);
                }
                __syncwarp();
                if (lane == 0)
                    ptx::mbar_arrive(&b_empty[bs]);
                if (++bs == kBiasSlots) {
                    bs = 0;
                    bphase ^= 1;
                }
            }
            if (!DUMP) {
                p.candCount[seg0] = cnt0;
                p.candCount[seg1] = cnt1;
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<512>(tmem_base);
    }
}

} // namespace tc
} // namespace fb200
