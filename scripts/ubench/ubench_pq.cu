// Micro-benchmark of the IVF-PQ scan inner loop on synthetic data (no index build: seconds per run).
// Stream `bytes` of rotated interleaved-by-32 codes (M = 32) through CTAs that hold a 64 KB LUT in shared
// memory, exactly like ivfpq_scan_interleaved_kernel (faiss_b200/csrc/ivfpq_scan.cu), with the top-k
// replaced by a rare-pass threshold so that only the code stream + lookups + accumulation are measured.
//
// Variants (argv[1], default: all):
//   0  product loop: PRMT + LDS + FADD, two accumulators, 16 warps, kU = 4 groups per warp iteration
//   1  FADD2: lookups land in register pairs, one add.f32x2 per two lookups
//   2  as 0 with ld.global.nc.L1::no_allocate code loads (the stream is read once)
//   3  as 0 with 8 groups per warp iteration (needs 128 registers: 1 CTA of 16 warps per SM)
//   4  as 0 with 32 warps per CTA, one CTA per SM (64 KB LUT shared by twice the warps)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ubench_pq ubench_pq.cu
// Run:   ./ubench_pq [variant] [GiB of codes, default 8] [vectors per "list", default 24416]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 32;
constexpr int kLutSlots = 64;

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void add2(float& a0, float& a1, float v0, float v1) {
    asm("{\n.reg .b64 ra, rv, rd;\nmov.b64 ra, {%0,%1};\nmov.b64 rv, {%2,%3};\nadd.rn.f32x2 rd, ra, rv;\nmov.b64 {%0,%1}, rd;\n}"
        : "+f"(a0), "+f"(a1) : "f"(v0), "f"(v1));
}

// one "list" = groupsPerList groups of 32 vectors; CTA b scans lists b, b + gridDim.x, ...
template <int kWarps, int kU, bool FADD2, bool STREAM>
__global__ void __launch_bounds__(kWarps * 32) scan_kernel(
        const uint8_t* __restrict__ codes, long long numLists, int groupsPerList, float thr, unsigned long long* out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* lut = reinterpret_cast<float*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned char* lutB = reinterpret_cast<const unsigned char*>(lut);
    const unsigned lane4 = (unsigned)lane << 2;
    unsigned long long hits = 0;
    for (long long l = blockIdx.x; l < numLists; l += gridDim.x) {
        __syncthreads();
        // stand-in for the per-probe LUT build: 2 stores per entry, values in (0.5, 1.5)
        for (int e = threadIdx.x; e < 256 * M; e += kWarps * 32) {
            const int c = e / M, m = e - c * M;
            const float val = 0.5f + (float)((e * 2654435761u + (unsigned)l) >> 8 & 0xffff) * (1.f / 65536.f);
            lut[c * kLutSlots + m] = val;
            lut[c * kLutSlots + M + m] = val;
        }
        __syncthreads();
        const uint8_t* base = codes + l * (long long)groupsPerList * 32 * M;
        for (int g0 = warp * kU; g0 < groupsPerList; g0 += kWarps * kU) {
            uint4 c4[kU][M / 16];
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const int g = min(g0 + u, groupsPerList - 1);
                const uint4* gp = reinterpret_cast<const uint4*>(base + (long long)g * 32 * M) + lane;
#pragma unroll
                for (int h = 0; h < M / 16; h++)
                    c4[u][h] = STREAM ? ldg_stream(gp + h * 32) : __ldg(gp + h * 32);
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int h = 0; h < M / 16; h++) {
                    const unsigned wds[4] = {c4[u][h].x, c4[u][h].y, c4[u][h].z, c4[u][h].w};
#pragma unroll
                    for (int wi = 0; wi < 4; wi++) {
                        if (FADD2) {
#pragma unroll
                            for (int b = 0; b < 4; b += 2) {
                                const int j = h * 16 + wi * 4 + b;
                                const unsigned R0 = __byte_perm(wds[wi], lane4, 0x6504 | (b << 4));
                                const unsigned R1 = __byte_perm(wds[wi], lane4, 0x6504 | ((b + 1) << 4));
                                const float v0 = *reinterpret_cast<const float*>(lutB + R0 + j * 4);
                                const float v1 = *reinterpret_cast<const float*>(lutB + R1 + (j + 1) * 4);
                                add2(a0, a1, v0, v1);
                            }
                        } else {
#pragma unroll
                            for (int b = 0; b < 4; b++) {
                                const int j = h * 16 + wi * 4 + b;
                                const unsigned R = __byte_perm(wds[wi], lane4, 0x6504 | (b << 4));
                                const float val = *reinterpret_cast<const float*>(lutB + R + j * 4);
                                if (j & 1)
                                    a1 += val;
                                else
                                    a0 += val;
                            }
                        }
                    }
                }
                const float key = a0 + a1;
                if (g0 + u < groupsPerList && key < thr)
                    hits++;
            }
        }
    }
    if (hits)
        atomicAdd(out, hits);
}

// ---- v2 product loop: XOR layout, [256][2 tables][32 slots] LUT, lists in pairs, LDS base in the immediate
template <unsigned SEL>
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b) {
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "n"(SEL));
    return d;
}
template <int IMM>
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
    return v;
}
// MODE 0: as the product; 1: FADD2; 2: rolling prefetch (slot u reloaded with the next unit's group right after
// its lookups); 3: rolling prefetch + FADD2; 4: 8 hoisted column registers (no LOP3 in the loop)
template <int kWarps, int kU, int MODE, int SBASE>
__global__ void __launch_bounds__(kWarps * 32, 1024 / (kWarps * 32)) scan_v2_kernel(
        const uint8_t* __restrict__ codes, long long numLists, int groupsPerList, float thr, unsigned long long* out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* lut = reinterpret_cast<float*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(lut);
    if (sbase != (unsigned)SBASE) {
        if (threadIdx.x == 0 && blockIdx.x == 0)
            printf("smem base %u != %d\n", sbase, SBASE);
        return;
    }
    constexpr bool FADD2 = MODE == 1 || MODE == 3;
    constexpr bool ROLL = MODE == 2 || MODE == 3;
    const unsigned t4 = (unsigned)lane << 2;
    const unsigned P0 = t4 | ((t4 ^ 4u) << 8) | ((t4 ^ 8u) << 16) | ((t4 ^ 12u) << 24);
    unsigned Pw[8];
#pragma unroll
    for (int w = 0; w < 8; w++)
        Pw[w] = P0 ^ ((unsigned)w * 0x10101010u);
    unsigned long long hits = 0;
    const int units = (groupsPerList + kU - 1) / kU;
    for (long long l0 = 2 * (long long)blockIdx.x; l0 < numLists; l0 += 2 * (long long)gridDim.x) {
        __syncthreads();
        for (int s = 0; s < 2; s++)
            for (int e = threadIdx.x; e < 256 * M; e += kWarps * 32) {
                const int c = e / M, m = e - c * M;
                lut[c * 64 + s * 32 + m] = 0.5f + (float)((e * 2654435761u + (unsigned)(l0 + s)) >> 8 & 0xffff) * (1.f / 65536.f);
            }
        __syncthreads();
        // unit stream of this warp over the two lists: global unit i -> warp i % kWarps
        int uGlobal = warp; // in [0, 2*units)
        auto unitBase = [&](int ug) -> const uint8_t* {
            const int s = ug >= units ? 1 : 0;
            const long long l = min(l0 + s, numLists - 1);
            return codes + l * (long long)groupsPerList * 32 * M;
        };
        auto loadGroup = [&](int ug, int u, uint4 (&dst)[M / 16]) {
            const int s = ug >= units ? 1 : 0;
            const int g = min((ug - s * units) * kU + u, groupsPerList - 1);
            const uint4* gp = reinterpret_cast<const uint4*>(unitBase(ug) + (long long)g * 32 * M) + lane;
#pragma unroll
            for (int h = 0; h < M / 16; h++)
                dst[h] = __ldg(gp + h * 32);
        };
        uint4 c4[kU][M / 16];
        if (ROLL && uGlobal < 2 * units) {
#pragma unroll
            for (int u = 0; u < kU; u++)
                loadGroup(uGlobal, u, c4[u]);
        }
        for (; uGlobal < 2 * units; uGlobal += kWarps) {
            const int s = uGlobal >= units ? 1 : 0;
            if (!ROLL) {
#pragma unroll
                for (int u = 0; u < kU; u++)
                    loadGroup(uGlobal, u, c4[u]);
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int h = 0; h < M / 16; h++) {
                    const unsigned wds[4] = {c4[u][h].x, c4[u][h].y, c4[u][h].z, c4[u][h].w};
#pragma unroll
                    for (int wi = 0; wi < 4; wi++) {
                        const unsigned P = MODE == 4 ? Pw[h * 4 + wi] : (P0 ^ ((unsigned)(h * 4 + wi) * 0x10101010u));
                        const unsigned R0 = prmt<0xCC04>(wds[wi], P), R1 = prmt<0xCC15>(wds[wi], P);
                        const unsigned R2 = prmt<0xCC26>(wds[wi], P), R3 = prmt<0xCC37>(wds[wi], P);
                        float v0, v1, v2, v3;
                        if (s == 0) {
                            v0 = lds_f32<SBASE>(R0), v1 = lds_f32<SBASE>(R1), v2 = lds_f32<SBASE>(R2), v3 = lds_f32<SBASE>(R3);
                        } else {
                            v0 = lds_f32<SBASE + 128>(R0), v1 = lds_f32<SBASE + 128>(R1), v2 = lds_f32<SBASE + 128>(R2), v3 = lds_f32<SBASE + 128>(R3);
                        }
                        if (h == 0 && wi == 0) {
                            a0 = v0;
                            a1 = v1;
                        } else if (FADD2) {
                            add2(a0, a1, v0, v1);
                        } else {
                            a0 += v0;
                            a1 += v1;
                        }
                        if (FADD2) {
                            add2(a0, a1, v2, v3);
                        } else {
                            a0 += v2;
                            a1 += v3;
                        }
                    }
                }
                if (ROLL && uGlobal + kWarps < 2 * units)
                    loadGroup(uGlobal + kWarps, u, c4[u]); // this slot's registers are free again
                const float key = a0 + a1;
                if (key < thr)
                    hits++;
            }
        }
    }
    if (hits)
        atomicAdd(out, hits);
}

template <int kWarps, int kU, int MODE>
static void run2(const char* name, const uint8_t* codes, long long numLists, int groupsPerList, unsigned long long* out) {
    auto kern = scan_v2_kernel<kWarps, kU, MODE, 1024>;
    const size_t smem = sizeof(float) * 256 * 64 + 10 * 1024;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int perSm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, kWarps * 32, smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    const int grid = 148 * 64;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const float thr = 0.6f * M;
    for (int it = 0; it < 2; it++)
        kern<<<grid, kWarps * 32, smem>>>(codes, numLists, groupsPerList, thr, out);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 3;
    for (int it = 0; it < reps; it++)
        kern<<<grid, kWarps * 32, smem>>>(codes, numLists, groupsPerList, thr, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double bytes = (double)numLists * groupsPerList * 32 * M;
    printf("%-52s regs %3d  CTAs/SM %d  %8.2f ms  %7.1f GB/s\n", name, fa.numRegs, perSm, ms, bytes / ms / 1e6);
}

template <int kWarps, int kU, bool FADD2, bool STREAM>
static void run(const char* name, const uint8_t* codes, long long numLists, int groupsPerList, unsigned long long* out) {
    auto kern = scan_kernel<kWarps, kU, FADD2, STREAM>;
    const size_t smem = sizeof(float) * 256 * kLutSlots + 16 * 1024; // + the room the product's top-k lists take
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int perSm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, kWarps * 32, smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    const int grid = 148 * 64; // many CTAs, each walks lists round-robin (like 10k query CTAs)
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const float thr = 0.6f * M; // mean key = M: passes are rare
    for (int it = 0; it < 2; it++)
        kern<<<grid, kWarps * 32, smem>>>(codes, numLists, groupsPerList, thr, out);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 3;
    for (int it = 0; it < reps; it++)
        kern<<<grid, kWarps * 32, smem>>>(codes, numLists, groupsPerList, thr, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double bytes = (double)numLists * groupsPerList * 32 * M;
    printf("%-44s regs %3d  CTAs/SM %d  %8.2f ms  %7.1f GB/s  %6.2f G lookups/s\n", name, fa.numRegs, perSm, ms, bytes / ms / 1e6,
           bytes / ms / 1e6);
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : -1;
    const double gib = argc > 2 ? atof(argv[2]) : 8.0;
    const int vecsPerList = argc > 3 ? atoi(argv[3]) : 24416;
    const int groupsPerList = (vecsPerList + 31) / 32;
    const long long listBytes = (long long)groupsPerList * 32 * M;
    const long long numLists = (long long)(gib * (1ull << 30)) / listBytes;
    uint8_t* codes;
    CK(cudaMalloc(&codes, numLists * listBytes));
    // pseudo-random bytes (cheap fill: 4-byte LCG per word)
    {
        const size_t words = (size_t)numLists * listBytes / 4;
        uint32_t* h = (uint32_t*)malloc(64 << 20);
        uint32_t x = 12345u;
        for (size_t i = 0; i < (64u << 20) / 4; i++) {
            x = x * 1664525u + 1013904223u;
            h[i] = x ^ (x >> 15);
        }
        for (size_t off = 0; off < words * 4; off += (64u << 20))
            CK(cudaMemcpy(codes + off, h, std::min<size_t>(64u << 20, words * 4 - off), cudaMemcpyHostToDevice));
        free(h);
    }
    unsigned long long* out;
    CK(cudaMalloc(&out, 8));
    CK(cudaMemset(out, 0, 8));
    printf("codes: %.2f GiB, %lld lists of %d vectors (M=%d)\n", numLists * listBytes / double(1ull << 30), numLists, vecsPerList, M);
    if (variant < 0 || variant == 0)
        run<16, 4, false, false>("0 product loop (16 warps, kU=4)", codes, numLists, groupsPerList, out);
    if (variant < 0 || variant == 1)
        run<16, 4, true, false>("1 FADD2 accumulation", codes, numLists, groupsPerList, out);
    if (variant < 0 || variant == 2)
        run<16, 4, false, true>("2 L1::no_allocate code loads", codes, numLists, groupsPerList, out);
    if (variant < 0 || variant == 3)
        run<16, 8, false, false>("3 kU=8 (more bytes in flight per warp)", codes, numLists, groupsPerList, out);
    if (variant < 0 || variant == 4)
        run<32, 4, false, false>("4 32 warps per CTA", codes, numLists, groupsPerList, out);
    if (variant < 0 || variant == 10)
        run2<16, 4, 0>("10 v2: xor layout, 2 tables, list pairs", codes, numLists, groupsPerList, out);
    if (variant < 0 || variant == 11)
        run2<16, 4, 1>("11 v2 + FADD2", codes, numLists, groupsPerList, out);
    if (variant < 0 || variant == 12)
        run2<16, 4, 2>("12 v2 + rolling prefetch", codes, numLists, groupsPerList, out);
    if (variant < 0 || variant == 13)
        run2<16, 4, 3>("13 v2 + rolling prefetch + FADD2", codes, numLists, groupsPerList, out);
    if (variant < 0 || variant == 14)
        run2<16, 4, 4>("14 v2 + hoisted column registers", codes, numLists, groupsPerList, out);
    if (variant < 0 || variant == 15)
        run2<8, 4, 2>("15 v2 rolling, 8 warps/CTA (4 CTAs/SM?)", codes, numLists, groupsPerList, out);
    if (variant < 0 || variant == 16)
        run2<32, 4, 2>("16 v2 rolling, 32 warps/CTA", codes, numLists, groupsPerList, out);
    unsigned long long h = 0;
    CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
    printf("(threshold passes: %llu)\n", h);
    return 0;
}
