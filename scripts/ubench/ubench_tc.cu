// micro-benchmarks: (1) tcgen05.ld bandwidth, (2) tcgen05.mma issue rate for SS/TS x N=128/256, single CTA per SM
#include <cstdio>
#include <cuda_runtime.h>
#include "../../faiss_b200/csrc/tc_ptx.cuh"
using namespace fb200;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: LDTM only (8 warps), mode 1: SS N=128, 2: SS N=256, 3: TS N=128, 4: TS N=256, 5: SS N=128 + LDTM concurrently, 6: SS N=256 + LDTM
__global__ void __launch_bounds__(320, 1) ubench(int mode, int iters, long long* out, int commitEvery, int ncommit) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint64_t dummy[4];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); for (int i = 0; i < 4; i++) ptx::mbar_init(&dummy[i], 1); ptx::fence_barrier_init(); }
    if (warp == 1) ptx::tmem_alloc<512>(&slot);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tb = slot;
    // fill smem with small fp16 values
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    __syncthreads();
    long long t0 = clock64();
    const bool doMma = mode >= 1;
    const bool doLd = mode == 0 || mode >= 5;
    if (warp == 1 && lane == 0 && doMma) {
        const int N = (mode == 2 || mode == 4 || mode == 6) ? 256 : 128;
        const bool ts = (mode == 3 || mode == 4);
        const uint32_t idesc = ptx::make_idesc_f16(128, N);
        const uint32_t sa = ptx::smem_u32(smem), sb = ptx::smem_u32(smem + 32768);
        for (int it = 0; it < iters; it++) {
            const uint32_t dcol = tb + 128 + (it & 1) * (N == 256 ? 0 : 128);
            for (int kb = 0; kb < 2; kb++)
                for (int k4 = 0; k4 < 4; k4++) {
                    uint64_t db = ptx::make_smem_desc_sw128(sb + kb * (N * 128) + k4 * 32);
                    if (ts) mma_ts(dcol, tb + kb * 32 + k4 * 8, db, idesc, (kb | k4) ? 1u : 0u);
                    else {
                        uint64_t da = ptx::make_smem_desc_sw128(sa + kb * 16384 + k4 * 32);
                        ptx::mma_f16_ss(dcol, da, db, idesc, (kb | k4) ? 1u : 0u);
                    }
                }
            if (commitEvery && ((it + 1) % commitEvery == 0))
                for (int c = 0; c < ncommit; c++) ptx::mma_commit(&dummy[c]);
        }
        ptx::mma_commit(&bar);
        ptx::mbar_wait(&bar, 0);
    }
    if (warp >= 2 && doLd) {
        const int q = warp & 3;
        uint32_t r[32];
        uint32_t acc = 0;
        for (int it = 0; it < iters * 2; it++) {
            ptx::tmem_ld_32x32b_x32(tb + ((uint32_t)(q * 32) << 16) + 128 + ((warp - 2) >> 2) * 64 + (it & 1) * 32, r);
            ptx::tmem_ld_wait();
            acc += r[0] ^ r[31];
        }
        if (acc == 0x12345) out[1] = acc;
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tb); }
}

int main() {
    long long* out; cudaMalloc(&out, 64);
    cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 20000;
    const char* names[] = {"LDTM only (8 warps, x32)", "MMA SS N=128", "MMA SS N=256", "MMA TS N=128", "MMA TS N=256", "MMA SS N=128 + LDTM", "MMA SS N=256 + LDTM"};
    struct Cfg { int mode, ce, nc; const char* name; } cfgs[] = {
        {1,0,0,"SS N=128 no commit"}, {1,1,1,"SS N=128 commit/8 MMAs x1"}, {1,1,2,"SS N=128 commit/8 MMAs x2"}, {1,2,2,"SS N=128 commit/16 MMAs x2"}, {1,4,2,"SS N=128 commit/32 MMAs x2"},
        {3,0,0,"TS N=128 no commit"}, {3,1,2,"TS N=128 commit/8 x2"}, {3,2,2,"TS N=128 commit/16 x2"},
        {2,0,0,"SS N=256 no commit"}, {2,1,2,"SS N=256 commit/8 x2"}, {4,0,0,"TS N=256 no commit"}, {4,1,2,"TS N=256 commit/8 x2"}, {0,0,0,"LDTM only"}};
    for (auto& c : cfgs) {
        int mode = c.mode;
        for (int rep = 0; rep < 2; rep++) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            ubench<<<148, 320, 200 * 1024>>>(mode, iters, out, c.ce, c.nc);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            long long cyc; cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaGetLastError();
            if (rep == 1) {
                double perIt = (double)cyc / iters;
                int N = (mode == 2 || mode == 4 || mode == 6) ? 256 : 128;
                double flops = mode >= 1 ? 2.0 * 128 * N * 128 * iters * 148 / (ms * 1e-3) / 1e12 : 0;
                // LDTM bytes per iteration: 8 warps * 2 loads * 32 lanes * 32 regs * 4 B = 64 KB
                double ldBytesPerCyc = (mode == 0 || mode >= 5) ? 65536.0 / perIt : 0;
                printf("%-30s %8.1f cycles/iter  %7.1f TFLOP/s  LDTM %.1f B/cyc  (%.2f ms, err %d)\n", c.name, perIt, flops, ldBytesPerCyc, ms, (int)e);
            }
        }
    }
    return 0;
}
