// Micro-benchmark: how many cycles does one tcgen05.mma (kind::f16, cta_group::1, M=128) cost when a single thread
// issues a long run of them?  Variants: operand A from shared memory (SS) or tensor memory (TS), N = 128 / 256, K-major
// no-swizzle operands exactly as reg_fused.cu lays them out.  One CTA per SM, everything else idle.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bench_micro/mma_rate tools/bench_micro/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../chromoformer_b200/csrc/umma_ptx.cuh"
using namespace chromo;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 SS N=128, 1 SS N=256, 2 TS N=128, 3 TS N=256, 4 SS N=128 with precomputed descriptors (adds only)
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int mode, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32768);       // A tile 32 KB, then 3-4 B stages
        const int N = (mode == 1 || mode == 3) ? 256 : 128;
        const uint32_t idesc = umma_idesc_bf16(128, N);
        const uint32_t sbo_b = (N == 256) ? 1024 : 2048;                        // [256 x 64] or [128 x 128] chunks of 32 KB
        const int ksteps = (N == 256) ? 4 : 8;
        long long t0 = clock64();
        int n = 0;
        if (mode == 4) {
            uint64_t ad = umma_smem_desc(a0, 128, 2048);
            for (int it = 0; it < iters; ++it) {
                uint64_t bd = umma_smem_desc(b0 + (it % 3) * 32768, 128, 2048);
#pragma unroll
                for (int k = 0; k < 8; ++k) { umma_bf16(tmem, ad + 16 * k, bd + 16 * k, idesc, k > 0); ++n; }
            }
        } else {
            for (int it = 0; it < iters; ++it) {
                const uint32_t b_addr = b0 + (it % 3) * 32768;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (k >= ksteps) break;
                    const uint64_t bd = umma_smem_desc(b_addr + k * 256, 128, sbo_b);
                    if (mode >= 2) umma_ts(tmem + 256 * 0, tmem + 448 + 8 * k, bd, idesc, k > 0);
                    else umma_bf16(tmem, umma_smem_desc(a0 + k * 256, 128, 2048), bd, idesc, k > 0);
                    ++n;
                }
            }
        }
        long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = n; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long* out;
    cudaMalloc(&out, 64);
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const char* names[] = {"SS N=128 (as reg_fused)", "SS N=256 ([256x64] chunks)", "TS N=128 (A in TMEM)", "TS N=256", "SS N=128, descriptors by add"};
    for (int grid : {1, 148}) {
        for (int mode = 0; mode < 5; ++mode) {
            long long h[3];
            for (int rep = 0; rep < 2; ++rep) {
                mma_rate_kernel<<<grid, 128, 160 * 1024>>>(mode, 64, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
            const double flops = 2.0 * 128 * ((mode == 1 || mode == 3) ? 256 : 128) * 16;
            printf("grid %3d  %-30s: %lld MMAs, issue %.1f cyc/MMA, complete %.1f cyc/MMA = %.0f FLOP/cyc/SM (floor %d cyc)\n", grid,
                   names[mode], h[2], (double)h[0] / h[2], (double)h[1] / h[2], flops * h[2] / h[1], (mode == 1 || mode == 3) ? 128 : 64);
        }
    }
    return 0;
}
