// Probe: tcgen05.mma with disable_output_lane.  D[128 x 16] (FP32, TMEM) = A[128 x 16] (BF16, TMEM) * B_g[16 x 16]^T, issued
// once per "gene" g (rows 9g..9g+8 enabled, everything else disabled) with a different B_g each time.  Checks the values and
// times a long run of such small MMAs from one thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bench_micro/mask_probe tools/bench_micro/mask_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../chromoformer_b200/csrc/umma_ptx.cuh"
using namespace chromo;

__device__ __forceinline__ void umma_ts_masked(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc,
                                               uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc), "r"(m0), "r"(m1), "r"(m2), "r"(m3) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16f(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// A: [128][16] floats, B: [14][16][16] floats (gene, key row, k), D out: [128][16]
__global__ void __launch_bounds__(128, 1) mask_probe_kernel(const float* A, const float* B, float* D, long long* timing, int reps) {
    __shared__ __align__(1024) __nv_bfloat16 sB[14 * 16 * 16];     // per gene: K-major [16 rows x 16 k] = 2 core matrices rows x 2 k-chunks
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&slot, 64);
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    // B_g in the canonical K-major no-swizzle layout: (r, k) at (r/8)*SBO + (k/8)*128 + (r%8)*16 + (k%8)*2, SBO = 256 (K = 16)
    for (int i = tid; i < 14 * 16 * 16; i += 128) {
        const int g = i / 256, r = (i / 16) % 16, k = i % 16;
        const int off = g * 256 + ((r / 8) * 256 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2) / 2;
        sB[off] = __float2bfloat16(B[i]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    {   // A row of this thread -> BF16 pairs in TMEM columns [32, 40); D columns [0, 16) start as 1000 + row
        uint32_t ap[8];
        for (int c = 0; c < 8; ++c) {
            __nv_bfloat162 v = __floats2bfloat162_rn(A[tid * 16 + 2 * c], A[tid * 16 + 2 * c + 1]);
            ap[c] = *reinterpret_cast<uint32_t*>(&v);
        }
        tmem_st8(trow + 32, ap);
        uint32_t init[8];
        for (int c = 0; c < 8; ++c) init[c] = __float_as_uint(1000.f + tid);
        tmem_st8(trow + 0, init);
        tmem_st8(trow + 8, init);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, 16);
        const uint64_t bd0 = umma_smem_desc(smem_u32(sB), 128, 256);
        long long t0 = clock64();
        for (int rep = 0; rep < reps; ++rep)
#pragma unroll
            for (int g = 0; g < 14; ++g) {
                const int lo = 9 * g, hi = lo + 9;                     // enabled lanes [lo, hi)
                uint32_t m[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    uint32_t en = 0;
                    const int a0 = max(lo, 32 * w), a1 = min(hi, 32 * w + 32);
                    if (a1 > a0) en = ((a1 - a0 == 32) ? 0xffffffffu : ((1u << (a1 - a0)) - 1u)) << (a0 - 32 * w);
                    m[w] = ~en;
                }
                umma_ts_masked(tmem, tmem + 32, bd0 + 32 * g, idesc, 0, m[0], m[1], m[2], m[3]);   // +512 B per gene = +32
            }
        long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        timing[0] = t1 - t0; timing[1] = t2 - t0; timing[2] = 14LL * reps;
    }
    __syncthreads();
    tc_fence_after();
    float d[16];
    tmem_ld16f(trow, d);
    for (int c = 0; c < 16; ++c) D[tid * 16 + c] = d[c];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
    std::vector<float> A(128 * 16), B(14 * 256), D(128 * 16);
    auto bf = [](float x) { return __bfloat162float(__float2bfloat16(x)); };
    for (auto& x : A) x = bf((rand() % 17 - 8) / 8.f);
    for (auto& x : B) x = bf((rand() % 13 - 6) / 4.f);
    float *dA, *dB, *dD; long long* dT;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dT, 64);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    for (int reps : {1, 64}) {
        mask_probe_kernel<<<1, 128>>>(dA, dB, dD, dT, reps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
        long long t[3];
        cudaMemcpy(t, dT, 24, cudaMemcpyDeviceToHost);
        printf("reps %d: %lld masked N=16 MMAs, issue %.1f cyc/MMA, complete %.1f cyc/MMA\n", reps, t[2], (double)t[0] / t[2], (double)t[1] / t[2]);
    }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 128; ++r) {
        const int g = r / 9;
        for (int c = 0; c < 16; ++c) {
            float want;
            if (g < 14) { want = 0; for (int k = 0; k < 16; ++k) want += A[r * 16 + k] * B[g * 256 + c * 16 + k]; }
            else want = 1000.f + r;                                   // rows 126, 127: never enabled
            if (fabsf(D[r * 16 + c] - want) > 1e-3f) { if (bad < 8) printf("row %d col %d: got %f want %f\n", r, c, D[r * 16 + c], want); ++bad; }
        }
    }
    printf(bad ? "MASK PROBE FAILED: %d mismatches\n" : "mask probe ok (%d mismatches)\n", bad);
    return bad != 0;
}
