// Hardware-semantics probes for the next step of the fused Regulation kernel (attention on the tensor
// pipe): (1) B operand in MN-major shared-memory layout, (2) A operand in TMEM written by tcgen05.st.
// D[128, N] = A[128, K] * B  with B given as [K, N] row-major FP32; M = 128, N % 16 == 0, K % 16 == 0.
// Test-only: built into tests/libchromo_probe.so (not into the product library) by __graft_entry__.build().
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../chromoformer_b200/csrc/umma_ptx.cuh"

namespace chromo {

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :
        : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// mode bit 0: B MN-major in shared memory (else K-major); mode bit 1: A in TMEM (else shared, K-major);
// mode bit 2 (with bit 0): the MN-major B is stored with its k-blocks far apart and its n-blocks adjacent
// (the bytes of a K-major [K, N] operand re-read as MN-major: LBO = N*16, SBO = 128)
__global__ void __launch_bounds__(128) umma_probe_kernel(int mode, const float* A, const float* B, float* D, int N, int K) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem);
    __nv_bfloat16* sB = reinterpret_cast<__nv_bfloat16*>(smem + 128 * K * 2);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (128 + N) * K * 2);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(slot, 512);
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *slot;
    const bool b_mn = mode & 1, a_tm = mode & 2, b_tr = (mode & 4) && b_mn;
    const int KC = K / 8;
    // A: shared K-major canonical, or TMEM (row = lane, two BF16 per 32-bit column, even k in the low half)
    if (!a_tm) {
        for (int i = tid; i < 128 * K; i += 128) {
            const int r = i / K, k = i % K;
            sA[((r >> 3) * KC + (k >> 3)) * 64 + (r & 7) * 8 + (k & 7)] = __float2bfloat16_rn(A[i]);
        }
    } else {
        const int r = warp * 32 + lane;
        for (int c0 = 0; c0 < K / 2; c0 += 32) {
            uint32_t regs[32];
            for (int j = 0; j < 32; ++j) {
                const int k = 2 * (c0 + j);
                float lo = k < K ? A[r * K + k] : 0.f, hi = k + 1 < K ? A[r * K + k + 1] : 0.f;
                __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
                regs[j] = *reinterpret_cast<uint32_t*>(&v);
            }
            tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c0, regs);
        }
    }
    // B: element (n, k) = B[k*N + n]
    for (int i = tid; i < N * K; i += 128) {
        const int k = i / N, n = i % N;
        const __nv_bfloat16 v = __float2bfloat16_rn(B[i]);
        if (!b_mn) sB[((n >> 3) * KC + (k >> 3)) * 64 + (n & 7) * 8 + (k & 7)] = v;          // K-major: LBO 128, SBO K*16
        else if (b_tr) sB[((k >> 3) * (N / 8) + (n >> 3)) * 64 + (k & 7) * 8 + (n & 7)] = v;   // k-blocks N*16 B apart
        else sB[((n >> 3) * KC + (k >> 3)) * 64 + (k & 7) * 8 + (n & 7)] = v;                 // MN-major: 8(k) x 8(n) cores
    }
    fence_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (tid == 0) {
        // idesc: + bit 16 = B MN-major
        const uint32_t idesc = umma_idesc_bf16(128, N) | (b_mn ? (1u << 16) : 0u);
        for (int k = 0; k < K / 16; ++k) {
            // K-major: advance 2 core matrices (256 B) per MMA.  MN-major: k-blocks are LBO = 128 B apart, n-blocks SBO = K*16
            const uint64_t bd = b_tr ? umma_smem_desc(smem_u32(sB) + k * 2 * N * 16, (uint32_t)N * 16, 128)
                                     : umma_smem_desc(smem_u32(sB) + k * 256, 128, (uint32_t)K * 16);
            if (!a_tm) umma_bf16(tmem, umma_smem_desc(smem_u32(sA) + k * 256, 128, (uint32_t)K * 16), bd, idesc, k > 0);
            else umma_bf16_ts(tmem, tmem + 256 + k * 8, bd, idesc, k > 0);
        }
        umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    tc_fence_after();
    float v[32];
    const int row = warp * 32 + lane;
    for (int c = 0; c < N; c += 32) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        for (int j = 0; j < 32 && c + j < N; ++j) D[row * N + c + j] = v[j];
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace chromo

using namespace chromo;

extern "C" int chromo_debug_umma_probe(int32_t mode, const float* A, const float* B, float* D, int32_t n, int32_t k,
                                       void* stream) {
    if (!A || !B || !D || n < 16 || n > 256 || n % 16 || k < 16 || k > 256 || k % 16) return -1;
    const size_t smem = (size_t)(128 + n) * k * 2 + 64;
    cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mode, A, B, D, n, k);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
