// Tensor-core contractions of the TRAINING step (train.py:195: the autograd of every nn.Linear on the path):
//
//   data gradient    dX[m, k]  (+)= sum_n dY[m, n] * W[n, k]          A = dY rows (K-major), B = W rows   (MN-major)
//   weight gradient  dW[n, k]   +=  sum_m dY[m, n] * X[m, k]          A = dY rows (MN-major), B = X rows  (MN-major)
//
// Both operands are FP32 row-major activations / parameters that change every step, so nothing is pre-packed: a CTA
// converts its [128 x 128] A chunk and [NT x 128] B chunk to BF16 while staging them into shared memory in the canonical
// no-swizzle UMMA layout of the wanted major-ness (8 x 8 core matrices, 16-byte rows), issues the chunk's
// tcgen05.mma (M = 128, N = NT <= 256, K = 16) into a TMEM accumulator and moves on to the next contraction chunk
// (two shared-memory stages: staging of chunk i+1 runs under the MMAs of chunk i).  The contraction can be split
// over CTAs (gridDim.z) for accumulating outputs: partial tiles are added with FP32 atomics, which is what gives the
// small training batch (a few hundred rows) enough CTAs to matter.
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma_ptx.cuh"
#include "umma_train.cuh"

namespace chromo {

namespace {

constexpr int TG_THREADS = 256;
constexpr int TG_KCH = 128;                    // contraction elements per chunk
constexpr uint32_t TG_A_BYTES = 128 * TG_KCH * 2;

__device__ __forceinline__ uint4 pack8(const float4& a, const float4& b) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
    pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
    return pk;
}

// Stage one operand chunk: `rows` MN-indices (tile) x TG_KCH contraction indices (chunk) -> BF16 in `dst`.
//   mn_major == 0: memory is [mn, kc] (kc contiguous)  -> K-major cores:  ((mn/8)*16 + kc/8)*128 + (mn%8)*16 + (kc%8)*2
//   mn_major == 1: memory is [kc, mn] (mn contiguous)  -> MN-major cores: ((mn/8)*16 + kc/8)*128 + (kc%8)*16 + (mn%8)*2
// Out-of-range elements are zeros.  `div` broadcasts memory ROWS (row index / div).
__device__ __forceinline__ void stage_operand(uint8_t* dst, const float* src, long long ld, int mn_major, int div, int rows,
                                              int mn0, int mn_lim, int kc0, int kc_lim, int warp, int lane) {
    if (!mn_major) {
        const int units = (rows / 8) * 4;                       // 8 mn rows x 32 kc columns per warp pass
        for (int u = warp; u < units; u += TG_THREADS / 32) {
            const int r = (u >> 2) * 8 + (lane >> 2), c8 = (u & 3) * 4 + (lane & 3);
            const int mn = mn0 + r, kc = kc0 + c8 * 8;
            float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
            if (mn < mn_lim && kc < kc_lim) {
                const float* p = src + (long long)(mn / div) * ld + kc;
                if (kc + 8 <= kc_lim) {
                    x0 = __ldg(reinterpret_cast<const float4*>(p));
                    x1 = __ldg(reinterpret_cast<const float4*>(p + 4));
                } else {
                    float t[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t[j] = kc + j < kc_lim ? __ldg(p + j) : 0.f;
                    x0 = make_float4(t[0], t[1], t[2], t[3]); x1 = make_float4(t[4], t[5], t[6], t[7]);
                }
            }
            *reinterpret_cast<uint4*>(dst + ((r >> 3) * 16 + c8) * 128 + (r & 7) * 16) = pack8(x0, x1);
        }
    } else {
        const int rg_n = (rows + 31) / 32;
        const int units = (TG_KCH / 8) * rg_n;                  // 8 kc rows x 32 mn columns per warp pass
        for (int u = warp; u < units; u += TG_THREADS / 32) {
            const int k = (u / rg_n) * 8 + (lane >> 2), r8 = (u % rg_n) * 4 + (lane & 3);
            const int kc = kc0 + k, mn = mn0 + r8 * 8;
            if (r8 * 8 >= rows) continue;                       // (tile narrower than the last 32-column group)
            float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
            if (kc < kc_lim && mn < mn_lim) {
                const float* p = src + (long long)(kc / div) * ld + mn;
                if (mn + 8 <= mn_lim) {
                    x0 = __ldg(reinterpret_cast<const float4*>(p));
                    x1 = __ldg(reinterpret_cast<const float4*>(p + 4));
                } else {
                    float t[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t[j] = mn + j < mn_lim ? __ldg(p + j) : 0.f;
                    x0 = make_float4(t[0], t[1], t[2], t[3]); x1 = make_float4(t[4], t[5], t[6], t[7]);
                }
            }
            *reinterpret_cast<uint4*>(dst + (r8 * 16 + (k >> 3)) * 128 + (k & 7) * 16) = pack8(x0, x1);
        }
    }
}

__global__ void __launch_bounds__(TG_THREADS) umma_staged_gemm_kernel(const TGemmArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int NT = a.NT;
    const uint32_t b_bytes = (uint32_t)NT * TG_KCH * 2, stage_bytes = TG_A_BYTES + b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_bytes);      // [0,1] stage free, [2] all done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * NT, m0 = blockIdx.y * 128;
    const int z = blockIdx.z / a.ksplit, split = blockIdx.z % a.ksplit;
    const float* A = a.A + z * a.a_z;
    const float* B = a.B + z * a.b_z;
    float* C = a.C + z * a.c_z;
    int tmem_cols = 32;
    while (tmem_cols < NT) tmem_cols *= 2;

    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int chunks = (a.Kc + TG_KCH - 1) / TG_KCH;
    const int per = (chunks + a.ksplit - 1) / a.ksplit;
    const int c_lo = split * per, c_hi = min(chunks, c_lo + per);
    // idesc: M = 128, N = NT, + bit 15 = A MN-major, bit 16 = B MN-major
    const uint32_t idesc = umma_idesc_bf16(128, NT) | (a.a_t ? (1u << 15) : 0u) | (a.b_t ? (1u << 16) : 0u);
    for (int ci = c_lo; ci < c_hi; ++ci) {
        const int i = ci - c_lo, st = i & 1;
        uint8_t* sA = smem + st * stage_bytes;
        uint8_t* sB = sA + TG_A_BYTES;
        if (i >= 2) mbar_wait(&bars[st], ((i >> 1) - 1) & 1);           // the MMAs of chunk i-2 have left this stage
        const int kc0 = ci * TG_KCH;
        stage_operand(sA, A, a.lda, a.a_t, a.a_div, 128, m0, a.M, kc0, a.Kc, warp, lane);
        stage_operand(sB, B, a.ldb, a.b_t, a.b_div, NT, n0, a.N, kc0, a.Kc, warp, lane);
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const int ksteps = (min(TG_KCH, a.Kc - kc0) + 15) / 16;
            const uint64_t ad = umma_smem_desc(smem_u32(sA), 128, 2048), bd = umma_smem_desc(smem_u32(sB), 128, 2048);
            for (int k = 0; k < ksteps; ++k) umma_bf16(tmem, ad + 16 * k, bd + 16 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
            umma_commit(&bars[st]);
            if (ci == c_hi - 1) umma_commit(&bars[2]);
        }
    }
    if (c_hi > c_lo) {
        mbar_wait(&bars[2], 0);
        tc_fence_after();
        const int lq = warp & 3, ch = warp >> 2;
        const int m = m0 + lq * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16);
        float v[32];
        for (int c = ch * 32; c < NT; c += 64) {
            tmem_ld32(trow + c, v);
            if (m < a.M) {
                float* out = C + (long long)m * a.ldc + n0 + c;
                const int ncols = min(32, min(NT, a.N - n0) - c);
                if (a.atomic) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncols) atomicAdd(out + j, v[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (j < ncols) *reinterpret_cast<float4*>(out + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

int choose_nt(int N) {
    if (N % 16 != 0) return 0;
    if (N <= 256) return N;
    for (int nt = 256; nt >= 16; nt -= 16)
        if (N % nt == 0) return nt;
    return 0;
}

}  // namespace

bool tgemm_supported(const TGemmArgs& a) {
    if (a.M < 16 || a.Kc < 16 || choose_nt(a.N) == 0) return false;
    if ((a.lda | a.ldb | a.ldc) % 4 != 0) return false;
    if ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.B) | reinterpret_cast<uintptr_t>(a.C)) & 15) return false;
    if ((a.a_z | a.b_z | a.c_z) % 4 != 0) return false;
    // vector loads run along the contiguous dimension of each operand: its extent has to be a multiple of 4 floats there
    if (a.a_t ? (a.M % 4 != 0) : (a.Kc % 4 != 0)) return false;
    if (a.b_t ? (a.N % 4 != 0) : (a.Kc % 4 != 0)) return false;
    return true;
}

int tgemm_launch(const TGemmArgs& in, int nz, cudaStream_t st) {
    TGemmArgs a = in;
    a.NT = choose_nt(a.N);
    if (a.NT == 0) { set_error("tgemm: N must be a multiple of 16"); return CHROMO_EINVAL; }
    if (a.ksplit < 1) a.ksplit = 1;
    const int chunks = (a.Kc + TG_KCH - 1) / TG_KCH;
    if (a.ksplit > chunks) a.ksplit = chunks;
    if (a.ksplit > 1 && !a.atomic) { set_error("tgemm: a split contraction needs an accumulating output"); return CHROMO_EINVAL; }
    const size_t smem = 2 * ((size_t)TG_A_BYTES + (size_t)a.NT * TG_KCH * 2) + 64;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(umma_staged_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("tgemm smem attribute: %s", cudaGetErrorString(e)); return CHROMO_ECUDA; }
        configured = smem;
    }
    dim3 grid(a.N / a.NT, (a.M + 127) / 128, nz * a.ksplit);
    umma_staged_gemm_kernel<<<grid, TG_THREADS, smem, st>>>(a);
    CHROMO_CHECK_LAUNCH("umma_staged_gemm");
    return CHROMO_OK;
}

}  // namespace chromo

// C ABI: the general contraction of the training step (see include/chromoformer_b200.h)
extern "C" int chromo_matmul(const float* A, int64_t lda, int32_t a_transposed, const float* B, int64_t ldb, int32_t b_transposed,
                             float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t accumulate, int32_t ksplit,
                             void* stream) {
    using namespace chromo;
    if (!A || !B || !C || M < 1 || N < 1 || K < 1) { set_error("chromo_matmul: bad argument"); return CHROMO_EINVAL; }
    TGemmArgs a;
    a.A = A; a.lda = lda; a.a_z = 0; a.a_t = a_transposed ? 1 : 0; a.a_div = 1;
    a.B = B; a.ldb = ldb; a.b_z = 0; a.b_t = b_transposed ? 1 : 0; a.b_div = 1;
    a.C = C; a.ldc = ldc; a.c_z = 0; a.M = M; a.N = N; a.Kc = K; a.NT = 0;
    a.ksplit = ksplit < 1 ? 1 : ksplit; a.atomic = accumulate ? 1 : 0;
    if (!tgemm_supported(a)) { set_error("chromo_matmul: shape / alignment not supported by the tensor-core path"); return CHROMO_EINVAL; }
    return tgemm_launch(a, 1, (cudaStream_t)stream);
}
