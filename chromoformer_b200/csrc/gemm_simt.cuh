// Generic batched FP32 GEMM on the CUDA cores with fused epilogues.
//
//   C[z][row(m), n] = epi( alpha * sum_k A[z](m, k) * B[z](k, n) )
//
// It is the strict-FP32 engine of every dense contraction of the model
// (nn.Linear call sites of modules.py:38,48,100-101,159-160 and net.py:42,118,
// 326-330) and of their weight/data gradients.  Operand layouts:
//   A_KC = true : A(m,k) = A[(m / a_div) * lda + k]     (K contiguous)
//   A_KC = false: A(m,k) = A[k * lda + m]               (M contiguous, "A^T")
//   B_KC = true : B(k,n) = B[n * ldb + k]               (nn.Linear weight [N,K])
//   B_KC = false: B(k,n) = B[k * ldb + n]
// blockIdx.z enumerates (z1, z2, ksplit) with two-level strides so that one
// launch covers e.g. all resolutions x all heads.
#pragma once
#include <string.h>

#include "common.cuh"

namespace chromo {

enum { EPI_PLAIN = 0, EPI_BIAS = 1, EPI_BIAS_RELU = 2, EPI_BIAS_RES_LN = 3 };

struct GemmArgs {
    const float* A; const float* B; float* C;
    int M, N, K;
    int lda, ldb, ldc;
    int zdiv;                                  // z = z1 * zdiv + z2
    long long sA1, sA2, sB1, sB2, sC1, sC2;
    int a_div;                                 // A row broadcast (>=1)
    int b_div;                                 // B k-index broadcast when !B_KC (>=1)
    int epi;
    const float* bias; long long sBias1, sBias2;
    const float* res; int ldres; int res_div; long long sRes1;
    const float* gamma; const float* beta; long long sLn1;
    float* pre; long long sPre1;               // optional pre-LayerNorm save [M,N]
    int c_div, c_mul, c_add;                   // C row = (m / c_div) * c_mul + m % c_div + c_add
    float alpha;
    int accumulate;                            // C += result
    int ksplit;                                // split-K factor; >1 => atomicAdd into C
    int c_bf16;                                // tensor path only: C is BF16 (ldc, strides in elements)
    int res_plain;                             // plain epilogues: C = acc (+ bias) + res[m / res_div, n]  (not with split-K)
    const float* mask; int ldmask; long long sMask1;   // zero where mask[m, n] <= 0 (ReLU backward; not with split-K)
    int c_sqa_tiles;                           // tensor path only: C [M, 256] = QK of (region m, head n / 128) written as the
                                               //   BF16 operand tiles of sqa_fused (row 2m + head, 128-row tiles in the
                                               //   canonical K-major layout); sC1 in BF16 elements
    // ragged plan (query GEMM of a pairwise stage, qk_tiles_kernel only): row m reads A row a_rows[m] / a_div, and only
    // the first *m_dev rows exist (both device pointers, nullptr = off)
    const int* a_rows; const int* m_dev;
};

static inline GemmArgs gemm_args() {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.zdiv = 1; g.a_div = 1; g.b_div = 1; g.res_div = 1; g.c_div = 1; g.c_mul = 1; g.c_add = 0;
    g.alpha = 1.f; g.ksplit = 1;
    return g;
}

constexpr int GEMM_BK = 16;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_PAD = 4;

template <int BD, bool KCONTIG>
__device__ __forceinline__ void gemm_load_tile(float (*S)[BD + GEMM_PAD], const float* __restrict__ src,
                                               int ld, int d0, int Dmax, int k0, int kend, int div,
                                               int tid) {
    if (KCONTIG) {
        constexpr int VPR = GEMM_BK / 4;
#pragma unroll
        for (int v = tid; v < BD * VPR; v += GEMM_THREADS) {
            const int d = v / VPR, kq = (v % VPR) * 4;
            const int gd = d0 + d, gk = k0 + kq;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gd < Dmax && gk < kend) {
                const float* p = src + (long long)(gd / div) * ld + gk;
                if (gk + 3 < kend && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
                    val = *reinterpret_cast<const float4*>(p);
                } else {
                    val.x = p[0];
                    if (gk + 1 < kend) val.y = p[1];
                    if (gk + 2 < kend) val.z = p[2];
                    if (gk + 3 < kend) val.w = p[3];
                }
            }
            S[kq + 0][d] = val.x; S[kq + 1][d] = val.y; S[kq + 2][d] = val.z; S[kq + 3][d] = val.w;
        }
    } else {
        constexpr int VPK = BD / 4;
#pragma unroll
        for (int v = tid; v < GEMM_BK * VPK; v += GEMM_THREADS) {
            const int k = v / VPK, dq = (v % VPK) * 4;
            const int gk = k0 + k, gd = d0 + dq;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gk < kend && gd < Dmax) {
                const float* p = src + (long long)(gk / div) * ld + gd;
                if (gd + 3 < Dmax && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
                    val = *reinterpret_cast<const float4*>(p);
                } else {
                    val.x = p[0];
                    if (gd + 1 < Dmax) val.y = p[1];
                    if (gd + 2 < Dmax) val.z = p[2];
                    if (gd + 3 < Dmax) val.w = p[3];
                }
            }
            *reinterpret_cast<float4*>(&S[k][dq]) = val;
        }
    }
}

// BM x BN tile, 4x4 register micro-tile, 256 threads.  LN = fused
// bias + residual + LayerNorm epilogue (requires BN == N == 128: one warp owns
// four complete rows).
template <int BM, int BN, bool A_KC, bool B_KC, bool LN>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_simt_kernel(const GemmArgs g) {
    CHROMO_PDL_ENTER();
    static_assert((BM / 4) * (BN / 4) == GEMM_THREADS, "tile/thread mismatch");
    __shared__ __align__(16) float As[GEMM_BK][BM + GEMM_PAD];
    __shared__ __align__(16) float Bs[GEMM_BK][BN + GEMM_PAD];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / 4), ty = tid / (BN / 4);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int zb = blockIdx.z / g.ksplit, ks = blockIdx.z % g.ksplit;
    const int z1 = zb / g.zdiv, z2 = zb % g.zdiv;

    const float* A = g.A + z1 * g.sA1 + z2 * g.sA2;
    const float* B = g.B + z1 * g.sB1 + z2 * g.sB2;
    float* C = g.C + z1 * g.sC1 + z2 * g.sC2;

    int kbeg = 0, kend = g.K;
    if (g.ksplit > 1) {
        int chunk = (g.K + g.ksplit - 1) / g.ksplit;
        chunk = (chunk + GEMM_BK - 1) / GEMM_BK * GEMM_BK;
        kbeg = ks * chunk;
        kend = min(g.K, kbeg + chunk);
    }

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = kbeg; k0 < kend; k0 += GEMM_BK) {
        gemm_load_tile<BM, A_KC>(As, A, g.lda, m0, g.M, k0, kend, A_KC ? g.a_div : 1, tid);
        gemm_load_tile<BN, B_KC>(Bs, B, g.ldb, n0, g.N, k0, kend, B_KC ? 1 : g.b_div, tid);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GEMM_BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ------------------------------------------------------------ epilogue --
    const float* bias = g.bias ? g.bias + z1 * g.sBias1 + z2 * g.sBias2 : nullptr;
    if (LN) {
        // N == BN == 128, tx == lane: columns tx*4..tx*4+3 of rows ty*4..ty*4+3.
        const float* res = g.res + z1 * g.sRes1;
        const float* gamma = g.gamma + z1 * g.sLn1;
        const float* beta = g.beta + z1 * g.sLn1;
        float* pre = g.pre ? g.pre + z1 * g.sPre1 : nullptr;
        const int n = n0 + tx * 4;
        const float4 bi = bias ? *reinterpret_cast<const float4*>(bias + n) : make_float4(0, 0, 0, 0);
        const float4 ga = *reinterpret_cast<const float4*>(gamma + n);
        const float4 be = *reinterpret_cast<const float4*>(beta + n);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + ty * 4 + i;
            const bool ok = m < g.M;   // warp-uniform
            float4 r = make_float4(0, 0, 0, 0);
            if (ok) r = *reinterpret_cast<const float4*>(res + (long long)(m / g.res_div) * g.ldres + n);
            float v0 = g.alpha * acc[i][0] + bi.x + r.x;
            float v1 = g.alpha * acc[i][1] + bi.y + r.y;
            float v2 = g.alpha * acc[i][2] + bi.z + r.z;
            float v3 = g.alpha * acc[i][3] + bi.w + r.w;
            if (pre && ok) *reinterpret_cast<float4*>(pre + (long long)m * g.N + n) = make_float4(v0, v1, v2, v3);
            float s = v0 + v1 + v2 + v3;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s * (1.f / 128.f);
            const float d0 = v0 - mean, d1 = v1 - mean, d2 = v2 - mean, d3 = v3 - mean;
            float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            const float rstd = rsqrtf(q * (1.f / 128.f) + 1e-5f);
            if (ok) {
                const long long crow = (long long)(m / g.c_div) * g.c_mul + (m % g.c_div) + g.c_add;
                *reinterpret_cast<float4*>(C + crow * g.ldc + n) =
                    make_float4(d0 * rstd * ga.x + be.x, d1 * rstd * ga.y + be.y,
                                d2 * rstd * ga.z + be.z, d3 * rstd * ga.w + be.w);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + ty * 4 + i;
            if (m >= g.M) continue;
            const long long crow = (long long)(m / g.c_div) * g.c_mul + (m % g.c_div) + g.c_add;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tx * 4 + j;
                if (n >= g.N) continue;
                float v = g.alpha * acc[i][j];
                if (g.epi != EPI_PLAIN && bias) v += bias[n];
                if (g.res_plain) v += g.res[z1 * g.sRes1 + (long long)(m / g.res_div) * g.ldres + n];
                if (g.epi == EPI_BIAS_RELU) v = fmaxf(v, 0.f);
                if (g.mask && !(g.mask[z1 * g.sMask1 + (long long)m * g.ldmask + n] > 0.f)) v = 0.f;
                float* c = C + crow * g.ldc + n;
                if (g.ksplit > 1) atomicAdd(c, v);
                else if (g.accumulate) *c += v;
                else *c = v;
            }
        }
    }
}

// Host launcher.  nz = number of (z1,z2) batches.
template <bool A_KC, bool B_KC>
static inline int gemm_launch_plain(const GemmArgs& g, int nz, cudaStream_t st) {
    dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, nz * g.ksplit);
    launch_pdl(gemm_simt_kernel<64, 64, A_KC, B_KC, false>, dim3(grid), dim3(GEMM_THREADS), 0, st, g);
    CHROMO_CHECK_LAUNCH("gemm_simt");
    return CHROMO_OK;
}

static inline int gemm_launch(const GemmArgs& g, bool a_kc, bool b_kc, int nz, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0) return CHROMO_OK;
    if (g.epi == EPI_BIAS_RES_LN) {
        if (g.N != 128 || !a_kc || !b_kc || g.ksplit != 1) {
            set_error("LayerNorm epilogue needs N == 128 and K-contiguous operands");
            return CHROMO_EINVAL;
        }
        dim3 grid(1, (g.M + 31) / 32, nz);
        launch_pdl(gemm_simt_kernel<32, 128, true, true, true>, dim3(grid), dim3(GEMM_THREADS), 0, st, g);
        CHROMO_CHECK_LAUNCH("gemm_simt_ln");
        return CHROMO_OK;
    }
    if (a_kc && b_kc) return gemm_launch_plain<true, true>(g, nz, st);
    if (a_kc && !b_kc) return gemm_launch_plain<true, false>(g, nz, st);
    if (!a_kc && b_kc) return gemm_launch_plain<false, true>(g, nz, st);
    return gemm_launch_plain<false, false>(g, nz, st);
}

}  // namespace chromo
