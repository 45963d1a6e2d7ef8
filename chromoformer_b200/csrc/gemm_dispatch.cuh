// One call site, two engines: the strict FP32 CUDA-core GEMM (gemm_simt.cuh) or, for the BF16 training step, the staged
// tcgen05 GEMM (train_gemm.cu) when the call maps onto it.
#pragma once
#include "gemm_simt.cuh"
#include "train_gemm.cuh"

namespace chromo {

// GemmArgs (A rows K-contiguous; B either [N,K] rows = nn.Linear weight, or [K,N]) -> TcGemm.  Batches: either the
// outer stride set (zdiv == 1) or pure head batching (nz == zdiv, outer strides unused).
static inline bool tc_from_gemm(const GemmArgs& g, bool a_kc, bool b_kc, int nz, TcGemm& t) {
    if (!a_kc || g.alpha != 1.f || g.c_bf16 || g.c_sqa_tiles) return false;
    const bool heads = g.zdiv > 1;
    if (heads && nz != g.zdiv) return false;
    t = tc_gemm_args();
    t.A = g.A; t.lda = g.lda; t.a_z = heads ? g.sA2 : g.sA1; t.a_div = g.a_div;
    t.B = g.B; t.ldb = g.ldb; t.b_z = heads ? g.sB2 : g.sB1; t.b_t = b_kc ? 0 : 1; t.b_div = b_kc ? 1 : g.b_div;
    t.C = g.C; t.ldc = g.ldc; t.c_z = heads ? g.sC2 : g.sC1; t.c_div = g.c_div; t.c_mul = g.c_mul; t.c_add = g.c_add;
    t.M = g.M; t.N = g.N; t.Kc = g.K;
    if (g.epi != EPI_PLAIN && g.bias) { t.epi |= TC_BIAS; t.bias = g.bias; t.bias_z = heads ? g.sBias2 : g.sBias1; }
    if (g.epi == EPI_BIAS_RELU) t.epi |= TC_RELU;
    if (g.epi == EPI_BIAS_RES_LN) {
        if (heads) return false;
        t.epi |= TC_RES | TC_LN;
        t.res = g.res; t.ldres = g.ldres; t.res_z = g.sRes1; t.res_div = g.res_div;
        t.gamma = g.gamma; t.beta = g.beta; t.ln_z = g.sLn1;
        t.pre = g.pre; t.pre_z = g.sPre1;
    } else if (g.accumulate) {          // C += : the output rows are their own residual
        if (g.c_div != 1 || g.c_add != 0 || g.res_plain) return false;
        t.epi |= TC_RES; t.res = g.C; t.ldres = (long long)g.ldc * g.c_mul; t.res_z = t.c_z;
    } else if (g.res_plain) {
        if (heads) return false;
        t.epi |= TC_RES; t.res = g.res; t.ldres = g.ldres; t.res_z = g.sRes1; t.res_div = g.res_div;
    }
    if (g.mask) {
        if (heads) return false;
        t.epi |= TC_MASK; t.mask = g.mask; t.ldmask = g.ldmask; t.mask_z = g.sMask1;
    }
    return true;
}

static inline int gemm_auto(const GemmArgs& g, bool a_kc, bool b_kc, int nz, cudaStream_t st, bool tc) {
    if (tc) {
        TcGemm t;
        if (tc_from_gemm(g, a_kc, b_kc, nz, t) && tc_gemm_supported(t)) return tc_gemm_launch(t, nz, st);
    }
    return gemm_launch(g, a_kc, b_kc, nz, st);
}

}  // namespace chromo
