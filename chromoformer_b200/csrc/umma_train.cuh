// Tensor-core contractions of the training step (umma_train.cu).
#pragma once
#include "common.cuh"

namespace chromo {

// C[z][m, n] (=|+=) sum_kc opA(m, kc) * opB(n, kc), FP32 in / FP32 out, BF16 operands, FP32 accumulation in TMEM.
struct TGemmArgs {
    const float* A; long long lda, a_z; int a_t, a_div;   // a_t = 0: memory [m, kc]; 1: memory [kc, m].  memory row / a_div
    const float* B; long long ldb, b_z; int b_t, b_div;   // b_t = 0: memory [n, kc]; 1: memory [kc, n].  memory row / b_div
    float* C; long long ldc, c_z;                          // [M, N] row-major
    int M, N, Kc;
    int NT;                                                // filled in by tgemm_launch
    int ksplit;                                            // contraction split over CTAs (needs atomic)
    int atomic;                                            // 1: C += (FP32 atomics), 0: C = (plain stores)
};
bool tgemm_supported(const TGemmArgs& a);
int tgemm_launch(const TGemmArgs& a, int nz, cudaStream_t st);

}  // namespace chromo
