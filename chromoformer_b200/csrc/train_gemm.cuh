// Tensor-core contractions of the TRAINING step (train_gemm.cu): every nn.Linear on the path, its data gradient and
// its weight / bias gradients (train.py:195).
#pragma once
#include <vector>

#include "common.cuh"

namespace chromo {

// C[z][row(m), n] = epi( sum_kc opA(m, kc) * opB(n, kc) )      FP32 in HBM, BF16 operands (converted while staged),
//                                                               FP32 accumulation in TMEM
enum { TC_BIAS = 1, TC_RELU = 2, TC_MASK = 4, TC_RES = 8, TC_LN = 16 };
struct TcGemm {
    const float* A; long long lda, a_z; int a_t, a_div;    // a_t = 0: memory [m, kc];  1: memory [kc, m].  memory row / a_div
    const float* B; long long ldb, b_z; int b_t, b_div;    // b_t = 0: memory [n, kc] (nn.Linear weight);  1: memory [kc, n]
    float* C; long long ldc, c_z; int c_div, c_mul, c_add; // C row = (m / c_div) * c_mul + m % c_div + c_add
    int M, N, Kc;
    int epi;                                               // TC_* bits, applied in this order:
    const float* bias; long long bias_z;                   //   + bias[n]
    const float* res; long long ldres, res_z; int res_div; //   + res[m / res_div, n]
    float* pre; long long pre_z;                           //   (TC_LN) pre[m, n] = the value so far, then LayerNorm over the row
    const float* gamma; const float* beta; long long ln_z; //          (N == 128)
                                                           //   ReLU
    const float* mask; long long ldmask, mask_z;           //   zero where mask[m, n] <= 0   (ReLU backward)
    int NT, ksplit;                                        // filled in by tc_gemm_launch (ksplit > 1: contraction split over CTAs,
                                                           //   partial tiles added to a zeroed C with FP32 atomics)
};
TcGemm tc_gemm_args();
bool tc_gemm_supported(const TcGemm& g);
int tc_gemm_launch(const TcGemm& g, int nz, cudaStream_t st);
void tc_gemm_set_trace(long long* buf);      // chromo_debug_trace: phase clocks of CTA 0 at buf[2048..2056]

// Deferred weight / bias gradients: every  dW += dY^T X  and  db += colsum(dY)  of a backward pass is queued here and
// executed by ONE persistent launch at the end (their operands stay alive until then: backward.cu keeps each layer's
// gradient tensors in its own buffer).
struct WgItem {
    const float* A;        // dY  [tokens, m_rows]   (pre-offset to the tile's first output row)
    const float* B;        // X   [tokens / b_div, n_cols]  (pre-offset to the tile's first output column);  unused for bias items
    float* C;              // dW tile [m_rows, n_cols] (ldc)   |   db [n_cols] for bias items
    int lda, ldb, ldc, tokens;
    short m_rows, n_cols, b_div, kind;      // kind bit 0: bias (column sums of A[:, 0..n_cols)) instead of a weight tile;
                                            //      bit 1: the output is shared with another item (FP32 atomics)
};
class WgradQueue {
public:
    // dW[z] (+)= dY[z]^T X[z]:  dY [tokens, N] (ld ldy), X [tokens / x_div, K] (ld ldx), dW [N, K] row-major (ld lddw)
    void add_weight(const float* dY, int ldy, long long dy_z, const float* X, int ldx, int x_div, long long x_z, float* dW,
                    int lddw, long long dw_z, int tokens, int N, int K, int nz);
    void add_bias(const float* dY, int ld, long long dy_z, float* db, long long db_z, int tokens, int N, int nz);
    int flush(cudaStream_t st, int max_ctas = 0);   // one launch (per 640 items); max_ctas > 0 leaves SMs to other streams
    size_t size() const { return items_.size(); }
private:
    std::vector<WgItem> items_;
};

}  // namespace chromo
