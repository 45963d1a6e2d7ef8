// Validation metrics on the device (train.py:198-251,286-320: sklearn accuracy / roc_auc / average_precision for the
// classifier, sklearn r2_score and scipy pearsonr for the regressor), computed from the logits where they already are.
// No sort: the rank statistics are exact integer pair counts.
//
//   AUROC = sum_{i in pos} ( #{neg j: s_j < s_i} + 0.5 #{neg j: s_j == s_i} ) / (P N)
//           (= the trapezoidal area sklearn.metrics.roc_auc_score integrates, ties included)
//   AP    = (1/P) sum_{i in pos} TP_i / (TP_i + FP_i),  TP_i = #{pos j: s_j >= s_i}, FP_i = #{neg j: s_j >= s_i}
//           (= sklearn's step-wise sum over the DISTINCT thresholds: the positives tied at one threshold each add that
//            threshold's precision, together (TP_t - TP_prev) / P * precision_t)
//
// n is a validation set (4.7 k genes per fold, 18,955 at most): n^2 comparisons = 3.6e8, ~30 us on 148 SMs.
#include "common.cuh"

namespace chromo {

namespace {

constexpr int RM_THREADS = 128;
constexpr int RM_TILE = 2048;      // candidates staged per pass (score + label = 10 KB)

// scores of the positive class (softmax over the 2 logits, FP32 like Tensor.softmax) + argmax hits
__global__ void clf_scores_kernel(const float* __restrict__ logits, const int64_t* __restrict__ label, int n, int C,
                                  float* __restrict__ score, unsigned long long* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int hit = 0, pos = 0;
    if (i < n) {
        const float* row = logits + (long long)i * C;
        float mx = row[0];
        int arg = 0;
        for (int c = 1; c < C; ++c)
            if (row[c] > mx) { mx = row[c]; arg = c; }
        float z = 0.f;
        for (int c = 0; c < C; ++c) z += expf(row[c] - mx);
        score[i] = expf(row[1] - mx) / z;
        hit = arg == (int)label[i];
        pos = label[i] != 0;
    }
    const unsigned hm = __ballot_sync(0xffffffffu, hit), pm = __ballot_sync(0xffffffffu, pos);
    if ((threadIdx.x & 31) == 0) {
        if (hm) atomicAdd(&counts[0], (unsigned long long)__popc(hm));
        if (pm) atomicAdd(&counts[1], (unsigned long long)__popc(pm));
    }
}

// out[0] += sum over this CTA's positives of (less + 0.5 ties) ; out[1] += sum of TP/(TP+FP)
__global__ void __launch_bounds__(RM_THREADS) rank_pairs_kernel(const float* __restrict__ score,
                                                                const int64_t* __restrict__ label, int n,
                                                                double* __restrict__ out) {
    __shared__ float s_s[RM_TILE];
    __shared__ unsigned char s_l[RM_TILE];
    __shared__ double red[2][RM_THREADS / 32];
    const int i = blockIdx.x * RM_THREADS + threadIdx.x;
    const float si = i < n ? score[i] : 0.f;
    const bool mine = i < n && label[i] != 0;
    unsigned lt_neg = 0, eq_neg = 0, ge_pos = 0, ge_neg = 0;
    for (int j0 = 0; j0 < n; j0 += RM_TILE) {
        const int m = min(RM_TILE, n - j0);
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += RM_THREADS) {
            s_s[j] = score[j0 + j];
            s_l[j] = label[j0 + j] != 0;
        }
        __syncthreads();
        if (mine) {
#pragma unroll 8
            for (int j = 0; j < m; ++j) {
                const float sj = s_s[j];
                const unsigned p = s_l[j], q = 1u - p;
                const unsigned ge = sj >= si, eq = sj == si;
                ge_pos += ge & p;
                ge_neg += ge & q;
                eq_neg += eq & q;
                lt_neg += (1u - ge) & q;
            }
        }
    }
    double a = 0.0, b = 0.0;
    if (mine) {
        a = (double)lt_neg + 0.5 * (double)eq_neg;
        b = (double)ge_pos / (double)(ge_pos + ge_neg);
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < RM_THREADS / 32; ++w) { a += red[0][w]; b += red[1][w]; }
        atomicAdd(&out[0], a);
        atomicAdd(&out[1], b);
    }
}

__global__ void clf_finish_kernel(const unsigned long long* __restrict__ counts, const double* __restrict__ sums, int n,
                                  double* __restrict__ out) {
    const double P = (double)counts[1], N = (double)n - P;
    out[0] = (double)counts[0] / (double)n;                 // accuracy
    out[1] = sums[0] / (P * N);                             // AUROC (NaN when one class is absent: sklearn raises there)
    out[2] = sums[1] / P;                                   // average precision
    out[3] = P;
}

__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;
}

// one CTA, two passes in FP64: means, then centred sums (the order scipy.stats.pearsonr and sklearn.r2_score use)
__global__ void __launch_bounds__(1024) reg_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ label,
                                                           int n, double* __restrict__ out) {
    __shared__ double red[32];
    double sp = 0.0, sl = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { sp += (double)pred[i]; sl += (double)label[i]; }
    const double mp = block_sum(sp, red) / n, ml = block_sum(sl, red) / n;
    double cpl = 0.0, cpp = 0.0, cll = 0.0, res = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double p = (double)pred[i], l = (double)label[i];
        cpl += (p - mp) * (l - ml);
        cpp += (p - mp) * (p - mp);
        cll += (l - ml) * (l - ml);
        res += (l - p) * (l - p);
    }
    cpl = block_sum(cpl, red); cpp = block_sum(cpp, red); cll = block_sum(cll, red); res = block_sum(res, red);
    if (threadIdx.x == 0) {
        out[0] = 1.0 - res / cll;                           // r2_score(label, pred)
        out[1] = cpl / (sqrt(cpp) * sqrt(cll));             // pearsonr(label, pred)[0]
        out[2] = res / n;                                   // MSE (the validation loss of train.py:286-320)
    }
}

}  // namespace

}  // namespace chromo

extern "C" int chromo_clf_metrics(const float* logits, const int64_t* labels, int32_t n, int32_t n_out, float* score,
                                  double* out, void* scratch, void* stream) {
    using namespace chromo;
    if (!logits || !labels || !score || !out || !scratch || n < 1 || n_out < 2) {
        set_error("chromo_clf_metrics: bad argument");
        return CHROMO_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* counts = reinterpret_cast<unsigned long long*>(scratch);
    double* sums = reinterpret_cast<double*>(counts + 2);
    if (cudaMemsetAsync(scratch, 0, 32, st) != cudaSuccess) { set_error("chromo_clf_metrics: memset failed"); return CHROMO_ECUDA; }
    clf_scores_kernel<<<(n + 255) / 256, 256, 0, st>>>(logits, labels, n, n_out, score, counts);
    CHROMO_CHECK_LAUNCH("clf_scores");
    rank_pairs_kernel<<<(n + RM_THREADS - 1) / RM_THREADS, RM_THREADS, 0, st>>>(score, labels, n, sums);
    CHROMO_CHECK_LAUNCH("rank_pairs");
    clf_finish_kernel<<<1, 1, 0, st>>>(counts, sums, n, out);
    CHROMO_CHECK_LAUNCH("clf_finish");
    return CHROMO_OK;
}

extern "C" int chromo_reg_metrics(const float* pred, const float* labels, int32_t n, double* out, void* stream) {
    using namespace chromo;
    if (!pred || !labels || !out || n < 2) { set_error("chromo_reg_metrics: bad argument"); return CHROMO_EINVAL; }
    reg_metrics_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pred, labels, n, out);
    CHROMO_CHECK_LAUNCH("reg_metrics");
    return CHROMO_OK;
}
