// Losses (train.py:156,193) and the fused AdamW update (train.py:157,196).
#include "common.cuh"

namespace chromo {

// MSELoss, mean reduction: loss = mean((y - t)^2); dy = 2 (y - t) / count * grad_scale.
__global__ void mse_loss_kernel(const float* __restrict__ y, const float* __restrict__ t, int count,
                                float grad_scale, float* loss, float* dy) {
    CHROMO_PDL_ENTER();
    __shared__ float red[32];
    float s = 0.f;
    const float inv = 1.f / (float)count;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const float d = y[i] - t[i];
        s = fmaf(d, d, s);
        dy[i] = 2.f * d * inv * grad_scale;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) loss[0] = s * inv;
    }
}

// CrossEntropyLoss, mean reduction over the batch: log-softmax + NLL.
__global__ void ce_loss_kernel(const float* __restrict__ y, const int64_t* __restrict__ lab, int batch,
                               int C, float grad_scale, float* loss, float* dy) {
    CHROMO_PDL_ENTER();
    __shared__ float red[32];
    float s = 0.f;
    const float inv = 1.f / (float)batch;
    for (int b = threadIdx.x; b < batch; b += blockDim.x) {
        const float* row = y + (long long)b * C;
        float mx = row[0];
        for (int c = 1; c < C; ++c) mx = fmaxf(mx, row[c]);
        float z = 0.f;
        for (int c = 0; c < C; ++c) z += expf(row[c] - mx);
        const float lse = mx + logf(z);
        const int t = (int)lab[b];
        s += lse - row[t];
        for (int c = 0; c < C; ++c) {
            const float p = expf(row[c] - lse);
            dy[(long long)b * C + c] = (p - (c == t ? 1.f : 0.f)) * inv * grad_scale;
        }
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) loss[0] = s * inv;
    }
}

// torch.optim.AdamW (decoupled weight decay, no amsgrad), one launch, float4:
//   p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
//   p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
// 28 B of HBM traffic per parameter (read p,g,m,v; write p,m,v).
__global__ void __launch_bounds__(256) adamw_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                    float4* __restrict__ m, float4* __restrict__ v,
                                                    long long n4, float lr, float b1, float b2, float eps,
                                                    float decay, float step_size, float bc2_sqrt,
                                                    float gs) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        float4 P = p[i], G = g[i], M = m[i], V = v[i];
        float* pp = reinterpret_cast<float*>(&P);
        float* gg = reinterpret_cast<float*>(&G);
        float* mm = reinterpret_cast<float*>(&M);
        float* vv = reinterpret_cast<float*>(&V);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = gg[k] * gs;
            float x = pp[k] * decay;
            mm[k] = mm[k] + (1.f - b1) * (gr - mm[k]);          // exp_avg.lerp_(grad, 1 - beta1)
            vv[k] = b2 * vv[k] + (1.f - b2) * gr * gr;
            const float denom = sqrtf(vv[k]) / bc2_sqrt + eps;
            pp[k] = x - step_size * (mm[k] / denom);
        }
        p[i] = P; m[i] = M; v[i] = V;
    }
}

}  // namespace chromo

using namespace chromo;

extern "C" {

int chromo_mse_loss(const float* logits, const float* target, int32_t count, float grad_scale, float* loss,
                    float* dlogits, void* stream) {
    if (!logits || !target || !loss || !dlogits || count < 1) { set_error("mse_loss: bad argument"); return CHROMO_EINVAL; }
    launch_pdl(mse_loss_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, logits, target, count, grad_scale, loss, dlogits);
    CHROMO_CHECK_LAUNCH("mse_loss");
    return CHROMO_OK;
}

int chromo_ce_loss(const float* logits, const int64_t* labels, int32_t batch, int32_t n_classes,
                   float grad_scale, float* loss, float* dlogits, void* stream) {
    if (!logits || !labels || !loss || !dlogits || batch < 1 || n_classes < 2) { set_error("ce_loss: bad argument"); return CHROMO_EINVAL; }
    launch_pdl(ce_loss_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, logits, labels, batch, n_classes, grad_scale, loss, dlogits);
    CHROMO_CHECK_LAUNCH("ce_loss");
    return CHROMO_OK;
}

int chromo_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t count,
                 float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                 float grad_scale, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || count < 0 || step < 1) { set_error("adamw: bad argument"); return CHROMO_EINVAL; }
    if (count % 4 != 0 || (reinterpret_cast<uintptr_t>(params) & 15) || (reinterpret_cast<uintptr_t>(grads) & 15) ||
        (reinterpret_cast<uintptr_t>(exp_avg) & 15) || (reinterpret_cast<uintptr_t>(exp_avg_sq) & 15)) {
        set_error("adamw: buffers must be 16-byte aligned and count a multiple of 4");
        return CHROMO_EINVAL;
    }
    if (count == 0) return CHROMO_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float decay = 1.f - lr * weight_decay;
    const long long n4 = count / 4;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    adamw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4*>(params), reinterpret_cast<const float4*>(grads),
        reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq), n4, lr, beta1, beta2, eps,
        decay, step_size, bc2_sqrt, grad_scale);
    CHROMO_CHECK_LAUNCH("adamw");
    return CHROMO_OK;
}

}  // extern "C"
