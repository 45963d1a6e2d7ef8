// tcgen05 BF16 GEMM engine (umma_gemm.cu).
#pragma once
#include <cuda_bf16.h>

#include "gemm_simt.cuh"

namespace chromo {

// tile width the engine uses for an N-wide weight (0 = unsupported)
int umma_tile_n(int N);
bool umma_supported(const GemmArgs& g);
bool umma_qk_persistent(const GemmArgs& g);   // the call goes to the persistent query GEMM (which takes a ragged plan)
// Bp: weights packed by pack_weights() with NT = umma_tile_n(g.N); same z strides as g.B
int umma_launch(const GemmArgs& g, const __nv_bfloat16* Bp, int nz, cudaStream_t st);
// FP32 [N,K] (or its transpose when `transposed`: element (n,k) at src[k*ld_src + n]) -> packed BF16
int pack_weights(const float* src, __nv_bfloat16* dst, int N, int K, int NT, long long z_stride, int nz,
                 bool transposed, int ld_src, cudaStream_t st, int valid = -1);

}  // namespace chromo
