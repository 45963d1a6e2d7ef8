// Fused Regulation-transformer layer (reg_fused.cu).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace chromo {

struct RegFusedArgs {
    int B, S, G, n_tiles;                   // genes, tokens per gene, genes per tile (G*S <= 128), tiles
    const float* x; long long x_z;          // layer input  [B*S, 128] FP32 (z = resolution stride)
    float* y; long long y_z;                // layer output [B*S, 128] FP32
    const __nv_bfloat16* wstream; long long w_z;   // this layer's 14 packed weight chunks
    const float* gamma_f; const float* bo; const float* ln1w; const float* ln1b;
    const float* b1; const float* b2; const float* ln2w; const float* ln2b; long long p_z;
    const float* freq;                      // [B,S,S]
    const uint8_t* imask[CHROMO_MAX_RES];   // [B,S,S] per resolution
};

struct RegStreamArgs {
    const float* params; long long p_z;     // flat FP32 parameters, per-resolution stride
    long long att[CHROMO_MAX_LAYERS], ffw[CHROMO_MAX_LAYERS], l1w[CHROMO_MAX_LAYERS], l2w[CHROMO_MAX_LAYERS];
    int n_layers;
    __nv_bfloat16* stream;                  // [res][layer][14][128*128]
};

long long reg_stream_elems_per_layer();
int pack_reg_stream(const RegStreamArgs& a, int n_res, cudaStream_t st);
int launch_reg_layer_fused(const RegFusedArgs& a, int n_res, cudaStream_t st);

}  // namespace chromo
