// Fused single-query attention core (BF16 inference path): scores against the position table, the rank-F feature
// term, masked softmax and the probability-weighted sums in ONE kernel per stage, all resolutions in one launch.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace chromo {

struct SqaFusedArgs {
    int n_res, regions;                       // regions (gene or gene x pCRE slot) per resolution; 2 heads each
    int n[CHROMO_MAX_RES];                    // bins per region
    int ns[CHROMO_MAX_RES];                   // bins padded to the packed position table (multiple of 16, <= 400)
    int order[CHROMO_MAX_RES];                // resolutions by descending n (long tiles first)
    const __nv_bfloat16* qk_tiles; long long qk_tz;   // W_k[h]^T q per (region, head): rows 2*region + head, BF16, tiles of
                                              //   128 rows in the canonical K-major operand layout (sqa_pack_qk)
    const float* x[CHROMO_MAX_RES];           // [regions, n, 7]
    const uint8_t* mask[CHROMO_MAX_RES]; long long mask_stride[CHROMO_MAX_RES], mask_row_offset[CHROMO_MAX_RES];
    const float* w_in; long long w_in_z;      // [128, 7]
    const __nv_bfloat16* w_in_pk; long long w_in_pk_z;   // the same as the B operand of u = QK W_in (sqa_pack_w_in, 2048 elements)
    const __nv_bfloat16* pe_pk[CHROMO_MAX_RES];   // position table, BF16, [ns/8][16][8][8] (pack_weights order)
    float scale;
    float* cbar; long long cbar_z;            // [regions*2, 128]  sum_j p_j (W_in x_j + PE_j)
    __nv_bfloat16* cbar_bf16 = nullptr;       // set: the result is written here in BF16 instead (dense rows; resolution stride 2 * cbar_z elements)
    // ragged plan (ragged.cu), all optional: tile row 2m + head belongs to region perm[m]; only the first live[0] rows /
    // live[1] tiles exist; tile t of resolution r walks the keys [tile_k0[r][t], + tile_ns[r][t]) (a multiple of 8 / of 16):
    // every key outside that window is masked for all 64 regions of the tile
    const int* perm = nullptr;
    const int* live = nullptr;
    const int* tile_k0[CHROMO_MAX_RES] = {};
    const int* tile_ns[CHROMO_MAX_RES] = {};
};

int sqa_pack_qk(const float* qk, long long qk_z, __nv_bfloat16* tiles, long long tz, int rows, int n_res, cudaStream_t st);
int sqa_pack_w_in(const float* w_in, long long w_z, __nv_bfloat16* out, long long out_z, int n_res, cudaStream_t st);
bool sqa_fused_supported(const SqaFusedArgs& a, int H, int F, int D);
int launch_sqa_fused(const SqaFusedArgs& a, cudaStream_t st);

}  // namespace chromo
