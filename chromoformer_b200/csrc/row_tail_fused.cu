// Tail of an Embedding / Pairwise-Interaction layer as ONE kernel (inference, folded weights):
//
//   U = LN(res + Cbar N^T + b_o)          N = [W_o[:,h] W_v[h]]_h  ([128, 256], modules.py:29-30 / 150-152)
//   Y = LN(U + W_2 relu(W_1 U + b_1) + b_2)                        (modules.py:100-101)
//
// Same machinery as reg_fused.cu (driver warp: TMA weight stage + tcgen05.mma; 8 compute warps:
// TMEM epilogues), without the attention: a 128-row tile of Cbar goes in, the layer output comes
// out; U and the FFN hidden activations never leave the SM.
//
// The phases of a tile are strictly serial (MMA -> epilogue -> MMA ...), so the tile is kept SMALL instead of fast: one
// 64 KB operand buffer that holds Cbar, then U (over its first half), then the hidden activations; one 32 KB weight stage;
// 256 TMEM columns (the out-projection / FFN-2 accumulator under the FFN-1 one).  Two CTAs then share an SM (104 KB of
// shared memory, 256 columns each) and one tile's epilogues run under the other's MMAs and weight copies.
#include "common.cuh"
#include "kernels.cuh"
#include "reg_fused.cuh"
#include "umma_ptx.cuh"

namespace chromo {

namespace {

constexpr int RT_THREADS = 288;
constexpr int RT_NSTAGE = 1;
constexpr int RT_CHUNK_ELEMS = 128 * 128;
constexpr uint32_t RT_CHUNK_BYTES = RT_CHUNK_ELEMS * 2;

constexpr uint32_t RT_OFF_A = 0;                              // [128 x 256] BF16 operand (Cbar, then F, then store staging)
constexpr uint32_t RT_OFF_U = RT_OFF_A;                       // [128 x 128] BF16 operand (U) over the dead Cbar tile
constexpr uint32_t RT_OFF_STAGE = RT_OFF_A + 65536;           // 32 KB weight stage
constexpr uint32_t RT_TMEM_COLS = 256;
constexpr uint32_t RT_OFF_PRM = RT_OFF_STAGE + RT_NSTAGE * RT_CHUNK_BYTES;   // 1024 floats of parameters
constexpr uint32_t RT_OFF_RED = RT_OFF_PRM + 4096;            // 2 x 512 floats of LayerNorm partials
constexpr uint32_t RT_OFF_CTL = RT_OFF_RED + 4096;
constexpr uint32_t RT_OFF_YROW = RT_OFF_CTL + 256;            // output row of every tile row (128 ints)
constexpr uint32_t RT_SMEM = RT_OFF_YROW + 512;

enum { T_FULL0 = 0, T_FREE0 = 3, T_AREADY = 6, T_UREADY, T_FREADY, T_ACCO, T_ACCF1, T_ACCF2, T_COUNT };

__device__ __forceinline__ void t_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void t_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ uint32_t t_pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t t_chunk(int row, int kc, int K) {
    return (uint32_t)(row >> 3) * (uint32_t)(K * 16) + (uint32_t)kc * 128u + (uint32_t)(row & 7) * 16u;
}

}  // namespace

__global__ void __maxnreg__(96) row_tail_fused_kernel(const RowTailArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RT_OFF_CTL);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + T_COUNT);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int z = blockIdx.y;
    const int m0 = blockIdx.x * 128;
    const int rows_valid = min(128, (a.m_dev ? *a.m_dev : a.M) - m0);
    if (rows_valid <= 0) return;                              // ragged plan: a tile past the live rows
    const int nff = a.dff / 128;                              // 1 or 2 chunks per FFN matrix

    if (warp == 0) tmem_alloc(tmem_slot, RT_TMEM_COLS);
    if (tid == 0) {
        for (int i = 0; i < RT_NSTAGE; ++i) { mbar_init(&bars[T_FULL0 + i], 1); mbar_init(&bars[T_FREE0 + i], 1); }
        mbar_init(&bars[T_AREADY], 8); mbar_init(&bars[T_UREADY], 8); mbar_init(&bars[T_FREADY], 8);
        mbar_init(&bars[T_ACCO], 1); mbar_init(&bars[T_ACCF1], 1); mbar_init(&bars[T_ACCF2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {
            const __nv_bfloat16* wsrc = a.wstream + z * a.w_z;
            const uint32_t idesc = umma_idesc_bf16(128, 128);
            const uint32_t s_u = smem_u32(smem + RT_OFF_U), s_a = smem_u32(smem + RT_OFF_A);
            auto issue_load = [&](int c) {
                const int st = c % RT_NSTAGE;
                if (c >= RT_NSTAGE) mbar_wait(&bars[T_FREE0 + st], ((c / RT_NSTAGE) - 1) & 1);
                mbar_expect_tx(&bars[T_FULL0 + st], RT_CHUNK_BYTES);
                tma_bulk_g2s(smem + RT_OFF_STAGE + st * RT_CHUNK_BYTES, wsrc + (long long)c * RT_CHUNK_ELEMS, RT_CHUNK_BYTES,
                             &bars[T_FULL0 + st]);
            };
            auto consume = [&](int c, uint32_t a_addr, uint32_t a_sbo, uint32_t col, bool accumulate) {
                const int st = c % RT_NSTAGE;
                mbar_wait(&bars[T_FULL0 + st], (c / RT_NSTAGE) & 1);
                tc_fence_after();
                const uint32_t b_addr = smem_u32(smem + RT_OFF_STAGE + st * RT_CHUNK_BYTES);
                const uint64_t ad = umma_smem_desc(a_addr, 128, a_sbo), bd = umma_smem_desc(b_addr, 128, 2048);   // (see reg_fused.cu)
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_bf16(tmem + col, ad + 16 * k, bd + 16 * k, idesc, (accumulate || k > 0) ? 1u : 0u);
                umma_commit(&bars[T_FREE0 + st]);
            };
            // one weight stage: the copy of chunk c+1 is issued as soon as the MMAs of chunk c have left the stage - after
            // the phase's commit, so that the epilogue warps are released first; the chunk that opens the next phase
            // travels under this phase's epilogue
            int next = 0;
            issue_load(next++);
            mbar_wait(&bars[T_AREADY], 0);
            consume(0, s_a, 4096, 0, false);                    // folded out-projection, K halves
            issue_load(next++);
            consume(1, s_a + 2048, 4096, 0, true);
            umma_commit(&bars[T_ACCO]);
            issue_load(next++);
            mbar_wait(&bars[T_UREADY], 0);
            for (int h = 0; h < nff; ++h) {                     // FFN-1, N halves, over the (read-out) out-projection accumulator
                consume(2 + h, s_u, 2048, 128 * h, false);
                if (h + 1 < nff) issue_load(next++);
            }
            umma_commit(&bars[T_ACCF1]);
            issue_load(next++);
            mbar_wait(&bars[T_FREADY], 0);
            for (int h = 0; h < nff; ++h) {                     // FFN-2, K halves, over the (read-out) FFN-1 accumulator
                consume(2 + nff + h, s_a + 2048 * h, (uint32_t)a.dff * 16, 0, h > 0);
                if (h + 1 < nff) issue_load(next++);
            }
            umma_commit(&bars[T_ACCF2]);
        }
    } else {
        const int lq = warp & 3, ch = warp >> 2;
        const int row = lq * 32 + lane;
        const bool valid = row < rows_valid;
        const int m = m0 + row;
        const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16);
        float v[32];
        float* prm = reinterpret_cast<float*>(smem + RT_OFF_PRM);
        float* red = reinterpret_cast<float*>(smem + RT_OFF_RED);
        int* yrow_s = reinterpret_cast<int*>(smem + RT_OFF_YROW);
        // residual row of this thread's tile row (a gather under a ragged plan): the lookup travels under phase 0
        const int res_m = valid ? (a.res_rows ? a.res_rows[m] : m) / a.res_div : 0;

        // ---- phase 0: parameters -> shared;  Cbar tile (FP32 [128 x 256]) -> BF16 operand
        {
            const float* srcs[6] = {a.bo, a.ln1w, a.ln1b, a.b2, a.ln2w, a.ln2b};
            const int ct = warp * 32 + lane;
#pragma unroll
            for (int k = 0; k < 6; ++k)
                if (ct < 128) prm[k * 128 + ct] = srcs[k][z * a.p_z + ct];
            if (ct < a.dff) prm[768 + ct] = a.b1[z * a.p_z + ct];
            if (ct < 128) {                                     // output rows (remapped / scattered), once per tile row
                const int mm = m0 + ct;
                yrow_s[ct] = ct < rows_valid ? (a.y_rows ? a.y_rows[mm] : (mm / a.c_div) * a.c_mul + (mm % a.c_div) + a.c_add) : 0;
            }
            const float* A = a.a + z * a.a_z;
            if (a.a_bf16) {                                     // BF16 rows: 16-byte chunks go straight into the operand
                const __nv_bfloat16* Ab = a.a_bf16 + 2 * z * a.a_z;   // (same byte stride as the FP32 view)
                uint4 xb[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int u = warp + (q >> 2) * 32 + (q & 3) * 8;
                    const int r = (u & 15) * 8 + (lane >> 2), kc = (u >> 4) * 4 + (lane & 3);
                    xb[q] = make_uint4(0u, 0u, 0u, 0u);
                    if (r < rows_valid) xb[q] = __ldg(reinterpret_cast<const uint4*>(Ab + (long long)(m0 + r) * a.lda + kc * 8));
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int u = warp + (q >> 2) * 32 + (q & 3) * 8;
                    const int r = (u & 15) * 8 + (lane >> 2), kc = (u >> 4) * 4 + (lane & 3);
                    *reinterpret_cast<uint4*>(smem + RT_OFF_A + t_chunk(r, kc, 256)) = xb[q];
                }
            } else
            for (int u0 = warp; u0 < 128; u0 += 32) {           // 16 row groups x 8 groups of four 8-column chunks
                float4 x[4][2];
                int r_[4], kc_[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int u = u0 + q * 8;
                    r_[q] = (u & 15) * 8 + (lane >> 2);
                    kc_[q] = (u >> 4) * 4 + (lane & 3);
                    x[q][0] = x[q][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r_[q] < rows_valid) {
                        const float* p = A + (long long)(m0 + r_[q]) * a.lda + kc_[q] * 8;
                        x[q][0] = __ldg(reinterpret_cast<const float4*>(p));
                        x[q][1] = __ldg(reinterpret_cast<const float4*>(p + 4));
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 pk;
                    pk.x = t_pack2(x[q][0].x, x[q][0].y); pk.y = t_pack2(x[q][0].z, x[q][0].w);
                    pk.z = t_pack2(x[q][1].x, x[q][1].y); pk.w = t_pack2(x[q][1].z, x[q][1].w);
                    *reinterpret_cast<uint4*>(smem + RT_OFF_A + t_chunk(r_[q], kc_[q], 256)) = pk;
                }
            }
            fence_async_smem();
            t_arrive(&bars[T_AREADY], lane);
        }
        t_barrier();                                            // parameters visible

        // ---- phase 1: out-projection epilogue: + bias + residual, LayerNorm -> U
        float u_keep[2][32];
        {
            // residual rows first: their (row-strided) loads overlap the out-projection MMA
            const float* res = a.res + z * a.res_z + (long long)res_m * 128;
#pragma unroll
            for (int ci = 0; ci < 2; ++ci) {
                const int c = (2 * ci + ch) * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid) r4 = __ldg(reinterpret_cast<const float4*>(res + c + j));
                    u_keep[ci][j] = r4.x; u_keep[ci][j + 1] = r4.y; u_keep[ci][j + 2] = r4.z; u_keep[ci][j + 3] = r4.w;
                }
            }
            mbar_wait(&bars[T_ACCO], 0);
            tc_fence_after();
            float sum = 0.f, sq = 0.f;
#pragma unroll
            for (int ci = 0; ci < 2; ++ci) {
                const int c = (2 * ci + ch) * 32;
                tmem_ld32(trow + c, v);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(prm + c + j);
                    const float t0 = v[j] + b4.x + u_keep[ci][j], t1 = v[j + 1] + b4.y + u_keep[ci][j + 1];
                    const float t2 = v[j + 2] + b4.z + u_keep[ci][j + 2], t3 = v[j + 3] + b4.w + u_keep[ci][j + 3];
                    u_keep[ci][j] = t0; u_keep[ci][j + 1] = t1; u_keep[ci][j + 2] = t2; u_keep[ci][j + 3] = t3;
                    sum += (t0 + t1) + (t2 + t3);
                    sq += (t0 * t0 + t1 * t1) + (t2 * t2 + t3 * t3);
                }
            }
            tc_fence_before();
            red[(ch * 128 + row) * 2] = sum;
            red[(ch * 128 + row) * 2 + 1] = sq;
            t_barrier();
            sum += red[((1 - ch) * 128 + row) * 2];
            sq += red[((1 - ch) * 128 + row) * 2 + 1];
            const float mean = sum * (1.f / 128.f);
            const float rstd = rsqrtf(fmaxf(sq * (1.f / 128.f) - mean * mean, 0.f) + 1e-5f);
            const float* lw = prm + 128;
            const float* lb = prm + 256;
#pragma unroll
            for (int ci = 0; ci < 2; ++ci) {
                const int c = (2 * ci + ch) * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    float r[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        r[e] = (u_keep[ci][j + e] - mean) * rstd * lw[c + j + e] + lb[c + j + e];
                        u_keep[ci][j + e] = r[e];
                    }
                    uint4 pk;
                    pk.x = t_pack2(r[0], r[1]); pk.y = t_pack2(r[2], r[3]); pk.z = t_pack2(r[4], r[5]); pk.w = t_pack2(r[6], r[7]);
                    *reinterpret_cast<uint4*>(smem + RT_OFF_U + t_chunk(row, (c + j) >> 3, 128)) = pk;
                }
            }
            fence_async_smem();
            t_arrive(&bars[T_UREADY], lane);
        }

        // ---- phase 2: FFN-1 epilogue: + bias, ReLU -> BF16 operand (over the dead Cbar tile)
        {
            mbar_wait(&bars[T_ACCF1], 0);
            tc_fence_after();
            const float* b1 = prm + 768;
            for (int ci = 0; ci < 2 * nff; ++ci) {
                const int c = (2 * ci + ch) * 32;
                tmem_ld32(trow + c, v);
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    float r[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) r[e] = fmaxf(v[j + e] + b1[c + j + e], 0.f);
                    uint4 pk;
                    pk.x = t_pack2(r[0], r[1]); pk.y = t_pack2(r[2], r[3]); pk.z = t_pack2(r[4], r[5]); pk.w = t_pack2(r[6], r[7]);
                    *reinterpret_cast<uint4*>(smem + RT_OFF_A + t_chunk(row, (c + j) >> 3, a.dff)) = pk;
                }
            }
            tc_fence_before();
            fence_async_smem();
            t_arrive(&bars[T_FREADY], lane);
        }

        // ---- phase 3: FFN-2 epilogue: + bias + U, LayerNorm -> Y (row-remapped, coalesced through shared)
        {
            mbar_wait(&bars[T_ACCF2], 0);
            tc_fence_after();
            const float* b2 = prm + 384;
            float* red2 = red + 512;
            float sum = 0.f, sq = 0.f;
#pragma unroll
            for (int ci = 0; ci < 2; ++ci) {
                const int c = (2 * ci + ch) * 32;
                tmem_ld32(trow + c, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float t0 = v[j] + b2[c + j] + u_keep[ci][j];
                    u_keep[ci][j] = t0;
                    sum += t0;
                    sq += t0 * t0;
                }
            }
            tc_fence_before();
            red2[(ch * 128 + row) * 2] = sum;
            red2[(ch * 128 + row) * 2 + 1] = sq;
            t_barrier();
            sum += red2[((1 - ch) * 128 + row) * 2];
            sq += red2[((1 - ch) * 128 + row) * 2 + 1];
            const float mean = sum * (1.f / 128.f);
            const float rstd = rsqrtf(fmaxf(sq * (1.f / 128.f) - mean * mean, 0.f) + 1e-5f);
            const float* lw = prm + 512;
            const float* lb = prm + 640;
            float* Y = a.y + z * a.y_z;
            float* stage = reinterpret_cast<float*>(smem + RT_OFF_A) + warp * (32 * 33);
#pragma unroll
            for (int ci = 0; ci < 2; ++ci) {
                const int c = (2 * ci + ch) * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    stage[lane * 33 + j] = (u_keep[ci][j] - mean) * rstd * lw[c + j] + lb[c + j];
                __syncwarp();
                const int cq = (lane & 7) * 4;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int r = it * 4 + (lane >> 3);
                    const float* sp = stage + r * 33 + cq;
                    const int trw = lq * 32 + r;
                    if (trw < rows_valid)
                        *reinterpret_cast<float4*>(Y + (long long)yrow_s[trw] * 128 + c + cq) = make_float4(sp[0], sp[1], sp[2], sp[3]);
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, RT_TMEM_COLS);
}

// Weight stream of one tail: [N K-half 0, N K-half 1, W1 N-halves.., W2 K-halves..], each [128 x 128] BF16 tiles.
__global__ void pack_tail_stream_kernel(TailStreamArgs a) {
    const int z = blockIdx.y;
    const int nff = a.dff / 128, nchunk = 2 + 2 * nff;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nchunk * 128 * 16) return;
    const int c = i / (128 * 16), n = (i / 16) % 128, kc = i % 16;
    const float* src;
    if (c < 2) src = a.nfold + z * a.nfold_z + (long long)n * 256 + c * 128 + kc * 8;
    else if (c < 2 + nff) src = a.params + z * a.p_z + a.l1w + (long long)((c - 2) * 128 + n) * 128 + kc * 8;
    else src = a.params + z * a.p_z + a.l2w + (long long)n * a.dff + (c - 2 - nff) * 128 + kc * 8;
    const float4 x0 = *reinterpret_cast<const float4*>(src), x1 = *reinterpret_cast<const float4*>(src + 4);
    uint4 pk;
    pk.x = t_pack2(x0.x, x0.y); pk.y = t_pack2(x0.z, x0.w); pk.z = t_pack2(x1.x, x1.y); pk.w = t_pack2(x1.z, x1.w);
    __nv_bfloat16* dst = a.stream + z * a.stream_z + (long long)c * RT_CHUNK_ELEMS + ((n >> 3) * 16 + kc) * 64 + (n & 7) * 8;
    *reinterpret_cast<uint4*>(dst) = pk;
}

int pack_tail_stream(const TailStreamArgs& a, int n_res, cudaStream_t st) {
    const int nchunk = 2 + 2 * (a.dff / 128);
    pack_tail_stream_kernel<<<dim3((nchunk * 128 * 16 + 255) / 256, n_res), 256, 0, st>>>(a);
    CHROMO_CHECK_LAUNCH("pack_tail_stream");
    return CHROMO_OK;
}

int launch_row_tail_fused(const RowTailArgs& a, int n_res, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(row_tail_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RT_SMEM);
        if (e != cudaSuccess) { set_error("row_tail smem attribute: %s", cudaGetErrorString(e)); return CHROMO_ECUDA; }
        // two CTAs per SM: 2 x 104 KB of shared memory, 2 x 256 TMEM columns, 18 warps x 96 registers (5 warps per scheduler)
        cudaFuncSetAttribute(row_tail_fused_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured = true;
    }
    if (a.dff != 128 && a.dff != 256) { set_error("row_tail_fused: d_ff must be 128 or 256"); return CHROMO_EINVAL; }
    row_tail_fused_kernel<<<dim3((a.M + 127) / 128, n_res), RT_THREADS, RT_SMEM, st>>>(a);
    CHROMO_CHECK_LAUNCH("row_tail_fused");
    return CHROMO_OK;
}

}  // namespace chromo
