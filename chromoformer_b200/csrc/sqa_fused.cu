// Single-query attention core of the Embedding / Pairwise Interaction transformers, fused (BF16 inference).
//
// Only the centre query row of these transformers is ever consumed, and for that row
//   score_j = qk . (W_in x_j + PE_j),    cbar = sum_j p_j (W_in x_j + PE_j) = W_in (sum_j p_j x_j) + sum_j p_j PE_j
// (modules.py:16-30 with net.py:31-59 / 105-139 substituted).  The PE terms are two GEMMs against the SAME table:
//   S = QK PE^T  [128 rows x n]   and   O = P PE  [128 rows x 128]
// and this kernel keeps S, P and O in tensor memory, so that neither the scores nor the probabilities (2 x 105 MB per
// stage at n = 400) ever reach HBM.  One persistent CTA per SM walks 128-row tiles (64 regions x 2 heads):
//   driver warp    QK tile (BF16, already in the canonical K-major operand layout: written that way by the query
//                  GEMM's epilogue, or by sqa_pack_qk) -> shared memory with ONE TMA bulk copy, a tile ahead
//   driver warp    MMA 1: S = QK PE^T (B = table, K-major) -> TMEM [0, ns);  u = QK W_in (N = 16) -> TMEM [400, 416)
//   compute warps  thread = (row, key half), ONE pass over the keys with an online softmax:
//                    s_j = (S_j + u.x_j) scale (base 2), mask;  p_j = 2^(s_j - m);  sum, sum p_j x_j;  P (BF16) -> TMEM
//                  m is the running reference maximum of the thread; it is only moved when a chunk exceeds it by
//                  more than 2^16 (then the P chunks written so far are rescaled in tensor memory - rare), so the
//                  features stream from HBM exactly once through a 3-stage cp.async ring (8 keys per stage and half)
//   driver warp    MMA 2: O_half = P_half PE (A = P in TMEM, B = the same table bytes read MN-major), one FP32
//                  accumulator per key half because the halves carry different reference maxima
//   compute warps  cbar = (O_A 2^(mA-m) + O_B 2^(mB-m)) / sum + W_in xbar, transposed through shared memory,
//                  512-byte row stores
// Tensor-memory map (512 columns): scores [0, ns); P of the first key half in place [0, 4*CA), P of the second half
// [416, 416 + 4*CB); u [400, 407); O_B [256, 384) and O_A [128, 256) once the scores under them are dead.
#include <stdlib.h>

#include "sqa_fused.cuh"
#include "umma_ptx.cuh"

namespace chromo {
namespace {

constexpr int SQ_THREADS = 288;                 // 8 compute warps + 1 driver warp
constexpr int SQ_NSTAGE = 3;
constexpr int SQ_F = 7;
constexpr int SQ_KC = 8;                        // keys per ring chunk
constexpr int SQ_XROW = SQ_KC * SQ_F;           // floats per region and chunk (224 B)
constexpr uint32_t SQ_HALF_X = 64 * SQ_XROW * 4 + 128;      // rows are shifted by 0/16/32 B (x_row_off) against bank conflicts
constexpr uint32_t SQ_STAGE_X = 2 * SQ_HALF_X;
constexpr uint32_t SQ_STAGE = SQ_STAGE_X + 2 * 64 * 8;      // + 8 mask bytes per region and half
constexpr int SQ_MAX_NS = 400;
constexpr uint32_t OFF_PE = 0;                               // position table, up to 400 x 128 BF16
constexpr uint32_t OFF_Q = OFF_PE + SQ_MAX_NS * 256;         // [128 x 128] BF16 QK tile (TMA destination)
constexpr uint32_t OFF_X = OFF_Q + 32768;                    // ring; after the pass: store staging + row exchange area
constexpr uint32_t OFF_WB = OFF_X + SQ_NSTAGE * SQ_STAGE;    // W_in as a BF16 operand [16 x 128] (rows 7.. are zero)
constexpr uint32_t OFF_CTL = OFF_WB + 4096;
constexpr uint32_t OFF_META = OFF_CTL + 128;                 // 2 x (64 region rows + key window) of the tile after this one
constexpr uint32_t SQ_META = 66 * 4;
constexpr uint32_t SQ_SMEM = OFF_META + 2 * SQ_META;
constexpr uint32_t SQ_XCH = 4 * 32 * 132 * 4;                // exchange area inside the ring, behind the store staging
static_assert(SQ_SMEM <= 227 * 1024, "shared memory budget");
constexpr uint32_t QX_MAX = 0, QX_SUM = 1024, QX_XB = 2048, QX_XBAR = 2048 + 8192;
static_assert(SQ_XCH + QX_XBAR + 4096 <= SQ_NSTAGE * SQ_STAGE, "store staging + exchange area fit in the ring");
// O_B is issued while the first key half may still be reading its scores, so it must land on columns of the SECOND
// half's (finished) scores: [256, 384) holds keys 256..383 at ns = 400.  O_A follows once everybody is through.
constexpr int P_B_COL = 416, OA_COL = 128, OB_COL = 256, U_COL = 400;
constexpr float SQ_TAU = 16.f;                  // the reference maximum moves when a chunk exceeds it by 2^16
constexpr float SQ_MINIT = -3.0e38f;

// byte offset of region row r inside a ring half: 224-byte rows, shifted so that the 16 rows a warp reads with one
// LDS.128 fall on every 4-bank group exactly twice (stride 224 B alone would put them on 4 groups, 4 rows each)
__device__ __forceinline__ uint32_t x_row_off(int r) {
    return (uint32_t)r * (SQ_XROW * 4) + 16u * (uint32_t)(((r >> 2) & 1) + ((r >> 3) & 1)) + 32u * (uint32_t)(r >> 4);
}

enum { B_PE = 0, B_QFULL, B_S, B_PA, B_PB, B_O, B_EPI, B_COUNT };

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void compute_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// the same load in two halves: issue early, complete (with the registers tied to the wait) just before the first use
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8_wait(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 :
                 : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// A operand in TMEM (lane = row, two BF16 per 32-bit column), B in shared memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;       // src-size 0: nothing is read, the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// key-chunk split of a table of ns keys (ns % 16 == 0): CA chunks for the first half (even, >= half), rest second
__device__ __forceinline__ int first_half_chunks(int ns) { return ((ns >> 4) + 1) & ~1; }

__global__ void __launch_bounds__(SQ_THREADS, 1) sqa_fused_kernel(const __grid_constant__ SqaFusedArgs a, int tiles_per_res, float tau) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_CTL);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 8) tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        mbar_init(&bars[B_PE], 1);
        mbar_init(&bars[B_QFULL], 1);
        mbar_init(&bars[B_S], 1);
        mbar_init(&bars[B_PA], 4);
        mbar_init(&bars[B_PB], 4);
        mbar_init(&bars[B_O], 1);
        mbar_init(&bars[B_EPI], 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (a.live) tiles_per_res = a.live[1];                    // ragged plan: the live tiles only
    const int total = tiles_per_res * a.n_res;
    const int rows_total = (a.live ? a.live[0] : a.regions) * 2;

    if (warp == 8) {
        // ===================================================== driver: table loads + MMA issue ====
        if (lane == 0) {
            const uint32_t s_pe = smem_u32(smem + OFF_PE), s_q = smem_u32(smem + OFF_Q);
            const uint32_t idesc_o = umma_idesc_bf16(128, 128) | (1u << 16);       // B MN-major
            int cur = -1, it = 0;
            uint32_t pe_par = 0;
            for (int g = blockIdx.x; g < total; g += gridDim.x, ++it) {
                const int res = a.order[g / tiles_per_res];
                // key window of the tile: keys [k0, k0 + ns) of the table (the whole table without a ragged plan)
                const int k0 = a.tile_k0[res] ? a.tile_k0[res][g % tiles_per_res] : 0;
                const int ns = a.tile_ns[res] ? a.tile_ns[res][g % tiles_per_res] : a.ns[res];
                const int CA = first_half_chunks(ns), CB = (ns >> 3) - CA;
                if (res != cur) {
                    if (it > 0) mbar_wait(&bars[B_O], (it - 1) & 1);               // MMA 2 of the last tile is done with the table
                    mbar_expect_tx(&bars[B_PE], (uint32_t)a.ns[res] * 256u + 4096u);
                    tma_bulk_g2s(smem + OFF_PE, a.pe_pk[res], (uint32_t)a.ns[res] * 256u, &bars[B_PE]);
                    tma_bulk_g2s(smem + OFF_WB, a.w_in_pk + res * a.w_in_pk_z, 4096u, &bars[B_PE]);
                    mbar_wait(&bars[B_PE], pe_par);
                    pe_par ^= 1;
                    cur = res;
                }
                if (it == 0) {                                                     // (later tiles are fetched a tile ahead)
                    mbar_expect_tx(&bars[B_QFULL], 32768u);
                    tma_bulk_g2s(smem + OFF_Q, a.qk_tiles + res * a.qk_tz + (long long)(g % tiles_per_res) * 16384, 32768u,
                                 &bars[B_QFULL]);
                }
                mbar_wait(&bars[B_QFULL], it & 1);
                if (it > 0) mbar_wait(&bars[B_EPI], (it - 1) & 1);                 // O of the last tile has been read
                tc_fence_after();
                // descriptors once, then adds in the address field (+256 B = +16): see reg_fused.cu / tools/bench_micro
                const uint64_t qd = umma_smem_desc(s_q, 128, 2048);
                for (int n0 = 0; n0 < ns; n0 += 256) {                             // S = QK PE^T, N in parts of <= 256
                    const int np = min(256, ns - n0);
                    const uint32_t idesc = umma_idesc_bf16(128, np);
                    const uint64_t pd = umma_smem_desc(s_pe + ((k0 + n0) >> 3) * 2048, 128, 2048);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        umma_bf16(tmem + n0, qd + 16 * k, pd + 16 * k, idesc, k > 0 ? 1u : 0u);
                }
                const uint64_t wd = umma_smem_desc(smem_u32(smem + OFF_WB), 128, 2048);
#pragma unroll
                for (int k = 0; k < 8; ++k)                                         // u = QK W_in (columns 7.. are zero)
                    umma_bf16(tmem + U_COL, qd + 16 * k, wd + 16 * k, umma_idesc_bf16(128, 16), k > 0 ? 1u : 0u);
                umma_commit(&bars[B_S]);
                {   // the QK tile of this CTA's next tile, as soon as MMA 1 no longer reads the buffer
                    const int g2 = g + gridDim.x;
                    if (g2 < total) {
                        mbar_wait(&bars[B_S], it & 1);
                        const int res2 = a.order[g2 / tiles_per_res];
                        mbar_expect_tx(&bars[B_QFULL], 32768u);
                        tma_bulk_g2s(smem + OFF_Q, a.qk_tiles + res2 * a.qk_tz + (long long)(g2 % tiles_per_res) * 16384,
                                     32768u, &bars[B_QFULL]);
                    }
                }
                // O = P PE per key half, as soon as that half's warps are through (the second half has fewer chunks and
                // finishes first): key blocks of the table are 2048 B apart (LBO), channel blocks 128 B (SBO)
                if (CB > 0) {
                    mbar_wait(&bars[B_PB], it & 1);
                    tc_fence_after();
                    const uint64_t ob = umma_smem_desc(s_pe + (k0 >> 3) * 2048 + (CA / 2) * 4096, 2048, 128);
                    for (int k = 0; k < CB / 2; ++k)
                        umma_bf16_ts(tmem + OB_COL, tmem + P_B_COL + 8 * k, ob + 256 * k, idesc_o, k > 0 ? 1u : 0u);
                }
                mbar_wait(&bars[B_PA], it & 1);
                tc_fence_after();
                const uint64_t oa = umma_smem_desc(s_pe + (k0 >> 3) * 2048, 2048, 128);
                for (int k = 0; k < CA / 2; ++k)
                    umma_bf16_ts(tmem + OA_COL, tmem + 8 * k, oa + 256 * k, idesc_o, k > 0 ? 1u : 0u);
                umma_commit(&bars[B_O]);
            }
        }
        __syncwarp();
    } else {
        // ===================================================== compute warps =======================
        const int half = warp >> 2, lq = warp & 3;
        const int row = lq * 32 + lane;                      // tile row == TMEM lane
        const int rit = row >> 1;                            // region within the tile
        const int ht = tid & 127;                            // thread within the key half
        const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16);
        const uint32_t s_ring = smem_u32(smem + OFF_X);
        float* red_max = reinterpret_cast<float*>(smem + OFF_X + SQ_XCH + QX_MAX);
        float* red_sum = reinterpret_cast<float*>(smem + OFF_X + SQ_XCH + QX_SUM);
        float* red_xb = reinterpret_cast<float*>(smem + OFF_X + SQ_XCH + QX_XB);
        float* xbar_s = reinterpret_cast<float*>(smem + OFF_X + SQ_XCH + QX_XBAR);
        const float scale2 = a.scale * 1.4426950408889634f;  // softmax in base 2
        const int fr0 = ht >> 4, fpart = ht & 15;             // copy plan of the ring (below)
        const int mreg = ht >> 1, modd = ht & 1;
        // What a tile needs before its first copy can be issued - key window and the 64 region rows of the tile (permuted
        // under a ragged plan) - is loaded a tile ahead into shared memory (two slots): dependent L2 round trips in front
        // of the first HBM access of the tile otherwise.
        auto tile_meta = [&](int g, int slot) {              // (threads 0..65 of the compute warps)
            int* m = reinterpret_cast<int*>(smem + OFF_META + slot * SQ_META);
            const int res = a.order[g / tiles_per_res], tile = g % tiles_per_res;
            if (tid < 64) {
                const int rr = tile * 64 + tid;
                m[tid] = 2 * rr < rows_total ? (a.perm ? a.perm[rr] : rr) : 0;
            } else if (tid == 64) m[64] = a.tile_k0[res] ? a.tile_k0[res][tile] : 0;
            else if (tid == 65) m[65] = a.tile_ns[res] ? a.tile_ns[res][tile] : a.ns[res];
        };
        if ((int)blockIdx.x < total) tile_meta(blockIdx.x, 0);
        compute_barrier();
        int it = 0;
        for (int g = blockIdx.x; g < total; g += gridDim.x, ++it) {
            const int res = a.order[g / tiles_per_res], tile = g % tiles_per_res;
            // key window [k0, k0 + ns) of the tile; n, nF, X and MK are relative to its first key
            const int* meta = reinterpret_cast<const int*>(smem + OFF_META + (it & 1) * SQ_META);
            const int k0 = meta[64], ns = meta[65];
            const int n = a.n[res] - k0;
            const int CA = first_half_chunks(ns), CB = (ns >> 3) - CA;
            const int row0 = tile * 128;
            const int rows_valid = min(128, rows_total - row0);
            const int region0 = row0 >> 1;
            const float* X = a.x[res] + k0 * SQ_F;
            const uint8_t* MK = a.mask[res] + a.mask_row_offset[res] + k0;
            const long long mstride = a.mask_stride[res];
            const float* W = a.w_in + res * a.w_in_z;
            const int c_begin = half ? CA : 0, c_cnt = half ? CB : CA;
            const int nF = n * SQ_F;
            const long long rstride = (long long)a.n[res] * SQ_F;     // floats between regions

            // ring: chunk idx of this half (8 keys x 7 features per region, and the 8 mask bytes).  Copy plan: the 16
            // threads ht/16*16.. move the fourteen 16-byte parts of ONE 224-byte chunk row (contiguous in HBM), regions
            // ht/16 + 8k for k = 0..7; x_row_off(r0 + 8k) = x_row_off(r0) + 1792 k + 16 (k & 1) + 32 (k >> 1).
            const bool fact = fpart < 14;
            uint32_t fsrc[8];                                  // the thread's eight region rows, in float4 units from X (n % 4 == 0)
#pragma unroll
            for (int k = 0; k < 8; ++k) fsrc[k] = (uint32_t)(((long long)meta[fr0 + 8 * k] * rstride) >> 2) + fpart;
            const uint32_t fdst = s_ring + half * SQ_HALF_X + x_row_off(fr0) + fpart * 16;
            const bool mvalid = 2 * mreg < rows_valid;
            const uint8_t* msrc = MK + (long long)meta[mreg] * mstride + modd * 4;
            const uint32_t mdst = s_ring + SQ_STAGE_X + half * 512 + mreg * 8 + modd * 4;
            auto fetch = [&](int idx, int stage) {
                if (idx < c_cnt) {
                    const int cg = c_begin + idx;
                    const uint32_t so = stage * SQ_STAGE;
                    const bool full = cg * 8 + 8 <= n;         // (uniform) every key of the chunk exists
                    const bool pok = fact && (full || cg * SQ_XROW + fpart * 4 + 3 < nF);
                    if (fact) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const bool ok = pok && 2 * (fr0 + 8 * k) < rows_valid;
                            cp_async16(fdst + so + k * 1792 + 16 * (k & 1) + 32 * (k >> 1), ok ? reinterpret_cast<const float4*>(X) + fsrc[k] + cg * (SQ_XROW / 4) : reinterpret_cast<const float4*>(X), ok);
                        }
                    }
                    const bool okm = mvalid && (full || cg * 8 + modd * 4 < n);
                    cp_async4(mdst + so, okm ? msrc + cg * 8 : MK, okm);
                }
                cp_async_commit();
            };
            fetch(0, 0);
            fetch(1, 1);
            // the tile after this one: its slot was last read at the top of the previous tile, a barrier ago
            if (g + (int)gridDim.x < total) tile_meta(g + gridDim.x, (it + 1) & 1);

            mbar_wait(&bars[B_S], it & 1);
            tc_fence_after();
            float u[SQ_F];                                    // u = W_in^T qk of this row (MMA of the driver, columns 400..406)
            {
                float uu[8];
                tmem_ld8(trow + U_COL, uu);
#pragma unroll
                for (int f = 0; f < SQ_F; ++f) u[f] = uu[f];
            }

            // ---- one pass over the keys: scores, online softmax, P (BF16 pairs) -> TMEM ----
            float m = SQ_MINIT, sum = 0.f, xb[SQ_F];
#pragma unroll
            for (int f = 0; f < SQ_F; ++f) xb[f] = 0.f;
            const uint32_t pcol = trow + (half ? P_B_COL : 0);
            const uint32_t xro = x_row_off(rit);
            int rs = 0, fs = 2;                               // ring stage being read / filled
            for (int idx = 0; idx < c_cnt; ++idx) {
                const int cg = c_begin + idx;
                uint32_t sraw[8];
                tmem_ld8_issue(trow + 8 * cg, sraw);           // the scores of the chunk travel while the ring is awaited
                cp_async_wait1();
                named_barrier(2 + half, 128);
                fetch(idx + 2, fs);
                const uint8_t* sb = smem + OFF_X + rs * SQ_STAGE;
                rs = rs == SQ_NSTAGE - 1 ? 0 : rs + 1;
                fs = fs == SQ_NSTAGE - 1 ? 0 : fs + 1;
                const float4* xs = reinterpret_cast<const float4*>(sb + half * SQ_HALF_X + xro);
                const uint2 mk = *reinterpret_cast<const uint2*>(sb + SQ_STAGE_X + half * 512 + rit * 8);
                float xv[SQ_XROW];
#pragma unroll
                for (int q = 0; q < 14; ++q) {
                    const float4 t = xs[q];
                    xv[4 * q] = t.x; xv[4 * q + 1] = t.y; xv[4 * q + 2] = t.z; xv[4 * q + 3] = t.w;
                }
                float s[8];
                tmem_ld8_wait(sraw);
#pragma unroll
                for (int i = 0; i < 8; ++i) s[i] = __uint_as_float(sraw[i]);
                float cmx = -INFINITY;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float t = s[i];
#pragma unroll
                    for (int f = 0; f < SQ_F; ++f) t = fmaf(u[f], xv[i * SQ_F + f], t);
                    t *= scale2;
                    const uint32_t mb = ((i < 4 ? mk.x : mk.y) >> (8 * (i & 3))) & 0xffu;
                    if (mb) t = -1e9f;
                    s[i] = t;
                }
                if (cg * 8 + 8 > n) {                          // (uniform) keys beyond n of a padded table
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (cg * 8 + i >= n) s[i] = -INFINITY;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) cmx = fmaxf(cmx, s[i]);
                const bool ev = cmx > m + tau;              // (always on the first chunk that holds a key)
                if (__any_sync(0xffffffffu, ev)) {
                    // move the reference maximum: rescale the running sums and the P chunks already in TMEM
                    const float fr = ev ? ex2(m - cmx) : 1.f;
                    sum *= fr;
#pragma unroll
                    for (int f = 0; f < SQ_F; ++f) xb[f] *= fr;
                    if (idx > 0) {
                        tmem_st_wait();
                        for (int cc = 0; cc < idx; ++cc) {
                            uint32_t w[4];
                            tmem_ld4(pcol + 4 * cc, w);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                w[k] = pack2(__uint_as_float(w[k] << 16) * fr, __uint_as_float(w[k] & 0xffff0000u) * fr);
                            tmem_st4(pcol + 4 * cc, w);
                        }
                    }
                    if (ev) m = cmx;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float p = ex2(s[i] - m);
                    s[i] = p;
                    sum += p;
#pragma unroll
                    for (int f = 0; f < SQ_F; ++f) xb[f] = fmaf(p, xv[i * SQ_F + f], xb[f]);
                }
                uint32_t pk[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) pk[i] = pack2(s[2 * i], s[2 * i + 1]);
                tmem_st4(pcol + 4 * idx, pk);
            }
            cp_async_wait0();
            tmem_st_wait();
            tc_fence_before();
            warp_arrive(&bars[half ? B_PB : B_PA], lane);
            // W_in rows of the thread's four output channels: used in the epilogue, in flight under the exchange and MMA 2
            float w4[4][SQ_F];
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int f = 0; f < SQ_F; ++f) w4[e][f] = __ldg(W + (4 * lane + e) * SQ_F + f);

            // ---- exchange the halves: common maximum, 1 / sum, xbar per row (in the ring, once every warp has left it) ----
            compute_barrier();
            red_max[half * 128 + row] = m;
            red_sum[half * 128 + row] = sum;
            {
                float4* dst = reinterpret_cast<float4*>(red_xb + (half * 128 + row) * 8);
                dst[0] = make_float4(xb[0], xb[1], xb[2], xb[3]);
                dst[1] = make_float4(xb[4], xb[5], xb[6], 0.f);
            }
            compute_barrier();
            float gA, gB;
            {
                const float mA = red_max[row], mB = red_max[128 + row];
                const float mt = fmaxf(mA, mB);
                const float fA = ex2(mA - mt), fB = ex2(mB - mt);        // (a half without keys keeps SQ_MINIT -> 0)
                const float inv = 1.f / (red_sum[row] * fA + red_sum[128 + row] * fB);
                gA = fA * inv; gB = fB * inv;
            }
            if (half == 0) {
                const float4* p0 = reinterpret_cast<const float4*>(red_xb + row * 8);
                const float4* p1 = reinterpret_cast<const float4*>(red_xb + (128 + row) * 8);
                const float4 a0 = p0[0], a1 = p0[1], b0 = p1[0], b1 = p1[1];
                float4* dst = reinterpret_cast<float4*>(xbar_s + row * 8);
                dst[0] = make_float4(a0.x * gA + b0.x * gB, a0.y * gA + b0.y * gB, a0.z * gA + b0.z * gB, a0.w * gA + b0.w * gB);
                dst[1] = make_float4(a1.x * gA + b1.x * gB, a1.y * gA + b1.y * gB, a1.z * gA + b1.z * gB, 0.f);
            }

            // ---- epilogue: cbar = (O_A gA + O_B gB) + W_in xbar ----
            mbar_wait(&bars[B_O], it & 1);
            tc_fence_after();
            float* stg = reinterpret_cast<float*>(smem + OFF_X) + lq * (32 * 132);
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int c0 = 64 * half + 32 * cc;
                float v[32], vb[32];
                tmem_ld32(trow + OA_COL + c0, v);
                if (CB > 0) {
                    tmem_ld32(trow + OB_COL + c0, vb);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = v[j] * gA + vb[j] * gB;
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= gA;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(stg + lane * 132 + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            tc_fence_before();
            warp_arrive(&bars[B_EPI], lane);
            compute_barrier();                                // staging + xbar_s complete
            {
                float* C = a.cbar + res * a.cbar_z + (long long)row0 * 128;
                __nv_bfloat16* Cb = a.cbar_bf16 ? a.cbar_bf16 + 2 * res * a.cbar_z + (long long)row0 * 128 : nullptr;   // (same byte stride as the FP32 view)
#pragma unroll 4
                for (int i = 0; i < 16; ++i) {
                    const int r = 32 * lq + 16 * half + i;
                    if (r < rows_valid) {
                        float4 o = *reinterpret_cast<const float4*>(stg + (16 * half + i) * 132 + 4 * lane);
                        const float4 x0 = *reinterpret_cast<const float4*>(xbar_s + r * 8);
                        const float4 x1 = *reinterpret_cast<const float4*>(xbar_s + r * 8 + 4);
                        const float xr[SQ_F] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z};
#pragma unroll
                        for (int f = 0; f < SQ_F; ++f) {
                            o.x = fmaf(w4[0][f], xr[f], o.x);
                            o.y = fmaf(w4[1][f], xr[f], o.y);
                            o.z = fmaf(w4[2][f], xr[f], o.z);
                            o.w = fmaf(w4[3][f], xr[f], o.w);
                        }
                        if (Cb) *reinterpret_cast<uint2*>(Cb + (long long)r * 128 + 4 * lane) = make_uint2(pack2(o.x, o.y), pack2(o.z, o.w));
                        else *reinterpret_cast<float4*>(C + (long long)r * 128 + 4 * lane) = o;
                    }
                }
            }
            compute_barrier();                                // exchange area and ring are reused by the next tile
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 512);
}


// ---- operand preparation -------------------------------------------------------------------------------------
// QK rows [rows, 128] FP32 -> BF16 tiles of 128 rows in the canonical K-major layout ([16 row groups][16 k chunks][8][8]);
// rows past the end of the last tile are zero.  (The query GEMM of the forward writes this layout itself.)
__global__ void sqa_pack_qk_kernel(const float* __restrict__ qk, long long qk_z, __nv_bfloat16* __restrict__ tiles,
                                   long long tz, int rows, int tiles_per_res) {
    const int z = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one 8-column chunk of one row
    if (i >= (long long)tiles_per_res * 128 * 16) return;
    const int row = (int)(i >> 4), kc = (int)(i & 15);
    uint4 pk = make_uint4(0u, 0u, 0u, 0u);
    if (row < rows) {
        const float4* src = reinterpret_cast<const float4*>(qk + z * qk_z + (long long)row * 128 + kc * 8);
        const float4 lo = src[0], hi = src[1];
        pk.x = pack2(lo.x, lo.y); pk.y = pack2(lo.z, lo.w); pk.z = pack2(hi.x, hi.y); pk.w = pack2(hi.z, hi.w);
    }
    const int r = row & 127;
    *reinterpret_cast<uint4*>(tiles + z * tz + (long long)(row >> 7) * 16384 + (r >> 3) * 1024 + kc * 64 + (r & 7) * 8) = pk;
}
// W_in [128, 7] FP32 -> B operand of u = QK W_in: [16 rows (feature, 7.. zero)][128 (channel)] BF16, canonical K-major
__global__ void sqa_pack_w_in_kernel(const float* __restrict__ w_in, long long w_z, __nv_bfloat16* __restrict__ out,
                                     long long out_z) {
    const int z = blockIdx.x, t = threadIdx.x;                                  // 256 threads: (feature row n, k chunk)
    const int n = t >> 4, kc = t & 15;
    __nv_bfloat16 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16_rn(n < SQ_F ? w_in[z * w_z + (kc * 8 + j) * SQ_F + n] : 0.f);
    *reinterpret_cast<uint4*>(out + z * out_z + ((n >> 3) * 16 + kc) * 64 + (n & 7) * 8) = *reinterpret_cast<const uint4*>(v);
}

}  // namespace

int sqa_pack_qk(const float* qk, long long qk_z, __nv_bfloat16* tiles, long long tz, int rows, int n_res, cudaStream_t st) {
    const int tpr = (rows + 127) / 128;
    const long long units = (long long)tpr * 128 * 16;
    sqa_pack_qk_kernel<<<dim3((unsigned)((units + 255) / 256), n_res), 256, 0, st>>>(qk, qk_z, tiles, tz, rows, tpr);
    CHROMO_CHECK_LAUNCH("sqa_pack_qk");
    return CHROMO_OK;
}
int sqa_pack_w_in(const float* w_in, long long w_z, __nv_bfloat16* out, long long out_z, int n_res, cudaStream_t st) {
    sqa_pack_w_in_kernel<<<n_res, 256, 0, st>>>(w_in, w_z, out, out_z);
    CHROMO_CHECK_LAUNCH("sqa_pack_w_in");
    return CHROMO_OK;
}

bool sqa_fused_supported(const SqaFusedArgs& a, int H, int F, int D) {
    if (H != 2 || F != SQ_F || D != 128 || a.n_res < 1 || a.regions < 1) return false;
    if (getenv("CHROMO_NO_SQA_FUSED")) return false;
    if (!a.qk_tiles || !a.w_in_pk || (reinterpret_cast<uintptr_t>(a.qk_tiles) & 15) || (a.qk_tz & 7) ||
        (reinterpret_cast<uintptr_t>(a.w_in_pk) & 15) || (a.w_in_pk_z & 7))
        return false;
    if ((reinterpret_cast<uintptr_t>(a.cbar) & 15) || (a.cbar_z & 3) || (reinterpret_cast<uintptr_t>(a.cbar_bf16) & 15)) return false;
    for (int r = 0; r < a.n_res; ++r) {
        if (a.n[r] < 4 || a.n[r] % 4 != 0 || a.ns[r] % 16 != 0 || a.ns[r] < a.n[r] || a.ns[r] > SQ_MAX_NS) return false;
        if (!a.pe_pk[r] || (reinterpret_cast<uintptr_t>(a.pe_pk[r]) & 15)) return false;
        if (reinterpret_cast<uintptr_t>(a.x[r]) & 15) return false;
        if (!a.w_in) return false;
        if ((reinterpret_cast<uintptr_t>(a.mask[r]) & 3) || (a.mask_stride[r] & 3) || (a.mask_row_offset[r] & 3)) return false;
    }
    return true;
}

int launch_sqa_fused(const SqaFusedArgs& a, cudaStream_t st) {
    static int sms = 0;
    if (!sms) {
        int dev = 0, count = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(sqa_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SQ_SMEM);
        if (e != cudaSuccess || count < 1) { set_error("sqa_fused: %s", cudaGetErrorString(e)); return CHROMO_ECUDA; }
        sms = count;
    }
    const int tiles = (a.regions * 2 + 127) / 128;
    const int total = tiles * a.n_res;
    float tau = SQ_TAU;                 // CHROMO_SQA_TAU=0 (tests): rescale on every new maximum
    if (const char* e = getenv("CHROMO_SQA_TAU")) tau = (float)atof(e);
    sqa_fused_kernel<<<total < sms ? total : sms, SQ_THREADS, SQ_SMEM, st>>>(a, tiles, tau);
    CHROMO_CHECK_LAUNCH("sqa_fused");
    return CHROMO_OK;
}

}  // namespace chromo
