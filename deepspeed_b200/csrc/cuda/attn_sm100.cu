// Training attention for sm_100a: causal / full softmax(Q K^T / sqrt(d)) V with grouped-query heads, head dim 128, bf16.
//
// Forward (this file, part 1) -- one CTA per (128 query rows, query head, batch):
//   warp 0      TMA producer: Q once, then a 2-stage ring of K and V tiles (128 keys x 128 dims = 32 KiB each)
//   warp 1      one elected thread issues tcgen05.mma: S_j = Q K_j^T into one of TWO 128-column TMEM accumulators (so the
//               QK^T of block j+1 runs while the softmax warps still work on block j), then O += P_j V_j
//   warps 2-5   128 threads = 128 query rows: read S from TMEM (tcgen05.ld, one row per thread -- no shuffles), online
//               softmax in the log2 domain (ex2.approx), rescale O in TMEM only when a row maximum moved
//               (tcgen05.ld / tcgen05.st), write P as bf16 into a 128-byte-swizzled K-major shared tile that the second
//               MMA consumes; at the end O / l -> bf16 -> global and the log-sum-exp for backward
//   S, P never touch global memory; K/V tiles are shared by the 4 query heads of a GQA group through L2.
// Backward (part 2) recomputes S / P per (query block, key block) pair from Q, K and the saved LSE with the same
// tile machinery and produces dQ, dK, dV.
//
// Reference role: the reference has no training attention kernel of its own for HF-style models (it calls flash-attn,
// sequence/fpdt_layer.py:235); its BERT-era fused layer is csrc/transformer/softmax_kernels.cu + cuBLAS strided batched GEMMs.
#include <cuda.h>
#include "dsb_tc.cuh"

namespace dsb {
namespace attn {
using namespace dsb::tc;

constexpr int BQ = 128, BKV = 128, D = 128;
// warp 0 TMA, warp 1 MMA, warps 2-9: TWO consumer warps per TMEM lane quarter (each owns half of the columns of "its" 32
// rows).  TMEM reads run at ~64 B/clk/SM: with one warp per scheduler the tcgen05.ld wait is dead time; with two, one warp's
// exp/convert work hides the other's accumulator reads.
constexpr int kThreads = 320;
constexpr int kConsumers = 8;
constexpr uint32_t TILE = BQ * D * 2;                   // 32 KiB: a [128 x 128] bf16 tile = 2 sub-tiles of [128 x 64]
constexpr uint32_t HALF = TILE / 2;                     // 16 KiB sub-tile (one 128-byte swizzle span of 64 bf16 per row)
constexpr uint32_t TM_COLS = 512;
namespace kf {  // forward shared-memory / TMEM plan
constexpr uint32_t SM_Q = 0;
constexpr uint32_t SM_K = SM_Q + TILE;                  // 2 stages
constexpr uint32_t SM_V = SM_K + 2 * TILE;              // 2 stages
constexpr uint32_t SM_P = SM_V + 2 * TILE;
constexpr uint32_t SM_X = SM_P + TILE;                  // row-statistics exchange between the two warps of a row: [3][2][128] f32
constexpr uint32_t SM_BAR = SM_X + 3 * 2 * BQ * 4;       // (two rotating slots for the block maxima + one for the final sums)
constexpr uint32_t SM_TOTAL = SM_BAR + 256 + 1024;
constexpr uint32_t TM_S = 0, TM_O = 256;                // S0 [0,128) S1 [128,256) O [256,384)
}  // namespace kf
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct FwdParams {
    __nv_bfloat16* o;   // [B*S, ld_o], head h at columns [h*D, (h+1)*D)
    float* lse;         // [B, Hq, S] natural-log sum-exp of the scaled scores
    int ld_o;
    int S, Hq, Hkv;
    int q_col0, k_col0, v_col0;  // first column of head 0 of Q / K / V inside their 2-D tensors (packed QKV: offsets)
    float scale_log2;   // softmax scale * log2(e)
    int causal;
};

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                const __grid_constant__ CUtensorMap map_v, const FwdParams p)
{
    using namespace kf;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + SM_BAR;
    // barrier map (8 bytes each)
    const uint32_t q_full = bar;
    auto k_full = [&](int s) { return bar + 8 * (1 + s); };
    auto v_full = [&](int s) { return bar + 8 * (3 + s); };
    auto k_empty = [&](int s) { return bar + 8 * (5 + s); };
    auto v_empty = [&](int s) { return bar + 8 * (7 + s); };
    auto s_full = [&](int b) { return bar + 8 * (9 + b); };
    auto s_empty = [&](int b) { return bar + 8 * (11 + b); };
    const uint32_t p_full = bar + 8 * 13, p_empty = bar + 8 * 14;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8 * 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // heavy (late) query blocks first: with causal masking block i does i + 1 key blocks of work
    const int qb = static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x);
    const int h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (p.Hq / p.Hkv);
    const int row0 = b * p.S + qb * BQ;  // first token row of this query block
    const int n_kv = p.causal ? qb + 1 : (p.S + BKV - 1) / BKV;  // S need not be a multiple of 128 (serving prompts)

    if (warp == 0 && lane == 0) {
        prefetch_map(&map_q);
        prefetch_map(&map_k);
        prefetch_map(&map_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(k_full(s), 1);
            mbar_init(v_full(s), 1);
            mbar_init(k_empty(s), 1);
            mbar_init(v_empty(s), 1);
            mbar_init(s_full(s), 1);
            mbar_init(s_empty(s), kConsumers);  // one arrival per softmax warp
        }
        mbar_init(p_full, kConsumers);
        mbar_init(p_empty, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (elect_one()) {
            mbar_expect_tx(q_full, TILE);
            tma_load_2d(sbase + SM_Q, &map_q, q_full, p.q_col0 + h * D, row0);
            tma_load_2d(sbase + SM_Q + HALF, &map_q, q_full, p.q_col0 + h * D + 64, row0);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j & 1;
                const uint32_t par = ((j >> 1) & 1) ^ 1;
                const int kv_row = b * p.S + j * BKV;
                mbar_wait(k_empty(st), par);
                mbar_expect_tx(k_full(st), TILE);
                tma_load_2d(sbase + SM_K + st * TILE, &map_k, k_full(st), p.k_col0 + hk * D, kv_row);
                tma_load_2d(sbase + SM_K + st * TILE + HALF, &map_k, k_full(st), p.k_col0 + hk * D + 64, kv_row);
                mbar_wait(v_empty(st), par);
                mbar_expect_tx(v_full(st), TILE);
                // V is the MN-major B operand of O += P V: four boxes of [64 keys x 64 dims]; key half kh at + kh * 16 KiB,
                // dim half at + 8 KiB
#pragma unroll
                for (int kh = 0; kh < 2; ++kh)
#pragma unroll
                    for (int nh = 0; nh < 2; ++nh)
                        tma_load_2d(sbase + SM_V + st * TILE + kh * HALF + nh * (HALF / 2), &map_v, v_full(st),
                                    p.v_col0 + hk * D + nh * 64, kv_row + kh * 64);
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (elect_one()) {
            constexpr uint32_t idesc_s = idesc_bf16(BQ, BKV, false, false);  // S = Q K^T, both K-major
            constexpr uint32_t idesc_o = idesc_bf16(BQ, D, false, true);     // O += P V, V MN-major
            auto issue_s = [&](int j) {
                const int st = j & 1, bs = j & 1;
                mbar_wait(k_full(st), (j >> 1) & 1);
                mbar_wait(s_empty(bs), ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t qa = sbase + SM_Q, kb = sbase + SM_K + st * TILE;
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint32_t off = (k >> 2) * HALF + (k & 3) * 32;  // 16 bf16 of K = 32 B inside the swizzle row
                    umma_bf16(tmem + TM_S + bs * BKV, desc_kmajor_sw128(qa + off), desc_kmajor_sw128(kb + off), idesc_s,
                              k > 0 ? 1u : 0u);
                }
                umma_commit(s_full(bs));
                umma_commit(k_empty(st));
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_kv; ++j) {
                if (j + 1 < n_kv) issue_s(j + 1);
                const int st = j & 1;
                mbar_wait(p_full, j & 1);
                mbar_wait(v_full(st), (j >> 1) & 1);
                tc_fence_after();
                const uint32_t pa = sbase + SM_P, vb = sbase + SM_V + st * TILE;
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k) {
                    const uint64_t da = desc_kmajor_sw128(pa + (k >> 2) * HALF + (k & 3) * 32);
                    const uint64_t db = desc_mnmajor_sw128(vb + (k >> 2) * HALF + (k & 3) * 2048);
                    umma_bf16(tmem + TM_O, da, db, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(v_empty(st));
                umma_commit(p_empty);
            }
        }
    } else {
        // ================================ softmax / correction / epilogue ================================
        const int q4 = warp & 3;                 // TMEM lane quarter this warp may touch
        const int half = (warp - 2) >> 2;        // which 64 of the block's 128 columns (and of O's 128) this warp owns
        const int r = q4 * 32 + lane;            // query row inside the block
        const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f;    // running maximum (log2 domain, scaled); l_run: this warp's half of the sum
        uint8_t* prow = smem + SM_P + half * HALF + r * 128;
        float* xch = reinterpret_cast<float*>(smem + SM_X);
        for (int j = 0; j < n_kv; ++j) {
            const int bs = j & 1;
            mbar_wait(s_full(bs), (j >> 1) & 1);
            tc_fence_after();
            const uint32_t ts = tmem + lane_addr + TM_S + bs * BKV + half * 64;
            // masking is needed on the causal diagonal block and on a ragged last key block (keys >= S are padding / the
            // next sequence's rows); a key is visible iff key < S and (not causal or key <= query)
            const bool diag = (p.causal && j == qb) || ((j + 1) * BKV > p.S);
            const int row_g = qb * BQ + r;
            const int lim = min(p.causal ? row_g : p.S - 1, p.S - 1) - j * BKV - half * 64;  // last visible own column
            // ---- this warp's 64 columns of the S row move TMEM -> registers once; the buffer goes back to the MMA warp ------
            uint32_t v[64];
            tmem_ld_32x32(ts, v);
            tmem_ld_32x32(ts + 32, v + 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty(bs));
            float mx = -INFINITY;
#pragma unroll
            for (int e = 0; e < 64; ++e)
                if (!diag || e <= lim) mx = fmaxf(mx, __uint_as_float(v[e]));
            // the row maximum needs the partner warp's half
            xch[(bs * 2 + half) * BQ + r] = mx;
            consumer_sync();
            mx = fmaxf(mx, xch[(bs * 2 + (half ^ 1)) * BQ + r]);
            const float m_new = fmaxf(m_run, mx * p.scale_log2);
            const float alpha = ex2(m_run - m_new);  // 0 on the first block (m_run = -inf)
            // ---- O correction (own 64 columns): after P_{j-1} V_{j-1} has landed, only if some row of this warp moved -------
            if (j > 0) {
                mbar_wait(p_empty, (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, m_new > m_run)) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t o[32];
                        tmem_ld_32x32(tmem + lane_addr + TM_O + half * 64 + c * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                        tmem_st_32x32(tmem + lane_addr + TM_O + half * 64 + c * 32, o);
                    }
                    tmem_st_wait();
                }
            }
            l_run *= alpha;
            m_run = m_new;
            // ---- P = 2^(s * scale - m), row sum, bf16 -> this warp's sub-tile of the swizzled K-major P tile ------------------
            float sum = 0.f;
#pragma unroll
            for (int chunk = 0; chunk < 8; ++chunk) {  // 8 x 16-byte chunks of 8 keys
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int col = chunk * 8 + e;
                    float pe = ex2(fmaf(__uint_as_float(v[col]), p.scale_log2, -m_new));
                    if (diag && col > lim) pe = 0.f;
                    sum += pe;
                    f[e] = pe;
                }
                *reinterpret_cast<Vec16*>(prow + ((chunk ^ (r & 7)) << 4)) = Elem<__nv_bfloat16>::pack(f);
            }
            l_run += sum;
            // P (and the corrected O) ready for the second MMA
            tc_fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // ---- epilogue: O / l -> bf16 -> global (own 64 columns), LSE ------------------------------------------------------------
        xch[(4 + half) * BQ + r] = l_run;
        consumer_sync();
        const float l_tot = l_run + xch[(4 + (half ^ 1)) * BQ + r];
        mbar_wait(p_empty, (n_kv - 1) & 1);
        tc_fence_after();
        const float inv_l = 1.f / l_tot;
        const bool row_ok = qb * BQ + r < p.S;
        __nv_bfloat16* orow = p.o + static_cast<int64_t>(row0 + r) * p.ld_o + h * D + half * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tmem + lane_addr + TM_O + half * 64 + c * 32, v);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[g * 8 + e]) * inv_l;
                    st_plain(orow + c * 32 + g * 8, Elem<__nv_bfloat16>::pack(f));
                }
            }
        }
        if (p.lse != nullptr && row_ok && half == 0)
            p.lse[(static_cast<int64_t>(b) * p.Hq + h) * p.S + qb * BQ + r] = (m_run + log2f(l_tot)) * 0.6931471805599453f;
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, TM_COLS);
}

// =============================================================================================================================
// Backward.  Two kernels, no atomics, every accumulator in TMEM:
//   (A) dK / dV: one CTA per (128 keys, KV head, batch) keeps K_j, V_j resident and streams (query head of the GQA group) x
//       (64-row query sub-block): S^T = K Q^T and dP^T = V dO^T (M = keys, so 128-row MMAs), one thread per KEY row turns
//       them into P^T = 2^(S^T*c - lse[q]) and dS^T = P^T (dP^T - delta[q]) * scale (the per-query statistics arrive in
//       shared memory with the tile), writes both as K-major bf16 tiles, and dV += P^T dO, dK += dS^T Q accumulate in TMEM
//       over ALL query heads of the group -- GQA needs no reduction pass.
//   (B) dQ: one CTA per (128 queries, query head, batch) keeps Q, dO resident and streams 64-key sub-blocks:
//       S = Q K^T, dP = dO V^T, one thread per QUERY row forms dS, dQ += dS K.
// S / P / dS are recomputed on chip; only Q, K, V, O, dO, LSE are read.  `delta = rowsum(dO * O)` and `lse * log2(e)` come
// from a small pre-pass.
// =============================================================================================================================
constexpr int BS = 64;                                   // streamed sub-block (queries in A, keys in B)
constexpr uint32_t SUB = BS * D * 2;                     // 16 KiB: [64 x 128] bf16 = 2 sub-tiles of [64 x 64]
constexpr uint32_t SUBH = SUB / 2;                       // 8 KiB
constexpr uint32_t PT = BQ * BS * 2;                     // 16 KiB: [128 x 64] bf16 K-major (one swizzle span per row)

struct BwdParams {
    __nv_bfloat16 *dq, *dk, *dv;  // outputs, 2-D [B*S, ld]: head h at columns [col0 + h*D, +D)
    int ld_dq, ld_dk, ld_dv;
    const float* lse2;   // [B, Hq, S]  log-sum-exp * log2(e)
    const float* delta;  // [B, Hq, S]  rowsum(dO * O)
    int S, Hq, Hkv;
    int q_col0, k_col0, v_col0;
    float scale, scale_log2;
    int causal;
};

// delta[b,h,s] = sum_d dO * O ; lse2 = lse * log2(e).  One warp per (token, head) row of 128 elements.
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, const float* __restrict__ lse,
                     float* __restrict__ delta, float* __restrict__ lse2, int B, int S, int Hq, int ld_o, int ld_do)
{
    const int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t rows = static_cast<int64_t>(B) * S * Hq;
    if (w >= rows) return;
    const int h = static_cast<int>(w % Hq);
    const int64_t t = w / Hq;  // token row b*S + s
    const __nv_bfloat16* po = o + t * ld_o + h * D + lane * 4;
    const __nv_bfloat16* pd = d_o + t * ld_do + h * D + lane * 4;
    const uint2 a = *reinterpret_cast<const uint2*>(po);
    const uint2 g = *reinterpret_cast<const uint2*>(pd);
    float acc = 0.f;
    acc = fmaf(__uint_as_float(a.x << 16), __uint_as_float(g.x << 16), acc);
    acc = fmaf(__uint_as_float(a.x & 0xffff0000u), __uint_as_float(g.x & 0xffff0000u), acc);
    acc = fmaf(__uint_as_float(a.y << 16), __uint_as_float(g.y << 16), acc);
    acc = fmaf(__uint_as_float(a.y & 0xffff0000u), __uint_as_float(g.y & 0xffff0000u), acc);
    acc = warp_reduce<SumOp>(acc);
    if (lane == 0) {
        const int b = static_cast<int>(t / S), s_ = static_cast<int>(t % S);
        const int64_t idx = (static_cast<int64_t>(b) * Hq + h) * S + s_;
        delta[idx] = acc;
        lse2[idx] = lse[idx] * 1.4426950408889634f;
    }
}

// ---- (A) dK, dV ---------------------------------------------------------------------------------------------------------------
namespace ka {
constexpr int kStages = 3;
constexpr uint32_t SM_K = 0, SM_V = TILE;
constexpr uint32_t SM_RING = 2 * TILE;                       // per stage: Q sub-block | dO sub-block
constexpr uint32_t SM_PT = SM_RING + kStages * 2 * SUB;      // 2 buffers x (P^T | dS^T)
constexpr uint32_t SM_STAT = SM_PT + 2 * 2 * PT;             // per stage: lse2[64] | delta[64]
constexpr uint32_t SM_BAR = SM_STAT + kStages * 2 * BS * 4;
constexpr uint32_t SM_TOTAL = SM_BAR + 256 + 1024;
constexpr uint32_t TM_ST = 0, TM_DPT = 128, TM_DV = 256, TM_DK = 384;
}  // namespace ka

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do, const BwdParams p)
{
    using namespace ka;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + SM_BAR;
    const uint32_t kv_full = bar;
    auto ring_full = [&](int s) { return bar + 8 * (1 + s); };
    auto ring_empty = [&](int s) { return bar + 8 * (4 + s); };
    auto sp_full = [&](int b) { return bar + 8 * (7 + b); };
    auto sp_empty = [&](int b) { return bar + 8 * (9 + b); };
    auto pd_full = [&](int b) { return bar + 8 * (11 + b); };
    auto pd_empty = [&](int b) { return bar + 8 * (13 + b); };
    const uint32_t acc_done = bar + 8 * 15;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8 * 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jb = blockIdx.x;  // key block: with causal masking block 0 has the most work and launches first
    const int g = blockIdx.y, b = blockIdx.z;
    const int rep = p.Hq / p.Hkv;
    const int kv_row0 = b * p.S + jb * BKV;
    const int n_sub = p.S / BS;
    const int sub0 = p.causal ? jb * (BKV / BS) : 0;  // first query sub-block that can see these keys
    const int per_head = n_sub - sub0;
    const int n_it = rep * per_head;

    if (warp == 0 && lane == 0) {
        prefetch_map(&map_q);
        prefetch_map(&map_k);
        prefetch_map(&map_v);
        prefetch_map(&map_do);
        mbar_init(kv_full, 1);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(ring_full(s), 1);
            mbar_init(ring_empty(s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(sp_full(i), 1);
            mbar_init(sp_empty(i), kConsumers);
            mbar_init(pd_full(i), kConsumers);
            mbar_init(pd_empty(i), 1);
        }
        mbar_init(acc_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(kv_full, 2 * TILE);
            tma_load_2d(sbase + SM_K, &map_k, kv_full, p.k_col0 + g * D, kv_row0);
            tma_load_2d(sbase + SM_K + HALF, &map_k, kv_full, p.k_col0 + g * D + 64, kv_row0);
            tma_load_2d(sbase + SM_V, &map_v, kv_full, p.v_col0 + g * D, kv_row0);
            tma_load_2d(sbase + SM_V + HALF, &map_v, kv_full, p.v_col0 + g * D + 64, kv_row0);
            for (int n = 0; n < n_it; ++n) {
                const int st = n % kStages;
                const int h = g * rep + n / per_head;
                const int qs = sub0 + n % per_head;
                const int q_row = b * p.S + qs * BS;
                mbar_wait(ring_empty(st), ((n / kStages) & 1) ^ 1);
                mbar_expect_tx(ring_full(st), 2 * SUB + 2 * BS * 4);
                const uint32_t dst = sbase + SM_RING + st * 2 * SUB;
                tma_load_2d(dst, &map_q, ring_full(st), p.q_col0 + h * D, q_row);
                tma_load_2d(dst + SUBH, &map_q, ring_full(st), p.q_col0 + h * D + 64, q_row);
                tma_load_2d(dst + SUB, &map_do, ring_full(st), h * D, q_row);
                tma_load_2d(dst + SUB + SUBH, &map_do, ring_full(st), h * D + 64, q_row);
                const int64_t soff = (static_cast<int64_t>(b) * p.Hq + h) * p.S + qs * BS;
                bulk_g2s(sbase + SM_STAT + st * 2 * BS * 4, p.lse2 + soff, BS * 4, ring_full(st));
                bulk_g2s(sbase + SM_STAT + st * 2 * BS * 4 + BS * 4, p.delta + soff, BS * 4, ring_full(st));
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc_st = idesc_bf16(BKV, BS, false, false);  // [128 keys x 64 queries], both K-major
            constexpr uint32_t idesc_acc = idesc_bf16(BKV, D, false, true);   // [128 keys x 128 dims], B MN-major
            auto issue_sp = [&](int n) {
                const int st = n % kStages, bs = n & 1;
                mbar_wait(ring_full(st), (n / kStages) & 1);
                mbar_wait(sp_empty(bs), ((n >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t qb = sbase + SM_RING + st * 2 * SUB, dob = qb + SUB;
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint32_t offa = (k >> 2) * HALF + (k & 3) * 32, offb = (k >> 2) * SUBH + (k & 3) * 32;
                    umma_bf16(tmem + TM_ST + bs * BS, desc_kmajor_sw128(sbase + SM_K + offa), desc_kmajor_sw128(qb + offb),
                              idesc_st, k > 0 ? 1u : 0u);
                }
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint32_t offa = (k >> 2) * HALF + (k & 3) * 32, offb = (k >> 2) * SUBH + (k & 3) * 32;
                    umma_bf16(tmem + TM_DPT + bs * BS, desc_kmajor_sw128(sbase + SM_V + offa), desc_kmajor_sw128(dob + offb),
                              idesc_st, k > 0 ? 1u : 0u);
                }
                umma_commit(sp_full(bs));
            };
            mbar_wait(kv_full, 0);
            if (n_it > 0) issue_sp(0);
            for (int n = 0; n < n_it; ++n) {
                if (n + 1 < n_it) issue_sp(n + 1);
                const int st = n % kStages, bs = n & 1;
                mbar_wait(pd_full(bs), (n >> 1) & 1);
                tc_fence_after();
                const uint32_t qb = sbase + SM_RING + st * 2 * SUB, dob = qb + SUB;
                const uint32_t ptb = sbase + SM_PT + bs * 2 * PT, dstb = ptb + PT;
#pragma unroll
                for (int k = 0; k < BS / 16; ++k)  // dV += P^T dO: A K-major over the 64 queries, B = dO [64 q x 128 d] MN-major
                    umma_bf16(tmem + TM_DV, desc_kmajor_sw128(ptb + k * 32), desc_mnmajor_sw128(dob + k * 2048, SUBH), idesc_acc,
                              (n > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < BS / 16; ++k)  // dK += dS^T Q
                    umma_bf16(tmem + TM_DK, desc_kmajor_sw128(dstb + k * 32), desc_mnmajor_sw128(qb + k * 2048, SUBH), idesc_acc,
                              (n > 0 || k > 0) ? 1u : 0u);
                umma_commit(pd_empty(bs));
                umma_commit(ring_empty(st));
            }
            umma_commit(acc_done);
        }
    } else {
        const int q4 = warp & 3;
        const int half = (warp - 2) >> 2;                  // which 32 of the sub-block's 64 query columns this warp owns
        const int r = q4 * 32 + lane;                      // key row inside the block
        const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
        const int key = jb * BKV + r;                      // key position inside the sequence
        for (int n = 0; n < n_it; ++n) {
            const int st = n % kStages, bs = n & 1;
            const int qs = sub0 + n % per_head;
            mbar_wait(sp_full(bs), (n >> 1) & 1);
            tc_fence_after();
            const float* stat = reinterpret_cast<const float*>(smem + SM_STAT + st * 2 * BS * 4) + half * 32;
            const bool diag = p.causal && (qs * BS < (jb + 1) * BKV);  // some (query, key) pairs of this tile are masked
            uint32_t sv[32], dv[32];
            tmem_ld_32x32(tmem + lane_addr + TM_ST + bs * BS + half * 32, sv);
            tmem_ld_32x32(tmem + lane_addr + TM_DPT + bs * BS + half * 32, dv);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sp_empty(bs));  // both accumulators are in registers: the next S^T / dP^T may start
            mbar_wait(pd_empty(bs), ((n >> 1) & 1) ^ 1);               // P^T / dS^T buffer free (MMAs of iteration n-2 done)
            uint8_t* prow = smem + SM_PT + bs * 2 * PT + r * 128;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float pf[8], df[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int col = ch * 8 + e;  // query inside this warp's 32
                    float pe = ex2(fmaf(__uint_as_float(sv[col]), p.scale_log2, -stat[col]));
                    if (diag && qs * BS + half * 32 + col < key) pe = 0.f;
                    pf[e] = pe;
                    df[e] = pe * (__uint_as_float(dv[col]) - stat[BS + col]) * p.scale;
                }
                const uint32_t off = static_cast<uint32_t>((((half * 4 + ch) ^ (r & 7)) << 4));
                *reinterpret_cast<Vec16*>(prow + off) = Elem<__nv_bfloat16>::pack(pf);
                *reinterpret_cast<Vec16*>(prow + PT + off) = Elem<__nv_bfloat16>::pack(df);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(pd_full(bs));
        }
        mbar_wait(acc_done, 0);
        tc_fence_after();
        const int64_t orow = static_cast<int64_t>(kv_row0 + r);
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            __nv_bfloat16* dst = (which == 0 ? p.dv + orow * p.ld_dv + p.v_col0 : p.dk + orow * p.ld_dk + p.k_col0) + g * D;
            const uint32_t tacc = tmem + lane_addr + (which == 0 ? TM_DV : TM_DK);
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int c = half * 2 + cc;  // this warp's 64 of the 128 dims
                uint32_t v[32];
                if (n_it > 0) {
                    tmem_ld_32x32(tacc + c * 32, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = 0u;
                }
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[gq * 8 + e]);
                    st_plain(dst + c * 32 + gq * 8, Elem<__nv_bfloat16>::pack(f));
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, TM_COLS);
}

// ---- (B) dQ -------------------------------------------------------------------------------------------------------------------
namespace kb {
constexpr int kStages = 3;
constexpr uint32_t SM_Q = 0, SM_DO = TILE;
constexpr uint32_t SM_RING = 2 * TILE;                       // per stage: K sub-block | V sub-block
constexpr uint32_t SM_DS = SM_RING + kStages * 2 * SUB;      // 2 buffers of dS [128 x 64]
constexpr uint32_t SM_BAR = SM_DS + 2 * PT;
constexpr uint32_t SM_TOTAL = SM_BAR + 256 + 1024;
constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DQ = 256;
}  // namespace kb

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do, const BwdParams p)
{
    using namespace kb;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + SM_BAR;
    const uint32_t q_full = bar;
    auto ring_full = [&](int s) { return bar + 8 * (1 + s); };
    auto ring_empty = [&](int s) { return bar + 8 * (4 + s); };
    auto sp_full = [&](int b) { return bar + 8 * (7 + b); };
    auto sp_empty = [&](int b) { return bar + 8 * (9 + b); };
    auto ds_full = [&](int b) { return bar + 8 * (11 + b); };
    auto ds_empty = [&](int b) { return bar + 8 * (13 + b); };
    const uint32_t acc_done = bar + 8 * 15;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8 * 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x);
    const int h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (p.Hq / p.Hkv);
    const int row0 = b * p.S + qb * BQ;
    const int n_it = p.causal ? (qb + 1) * (BQ / BS) : p.S / BS;  // 64-key sub-blocks this query block attends to

    if (warp == 0 && lane == 0) {
        prefetch_map(&map_q);
        prefetch_map(&map_k);
        prefetch_map(&map_v);
        prefetch_map(&map_do);
        mbar_init(q_full, 1);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(ring_full(s), 1);
            mbar_init(ring_empty(s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(sp_full(i), 1);
            mbar_init(sp_empty(i), kConsumers);
            mbar_init(ds_full(i), kConsumers);
            mbar_init(ds_empty(i), 1);
        }
        mbar_init(acc_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(q_full, 2 * TILE);
            tma_load_2d(sbase + SM_Q, &map_q, q_full, p.q_col0 + h * D, row0);
            tma_load_2d(sbase + SM_Q + HALF, &map_q, q_full, p.q_col0 + h * D + 64, row0);
            tma_load_2d(sbase + SM_DO, &map_do, q_full, h * D, row0);
            tma_load_2d(sbase + SM_DO + HALF, &map_do, q_full, h * D + 64, row0);
            for (int n = 0; n < n_it; ++n) {
                const int st = n % kStages;
                const int kv_row = b * p.S + n * BS;
                mbar_wait(ring_empty(st), ((n / kStages) & 1) ^ 1);
                mbar_expect_tx(ring_full(st), 2 * SUB);
                const uint32_t dst = sbase + SM_RING + st * 2 * SUB;
                tma_load_2d(dst, &map_k, ring_full(st), p.k_col0 + hk * D, kv_row);
                tma_load_2d(dst + SUBH, &map_k, ring_full(st), p.k_col0 + hk * D + 64, kv_row);
                tma_load_2d(dst + SUB, &map_v, ring_full(st), p.v_col0 + hk * D, kv_row);
                tma_load_2d(dst + SUB + SUBH, &map_v, ring_full(st), p.v_col0 + hk * D + 64, kv_row);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc_s = idesc_bf16(BQ, BS, false, false);  // [128 queries x 64 keys]
            constexpr uint32_t idesc_dq = idesc_bf16(BQ, D, false, true);   // dQ += dS K, K MN-major
            auto issue_sp = [&](int n) {
                const int st = n % kStages, bs = n & 1;
                mbar_wait(ring_full(st), (n / kStages) & 1);
                mbar_wait(sp_empty(bs), ((n >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t kb_ = sbase + SM_RING + st * 2 * SUB, vb = kb_ + SUB;
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint32_t offa = (k >> 2) * HALF + (k & 3) * 32, offb = (k >> 2) * SUBH + (k & 3) * 32;
                    umma_bf16(tmem + TM_S + bs * BS, desc_kmajor_sw128(sbase + SM_Q + offa), desc_kmajor_sw128(kb_ + offb),
                              idesc_s, k > 0 ? 1u : 0u);
                }
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint32_t offa = (k >> 2) * HALF + (k & 3) * 32, offb = (k >> 2) * SUBH + (k & 3) * 32;
                    umma_bf16(tmem + TM_DP + bs * BS, desc_kmajor_sw128(sbase + SM_DO + offa), desc_kmajor_sw128(vb + offb),
                              idesc_s, k > 0 ? 1u : 0u);
                }
                umma_commit(sp_full(bs));
            };
            mbar_wait(q_full, 0);
            issue_sp(0);
            for (int n = 0; n < n_it; ++n) {
                if (n + 1 < n_it) issue_sp(n + 1);
                const int st = n % kStages, bs = n & 1;
                mbar_wait(ds_full(bs), (n >> 1) & 1);
                tc_fence_after();
                const uint32_t kb_ = sbase + SM_RING + st * 2 * SUB;
                const uint32_t dsb = sbase + SM_DS + bs * PT;
#pragma unroll
                for (int k = 0; k < BS / 16; ++k)
                    umma_bf16(tmem + TM_DQ, desc_kmajor_sw128(dsb + k * 32), desc_mnmajor_sw128(kb_ + k * 2048, SUBH), idesc_dq,
                              (n > 0 || k > 0) ? 1u : 0u);
                umma_commit(ds_empty(bs));
                umma_commit(ring_empty(st));
            }
            umma_commit(acc_done);
        }
    } else {
        const int q4 = warp & 3;
        const int half = (warp - 2) >> 2;  // which 32 of the sub-block's 64 key columns this warp owns
        const int r = q4 * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
        const int64_t sidx = (static_cast<int64_t>(b) * p.Hq + h) * p.S + qb * BQ + r;
        const float lse2 = p.lse2[sidx], delta = p.delta[sidx];
        const int qpos = qb * BQ + r;
        for (int n = 0; n < n_it; ++n) {
            const int bs = n & 1;
            mbar_wait(sp_full(bs), (n >> 1) & 1);
            tc_fence_after();
            const bool diag = p.causal && ((n + 1) * BS > qb * BQ);
            uint32_t sv[32], dv[32];
            tmem_ld_32x32(tmem + lane_addr + TM_S + bs * BS + half * 32, sv);
            tmem_ld_32x32(tmem + lane_addr + TM_DP + bs * BS + half * 32, dv);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sp_empty(bs));
            mbar_wait(ds_empty(bs), ((n >> 1) & 1) ^ 1);
            uint8_t* drow = smem + SM_DS + bs * PT + r * 128;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float df[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int col = half * 32 + ch * 8 + e;
                    float pe = ex2(fmaf(__uint_as_float(sv[ch * 8 + e]), p.scale_log2, -lse2));
                    if (diag && n * BS + col > qpos) pe = 0.f;
                    df[e] = pe * (__uint_as_float(dv[ch * 8 + e]) - delta) * p.scale;
                }
                *reinterpret_cast<Vec16*>(drow + (((half * 4 + ch) ^ (r & 7)) << 4)) = Elem<__nv_bfloat16>::pack(df);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(ds_full(bs));
        }
        mbar_wait(acc_done, 0);
        tc_fence_after();
        __nv_bfloat16* dst = p.dq + static_cast<int64_t>(row0 + r) * p.ld_dq + p.q_col0 + h * D;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int c = half * 2 + cc;
            uint32_t v[32];
            tmem_ld_32x32(tmem + lane_addr + TM_DQ + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[gq * 8 + e]);
                st_plain(dst + c * 32 + gq * 8, Elem<__nv_bfloat16>::pack(f));
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, TM_COLS);
}

// ---- host ----------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &st) != cudaSuccess ||
        st != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
    return fn;
}
// row-major [rows, cols] bf16 with leading dimension ld (elements), box [box_rows x 64 cols], 128-byte swizzle
static int make_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return -3;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 100;
}

}  // namespace attn
}  // namespace dsb

using namespace dsb::attn;

// q / k / v: 2-D bf16 tensors [B*S, *] (row stride ld_*), head `i` at columns [col0 + i*128, +128) -- for a packed QKV
// projection output all three point into the same buffer.  o: [B*S, ld_o] (head h at h*128).  lse: fp32 [B, Hq, S] or null.
DSB_EXPORT int dsb_attn_fwd_bf16(const void* q, const void* k, const void* v, void* o, float* lse, int B, int S, int Hq,
                                 int Hkv, int head_dim, int ld_q, int ld_k, int ld_v, int ld_o, int q_cols, int k_cols,
                                 int v_cols, float scale, int causal, cudaStream_t stream)
{
    if (head_dim != D || S <= 0 || Hq % Hkv || ld_q % 8 || ld_k % 8 || ld_v % 8 || ld_o % 8) return -2;
    if (B > 1 && S % BQ) return -2;  // ragged lengths: one sequence per launch (rows of the next one would be read as keys)
    if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
         reinterpret_cast<uintptr_t>(o)) & 15)
        return -2;
    CUtensorMap mq, mk, mv;
    int rc;
    const uint64_t rows = static_cast<uint64_t>(B) * S;
    if ((rc = make_map(&mq, q, rows, q_cols, ld_q, BQ))) return rc;
    if ((rc = make_map(&mk, k, rows, k_cols, ld_k, BKV))) return rc;
    if ((rc = make_map(&mv, v, rows, v_cols, ld_v, 64))) return rc;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kf::SM_TOTAL);
        if (e != cudaSuccess) return static_cast<int>(e);
        attr = true;
    }
    FwdParams p;
    p.o = static_cast<__nv_bfloat16*>(o);
    p.lse = lse;
    p.ld_o = ld_o;
    p.S = S;
    p.Hq = Hq;
    p.Hkv = Hkv;
    p.q_col0 = p.k_col0 = p.v_col0 = 0;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.causal = causal;
    dim3 grid((S + BQ - 1) / BQ, Hq, B);
    attn_fwd_kernel<<<grid, kThreads, kf::SM_TOTAL, stream>>>(mq, mk, mv, p);
    DSB_CHECK_LAUNCH();
    return 0;
}

// Backward.  q/k/v as in the forward; o, d_o: [B*S, ld] with head h at h*128; lse from the forward.
// dq / dk / dv: 2-D outputs (may be column ranges of ONE packed buffer: pass the buffer base for all three with the
// q_col0 / k_col0 / v_col0 offsets folded into the pointers by the caller).  work: fp32 scratch of 2 * B * Hq * S elements.
DSB_EXPORT int dsb_attn_bwd_bf16(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                                 void* dq, void* dk, void* dv, float* work, int B, int S, int Hq, int Hkv, int head_dim,
                                 int ld_q, int ld_k, int ld_v, int ld_o, int ld_do, int ld_dq, int ld_dk, int ld_dv, int q_cols,
                                 int k_cols, int v_cols, float scale, int causal, cudaStream_t stream)
{
    if (head_dim != D || S % BQ || Hq % Hkv || (ld_q | ld_k | ld_v | ld_o | ld_do | ld_dq | ld_dk | ld_dv) % 8) return -2;
    CUtensorMap mq, mk, mv, mdo, mq64, mk64, mv64, mdo64;
    int rc;
    const uint64_t rows = static_cast<uint64_t>(B) * S;
    // 128-row boxes for the resident operands, 64-row boxes for the streamed sub-blocks
    if ((rc = make_map(&mq, q, rows, q_cols, ld_q, BQ))) return rc;
    if ((rc = make_map(&mk, k, rows, k_cols, ld_k, BKV))) return rc;
    if ((rc = make_map(&mv, v, rows, v_cols, ld_v, BKV))) return rc;
    if ((rc = make_map(&mdo, d_o, rows, static_cast<uint64_t>(Hq) * D, ld_do, BQ))) return rc;
    if ((rc = make_map(&mq64, q, rows, q_cols, ld_q, BS))) return rc;
    if ((rc = make_map(&mk64, k, rows, k_cols, ld_k, BS))) return rc;
    if ((rc = make_map(&mv64, v, rows, v_cols, ld_v, BS))) return rc;
    if ((rc = make_map(&mdo64, d_o, rows, static_cast<uint64_t>(Hq) * D, ld_do, BS))) return rc;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ka::SM_TOTAL);
        if (e != cudaSuccess) return static_cast<int>(e);
        e = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kb::SM_TOTAL);
        if (e != cudaSuccess) return static_cast<int>(e);
        attr = true;
    }
    float* delta = work;
    float* lse2 = work + static_cast<int64_t>(B) * Hq * S;
    const int64_t prep_rows = static_cast<int64_t>(B) * S * Hq;
    attn_bwd_prep_kernel<<<static_cast<unsigned>((prep_rows * 32 + 255) / 256), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(o), static_cast<const __nv_bfloat16*>(d_o), lse, delta, lse2, B, S, Hq, ld_o, ld_do);
    BwdParams p;
    p.dq = static_cast<__nv_bfloat16*>(dq);
    p.dk = static_cast<__nv_bfloat16*>(dk);
    p.dv = static_cast<__nv_bfloat16*>(dv);
    p.ld_dq = ld_dq;
    p.ld_dk = ld_dk;
    p.ld_dv = ld_dv;
    p.lse2 = lse2;
    p.delta = delta;
    p.S = S;
    p.Hq = Hq;
    p.Hkv = Hkv;
    p.q_col0 = p.k_col0 = p.v_col0 = 0;
    p.scale = scale;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.causal = causal;
    attn_bwd_dkdv_kernel<<<dim3(S / BKV, Hkv, B), kThreads, ka::SM_TOTAL, stream>>>(mq64, mk, mv, mdo64, p);
    attn_bwd_dq_kernel<<<dim3(S / BQ, Hq, B), kThreads, kb::SM_TOTAL, stream>>>(mq, mk64, mv64, mdo, p);
    DSB_CHECK_LAUNCH();
    return 0;
}
