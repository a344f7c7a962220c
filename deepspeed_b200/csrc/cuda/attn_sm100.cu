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
constexpr int kThreads = 192;                           // warp 0 TMA, warp 1 MMA, warps 2-5 softmax (one row per thread)
constexpr uint32_t TILE = BQ * D * 2;                   // 32 KiB: a [128 x 128] bf16 tile = 2 sub-tiles of [128 x 64]
constexpr uint32_t HALF = TILE / 2;                     // 16 KiB sub-tile (one 128-byte swizzle span of 64 bf16 per row)
constexpr uint32_t SM_Q = 0;
constexpr uint32_t SM_K = SM_Q + TILE;                  // 2 stages
constexpr uint32_t SM_V = SM_K + 2 * TILE;              // 2 stages
constexpr uint32_t SM_P = SM_V + 2 * TILE;
constexpr uint32_t SM_BAR = SM_P + TILE;
constexpr uint32_t SM_TOTAL = SM_BAR + 256 + 1024;
constexpr uint32_t TM_S = 0, TM_O = 256, TM_COLS = 512;  // S0 [0,128) S1 [128,256) O [256,384)

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
    const int n_kv = p.causal ? qb + 1 : p.S / BKV;

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
            mbar_init(s_empty(s), 4);  // one arrival per softmax warp
        }
        mbar_init(p_full, 4);
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
        const int r = q4 * 32 + lane;            // query row inside the block
        const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f;    // running maximum (log2 domain, scaled) and denominator
        uint8_t* prow = smem + SM_P + r * 128;
        for (int j = 0; j < n_kv; ++j) {
            const int bs = j & 1;
            mbar_wait(s_full(bs), (j >> 1) & 1);
            tc_fence_after();
            const uint32_t ts = tmem + lane_addr + TM_S + bs * BKV;
            const bool diag = p.causal && j == qb;
            // ---- pass 1: row maximum --------------------------------------------------------------------------------
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < BKV / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(ts + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float s = __uint_as_float(v[e]);
                    if (!diag || c * 32 + e <= r) mx = fmaxf(mx, s);
                }
            }
            const float m_new = fmaxf(m_run, mx * p.scale_log2);
            const float alpha = ex2(m_run - m_new);  // 0 on the first block (m_run = -inf)
            // ---- O correction: only after P_{j-1} V_{j-1} has landed, and only if some row of this warp moved ----------
            if (j > 0) {
                mbar_wait(p_empty, (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, m_new > m_run)) {
#pragma unroll
                    for (int c = 0; c < D / 32; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32(tmem + lane_addr + TM_O + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
                        tmem_st_32x32(tmem + lane_addr + TM_O + c * 32, v);
                    }
                    tmem_st_wait();
                }
            }
            l_run *= alpha;
            m_run = m_new;
            // ---- pass 2: P = 2^(s * scale - m), row sum, bf16 -> swizzled K-major smem tile -------------------------------
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < BKV / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(ts + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int g = 0; g < 4; ++g) {  // 4 x 8 columns -> 4 x 16-byte chunks
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int col = c * 32 + g * 8 + e;
                        float pe = ex2(fmaf(__uint_as_float(v[g * 8 + e]), p.scale_log2, -m_new));
                        if (diag && col > r) pe = 0.f;
                        sum += pe;
                        f[e] = pe;
                    }
                    const int chunk = c * 4 + g;  // 16-byte chunk index along the 128 keys: sub-tile chunk >> 3
                    const Vec16 pk = Elem<__nv_bfloat16>::pack(f);
                    *reinterpret_cast<Vec16*>(prow + (chunk >> 3) * HALF + (((chunk & 7) ^ (r & 7)) << 4)) = pk;
                }
            }
            l_run += sum;
            // S buffer free for block j + 2; P (and the corrected O) ready for the second MMA
            tc_fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(s_empty(bs));
                mbar_arrive(p_full);
            }
        }
        // ---- epilogue: O / l -> bf16 -> global, LSE ----------------------------------------------------------------------
        mbar_wait(p_empty, (n_kv - 1) & 1);
        tc_fence_after();
        const float inv_l = 1.f / l_run;
        __nv_bfloat16* orow = p.o + static_cast<int64_t>(row0 + r) * p.ld_o + h * D;
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tmem + lane_addr + TM_O + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[g * 8 + e]) * inv_l;
                st_plain(orow + c * 32 + g * 8, Elem<__nv_bfloat16>::pack(f));
            }
        }
        if (p.lse != nullptr)
            p.lse[(static_cast<int64_t>(b) * p.Hq + h) * p.S + qb * BQ + r] = (m_run + log2f(l_run)) * 0.6931471805599453f;
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
    if (head_dim != D || S % BQ || Hq % Hkv || ld_q % 8 || ld_k % 8 || ld_v % 8 || ld_o % 8) return -2;
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
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
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
    dim3 grid(S / BQ, Hq, B);
    attn_fwd_kernel<<<grid, kThreads, SM_TOTAL, stream>>>(mq, mk, mv, p);
    DSB_CHECK_LAUNCH();
    return 0;
}
