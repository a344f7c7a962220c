// Hand-written sm_100a GEMM: C[M,N] (bf16) = A[M,K] (bf16, K-major) x B[N,K]^T (bf16, K-major), fp32 accumulate.
//
//   * operands staged global -> shared by TMA (cp.async.bulk.tensor, 128-byte swizzle), 4-stage mbarrier ring
//   * tcgen05.mma (cta_group::1, kind::f16, M=128 N=256 K=16) issued by ONE elected thread, accumulators
//     in TMEM (2 x 256 columns: the epilogue of tile i overlaps the main loop of tile i+1)
//   * warp-specialised persistent CTAs (one per SM): warp0 = TMA producer, warp1 = MMA issuer,
//     warp2 = TMEM allocator, warps4-7 = epilogue (tcgen05.ld -> bf16 -> swizzled smem -> TMA store)
//   * OPTIONAL fused all-gather: when `ag` is given, the weight matrix B is being assembled in local
//     memory from the other ranks' ZeRO shards *by this same kernel*: `ag.comm_ctas` CTAs pull the
//     peers' shard bytes over NVLink (peer-mapped addresses) chunk by chunk and publish a per-chunk
//     ready flag (st.release.gpu); the TMA producer of each MMA CTA acquires the flag of the chunk(s)
//     backing a B tile before loading it, and tiles are visited starting with the locally owned rows,
//     so transfer and math overlap tile by tile with no NCCL call.
//
// The reference has no tensor-core GEMM of its own for training (cuBLAS via csrc/transformer/
// cublas_wrappers.cu:65 or torch.matmul); this kernel is the framework's native path.
#include <cuda.h>
#include "dsb_common.cuh"

namespace dsb {
namespace gemm {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 4;
constexpr int UMMA_K = 16;
constexpr int CCHUNK = 64;  // epilogue column chunk (64 bf16 = 128 B = one swizzle row)
constexpr int kThreads = 256;
constexpr int kEpiThreads = 128;
constexpr uint32_t A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr uint32_t B_STAGE_BYTES = BN * BK * 2;  // 32 KiB
constexpr uint32_t C_BUF_BYTES = BM * CCHUNK * 2;  // 16 KiB
constexpr uint32_t SMEM_A = 0;
constexpr uint32_t SMEM_B = SMEM_A + STAGES * A_STAGE_BYTES;
constexpr uint32_t SMEM_C = SMEM_B + STAGES * B_STAGE_BYTES;
constexpr uint32_t SMEM_BAR = SMEM_C + 2 * C_BUF_BYTES;
constexpr uint32_t SMEM_TOTAL = SMEM_BAR + 256 + 1024;  // + barriers + alignment slack
constexpr uint32_t TMEM_COLS = 512;

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
        : "memory");
}
// 32 lanes x 32 columns of fp32: thread `lane` receives row `lane`, registers = consecutive columns.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t* r)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
        "%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// K-major, 128B-swizzled operand tile: rows are 128 B apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_desc_kmajor_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4);          // start address
    d |= static_cast<uint64_t>(1) << 16;                              // LBO (unused for swizzled K-major) = 1
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                      // SBO = 1024 B
    d |= static_cast<uint64_t>(1) << 46;                              // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;                              // SWIZZLE_128B
    return d;
}

// c=F32, a=b=BF16, both K-major, N=256, M=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((BN >> 3) << 17) | ((BM >> 4) << 24);

// ---- fused all-gather descriptor ------------------------------------------------------------------------
struct AgParams {
    // B (= this unit's gathered weights) lives in `local_full`; rank p's shard is `shard_bytes` long and is
    // found at peers[p] (peer-mapped VA).  Chunk c covers bytes [c*chunk_bytes, (c+1)*chunk_bytes) of the
    // unit's flat buffer; flags[c] becomes `epoch` once the chunk is resident in local_full.
    const void* peers[8];
    void* local_full;
    uint32_t* flags;
    int64_t shard_bytes;
    int64_t chunk_bytes;
    int64_t b_offset_bytes;  // byte offset of B's first element inside the unit flat buffer
    int64_t b_bytes;         // extent of B: its chunks are gathered FIRST so the multiply can start while the rest streams
    int n_chunks;
    int world;
    int rank;
    int comm_ctas;  // trailing CTAs of the grid that do the gathering
    uint32_t epoch;
    int enabled;
};

__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Gather role: CTA `ci` of `nc` copies chunks ci, ci+nc, ... in the order "own shard first, then rank+1, ...".
// The copy is done by the TMA engine, not by SM load/store instructions: one elected thread streams each chunk through a
// ring of 3 x 64 KiB shared-memory buffers with  cp.async.bulk (peer global -> smem, mbarrier complete_tx)  followed by
// cp.async.bulk (smem -> local global, bulk_group);  192 KiB in flight per CTA hides the NVLink round trip, so a dozen
// gather CTAs saturate the links while leaving their SMs' issue slots idle.
constexpr uint32_t AG_PIECE = 64 * 1024;
constexpr int AG_BUFS = 3;

__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
                 : "memory");
}

__device__ void ag_gather_role(const AgParams& ag, int ci, int nc, uint8_t* smem)
{
    if (threadIdx.x != 0) return;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + AG_BUFS * AG_PIECE;  // AG_BUFS mbarriers after the ring
    for (int b = 0; b < AG_BUFS; ++b) mbar_init(bar0 + 8 * b, 1);
    fence_barrier_init();
    const int chunks_per_shard = static_cast<int>(ag.shard_bytes / ag.chunk_bytes);
    // chunks holding the matrix this very kernel multiplies by go first (pass 0), the rest of the unit follows (pass 1)
    const int b_c0 = static_cast<int>(ag.b_offset_bytes / ag.chunk_bytes);
    const int b_c1 = static_cast<int>((ag.b_offset_bytes + ag.b_bytes - 1) / ag.chunk_bytes);
    uint32_t issued = 0;           // pieces issued so far (ring position = issued % AG_BUFS)
    for (int pass = 0; pass < 2; ++pass)
    for (int j = ci; j < ag.n_chunks; j += nc) {
        // visit order: shards rotated so that every rank starts pulling from a different peer
        const int k = j / chunks_per_shard;
        const int within = j - k * chunks_per_shard;
        const int peer = (ag.rank + k) % ag.world;
        const int chunk = peer * chunks_per_shard + within;
        if ((chunk >= b_c0 && chunk <= b_c1) != (pass == 0)) continue;
        const int64_t off = static_cast<int64_t>(within) * ag.chunk_bytes;
        const char* src = static_cast<const char*>(ag.peers[peer]) + off;
        char* dst = static_cast<char*>(ag.local_full) + static_cast<int64_t>(peer) * ag.shard_bytes + off;
        const int n_pieces = static_cast<int>((ag.chunk_bytes + AG_PIECE - 1) / AG_PIECE);
        // software pipeline inside the chunk: loads run AG_BUFS-1 pieces ahead of the stores
        int loaded = 0, stored = 0;
        uint32_t first = issued;
        while (stored < n_pieces) {
            while (loaded < n_pieces && loaded - stored < AG_BUFS) {
                const uint32_t slot = (first + loaded) % AG_BUFS;
                // the store that last read this slot must have finished reading shared memory
                if (first + loaded >= static_cast<uint32_t>(AG_BUFS)) tma_store_wait_read<0>();
                const int64_t po = static_cast<int64_t>(loaded) * AG_PIECE;
                const int64_t rem = ag.chunk_bytes - po;
                const uint32_t bytes = static_cast<uint32_t>(rem < static_cast<int64_t>(AG_PIECE) ? rem : AG_PIECE);
                mbar_expect_tx(bar0 + 8 * slot, bytes);
                bulk_g2s(sbase + slot * AG_PIECE, src + po, bytes, bar0 + 8 * slot);
                ++loaded;
            }
            const uint32_t n = first + stored;
            const uint32_t slot = n % AG_BUFS;
            mbar_wait(bar0 + 8 * slot, (n / AG_BUFS) & 1);
            const int64_t po = static_cast<int64_t>(stored) * AG_PIECE;
            const int64_t rem2 = ag.chunk_bytes - po;
            const uint32_t bytes = static_cast<uint32_t>(rem2 < static_cast<int64_t>(AG_PIECE) ? rem2 : AG_PIECE);
            bulk_s2g(dst + po, sbase + slot * AG_PIECE, bytes);
            tma_store_commit();
            ++stored;
        }
        issued = first + n_pieces;
        tma_store_wait_all();  // the chunk's bytes are written before the flag is published
        __threadfence();
        st_release_gpu_u32(ag.flags + chunk, ag.epoch);
    }
}

// Wait until every chunk overlapping B rows [n0, n0+rows) is resident.
__device__ __forceinline__ void ag_wait_rows(const AgParams& ag, int n0, int rows, int K)
{
    const int64_t lo = ag.b_offset_bytes + static_cast<int64_t>(n0) * K * 2;
    const int64_t hi = lo + static_cast<int64_t>(rows) * K * 2;
    int c0 = static_cast<int>(lo / ag.chunk_bytes);
    int c1 = static_cast<int>((hi - 1) / ag.chunk_bytes);
    if (c1 >= ag.n_chunks) c1 = ag.n_chunks - 1;
    for (int c = c0; c <= c1; ++c) {
        while (ld_acquire_gpu_u32(ag.flags + c) != ag.epoch) __nanosleep(64);
    }
    // the chunk was written through the async proxy (bulk copies) and is about to be read through it (TMA loads)
    asm volatile("fence.proxy.async;" ::: "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, int M, int N, int K, const AgParams ag)
{
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment required by SWIZZLE_128B
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_base = sbase + SMEM_BAR;
    auto full_bar = [&](int s) { return bar_base + 8 * s; };
    auto empty_bar = [&](int s) { return bar_base + 8 * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8 * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8 * (2 * STAGES + 2 + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SMEM_BAR + 8 * (2 * STAGES + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int mma_ctas = ag.enabled ? static_cast<int>(gridDim.x) - ag.comm_ctas : static_cast<int>(gridDim.x);

    if (ag.enabled && static_cast<int>(blockIdx.x) >= mma_ctas) {
        ag_gather_role(ag, static_cast<int>(blockIdx.x) - mma_ctas, ag.comm_ctas, smem);
        return;
    }

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 4);  // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_m = (M + BM - 1) / BM;
    const int num_n = (N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_k = (K + BK - 1) / BK;
    // With the fused gather, n-blocks are visited starting at the locally owned rows.
    int n_rot = 0;
    if (ag.enabled) {
        const int64_t own_lo = static_cast<int64_t>(ag.rank) * ag.shard_bytes - ag.b_offset_bytes;
        int64_t row = own_lo > 0 ? (own_lo + static_cast<int64_t>(K) * 2 - 1) / (static_cast<int64_t>(K) * 2) : 0;
        n_rot = static_cast<int>((row + BN - 1) / BN) % num_n;
    }

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += mma_ctas) {
                const int m_blk = tile % num_m;
                const int n_blk = (tile / num_m + n_rot) % num_n;
                if (ag.enabled) ag_wait_rows(ag, n_blk * BN, (N - n_blk * BN) < BN ? (N - n_blk * BN) : BN, K);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), A_STAGE_BYTES + B_STAGE_BYTES);
                    tma_load_2d(sbase + SMEM_A + stage * A_STAGE_BYTES, &map_a, full_bar(stage), kb * BK, m_blk * BM);
                    tma_load_2d(sbase + SMEM_B + stage * B_STAGE_BYTES, &map_b, full_bar(stage), kb * BK, n_blk * BN);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += mma_ctas, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(tempty_bar(as), aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint64_t da = make_desc_kmajor_sw128(sbase + SMEM_A + stage * A_STAGE_BYTES);
                    const uint64_t db = make_desc_kmajor_sw128(sbase + SMEM_B + stage * B_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 32 bytes (16 bf16) along K inside the 128-byte swizzle row
                        umma_bf16(tmem_d, da + static_cast<uint64_t>((k * UMMA_K * 2) >> 4),
                                  db + static_cast<uint64_t>((k * UMMA_K * 2) >> 4), kIdesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(tfull_bar(as));  // accumulator complete -> epilogue
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;  // TMEM lane group == warp_id % 4
        const int etid = threadIdx.x - 128;
        const int row = ew * 32 + lane;  // row inside the 128-row tile
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += mma_ctas, ++it) {
            const int m_blk = tile % num_m;
            const int n_blk = (tile / num_m + n_rot) % num_n;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
#pragma unroll 1
            for (int c = 0; c < BN / CCHUNK; ++c) {
                const int buf = c & 1;
                if (etid == 0) tma_store_wait_read<1>();  // the store that last used `buf` has drained
                epi_bar_sync();
                uint32_t r[64];
                tmem_ld_32x32(taddr + c * CCHUNK, r);
                tmem_ld_32x32(taddr + c * CCHUNK + 32, r + 32);
                tmem_ld_wait();
                if (c == BN / CCHUNK - 1) {
                    // all TMEM reads of this accumulator are done: hand it back to the MMA warp early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(as));
                }
                uint8_t* crow = smem + SMEM_C + buf * C_BUF_BYTES + row * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[j * 8 + e]);
                    const Vec16 v = Elem<__nv_bfloat16>::pack(f);
                    const uint32_t dst = smem_u32(crow + ((j ^ (row & 7)) << 4));  // 128B swizzle
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]),
                                 "r"(v.w[3])
                                 : "memory");
                }
                fence_async_smem();
                epi_bar_sync();
                if (etid == 0) {
                    tma_store_2d(&map_c, sbase + SMEM_C + buf * C_BUF_BYTES, n_blk * BN + c * CCHUNK, m_blk * BM);
                    tma_store_commit();
                }
            }
        }
        if (etid == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}


// ==================================================================================================================
// Grouped (MoE) variant: C[rows, N] = A[rows, K] x W[e][N, K]^T where the rows of A are sorted by expert and expert e owns
// rows offsets[e] .. offsets[e+1].  One persistent launch covers every expert: the tile list is built on the device from
// `offsets` (no host synchronisation, CUDA-graph capturable), m-tiles of an expert start at its first row (so a tile never
// mixes two experts' weights) and the rows of a tail tile that belong to the next expert are computed and discarded by
// the epilogue, which therefore writes global memory directly (predicated 16-byte stores of full 128-byte row segments)
// instead of using a TMA store.  Same TMA -> tcgen05 -> TMEM pipeline as gemm_nt_kernel.
// Reference role: inference/v2/kernels/cutlass_ops/moe_gemm (CUTLASS grouped GEMM).
// ==================================================================================================================
constexpr int kMaxGroups = 256;
constexpr uint32_t SMEM_GRP = SMEM_BAR + 256;                        // int tile_prefix[kMaxGroups + 1]
constexpr uint32_t SMEM_GRP_TOTAL = SMEM_GRP + (kMaxGroups + 1) * 4 + 1024;

__global__ void __launch_bounds__(kThreads, 1)
gemm_grouped_nt_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       __nv_bfloat16* __restrict__ C, const int* __restrict__ offsets, int E, int N, int K, int ldc)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_base = sbase + SMEM_BAR;
    auto full_bar = [&](int s) { return bar_base + 8 * s; };
    auto empty_bar = [&](int s) { return bar_base + 8 * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8 * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8 * (2 * STAGES + 2 + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SMEM_BAR + 8 * (2 * STAGES + 4));
    int* tile_prefix = reinterpret_cast<int*>(smem + SMEM_GRP);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_n = (N + BN - 1) / BN;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 4);
        }
        fence_barrier_init();
    }
    if (warp == 3 && lane == 0) {
        int acc = 0;
        tile_prefix[0] = 0;
        for (int e = 0; e < E; ++e) {
            const int cnt = offsets[e + 1] - offsets[e];
            acc += ((cnt + BM - 1) / BM) * num_n;
            tile_prefix[e + 1] = acc;
        }
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int num_tiles = tile_prefix[E];
    const int num_k = (K + BK - 1) / BK;

    // tile id -> (expert, first row of the m-tile, rows of the expert left from there, n block); `e` only moves forward
    auto locate = [&](int tile, int& e, int& row0, int& rows_left, int& n_blk) {
        while (tile >= tile_prefix[e + 1]) ++e;
        const int lo = offsets[e], cnt = offsets[e + 1] - lo;
        const int num_m = (cnt + BM - 1) / BM;
        const int local = tile - tile_prefix[e];
        const int m_blk = local % num_m;
        n_blk = local / num_m;
        row0 = lo + m_blk * BM;
        rows_left = cnt - m_blk * BM;
    };

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0, e = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int row0, rows_left, n_blk;
                locate(tile, e, row0, rows_left, n_blk);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), A_STAGE_BYTES + B_STAGE_BYTES);
                    tma_load_2d(sbase + SMEM_A + stage * A_STAGE_BYTES, &map_a, full_bar(stage), kb * BK, row0);
                    tma_load_2d(sbase + SMEM_B + stage * B_STAGE_BYTES, &map_b, full_bar(stage), kb * BK, e * N + n_blk * BN);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(tempty_bar(as), aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint64_t da = make_desc_kmajor_sw128(sbase + SMEM_A + stage * A_STAGE_BYTES);
                    const uint64_t db = make_desc_kmajor_sw128(sbase + SMEM_B + stage * B_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        umma_bf16(tmem_d, da + static_cast<uint64_t>((k * UMMA_K * 2) >> 4),
                                  db + static_cast<uint64_t>((k * UMMA_K * 2) >> 4), kIdesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(tfull_bar(as));
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int row = ew * 32 + lane;
        int it = 0, e = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            int row0, rows_left, n_blk;
            locate(tile, e, row0, rows_left, n_blk);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
            const bool row_ok = row < rows_left;
            __nv_bfloat16* crow = C + static_cast<int64_t>(row0 + row) * ldc + n_blk * BN;
#pragma unroll 1
            for (int c = 0; c < BN / CCHUNK; ++c) {
                uint32_t r[64];
                tmem_ld_32x32(taddr + c * CCHUNK, r);
                tmem_ld_32x32(taddr + c * CCHUNK + 32, r + 32);
                tmem_ld_wait();
                if (c == BN / CCHUNK - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(as));
                }
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int col = n_blk * BN + c * CCHUNK + j * 8;
                        if (col < N) {  // N % 8 == 0: a vector of 8 is entirely inside or outside
                            float f[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) f[q] = __uint_as_float(r[j * 8 + q]);
                            *reinterpret_cast<Vec16*>(crow + c * CCHUNK + j * 8) = Elem<__nv_bfloat16>::pack(f);
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}


// ==================================================================================================================
// 2-CTA variant (cta_group::2): a CTA *pair* on one TPC computes a 256x256 tile.  Each CTA stages its own 128 rows of A
// and HALF of the B tile (128 of the 256 N rows), so the L2->SM operand traffic per flop drops by a third versus the
// 1-CTA kernel (32 KiB instead of 48 KiB per CTA per k-block); the leader CTA's elected thread issues
// tcgen05.mma.cta_group::2 (M=256) which reads both CTAs' shared memory and writes each CTA's half of the accumulator
// into that CTA's TMEM.  Barrier protocol:
//   full[s]   (leader only, count 2): each CTA's producer does a *remote* arrive.expect_tx(own bytes) on it and both
//             CTAs' TMA loads (cp.async.bulk.tensor...cta_group::2) complete_tx on it (peer bit of the address cleared)
//   empty[s]  (both CTAs): tcgen05.commit.cta_group::2 ... multicast::cluster mask 0b11 from the leader's MMA thread
//   tfull[a]  (both CTAs): multicast commit after the last k-block -> each CTA's epilogue drains its own TMEM half
//   tempty[a] (leader only, count 8): remote arrives from the 4 epilogue warps of each CTA
// ==================================================================================================================
constexpr int STAGES2 = 6;
constexpr uint32_t B2_STAGE_BYTES = (BN / 2) * BK * 2;  // 16 KiB: this CTA's half of the B tile
constexpr uint32_t SMEM2_A = 0;
constexpr uint32_t SMEM2_B = SMEM2_A + STAGES2 * A_STAGE_BYTES;
constexpr uint32_t SMEM2_C = SMEM2_B + STAGES2 * B2_STAGE_BYTES;
constexpr uint32_t SMEM2_BAR = SMEM2_C + 2 * C_BUF_BYTES;
constexpr uint32_t SMEM2_TOTAL = SMEM2_BAR + 256 + 1024;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> even CTA of the pair
constexpr uint32_t kIdesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((BN >> 3) << 17) | (((2 * BM) >> 4) << 24);
// MN-major operand (the matrix is stored [k, mn] row-major, i.e. mn is the contiguous dimension).  A stage holds the
// [128 mn x 64 k] tile as TWO TMA boxes of [64 k rows x 64 mn] (128-byte swizzled rows, 8 KiB each):
//   canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -> LBO = 8192 B between the two 64-wide mn atoms,
//   SBO = 1024 B between groups of 8 k rows; one UMMA_K=16 step advances the start address by 2048 B.
constexpr uint32_t MN_BOX_BYTES = 64 * 64 * 2;
__device__ __forceinline__ uint64_t make_desc_mnmajor_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4);
    d |= static_cast<uint64_t>(MN_BOX_BYTES >> 4) << 16;  // LBO
    d |= static_cast<uint64_t>(1024 >> 4) << 32;          // SBO
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx_leader(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(bar & kPeerMask), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerMask) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(dst),
        "l"(map), "r"(bar & kPeerMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit2(uint32_t bar)
{
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accum)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
        : "memory");
}

// ---- fused epilogues ------------------------------------------------------------------------------------------
// EPI_STORE    C = acc
// EPI_ACCUM    C = C + acc                                   (dW accumulation over micro-batches: no addmm, no memset)
// EPI_SWIGLU   B = [gate; up] stacked ([2I, K]); a tile multiplies 128 gate rows (leader CTA's half of B) and the
//              matching 128 up rows (peer CTA's half) so accumulator columns 0-127 / 128-255 are gate / up of the SAME
//              128 features:  C[M, I] = silu(gate) * up,  optionally gate|up saved to out2 [M, 2I] for backward.
// EPI_DSWIGLU  acc = d_act tile (dY W_down); aux = saved gate|up [M, 2I]:
//              dgate = acc * up * silu'(gate) -> C (the [M, I] left half of dgate_up),  dup = acc * silu(gate) -> out2 + I.
// (cuBLAS cannot express the last two: they remove one full pass over the [tokens, 2I] / [tokens, I] activations each.)
enum : int { EPI_STORE = 0, EPI_ACCUM = 1, EPI_SWIGLU = 2, EPI_DSWIGLU = 3 };

struct EpiParams {
    const __nv_bfloat16* aux;  // DSWIGLU: saved gate|up
    __nv_bfloat16* out2;       // SWIGLU: gate|up save (nullable); DSWIGLU: dgate_up base (dup goes to + inter)
    __nv_bfloat16* c;          // ACCUM: C base for the read-modify-write
    int ld_aux, ld_out2, ldc;
    int inter;                 // I
    int group_m;               // rasterisation: m-blocks per L2 super-group
};

__device__ __forceinline__ float silu_f(float g) { return g / (1.f + __expf(-g)); }
__device__ __forceinline__ float dsilu_f(float g)
{
    const float sg = 1.f / (1.f + __expf(-g));
    return sg * (1.f + g * (1.f - sg));
}

// Grouped rasterisation: tiles are numbered so that `group_m` consecutive m-blocks share every n-block before the next
// group starts -- one wave of CTA pairs then touches ~group_m A-blocks and ~pairs/group_m B-blocks instead of all A-blocks
// and two B-blocks, which keeps the wave's operand set inside L2 for the wide (N = 28672 / 128256) problems.
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int group_m, int& m_blk, int& n_blk)
{
    const int per_group = group_m * num_n;
    const int g = tile / per_group;
    const int first_m = g * group_m;
    const int gm = (num_m - first_m) < group_m ? (num_m - first_m) : group_m;
    const int local = tile - g * per_group;
    m_blk = first_m + local % gm;
    n_blk = local / gm;
}

// Shared-memory plan of the CTA-pair kernel.  The SwiGLU / dSwiGLU epilogues trade two pipeline stages for a double-buffered
// staging area of gate|up tiles: dSwiGLU TMA-LOADS the saved activations there (an otherwise idle warp, while the MMAs of
// the tile still run), SwiGLU stages the gate and up chunks it saves for backward and TMA-STOREs them (full 128-byte lines,
// asynchronous) next to the activation chunk.
template <int EPI>
struct Smem2 {
    static constexpr int kStages = (EPI == EPI_DSWIGLU || EPI == EPI_SWIGLU) ? 4 : STAGES2;
    static constexpr uint32_t A = 0;
    static constexpr uint32_t B = A + kStages * A_STAGE_BYTES;
    static constexpr uint32_t C = B + kStages * B2_STAGE_BYTES;
    static constexpr uint32_t AUX = C + 2 * C_BUF_BYTES;  // [2 buffers][gate chunk | up chunk] of 128 rows x 64 columns
    static constexpr uint32_t AUX_BYTES = (EPI == EPI_DSWIGLU || EPI == EPI_SWIGLU) ? 2 * 2 * C_BUF_BYTES : 0;
    static constexpr uint32_t BAR = AUX + AUX_BYTES;
    static constexpr uint32_t TOTAL = BAR + 256 + 1024;
};

template <bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_aux, int M, int N,
                 int K, const EpiParams ep)
{
    using L = Smem2<EPI>;
    constexpr int STAGES2 = L::kStages;  // shadows the namespace constant: the ring depth of THIS instantiation
    constexpr uint32_t SMEM2_A = L::A, SMEM2_B = L::B, SMEM2_C = L::C, SMEM2_BAR = L::BAR;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_base = sbase + SMEM2_BAR;
    auto full_bar = [&](int s) { return bar_base + 8 * s; };
    auto empty_bar = [&](int s) { return bar_base + 8 * (STAGES2 + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8 * (2 * STAGES2 + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8 * (2 * STAGES2 + 2 + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SMEM2_BAR + 8 * (2 * STAGES2 + 4));
    auto aux_full_bar = [&](int b) { return bar_base + 8 * (2 * STAGES2 + 5 + b); };
    auto aux_empty_bar = [&](int b) { return bar_base + 8 * (2 * STAGES2 + 7 + b); };

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n_pairs = static_cast<int>(gridDim.x) >> 1;
    const int pair = static_cast<int>(blockIdx.x) >> 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
        if constexpr (EPI == EPI_DSWIGLU || EPI == EPI_SWIGLU) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_aux) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES2; ++s) {
            mbar_init(full_bar(s), 2);   // one arrive.expect_tx per CTA of the pair (only the leader's copy is used)
            mbar_init(empty_bar(s), 1);  // multicast commit from the leader's MMA thread
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 8);  // 4 epilogue warps x 2 CTAs (only the leader's copy is used)
        }
        if constexpr (EPI == EPI_DSWIGLU) {
            for (int b = 0; b < 2; ++b) {
                mbar_init(aux_full_bar(b), 1);   // the loader's arrive.expect_tx (+ the TMA bytes)
                mbar_init(aux_empty_bar(b), 4);  // one arrival per epilogue warp of THIS CTA
            }
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc2(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // output-tile geometry: SWIGLU tiles are 128 output features wide (128 gate + 128 up accumulator columns)
    constexpr int TN = (EPI == EPI_SWIGLU) ? BN / 2 : BN;
    const int num_m = (M + 2 * BM - 1) / (2 * BM);
    const int num_n = (N + TN - 1) / TN;
    const int num_tiles = num_m * num_n;
    const int num_k = (K + BK - 1) / BK;
    const int group_m = ep.group_m > 0 ? ep.group_m : num_m;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += n_pairs) {
                int m_blk, n_blk;
                tile_coords(tile, num_m, num_n, group_m, m_blk, n_blk);
                const int m0 = m_blk * 2 * BM + static_cast<int>(rank) * BM;
                // this CTA's half of the B tile: plain = columns [rank*128, +128) of the tile; SWIGLU = gate rows for the
                // leader, the matching up rows (I further down the stacked weight) for the peer
                const int n0 = (EPI == EPI_SWIGLU) ? static_cast<int>(rank) * ep.inter + n_blk * TN
                                                   : n_blk * BN + static_cast<int>(rank) * (BN / 2);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx_leader(full_bar(stage), A_STAGE_BYTES + B2_STAGE_BYTES);
                    const uint32_t sa = sbase + SMEM2_A + stage * A_STAGE_BYTES;
                    const uint32_t sb = sbase + SMEM2_B + stage * B2_STAGE_BYTES;
                    if constexpr (A_MN) {
                        tma_load_2d_2sm(sa, &map_a, full_bar(stage), m0, kb * BK);
                        tma_load_2d_2sm(sa + MN_BOX_BYTES, &map_a, full_bar(stage), m0 + 64, kb * BK);
                    } else {
                        tma_load_2d_2sm(sa, &map_a, full_bar(stage), kb * BK, m0);
                    }
                    if constexpr (B_MN) {
                        tma_load_2d_2sm(sb, &map_b, full_bar(stage), n0, kb * BK);
                        tma_load_2d_2sm(sb + MN_BOX_BYTES, &map_b, full_bar(stage), n0 + 64, kb * BK);
                    } else {
                        tma_load_2d_2sm(sb, &map_b, full_bar(stage), kb * BK, n0);
                    }
                    if (++stage == STAGES2) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = pair; tile < num_tiles; tile += n_pairs, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(tempty_bar(as), aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = sbase + SMEM2_A + stage * A_STAGE_BYTES;
                    const uint32_t sb = sbase + SMEM2_B + stage * B2_STAGE_BYTES;
                    const uint64_t da = A_MN ? make_desc_mnmajor_sw128(sa) : make_desc_kmajor_sw128(sa);
                    const uint64_t db = B_MN ? make_desc_mnmajor_sw128(sb) : make_desc_kmajor_sw128(sb);
                    // per UMMA_K (16 elements of k): K-major advances 32 B inside the swizzle row, MN-major 16 k-rows = 2 KiB
                    constexpr uint64_t ka = A_MN ? (2048 >> 4) : ((UMMA_K * 2) >> 4);
                    constexpr uint64_t kbs = B_MN ? (2048 >> 4) : ((UMMA_K * 2) >> 4);
                    constexpr uint32_t idesc = kIdesc2 | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        umma_bf16_2cta(tmem_d, da + ka * k, db + kbs * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit2(empty_bar(stage));
                    if (++stage == STAGES2) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit2(tfull_bar(as));
            }
        }
    } else if (warp == 3) {
        // ===================== dSwiGLU: stage the saved gate|up tiles of upcoming chunks (both CTAs) =====================
        if constexpr (EPI == EPI_DSWIGLU) {
            if (elect_one()) {
                uint32_t cnt = 0;
                for (int tile = pair; tile < num_tiles; tile += n_pairs) {
                    int m_blk, n_blk;
                    tile_coords(tile, num_m, num_n, group_m, m_blk, n_blk);
                    const int m0 = m_blk * 2 * BM + static_cast<int>(rank) * BM;
                    for (int c = 0; c < TN / CCHUNK; ++c, ++cnt) {
                        const int b = cnt & 1;
                        mbar_wait(aux_empty_bar(b), ((cnt >> 1) & 1) ^ 1);
                        mbar_expect_tx(aux_full_bar(b), 2 * C_BUF_BYTES);
                        const uint32_t dst = sbase + L::AUX + b * 2 * C_BUF_BYTES;
                        const int col0 = n_blk * TN + c * CCHUNK;
                        tma_load_2d(dst, &map_aux, aux_full_bar(b), col0, m0);                         // gate chunk
                        tma_load_2d(dst + C_BUF_BYTES, &map_aux, aux_full_bar(b), ep.inter + col0, m0);  // up chunk
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs, own 128 rows) =====================
        const int ew = warp - 4;
        const int etid = threadIdx.x - 128;
        const int row = ew * 32 + lane;
        constexpr int kChunks = TN / CCHUNK;  // 64-column output chunks per tile (4, SWIGLU: 2)
        uint32_t aux_cnt = 0;
        int it = 0;
        for (int tile = pair; tile < num_tiles; tile += n_pairs, ++it) {
            int m_blk, n_blk;
            tile_coords(tile, num_m, num_n, group_m, m_blk, n_blk);
            const int m0 = m_blk * 2 * BM + static_cast<int>(rank) * BM;
            const int grow = m0 + row;  // this thread's global row
            const bool row_ok = grow < M;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c) {
                const int buf = c & 1;
                const int col0 = n_blk * TN + c * CCHUNK;  // first output column of this chunk
                if (etid == 0) tma_store_wait_read<1>();
                epi_bar_sync();
                uint8_t* crow = smem + SMEM2_C + buf * C_BUF_BYTES + row * 128;
                const uint8_t* arow = nullptr;  // dSwiGLU: this thread's row of the staged gate chunk (up: + C_BUF_BYTES)
                if constexpr (EPI == EPI_DSWIGLU) {
                    const int ab = aux_cnt & 1;
                    mbar_wait(aux_full_bar(ab), (aux_cnt >> 1) & 1);
                    arow = smem + L::AUX + ab * 2 * C_BUF_BYTES + row * 128;
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {  // two 32-column halves: bounds the live registers of the fused modes
                    const int colh = col0 + h * 32;
                    Vec16 x0[4], x1[4];  // operands fetched from global / shared memory (issued before the TMEM wait)
                    if constexpr (EPI == EPI_ACCUM) {
                        const __nv_bfloat16* src = ep.c + static_cast<int64_t>(grow) * ep.ldc + colh;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (row_ok && colh + j * 8 < N) x0[j] = ld_plain(src + j * 8);
                            else x0[j] = Vec16{{0, 0, 0, 0}};
                        }
                    }
                    if constexpr (EPI == EPI_DSWIGLU) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {  // same 128-byte swizzle as the C staging rows (TMA wrote them)
                            const uint32_t off = static_cast<uint32_t>(((h * 4 + j) ^ (row & 7)) << 4);
                            x0[j] = *reinterpret_cast<const Vec16*>(arow + off);
                            x1[j] = *reinterpret_cast<const Vec16*>(arow + C_BUF_BYTES + off);
                        }
                    }
                    uint32_t r[32], r2[32];
                    tmem_ld_32x32(taddr + c * CCHUNK + h * 32, r);
                    if constexpr (EPI == EPI_SWIGLU) tmem_ld_32x32(taddr + BN / 2 + c * CCHUNK + h * 32, r2);
                    tmem_ld_wait();
                    if (c == kChunks - 1 && h == 1) {
                        // all TMEM reads of this accumulator are done: hand it back to the MMA warp early
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader(tempty_bar(as));
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[j * 8 + e]);
                        if constexpr (EPI == EPI_ACCUM) {
                            float o[8];
                            Elem<__nv_bfloat16>::unpack(x0[j], o);
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] += o[e];
                        }
                        if constexpr (EPI == EPI_SWIGLU) {
                            float u[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) u[e] = __uint_as_float(r2[j * 8 + e]);
                            if (ep.out2 != nullptr) {
                                // gate / up chunks saved for backward: staged like the output chunk, stored by TMA below
                                uint8_t* grow_s = smem + L::AUX + buf * 2 * C_BUF_BYTES + row * 128;
                                const uint32_t off = static_cast<uint32_t>(((h * 4 + j) ^ (row & 7)) << 4);
                                *reinterpret_cast<Vec16*>(grow_s + off) = Elem<__nv_bfloat16>::pack(f);
                                *reinterpret_cast<Vec16*>(grow_s + C_BUF_BYTES + off) = Elem<__nv_bfloat16>::pack(u);
                            }
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] = silu_f(f[e]) * u[e];
                        }
                        if constexpr (EPI == EPI_DSWIGLU) {
                            float g[8], u[8], du[8];
                            Elem<__nv_bfloat16>::unpack(x0[j], g);
                            Elem<__nv_bfloat16>::unpack(x1[j], u);
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                du[e] = f[e] * silu_f(g[e]);
                                f[e] = f[e] * u[e] * dsilu_f(g[e]);
                            }
                            if (row_ok && colh + j * 8 < N)
                                st_plain(ep.out2 + static_cast<int64_t>(grow) * ep.ld_out2 + ep.inter + colh + j * 8,
                                         Elem<__nv_bfloat16>::pack(du));
                        }
                        const Vec16 v = Elem<__nv_bfloat16>::pack(f);
                        const uint32_t dst = smem_u32(crow + (((h * 4 + j) ^ (row & 7)) << 4));  // 128B swizzle
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]),
                                     "r"(v.w[3])
                                     : "memory");
                    }
                }
                if constexpr (EPI == EPI_DSWIGLU) {
                    // this warp's reads of the staged gate|up chunk are complete: hand the buffer back to the loader
                    __syncwarp();
                    if (lane == 0) mbar_arrive(aux_empty_bar(aux_cnt & 1));
                    ++aux_cnt;
                }
                fence_async_smem();
                epi_bar_sync();
                if (etid == 0 && m0 < M) {
                    tma_store_2d(&map_c, sbase + SMEM2_C + buf * C_BUF_BYTES, col0, m0);
                    if constexpr (EPI == EPI_SWIGLU) {
                        if (ep.out2 != nullptr) {
                            const uint32_t sg = sbase + L::AUX + buf * 2 * C_BUF_BYTES;
                            tma_store_2d(&map_aux, sg, col0, m0);
                            tma_store_2d(&map_aux, sg + C_BUF_BYTES, ep.inter + col0, m0);
                        }
                    }
                    tma_store_commit();
                }
            }
        }
        if (etid == 0) tma_store_wait_all();
    }

    tc_fence_before();
    cluster_sync_all();  // the peer may still be reading our smem / arriving on our barriers until here
    if (warp == 2) tmem_dealloc2(tmem_base, TMEM_COLS);
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) != cudaSuccess ||
        st != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// Row-major [rows, cols] bf16 matrix with leading dimension `ld` (elements); box = [box_rows, box_cols].
static int make_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                    uint32_t box_cols)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return -3;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 100;
}

}  // namespace gemm
}  // namespace dsb

using namespace dsb::gemm;

static int g_sm_count = 0;
static bool g_attr_set = false;

static int launch(const void* a, const void* b, void* c, int M, int N, int K, int lda, int ldb, int ldc, const AgParams& ag,
                  int sms, cudaStream_t stream)
{
    if (K % 8 || lda % 8 || ldb % 8 || ldc % 8) return -2;
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) return -2;
    CUtensorMap ma, mb, mc;
    int rc;
    if ((rc = make_map(&ma, a, M, K, lda, BM, BK))) return rc;
    if ((rc = make_map(&mb, b, N, K, ldb, BN, BK))) return rc;
    if ((rc = make_map(&mc, c, M, N, ldc, BM, CCHUNK))) return rc;
    if (!g_attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
        if (e != cudaSuccess) return static_cast<int>(e);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        g_attr_set = true;
    }
    int grid = sms > 0 ? sms : g_sm_count;
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    int mma = ag.enabled ? grid - ag.comm_ctas : grid;
    if (mma < 1) return -2;
    if (mma > tiles) {
        mma = tiles;
        grid = ag.enabled ? mma + ag.comm_ctas : mma;
    }
    gemm_nt_kernel<<<grid, kThreads, SMEM_TOTAL, stream>>>(ma, mb, mc, M, N, K, ag);
    DSB_CHECK_LAUNCH();
    return 0;
}

static bool g_attr2_set = false;

template <bool A_MN, bool B_MN, int EPI>
static int launch_2cta(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mx, int M,
                       int N, int K, int grid, const EpiParams& ep, cudaStream_t stream)
{
    static bool attr = false;
    constexpr uint32_t smem = Smem2<EPI>::TOTAL;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(gemm_2cta_kernel<A_MN, B_MN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        attr = true;
    }
    gemm_2cta_kernel<A_MN, B_MN, EPI><<<grid, kThreads, smem, stream>>>(ma, mb, mc, mx, M, N, K, ep);
    return 0;
}

template <int EPI>
static int dispatch_major(int a_mn, int b_mn, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc,
                          const CUtensorMap& mx, int M, int N, int K, int grid, const EpiParams& ep, cudaStream_t stream)
{
    if (a_mn && b_mn) return launch_2cta<true, true, EPI>(ma, mb, mc, mx, M, N, K, grid, ep, stream);
    if (a_mn) return launch_2cta<true, false, EPI>(ma, mb, mc, mx, M, N, K, grid, ep, stream);
    if (b_mn) return launch_2cta<false, true, EPI>(ma, mb, mc, mx, M, N, K, grid, ep, stream);
    return launch_2cta<false, false, EPI>(ma, mb, mc, mx, M, N, K, grid, ep, stream);
}

// 2-CTA (cta_group::2) GEMM, 256x256 tiles per CTA pair:  C[M,N] = epi(op(A) x op(B)), bf16 in/out, fp32 accumulate.
//   a_mn == 0: A is stored [M, K] row-major (K-major)      a_mn == 1: A is stored [K, M] row-major (MN-major)
//   b_mn == 0: B is stored [N, K] row-major (K-major)      b_mn == 1: B is stored [K, N] row-major (MN-major)
// so (0,0) = "NT" (y = x W^T), (0,1) = "NN" (dx = dy W), (1,1) = "TN" (dW = dy^T x).  lda/ldb are row strides in elements.
//   epi: 0 store, 1 accumulate into C, 2 SwiGLU (B = [gate; up] stacked [2*inter, K], N = inter, C = [M, inter], out2 =
//   optional gate|up save [M, 2*inter]), 3 dSwiGLU (N = inter, aux = saved gate|up, C = out2 = dgate|dup [M, 2*inter]).
//   group_m: m-blocks (of 256 rows) per rasterisation super-group, <= 0 = column-major tile order.
DSB_EXPORT int dsb_gemm_bf16_2cta_ex(const void* a, const void* b, void* c, int M, int N, int K, int lda, int ldb, int ldc,
                                     int a_mn, int b_mn, int epi, const void* aux, int ld_aux, void* out2, int ld_out2,
                                     int inter, int group_m, int sms, cudaStream_t stream)
{
    if (lda % 8 || ldb % 8 || ldc % 8) return -2;
    if ((a_mn ? M : K) % 8 || (b_mn ? N : K) % 8 || N % 8) return -2;
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) return -2;
    if (epi < 0 || epi > 3) return -2;
    if (epi == EPI_SWIGLU && (a_mn || b_mn || inter != N || inter % (BN / 2) || (out2 && (ld_out2 % 8)))) return -2;
    if (epi == EPI_DSWIGLU && (a_mn || !aux || !out2 || inter != N || ld_aux % 8 || ld_out2 % 8)) return -2;
    if ((reinterpret_cast<uintptr_t>(aux) | reinterpret_cast<uintptr_t>(out2)) & 15) return -2;
    CUtensorMap ma, mb, mc;
    int rc;
    if (a_mn) {
        if ((rc = make_map(&ma, a, K, M, lda, 64, 64))) return rc;  // rows = k, cols = mn; box [64 k x 64 mn]
    } else if ((rc = make_map(&ma, a, M, K, lda, BM, BK))) {
        return rc;
    }
    const uint64_t b_rows = (epi == EPI_SWIGLU) ? static_cast<uint64_t>(2) * inter : static_cast<uint64_t>(N);
    if (b_mn) {
        if ((rc = make_map(&mb, b, K, N, ldb, 64, 64))) return rc;
    } else if ((rc = make_map(&mb, b, b_rows, K, ldb, BN / 2, BK))) {
        return rc;
    }
    // C map spans exactly [M, N]: a partial last tile is clipped by the TMA unit (dSwiGLU: N = inter, i.e. the dgate half)
    if ((rc = make_map(&mc, c, M, N, ldc, BM, CCHUNK))) return rc;
    CUtensorMap mx = mc;  // dSwiGLU: the saved gate|up activations [M, 2*inter], staged chunk by chunk by the loader warp
    if (epi == EPI_DSWIGLU && (rc = make_map(&mx, aux, M, static_cast<uint64_t>(2) * inter, ld_aux, BM, CCHUNK))) return rc;
    if (epi == EPI_SWIGLU && out2 && (rc = make_map(&mx, out2, M, static_cast<uint64_t>(2) * inter, ld_out2, BM, CCHUNK)))
        return rc;
    if (!g_attr2_set) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        g_attr2_set = true;
    }
    EpiParams ep;
    ep.aux = static_cast<const __nv_bfloat16*>(aux);
    ep.out2 = static_cast<__nv_bfloat16*>(out2);
    ep.c = static_cast<__nv_bfloat16*>(c);
    ep.ld_aux = ld_aux;
    ep.ld_out2 = ld_out2;
    ep.ldc = ldc;
    ep.inter = inter;
    ep.group_m = group_m;
    int grid = (sms > 0 ? sms : g_sm_count) & ~1;
    const int tn = (epi == EPI_SWIGLU) ? BN / 2 : BN;
    const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + tn - 1) / tn);
    if (grid / 2 > tiles) grid = tiles * 2;
    if (grid < 2) return -2;
    switch (epi) {
        case EPI_STORE: rc = dispatch_major<EPI_STORE>(a_mn, b_mn, ma, mb, mc, mx, M, N, K, grid, ep, stream); break;
        case EPI_ACCUM: rc = dispatch_major<EPI_ACCUM>(a_mn, b_mn, ma, mb, mc, mx, M, N, K, grid, ep, stream); break;
        case EPI_SWIGLU: rc = launch_2cta<false, false, EPI_SWIGLU>(ma, mb, mc, mx, M, N, K, grid, ep, stream); break;
        default:
            rc = b_mn ? launch_2cta<false, true, EPI_DSWIGLU>(ma, mb, mc, mx, M, N, K, grid, ep, stream)
                      : launch_2cta<false, false, EPI_DSWIGLU>(ma, mb, mc, mx, M, N, K, grid, ep, stream);
            break;
    }
    if (rc) return rc;
    DSB_CHECK_LAUNCH();
    return 0;
}

DSB_EXPORT int dsb_gemm_bf16_2cta(const void* a, const void* b, void* c, int M, int N, int K, int lda, int ldb, int ldc,
                                  int a_mn, int b_mn, int sms, cudaStream_t stream)
{
    return dsb_gemm_bf16_2cta_ex(a, b, c, M, N, K, lda, ldb, ldc, a_mn, b_mn, EPI_STORE, nullptr, 0, nullptr, 0, 0, 8, sms,
                                 stream);
}

DSB_EXPORT int dsb_gemm_nt_bf16_2cta(const void* a, const void* b, void* c, int M, int N, int K, int lda, int ldb, int ldc,
                                     int sms, cudaStream_t stream)
{
    return dsb_gemm_bf16_2cta(a, b, c, M, N, K, lda, ldb, ldc, 0, 0, sms, stream);
}

// C[M,N] = A[M,K] @ B[N,K]^T   (bf16 in/out, fp32 accumulate).  sms <= 0 -> all SMs.
DSB_EXPORT int dsb_gemm_nt_bf16(const void* a, const void* b, void* c, int M, int N, int K, int lda, int ldb, int ldc,
                                int sms, cudaStream_t stream)
{
    AgParams ag;
    memset(&ag, 0, sizeof(ag));
    return launch(a, b, c, M, N, K, lda, ldb, ldc, ag, sms, stream);
}

// Fused all-gather + GEMM.  `b` points at B inside `local_full` (the unit's gathered buffer); the kernel
// fills local_full from peers[] (each `shard_bytes` long) while multiplying.  `flags` holds n_chunks words.
DSB_EXPORT int dsb_gemm_nt_bf16_allgather(const void* a, const void* b, void* c, int M, int N, int K, int lda, int ldb,
                                          int ldc, void* const* peers, void* local_full, uint32_t* flags,
                                          int64_t shard_bytes, int64_t chunk_bytes, int64_t b_offset_bytes, int world,
                                          int rank, int comm_ctas, uint32_t epoch, int sms, cudaStream_t stream)
{
    if (world > 8 || chunk_bytes <= 0 || shard_bytes % chunk_bytes || chunk_bytes % 16) return -2;
    AgParams ag;
    memset(&ag, 0, sizeof(ag));
    for (int i = 0; i < world; ++i) ag.peers[i] = peers[i];
    ag.local_full = local_full;
    ag.flags = flags;
    ag.shard_bytes = shard_bytes;
    ag.chunk_bytes = chunk_bytes;
    ag.b_offset_bytes = b_offset_bytes;
    ag.b_bytes = static_cast<int64_t>(N) * ldb * 2;
    ag.n_chunks = static_cast<int>(shard_bytes / chunk_bytes) * world;
    ag.world = world;
    ag.rank = rank;
    ag.comm_ctas = comm_ctas;
    ag.epoch = epoch;
    ag.enabled = 1;
    return launch(a, b, c, M, N, K, lda, ldb, ldc, ag, sms, stream);
}

// Grouped GEMM for expert-sorted rows: c[r, :] = a[r, :] @ w[e]^T for offsets[e] <= r < offsets[e+1].
// a [rows, K] (lda), w [E, N, K] contiguous, c [rows, N] (ldc), offsets int32 [E+1] ON THE DEVICE.
static bool g_attr_grp_set = false;
DSB_EXPORT int dsb_gemm_grouped_nt_bf16(const void* a, const void* w, void* c, const int* offsets, int rows, int E, int N, int K,
                                        int lda, int ldc, int sms, cudaStream_t stream)
{
    if (rows <= 0 || E <= 0) return 0;
    if (E > kMaxGroups || K % 8 || N % 8 || lda % 8 || ldc % 8) return -3;
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(c)) & 15) return -3;
    CUtensorMap ma, mb;
    int rc;
    if ((rc = make_map(&ma, a, rows, K, lda, BM, BK))) return rc;
    if ((rc = make_map(&mb, w, static_cast<uint64_t>(E) * N, K, K, BN, BK))) return rc;
    if (!g_attr_grp_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_grouped_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_GRP_TOTAL);
        if (e != cudaSuccess) return static_cast<int>(e);
        if (!g_sm_count) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        }
        g_attr_grp_set = true;
    }
    int grid = sms > 0 ? sms : g_sm_count;
    // upper bound of the tile count (each expert adds at most one partial m-tile) so tiny problems do not launch idle CTAs
    const int64_t max_tiles = (static_cast<int64_t>(rows + BM - 1) / BM + E) * ((N + BN - 1) / BN);
    if (grid > max_tiles) grid = static_cast<int>(max_tiles);
    gemm_grouped_nt_kernel<<<grid, kThreads, SMEM_GRP_TOTAL, stream>>>(ma, mb, static_cast<__nv_bfloat16*>(c), offsets, E, N, K, ldc);
    DSB_CHECK_LAUNCH();
    return 0;
}

DSB_EXPORT int dsb_gemm_tile_m() { return BM; }
DSB_EXPORT int dsb_gemm_tile_n() { return BN; }
DSB_EXPORT int dsb_gemm_tile_k() { return BK; }
