// In-kernel collectives over NVLink peer memory (symmetric arena) for the ZeRO hot paths.
//
//   * all_gather        : pull every peer's parameter shard straight into the local unit buffer
//   * reduce_scatter    : sum this rank's slice of every peer's gradient buffer in fp32, scale, and
//                         either ACCUMULATE into the gradient-shard arena or run the ADAM update on
//                         the fp32 master / moments and write the bf16 parameter shard -- one pass,
//                         no NCCL call, no intermediate reduced-gradient tensor
//   * all_reduce (one-shot, small tensors: TP inference activations, norms, flags)
//   * device-side barrier built from system-scope release/acquire flags in each rank's signal pad
//
// Every data kernel brackets itself with the cross-rank barrier: block 0 performs the flag exchange
// and opens a local gate that the other blocks spin on; the last block to finish performs the
// closing barrier, so one launch = sync + transfer/compute + sync.
// NVLS variants use multimem.ld_reduce / multimem.st on the multicast mapping when available.
//
// The reference implements these paths as torch.distributed calls + separate elementwise kernels
// (stage3.py:1372-1398, partition_parameters.py:1200, stage3.py:2178-2187).
#include "dsb_common.cuh"

namespace dsb {

constexpr int kMaxRanks = 8;
constexpr int kMaxSegs = 16;

struct PeerPtrs {
    void* p[kMaxRanks];
};

// Signal pad layout (uint32 words, per rank, lives in symmetric memory):
//   [channel][src_rank]  -- rank `src` writes its epoch into pad_of_dst[channel][src]
// plus local-only words used as the intra-kernel gate / completion counter.
constexpr int kChannels = 8;
constexpr int kPadWords = kChannels * kMaxRanks;      // remote-written region
constexpr int kLocalGate = kPadWords;                 // + channel  : gate epoch (local)
constexpr int kLocalDone = kPadWords + kChannels;     // + channel  : finished-block counter (local)

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Cross-rank barrier executed by (at least) `world` threads of ONE block.
__device__ __forceinline__ void rank_barrier(PeerPtrs pads, int rank, int world, int channel, uint32_t epoch)
{
    __threadfence_system();
    if (threadIdx.x < world) {
        const int peer = threadIdx.x;
        uint32_t* remote = static_cast<uint32_t*>(pads.p[peer]) + channel * kMaxRanks + rank;
        st_release_sys(remote, epoch);
        const uint32_t* mine = static_cast<const uint32_t*>(pads.p[rank]) + channel * kMaxRanks + peer;
        // epochs are monotonically increasing per channel; wrap-safe signed compare
        while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) __nanosleep(40);
    }
    __syncthreads();
}

// Entry gate: block 0 synchronises with the other ranks, then releases the local blocks.
__device__ __forceinline__ void gate_enter(PeerPtrs pads, int rank, int world, int channel, uint32_t epoch)
{
    uint32_t* local = static_cast<uint32_t*>(pads.p[rank]);
    if (blockIdx.x == 0) {
        rank_barrier(pads, rank, world, channel, epoch);
        if (threadIdx.x == 0) st_release_gpu(local + kLocalGate + channel, epoch);
    } else {
        if (threadIdx.x == 0) {
            while (static_cast<int32_t>(ld_acquire_gpu(local + kLocalGate + channel) - epoch) < 0) __nanosleep(20);
        }
        __syncthreads();
    }
}

// Exit: the last block to arrive runs the closing barrier (epoch + 1) so peers know our reads are done.
__device__ __forceinline__ void gate_exit(PeerPtrs pads, int rank, int world, int channel, uint32_t epoch)
{
    uint32_t* local = static_cast<uint32_t*>(pads.p[rank]);
    __shared__ int is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t prev = atomicAdd(local + kLocalDone + channel, 1u);
        is_last = (prev + 1 == gridDim.x);
        if (is_last) local[kLocalDone + channel] = 0;  // reset for the next launch on this channel
    }
    __syncthreads();
    if (is_last) rank_barrier(pads, rank, world, channel, epoch + 1);
}

// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) barrier_kernel(PeerPtrs pads, int rank, int world, int channel, uint32_t epoch)
{
    rank_barrier(pads, rank, world, channel, epoch);
}

// ------------------------------------------------------------------------------------------------------
// all-gather: full[p*S + i] = shard_of_peer_p[i].  `shards.p[p]` points at peer p's shard (peer VA).
// Peers are visited starting at rank+1 so all links carry traffic at the same time.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
all_gather_kernel(PeerPtrs shards, void* __restrict__ full, int64_t shard_bytes, PeerPtrs pads, int rank, int world,
                  int channel, uint32_t epoch, int sync_mode)
{
    if (sync_mode & 1) gate_enter(pads, rank, world, channel, epoch);
    const int64_t nvec = shard_bytes >> 4;
    const int64_t total = nvec * world;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    char* out = static_cast<char*>(full);
    constexpr int kU = 4;
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; i + (kU - 1) * stride < total; i += kU * stride) {
        Vec16 v[kU];
        int64_t dst[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int64_t j = i + u * stride;
            const int k = static_cast<int>(j / nvec);           // which peer slot (rotated)
            const int64_t e = j - static_cast<int64_t>(k) * nvec;
            const int peer = (rank + 1 + k) % world;
            dst[u] = (static_cast<int64_t>(peer) * nvec + e) << 4;
            v[u] = (peer == rank) ? ld_plain(static_cast<const char*>(shards.p[peer]) + (e << 4))
                                  : ld_peer(static_cast<const char*>(shards.p[peer]) + (e << 4));
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) st_plain(out + dst[u], v[u]);
    }
    for (; i < total; i += stride) {
        const int k = static_cast<int>(i / nvec);
        const int64_t e = i - static_cast<int64_t>(k) * nvec;
        const int peer = (rank + 1 + k) % world;
        const Vec16 v = (peer == rank) ? ld_plain(static_cast<const char*>(shards.p[peer]) + (e << 4))
                                       : ld_peer(static_cast<const char*>(shards.p[peer]) + (e << 4));
        st_plain(out + ((static_cast<int64_t>(peer) * nvec + e) << 4), v);
    }
    if (sync_mode & 2) gate_exit(pads, rank, world, channel, epoch);
}

// ------------------------------------------------------------------------------------------------------
// reduce-scatter + consumer.  grads.p[p] = base of peer p's FULL gradient buffer; this rank reduces
// elements [rank*S, (rank+1)*S).  T = gradient dtype (bf16 / fp16 / fp32).
// ------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void reduce8(const PeerPtrs& grads, int rank, int world, int64_t elem, float* acc)
{
    // 8 consecutive elements for 16-bit types (one 16-byte vector); 2 vectors for fp32.
    constexpr int kPer = Elem<T>::kPerVec;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    constexpr int kVecs = 8 / kPer;
    Vec16 v[kMaxRanks][kVecs];
#pragma unroll
    for (int k = 0; k < kMaxRanks; ++k) {
        if (k < world) {
            const int peer = (rank + k) % world;
            const T* src = static_cast<const T*>(grads.p[peer]) + elem;
#pragma unroll
            for (int q = 0; q < kVecs; ++q)
                v[k][q] = (peer == rank) ? ld_plain(src + q * kPer) : ld_peer(src + q * kPer);
        }
    }
#pragma unroll
    for (int k = 0; k < kMaxRanks; ++k) {
        if (k < world) {
#pragma unroll
            for (int q = 0; q < kVecs; ++q) {
                float f[kPer];
                Elem<T>::unpack(v[k][q], f);
#pragma unroll
                for (int e = 0; e < kPer; ++e) acc[q * kPer + e] += f[e];
            }
        }
    }
}

// NVLS: one multimem.ld_reduce returns the switch-computed sum over all ranks (bf16 inputs, fp32 acc).
__device__ __forceinline__ void reduce8_mc_bf16(const void* mc_addr, float* acc)
{
    uint32_t r0, r1, r2, r3;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "l"(mc_addr)
                 : "memory");
    Vec16 v;
    v.w[0] = r0; v.w[1] = r1; v.w[2] = r2; v.w[3] = r3;
    Elem<__nv_bfloat16>::unpack(v, acc);
}

// raw form of the above (the sum as packed bf16): lets a thread keep several switch reductions in flight before it
// starts consuming them -- one 16-byte request per thread cannot cover the NVLink round trip
__device__ __forceinline__ Vec16 mc_ld_reduce_raw(const void* mc_addr)
{
    Vec16 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
                 : "l"(mc_addr)
                 : "memory");
    return v;
}
constexpr int kRsUnroll = 4;

template <typename T, typename TD>
__global__ void __launch_bounds__(512)
reduce_scatter_acc_kernel(PeerPtrs grads, const void* __restrict__ mc_grads, TD* __restrict__ dst, int64_t shard_elems,
                          float scale, int accumulate, PeerPtrs pads, int rank, int world, int channel, uint32_t epoch,
                          float* __restrict__ sumsq_partials)
{
    __shared__ float scratch[32];
    gate_enter(pads, rank, world, channel, epoch);
    const int64_t base = static_cast<int64_t>(rank) * shard_elems;
    const int64_t n8 = shard_elems >> 3;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    float ss = 0.f;
    auto consume = [&](int64_t i, float* acc) {
        TD* d = dst + (i << 3);
        if (accumulate) {
            float old[8];
            if (sizeof(TD) == 4) {
                Elem<float>::unpack(ld_plain(d), old);
                Elem<float>::unpack(ld_plain(reinterpret_cast<const float*>(d) + 4), old + 4);
            } else {
                Elem<TD>::unpack(ld_plain(d), old);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(acc[e], scale, old[e]);
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] *= scale;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) ss = fmaf(acc[e], acc[e], ss);
        if (sizeof(TD) == 4) {
            st_plain(d, Elem<float>::pack(acc));
            st_plain(reinterpret_cast<float*>(d) + 4, Elem<float>::pack(acc + 4));
        } else {
            st_plain(d, Elem<TD>::pack(acc));
        }
    };
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (mc_grads != nullptr && sizeof(T) == 2) {
        const char* mc = static_cast<const char*>(mc_grads) + base * sizeof(T);
        for (; i + (kRsUnroll - 1) * stride < n8; i += kRsUnroll * stride) {
            Vec16 raw[kRsUnroll];
#pragma unroll
            for (int u = 0; u < kRsUnroll; ++u) raw[u] = mc_ld_reduce_raw(mc + ((i + u * stride) << 4));
#pragma unroll
            for (int u = 0; u < kRsUnroll; ++u) {
                float acc[8];
                Elem<__nv_bfloat16>::unpack(raw[u], acc);
                consume(i + u * stride, acc);
            }
        }
    }
    for (; i < n8; i += stride) {
        float acc[8];
        if (mc_grads != nullptr && sizeof(T) == 2)
            reduce8_mc_bf16(static_cast<const char*>(mc_grads) + ((base + (i << 3)) * sizeof(T)), acc);
        else
            reduce8<T>(grads, rank, world, base + (i << 3), acc);
        consume(i, acc);
    }
    if (sumsq_partials != nullptr) {
        ss = block_reduce<SumOp>(ss, scratch);
        if (threadIdx.x == 0) sumsq_partials[blockIdx.x] = ss;
    }
    gate_exit(pads, rank, world, channel, epoch);
}

struct AdamSeg {
    int64_t start, end;  // element range inside the shard
    float lr, beta1, beta2, eps, wd, bc1, bc2;
    int adamw;
};
struct AdamSegs {
    AdamSeg s[kMaxSegs];
    int n;
};

template <typename T, typename TO>
__global__ void __launch_bounds__(512)
reduce_scatter_adam_kernel(PeerPtrs grads, const void* __restrict__ mc_grads, float* __restrict__ master,
                           float* __restrict__ m, float* __restrict__ v, TO* __restrict__ lp_out,
                           int64_t shard_elems, float scale, AdamSegs segs, PeerPtrs pads, int rank, int world,
                           int channel, uint32_t epoch)
{
    gate_enter(pads, rank, world, channel, epoch);
    const int64_t base = static_cast<int64_t>(rank) * shard_elems;
    const int64_t n8 = shard_elems >> 3;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    auto adam = [&](int64_t o, const float* g) {
        // segments are 8-aligned (PARAM_ALIGN), so one lookup per vector
        int si = -1;
#pragma unroll 1
        for (int k = 0; k < segs.n; ++k)
            if (o >= segs.s[k].start && o < segs.s[k].end) si = k;
        if (si < 0) return;  // frozen / unmanaged range
        const AdamSeg a = segs.s[si];
        float p[8], mm[8], vv[8];
        Elem<float>::unpack(ld_stream(master + o), p);
        Elem<float>::unpack(ld_stream(master + o + 4), p + 4);
        Elem<float>::unpack(ld_stream(m + o), mm);
        Elem<float>::unpack(ld_stream(m + o + 4), mm + 4);
        Elem<float>::unpack(ld_stream(v + o), vv);
        Elem<float>::unpack(ld_stream(v + o + 4), vv + 4);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float ge = g[e] * scale;
            if (!a.adamw) ge = fmaf(a.wd, p[e], ge);
            mm[e] = fmaf(a.beta1, mm[e], (1.f - a.beta1) * ge);
            vv[e] = fmaf(a.beta2, vv[e], (1.f - a.beta2) * ge * ge);
            float upd = (mm[e] / a.bc1) / (sqrtf(vv[e] / a.bc2) + a.eps);
            if (a.adamw) upd = fmaf(a.wd, p[e], upd);
            p[e] = fmaf(-a.lr, upd, p[e]);
        }
        st_stream(master + o, Elem<float>::pack(p));
        st_stream(master + o + 4, Elem<float>::pack(p + 4));
        st_stream(m + o, Elem<float>::pack(mm));
        st_stream(m + o + 4, Elem<float>::pack(mm + 4));
        st_stream(v + o, Elem<float>::pack(vv));
        st_stream(v + o + 4, Elem<float>::pack(vv + 4));
        st_plain(lp_out + o, Elem<TO>::pack(p));
    };
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (mc_grads != nullptr && sizeof(T) == 2) {
        // NVLS: keep kRsUnroll switch reductions per thread in flight (the NVLink round trip is several microseconds)
        const char* mc = static_cast<const char*>(mc_grads) + base * sizeof(T);
        for (; i + (kRsUnroll - 1) * stride < n8; i += kRsUnroll * stride) {
            Vec16 raw[kRsUnroll];
#pragma unroll
            for (int u = 0; u < kRsUnroll; ++u) raw[u] = mc_ld_reduce_raw(mc + ((i + u * stride) << 4));
#pragma unroll
            for (int u = 0; u < kRsUnroll; ++u) {
                float g[8];
                Elem<__nv_bfloat16>::unpack(raw[u], g);
                adam((i + u * stride) << 3, g);
            }
        }
    }
    for (; i < n8; i += stride) {
        const int64_t o = i << 3;
        float g[8];
        if (mc_grads != nullptr && sizeof(T) == 2)
            reduce8_mc_bf16(static_cast<const char*>(mc_grads) + ((base + o) * sizeof(T)), g);
        else
            reduce8<T>(grads, rank, world, base + o, g);
        adam(o, g);
    }
    gate_exit(pads, rank, world, channel, epoch);
}

// ------------------------------------------------------------------------------------------------------
// one-shot all-reduce (small messages): every rank reads all peers' buffers and writes the sum in place
// into its own copy after the closing barrier protects the inputs.  out may alias the local input.
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(512)
all_reduce_oneshot_kernel(PeerPtrs bufs, T* __restrict__ out, int64_t n, PeerPtrs pads, int rank, int world,
                          int channel, uint32_t epoch)
{
    gate_enter(pads, rank, world, channel, epoch);
    const int64_t n8 = n >> 3;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    // results are staged in registers per thread chunk, written after the closing barrier would be
    // ideal; instead write to `out` which must NOT alias any rank's input when world > 1 and the
    // caller passes a separate output (python side provides a staging buffer).
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
        float acc[8];
        reduce8<T>(bufs, rank, world, i << 3, acc);
        if (sizeof(T) == 4) {
            st_plain(out + (i << 3), Elem<float>::pack(acc));
            st_plain(reinterpret_cast<float*>(out + (i << 3)) + 4, Elem<float>::pack(acc + 4));
        } else {
            st_plain(out + (i << 3), Elem<T>::pack(acc));
        }
    }
    gate_exit(pads, rank, world, channel, epoch);
}

}  // namespace dsb

using namespace dsb;

static inline PeerPtrs to_peers(void* const* arr, int world)
{
    PeerPtrs p;
    for (int i = 0; i < kMaxRanks; ++i) p.p[i] = (i < world) ? arr[i] : nullptr;
    return p;
}

DSB_EXPORT int dsb_symm_pad_bytes() { return (kPadWords + 2 * kChannels) * 4; }
DSB_EXPORT int dsb_symm_max_ranks() { return kMaxRanks; }
DSB_EXPORT int dsb_symm_channels() { return kChannels; }

DSB_EXPORT int dsb_symm_barrier(void* const* pads, int rank, int world, int channel, uint32_t epoch,
                                cudaStream_t stream)
{
    if (world > kMaxRanks) return -2;
    barrier_kernel<<<1, 32, 0, stream>>>(to_peers(pads, world), rank, world, channel, epoch);
    DSB_CHECK_LAUNCH();
    return 0;
}

// sync_mode: bit0 = barrier before, bit1 = barrier after (epoch, epoch+1 consumed when set)
DSB_EXPORT int dsb_symm_all_gather(void* const* shards, void* full, int64_t shard_bytes, void* const* pads, int rank,
                                   int world, int channel, uint32_t epoch, int sync_mode, int ctas,
                                   cudaStream_t stream)
{
    if (world > kMaxRanks || (shard_bytes & 15)) return -2;
    all_gather_kernel<<<ctas, 512, 0, stream>>>(to_peers(shards, world), full, shard_bytes, to_peers(pads, world), rank,
                                                 world, channel, epoch, sync_mode);
    DSB_CHECK_LAUNCH();
    return 0;
}

#define DISPATCH_G(code, T, ...)   \
    if ((code) == kBF16) {         \
        using T = __nv_bfloat16;   \
        __VA_ARGS__                \
    } else if ((code) == kF16) {   \
        using T = __half;          \
        __VA_ARGS__                \
    } else if ((code) == kF32) {   \
        using T = float;           \
        __VA_ARGS__                \
    } else {                       \
        return -1;                 \
    }

DSB_EXPORT int dsb_symm_reduce_scatter_acc(void* const* grads, const void* mc_grads, void* dst, int64_t shard_elems,
                                           int g_dtype, int d_dtype, float scale, int accumulate, void* const* pads,
                                           int rank, int world, int channel, uint32_t epoch, float* sumsq_partials,
                                           int ctas, cudaStream_t stream)
{
    if (world > kMaxRanks || (shard_elems & 7)) return -2;
    const PeerPtrs g = to_peers(grads, world), pd = to_peers(pads, world);
    if (g_dtype != kBF16) mc_grads = nullptr;
    DISPATCH_G(g_dtype, T, DISPATCH_G(d_dtype, TD, {
        reduce_scatter_acc_kernel<T, TD><<<ctas, 512, 0, stream>>>(g, mc_grads, (TD*)dst, shard_elems, scale, accumulate,
                                                                    pd, rank, world, channel, epoch, sumsq_partials);
    }))
    DSB_CHECK_LAUNCH();
    return 0;
}

// seg_data: n_segs rows of {start, end (int64 as 2 floats each via memcpy), lr, b1, b2, eps, wd, bc1, bc2, adamw}
// passed as packed struct array from the host (see comm/symm_impl.py).
DSB_EXPORT int dsb_symm_reduce_scatter_adam(void* const* grads, const void* mc_grads, float* master, float* m, float* v,
                                            void* lp_out, int64_t shard_elems, int g_dtype, int o_dtype, float scale,
                                            const AdamSeg* seg_host, int n_segs, void* const* pads, int rank, int world,
                                            int channel, uint32_t epoch, int ctas, cudaStream_t stream)
{
    if (world > kMaxRanks || (shard_elems & 7) || n_segs > kMaxSegs) return -2;
    AdamSegs segs;
    segs.n = n_segs;
    for (int i = 0; i < n_segs; ++i) segs.s[i] = seg_host[i];
    const PeerPtrs g = to_peers(grads, world), pd = to_peers(pads, world);
    if (g_dtype != kBF16) mc_grads = nullptr;
    DISPATCH_G(g_dtype, T, {
        if (o_dtype == kBF16)
            reduce_scatter_adam_kernel<T, __nv_bfloat16><<<ctas, 512, 0, stream>>>(
                g, mc_grads, master, m, v, (__nv_bfloat16*)lp_out, shard_elems, scale, segs, pd, rank, world, channel,
                epoch);
        else if (o_dtype == kF16)
            reduce_scatter_adam_kernel<T, __half><<<ctas, 512, 0, stream>>>(
                g, mc_grads, master, m, v, (__half*)lp_out, shard_elems, scale, segs, pd, rank, world, channel, epoch);
        else
            return -1;
    })
    DSB_CHECK_LAUNCH();
    return 0;
}

DSB_EXPORT int dsb_symm_all_reduce(void* const* bufs, void* out, int64_t n, int dtype, void* const* pads, int rank,
                                   int world, int channel, uint32_t epoch, int ctas, cudaStream_t stream)
{
    if (world > kMaxRanks || (n & 7)) return -2;
    const PeerPtrs b = to_peers(bufs, world), pd = to_peers(pads, world);
    DISPATCH_G(dtype, T, {
        all_reduce_oneshot_kernel<T><<<ctas, 512, 0, stream>>>(b, (T*)out, n, pd, rank, world, channel, epoch);
    })
    DSB_CHECK_LAUNCH();
    return 0;
}
