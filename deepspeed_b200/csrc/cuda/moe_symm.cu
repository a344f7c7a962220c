// Expert-parallel dispatch / combine fused with the all-to-all, over NVLink peer memory.
//
// The reference (and this repo's NCCL path) does   scatter -> all_to_all_single -> experts -> all_to_all_single -> gather
// (deepspeed/moe/sharded_moe.py:609,669 `_AllToAll`).  Here the permutation kernels address the *peers'* symmetric
// buffers directly, so there is no staging buffer and no NCCL call:
//
//   dispatch:  token row x[t] is stored straight into the receive buffer of the rank that owns its expert,
//              at   row = (e_local * ep + src_rank) * C + pos      (the [E_local, ep*C, H] layout the experts consume)
//   combine:   y[t] = sum_k w[t,k] * out_of_owner_rank[row]        (peer loads, fp32 accumulation)
//
// and the same two kernels run the backward pass with roles swapped (combine-backward scatters w*dy into the peers and
// dots dy with the peers' expert outputs for the gate-weight gradient; dispatch-backward gathers).  Ordering between ranks
// is provided by the symmetric-memory barrier kernels around these launches (symm_coll.cu).
#include "dsb_common.cuh"

namespace dsb {
namespace moesymm {

constexpr int kMaxRanks = 8;
struct Peers {
    void* p[kMaxRanks];
};

// One block per assignment a = t*K + k.  value = scale * x[t]  (scale = weights[a] when given).
// If `eo_peers` is given, additionally dweights[a] = <x[t], eo_peer_row>  (combine backward).
template <typename T>
__global__ void __launch_bounds__(256)
scatter_peer_kernel(const T* __restrict__ x, Peers dst, const int32_t* __restrict__ expert_ids,
                    const int32_t* __restrict__ positions, const float* __restrict__ weights, Peers eo_peers,
                    float* __restrict__ dweights, int n_assign, int K, int hidden, int capacity, int e_local, int ep,
                    int my_rank, int has_eo)
{
    __shared__ float scratch[32];
    constexpr int kPer = Elem<T>::kPerVec;
    const int a = blockIdx.x;
    if (a >= n_assign) return;
    const int e = expert_ids[a];
    const int pos = positions[a];
    if (pos >= capacity) {
        if (threadIdx.x == 0 && dweights) dweights[a] = 0.f;
        return;
    }
    const int r = e / e_local;
    const int64_t row = (static_cast<int64_t>(e % e_local) * ep + my_rank) * capacity + pos;
    const T* src = x + static_cast<int64_t>(a / K) * hidden;
    T* d = static_cast<T*>(dst.p[r]) + row * hidden;
    const T* eo = has_eo ? static_cast<const T*>(eo_peers.p[r]) + row * hidden : nullptr;
    const float w = weights ? weights[a] : 1.f;
    const bool local = r == my_rank;
    float dot = 0.f;
    for (int v = threadIdx.x; v < hidden / kPer; v += blockDim.x) {
        const Vec16 raw = ld_stream(src + v * kPer);
        if (weights || has_eo) {
            float f[kPer], o[kPer];
            Elem<T>::unpack(raw, f);
            if (has_eo) {
                float g[kPer];
                Elem<T>::unpack(local ? ld_plain(eo + v * kPer) : ld_peer(eo + v * kPer), g);
#pragma unroll
                for (int i = 0; i < kPer; ++i) dot = fmaf(f[i], g[i], dot);
            }
#pragma unroll
            for (int i = 0; i < kPer; ++i) o[i] = w * f[i];
            const Vec16 out = Elem<T>::pack(o);
            if (local) st_plain(d + v * kPer, out); else st_peer(d + v * kPer, out);
        } else {
            if (local) st_plain(d + v * kPer, raw); else st_peer(d + v * kPer, raw);
        }
    }
    if (dweights) {
        dot = block_reduce<SumOp>(dot, scratch);
        if (threadIdx.x == 0) dweights[a] = dot;
    }
}

// One block per token: y[t] = sum_k w[t,k] * src_peer[row(t,k)]   (w == 1 when weights is null).
template <typename T>
__global__ void __launch_bounds__(256)
gather_peer_kernel(Peers src, T* __restrict__ y, const int32_t* __restrict__ expert_ids,
                   const int32_t* __restrict__ positions, const float* __restrict__ weights, int tokens, int K, int hidden,
                   int capacity, int e_local, int ep, int my_rank)
{
    constexpr int kPer = Elem<T>::kPerVec;
    constexpr int kMaxK = 8;
    const int t = blockIdx.x;
    if (t >= tokens) return;
    const T* rowp[kMaxK];
    float w[kMaxK];
    bool loc[kMaxK];
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) {
        rowp[k] = nullptr;
        w[k] = 0.f;
        loc[k] = false;
        if (k < K) {
            const int a = t * K + k;
            const int pos = positions[a];
            if (pos < capacity) {
                const int e = expert_ids[a];
                const int r = e / e_local;
                rowp[k] = static_cast<const T*>(src.p[r]) +
                          ((static_cast<int64_t>(e % e_local) * ep + my_rank) * capacity + pos) * hidden;
                w[k] = weights ? weights[a] : 1.f;
                loc[k] = r == my_rank;
            }
        }
    }
    for (int v = threadIdx.x; v < hidden / kPer; v += blockDim.x) {
        float acc[kPer];
#pragma unroll
        for (int i = 0; i < kPer; ++i) acc[i] = 0.f;
#pragma unroll
        for (int k = 0; k < kMaxK; ++k) {
            if (k < K && rowp[k]) {
                float f[kPer];
                Elem<T>::unpack(loc[k] ? ld_plain(rowp[k] + v * kPer) : ld_peer(rowp[k] + v * kPer), f);
#pragma unroll
                for (int i = 0; i < kPer; ++i) acc[i] = fmaf(w[k], f[i], acc[i]);
            }
        }
        st_plain(y + static_cast<int64_t>(t) * hidden + v * kPer, Elem<T>::pack(acc));
    }
}

}  // namespace moesymm
}  // namespace dsb

using namespace dsb;
using namespace dsb::moesymm;

static Peers to_peers(void* const* p, int world)
{
    Peers o;
    for (int i = 0; i < kMaxRanks; ++i) o.p[i] = i < world ? p[i] : nullptr;
    return o;
}

#define DISPATCH_MS(code, T, ...)  \
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

// dst_peers[r]: base of rank r's receive buffer ([E_local, ep, C, H]).  eo_peers (optional): peers' expert outputs for the
// gate-weight gradient.  weights (optional): per-assignment scale.
DSB_EXPORT int dsb_moe_scatter_peer(const void* x, void* const* dst_peers, const int32_t* expert_ids,
                                    const int32_t* positions, const float* weights, void* const* eo_peers, float* dweights,
                                    int n_assign, int K, int hidden, int capacity, int e_local, int ep, int my_rank,
                                    int dtype, cudaStream_t stream)
{
    if (n_assign <= 0) return 0;
    if (ep > kMaxRanks) return -2;
    const int per = dtype == kF32 ? 4 : 8;
    if (hidden % per) return -2;
    const Peers d = to_peers(dst_peers, ep);
    const Peers e = eo_peers ? to_peers(eo_peers, ep) : Peers{};
    DISPATCH_MS(dtype, T, {
        scatter_peer_kernel<T><<<n_assign, 256, 0, stream>>>((const T*)x, d, expert_ids, positions, weights, e, dweights,
                                                              n_assign, K, hidden, capacity, e_local, ep, my_rank,
                                                              eo_peers != nullptr);
    })
    DSB_CHECK_LAUNCH();
    return 0;
}

DSB_EXPORT int dsb_moe_gather_peer(void* const* src_peers, void* y, const int32_t* expert_ids, const int32_t* positions,
                                   const float* weights, int tokens, int K, int hidden, int capacity, int e_local, int ep,
                                   int my_rank, int dtype, cudaStream_t stream)
{
    if (tokens <= 0) return 0;
    if (ep > kMaxRanks || K > 8) return -2;
    const int per = dtype == kF32 ? 4 : 8;
    if (hidden % per) return -2;
    const Peers s = to_peers(src_peers, ep);
    DISPATCH_MS(dtype, T, {
        gather_peer_kernel<T><<<tokens, 256, 0, stream>>>(s, (T*)y, expert_ids, positions, weights, tokens, K, hidden,
                                                           capacity, e_local, ep, my_rank);
    })
    DSB_CHECK_LAUNCH();
    return 0;
}
