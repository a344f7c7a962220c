// Paged-KV attention for ragged (continuous-batching) inference on sm_100a.
//
// One CTA per (query token, query head).  K/V live in a blocked cache laid out by `kv_rotary_append`
// (moe_ragged.cu): cache[block][slot][k|v][kv_head][d].  Each warp walks a strided subset of the visible
// keys with an online softmax (running max / sum, fp32 accumulators), lanes split the head dimension;
// the four per-warp partial states are merged through shared memory.  GQA: query head h reads kv head
// h / (hq / hkv).  Causality: query token with absolute position p sees keys [0, p].
//
// Role parity: reference inference/v2/kernels/ragged_ops/blocked_flash (N9b, a wrapper over the external
// `dskernels` flash-attention build) and the v1 `softmax_context` decode path (N8).  Prefill-sized work goes
// through the training attention path (ops/attention.py); this kernel is the decode / short-chunk path.
#include "dsb_common.cuh"

namespace dsb {
namespace pattn {

constexpr int kWarps = 4;
constexpr int kMaxPerLane = 8;  // head_dim <= 256

template <typename T>
__global__ void __launch_bounds__(kWarps * 32)
paged_attention_kernel(const T* __restrict__ q, const T* __restrict__ cache, T* __restrict__ out,
                       const int32_t* __restrict__ seq_of, const int32_t* __restrict__ pos_of,
                       const int32_t* __restrict__ block_table, int hq, int hkv, int d, int q_stride, int block_size,
                       int max_blocks, float scale)
{
    __shared__ float sm_m[kWarps], sm_l[kWarps];
    __shared__ float sm_o[kWarps][256];
    const int t = blockIdx.x, h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seq = seq_of[t];
    const int kv_len = pos_of[t] + 1;
    const int kvh = h / (hq / hkv);
    const int per = (d + 31) / 32;
    float qv[kMaxPerLane], acc[kMaxPerLane];
    const T* qrow = q + static_cast<int64_t>(t) * q_stride + static_cast<int64_t>(h) * d;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
        const int e = lane + i * 32;
        qv[i] = (i < per && e < d) ? Elem<T>::to_f(qrow[e]) * scale : 0.f;
        acc[i] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    const int64_t tok_stride = static_cast<int64_t>(2) * hkv * d;
    for (int j = warp; j < kv_len; j += kWarps) {
        const int blk = block_table[seq * max_blocks + j / block_size];
        const T* kb = cache + (static_cast<int64_t>(blk) * block_size + (j % block_size)) * tok_stride +
                      static_cast<int64_t>(kvh) * d;
        const T* vb = kb + static_cast<int64_t>(hkv) * d;
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxPerLane; ++i) {
            const int e = lane + i * 32;
            if (i < per && e < d) dot = fmaf(qv[i], Elem<T>::to_f(kb[e]), dot);
        }
        dot = warp_reduce<SumOp>(dot);
        const float nm = fmaxf(m, dot);
        const float corr = __expf(m - nm);
        const float p = __expf(dot - nm);
        l = l * corr + p;
#pragma unroll
        for (int i = 0; i < kMaxPerLane; ++i) {
            const int e = lane + i * 32;
            if (i < per && e < d) acc[i] = fmaf(p, Elem<T>::to_f(vb[e]), acc[i] * corr);
        }
        m = nm;
    }
    if (lane == 0) {
        sm_m[warp] = m;
        sm_l[warp] = l;
    }
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
        const int e = lane + i * 32;
        if (i < per && e < d) sm_o[warp][e] = acc[i];
    }
    __syncthreads();
    float gm = -INFINITY;
    for (int w = 0; w < kWarps; ++w) gm = fmaxf(gm, sm_m[w]);
    float gl = 0.f;
    for (int w = 0; w < kWarps; ++w) gl += (sm_m[w] == -INFINITY) ? 0.f : sm_l[w] * __expf(sm_m[w] - gm);
    const float inv = gl > 0.f ? 1.f / gl : 0.f;
    T* orow = out + static_cast<int64_t>(t) * hq * d + static_cast<int64_t>(h) * d;
    for (int e = threadIdx.x; e < d; e += blockDim.x) {
        float o = 0.f;
        for (int w = 0; w < kWarps; ++w)
            if (sm_m[w] != -INFINITY) o += sm_o[w][e] * __expf(sm_m[w] - gm);
        orow[e] = Elem<T>::from_f(o * inv);
    }
}

}  // namespace pattn
}  // namespace dsb

using namespace dsb;

// q: [tokens, q_stride] with head h at offset h*d (a packed qkv buffer works with q_stride = (hq+2hkv)*d).
DSB_EXPORT int dsb_paged_attention(const void* q, const void* cache, void* out, const int32_t* seq_of,
                                   const int32_t* pos_of, const int32_t* block_table, int tokens, int hq, int hkv, int d,
                                   int q_stride, int block_size, int max_blocks, float scale, int dtype,
                                   cudaStream_t stream)
{
    if (tokens <= 0) return 0;
    if (d > 256 || hq % hkv) return -2;
    dim3 grid(tokens, hq);
    if (dtype == kBF16)
        pattn::paged_attention_kernel<__nv_bfloat16><<<grid, pattn::kWarps * 32, 0, stream>>>(
            (const __nv_bfloat16*)q, (const __nv_bfloat16*)cache, (__nv_bfloat16*)out, seq_of, pos_of, block_table, hq, hkv,
            d, q_stride, block_size, max_blocks, scale);
    else if (dtype == kF16)
        pattn::paged_attention_kernel<__half><<<grid, pattn::kWarps * 32, 0, stream>>>(
            (const __half*)q, (const __half*)cache, (__half*)out, seq_of, pos_of, block_table, hq, hkv, d, q_stride,
            block_size, max_blocks, scale);
    else if (dtype == kF32)
        pattn::paged_attention_kernel<float><<<grid, pattn::kWarps * 32, 0, stream>>>(
            (const float*)q, (const float*)cache, (float*)out, seq_of, pos_of, block_table, hq, hkv, d, q_stride,
            block_size, max_blocks, scale);
    else
        return -1;
    DSB_CHECK_LAUNCH();
    return 0;
}
