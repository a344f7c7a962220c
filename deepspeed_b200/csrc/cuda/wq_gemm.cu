// Weight-only-quantised linear for decode-sized inputs:  out[M, N] = x[M, K] @ dequant(Wq)[N, K]^T,  M <= 16.
//
// Role parity: reference inference/v2/kernels/core_ops/cuda_linear (FP6-LLM `QUANT_GEMM_Kernel`, N9a) and the cutlass
// `mixed_gemm` (N9c): weights stay packed in HBM (int8 or int4 + one fp32 scale per group of `group_size` consecutive K
// elements, the layout produced by quant.cu / ops/quantizer) and are dequantised in registers on their way to the FMAs, so a
// decode step streams 1 (int8) or 0.5 (int4) bytes per weight instead of 2.  With M <= 16 the op is pure weight-bandwidth:
// one warp owns an output feature n, its lanes stride over K in 16-byte packets (16 int8 / 32 int4 weights), the M
// activation rows come from L1/L2, partial sums are reduced with shuffles.  Larger M goes through dequantise + tensor cores.
#include "dsb_common.cuh"

namespace dsb {
namespace wq {

template <typename T>
__device__ __forceinline__ void load_x16(const T* p, float* f)  // 16 consecutive activations -> fp32
{
    Elem<T>::unpack(ld_plain(p), f);
    Elem<T>::unpack(ld_plain(p + 8), f + 8);
}

template <typename T, int BITS, int MT>
__global__ void __launch_bounds__(128)
wq_gemv_kernel(const T* __restrict__ x, const int8_t* __restrict__ wq, const float* __restrict__ scales,
               const T* __restrict__ bias, T* __restrict__ out, int M, int N, int K, int group_size)
{
    constexpr int WPP = BITS == 8 ? 16 : 32;  // weights per 16-byte packet
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps = (blockDim.x >> 5) * gridDim.x;
    const int64_t row_bytes = static_cast<int64_t>(K) * BITS / 8;
    for (int n = blockIdx.x * (blockDim.x >> 5) + warp; n < N; n += warps) {
        float acc[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) acc[m] = 0.f;
        const int8_t* wrow = wq + static_cast<int64_t>(n) * row_bytes;
        const int64_t gbase = static_cast<int64_t>(n) * K;
        for (int k0 = lane * WPP; k0 < K; k0 += 32 * WPP) {
            const Vec16 pk = ld_stream(wrow + static_cast<int64_t>(k0) * BITS / 8);
            float wf[WPP];
            if constexpr (BITS == 8) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t u = pk.w[i];
#pragma unroll
                    for (int b = 0; b < 4; ++b) wf[i * 4 + b] = static_cast<float>(static_cast<int8_t>((u >> (8 * b)) & 0xff));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t u = pk.w[i];
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        int v = static_cast<int>((u >> (4 * b)) & 0xf);
                        wf[i * 8 + b] = static_cast<float>(v >= 8 ? v - 16 : v);
                    }
                }
            }
            // group scales: group_size is a multiple of 16, so a packet spans 1 (int8) or at most 2 (int4, gs=16) groups
            const float s0 = scales[(gbase + k0) / group_size];
            const float s1 = (BITS == 4) ? scales[(gbase + k0 + 16) / group_size] : s0;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                if (m < M) {
                    const T* xr = x + static_cast<int64_t>(m) * K + k0;
                    float xf[16];
                    load_x16(xr, xf);
                    float p0 = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) p0 = fmaf(wf[i], xf[i], p0);
                    acc[m] = fmaf(p0, s0, acc[m]);
                    if constexpr (BITS == 4) {
                        load_x16(xr + 16, xf);
                        float p1 = 0.f;
#pragma unroll
                        for (int i = 0; i < 16; ++i) p1 = fmaf(wf[16 + i], xf[i], p1);
                        acc[m] = fmaf(p1, s1, acc[m]);
                    }
                }
            }
        }
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            float v = acc[m];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && m < M) {
                if (bias) v += Elem<T>::to_f(bias[n]);
                out[static_cast<int64_t>(m) * N + n] = Elem<T>::from_f(v);
            }
        }
    }
}

}  // namespace wq
}  // namespace dsb

using namespace dsb;

#define WQ_LAUNCH(TT, BB, MM)                                                                                          \
    wq::wq_gemv_kernel<TT, BB, MM><<<grid, 128, 0, stream>>>((const TT*)x, (const int8_t*)wq, scales, (const TT*)bias, \
                                                              (TT*)out, M, N, K, group_size)
#define WQ_M(TT, BB)                          \
    if (M <= 1) WQ_LAUNCH(TT, BB, 1);         \
    else if (M <= 2) WQ_LAUNCH(TT, BB, 2);    \
    else if (M <= 4) WQ_LAUNCH(TT, BB, 4);    \
    else if (M <= 8) WQ_LAUNCH(TT, BB, 8);    \
    else WQ_LAUNCH(TT, BB, 16);

// x [M, K] (bf16/fp16, contiguous), wq packed [N, K*bits/8], scales fp32 [N*K/group_size], out [M, N].
// Returns -3 when the shape is not eligible (caller falls back to dequantise + GEMM).
DSB_EXPORT int dsb_wq_gemv(const void* x, const void* wq, const float* scales, const void* bias, void* out, int M, int N, int K,
                           int bits, int group_size, int dtype, cudaStream_t stream)
{
    if (M <= 0 || N <= 0) return 0;
    if (M > 16 || (bits != 8 && bits != 4) || group_size % 16 || K % (bits == 8 ? 16 : 32) || K % group_size) return -3;
    if (dtype != kBF16 && dtype != kF16) return -3;
    int grid = (N + 3) / 4;
    const int cap = kSmCountB200 * 16;
    if (grid > cap) grid = cap;
    if (dtype == kBF16) {
        if (bits == 8) { WQ_M(__nv_bfloat16, 8) } else { WQ_M(__nv_bfloat16, 4) }
    } else {
        if (bits == 8) { WQ_M(__half, 8) } else { WQ_M(__half, 4) }
    }
    DSB_CHECK_LAUNCH();
    return 0;
}
