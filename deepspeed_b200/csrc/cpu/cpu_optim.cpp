// Host optimizers for the ZeRO-Offload tier: Adam/AdamW, Lion, Adagrad over flat fp32 buffers with
// AVX-512 / AVX2 / scalar code paths chosen at run time, OpenMP across cores.
//
// Role parity: reference csrc/adam/cpu_adam_impl.cpp + csrc/includes/{cpu_adam.h,simd.h} (N2),
// csrc/lion/cpu_lion_impl.cpp and csrc/adagrad/cpu_adagrad.cpp (N3).  Differences: one flat call
// per shard (the ZeRO arenas are flat), gradients may be fp32/bf16/fp16, and the bf16/fp16 copy of
// the updated parameters is produced in the same pass into a pinned staging buffer that the
// offload engine H2D-copies (the reference does a separate fp32->fp16 copy + H2D per tile).
#include <immintrin.h>
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <string.h>

#define DSB_EXPORT extern "C" __attribute__((visibility("default")))

enum { kF32 = 0, kF16 = 1, kBF16 = 2 };

// ---- scalar helpers -------------------------------------------------------------------------------
static inline float bf16_to_f32(uint16_t h)
{
    uint32_t u = static_cast<uint32_t>(h) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline uint16_t f32_to_bf16(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);  // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return static_cast<uint16_t>(u >> 16);
}
static inline float f16_to_f32(uint16_t h)
{
    const uint32_t s = (h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1f, m = h & 0x3ff, u;
    if (e == 0) {
        if (m == 0) {
            u = s;
        } else {
            e = 127 - 15 + 1;
            while (!(m & 0x400)) {
                m <<= 1;
                --e;
            }
            m &= 0x3ff;
            u = s | (e << 23) | (m << 13);
        }
    } else if (e == 31) {
        u = s | 0x7f800000u | (m << 13);
    } else {
        u = s | ((e + 112) << 23) | (m << 13);
    }
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline uint16_t f32_to_f16(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t s = (u >> 16) & 0x8000u;
    int32_t e = static_cast<int32_t>((u >> 23) & 0xff) - 127 + 15;
    uint32_t m = u & 0x7fffffu;
    if (((u >> 23) & 0xff) == 0xff) return static_cast<uint16_t>(s | 0x7c00u | (m ? 0x200u : 0));
    if (e >= 31) return static_cast<uint16_t>(s | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return static_cast<uint16_t>(s);
        m |= 0x800000u;
        const uint32_t shift = static_cast<uint32_t>(14 - e);
        uint32_t r = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (r & 1))) ++r;
        return static_cast<uint16_t>(s | r);
    }
    uint32_t r = (static_cast<uint32_t>(e) << 10) | (m >> 13);
    const uint32_t rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1))) ++r;
    return static_cast<uint16_t>(s | r);
}

static inline float load_g(const void* g, int dt, int64_t i)
{
    if (dt == kF32) return static_cast<const float*>(g)[i];
    if (dt == kBF16) return bf16_to_f32(static_cast<const uint16_t*>(g)[i]);
    return f16_to_f32(static_cast<const uint16_t*>(g)[i]);
}
static inline void store_o(void* o, int dt, int64_t i, float v)
{
    if (!o) return;
    if (dt == kBF16)
        static_cast<uint16_t*>(o)[i] = f32_to_bf16(v);
    else if (dt == kF16)
        static_cast<uint16_t*>(o)[i] = f32_to_f16(v);
    else
        static_cast<float*>(o)[i] = v;
}

struct AdamH {
    float lr, b1, b2, eps, wd, bc1, bc2, gs;
    int adamw;
};

static void adam_scalar(float* p, const void* g, float* m, float* v, void* out, int64_t lo, int64_t hi, int gdt,
                        int odt, const AdamH& h)
{
    for (int64_t i = lo; i < hi; ++i) {
        float gi = load_g(g, gdt, i) * h.gs;
        float pi = p[i];
        if (!h.adamw) gi += h.wd * pi;
        float mi = h.b1 * m[i] + (1.f - h.b1) * gi;
        float vi = h.b2 * v[i] + (1.f - h.b2) * gi * gi;
        float upd = (mi / h.bc1) / (sqrtf(vi / h.bc2) + h.eps);
        if (h.adamw) upd += h.wd * pi;
        pi -= h.lr * upd;
        p[i] = pi;
        m[i] = mi;
        v[i] = vi;
        store_o(out, odt, i, pi);
    }
}

// ---- AVX-512 ------------------------------------------------------------------------------------------
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) static inline __m512 ld_g512(const void* g, int dt,
                                                                                           int64_t i)
{
    if (dt == kF32) return _mm512_loadu_ps(static_cast<const float*>(g) + i);
    const __m256i h = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(static_cast<const uint16_t*>(g) + i));
    if (dt == kBF16) return _mm512_castsi512_ps(_mm512_slli_epi32(_mm512_cvtepu16_epi32(h), 16));
    return _mm512_cvtph_ps(h);
}
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) static inline void st_o512(void* o, int dt, int64_t i,
                                                                                         __m512 x)
{
    if (!o) return;
    if (dt == kF32) {
        _mm512_storeu_ps(static_cast<float*>(o) + i, x);
    } else if (dt == kBF16) {
        __m512i u = _mm512_castps_si512(x);
        __m512i lsb = _mm512_and_si512(_mm512_srli_epi32(u, 16), _mm512_set1_epi32(1));
        u = _mm512_add_epi32(u, _mm512_add_epi32(lsb, _mm512_set1_epi32(0x7fff)));
        __m256i r = _mm512_cvtepi32_epi16(_mm512_srli_epi32(u, 16));
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(static_cast<uint16_t*>(o) + i), r);
    } else {
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(static_cast<uint16_t*>(o) + i),
                            _mm512_cvtps_ph(x, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
    }
}

__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) static void adam_avx512(
    float* p, const void* g, float* m, float* v, void* out, int64_t lo, int64_t hi, int gdt, int odt, const AdamH& h)
{
    const __m512 b1 = _mm512_set1_ps(h.b1), b1m = _mm512_set1_ps(1.f - h.b1);
    const __m512 b2 = _mm512_set1_ps(h.b2), b2m = _mm512_set1_ps(1.f - h.b2);
    const __m512 eps = _mm512_set1_ps(h.eps), wd = _mm512_set1_ps(h.wd), gs = _mm512_set1_ps(h.gs);
    const __m512 ibc1 = _mm512_set1_ps(1.f / h.bc1), ibc2 = _mm512_set1_ps(1.f / h.bc2);
    const __m512 nlr = _mm512_set1_ps(-h.lr);
    int64_t i = lo;
    for (; i + 16 <= hi; i += 16) {
        __m512 gi = _mm512_mul_ps(ld_g512(g, gdt, i), gs);
        __m512 pi = _mm512_loadu_ps(p + i);
        if (!h.adamw) gi = _mm512_fmadd_ps(wd, pi, gi);
        __m512 mi = _mm512_fmadd_ps(b1, _mm512_loadu_ps(m + i), _mm512_mul_ps(b1m, gi));
        __m512 vi = _mm512_fmadd_ps(b2, _mm512_loadu_ps(v + i), _mm512_mul_ps(b2m, _mm512_mul_ps(gi, gi)));
        __m512 den = _mm512_add_ps(_mm512_sqrt_ps(_mm512_mul_ps(vi, ibc2)), eps);
        __m512 upd = _mm512_div_ps(_mm512_mul_ps(mi, ibc1), den);
        if (h.adamw) upd = _mm512_fmadd_ps(wd, pi, upd);
        pi = _mm512_fmadd_ps(nlr, upd, pi);
        _mm512_storeu_ps(p + i, pi);
        _mm512_storeu_ps(m + i, mi);
        _mm512_storeu_ps(v + i, vi);
        st_o512(out, odt, i, pi);
    }
    if (i < hi) adam_scalar(p, g, m, v, out, i, hi, gdt, odt, h);
}

// ---- AVX2 -----------------------------------------------------------------------------------------------
__attribute__((target("avx2,fma,f16c"))) static inline __m256 ld_g256(const void* g, int dt, int64_t i)
{
    if (dt == kF32) return _mm256_loadu_ps(static_cast<const float*>(g) + i);
    const __m128i h = _mm_loadu_si128(reinterpret_cast<const __m128i*>(static_cast<const uint16_t*>(g) + i));
    if (dt == kBF16) return _mm256_castsi256_ps(_mm256_slli_epi32(_mm256_cvtepu16_epi32(h), 16));
    return _mm256_cvtph_ps(h);
}
__attribute__((target("avx2,fma,f16c"))) static inline void st_o256(void* o, int dt, int64_t i, __m256 x)
{
    if (!o) return;
    if (dt == kF32) {
        _mm256_storeu_ps(static_cast<float*>(o) + i, x);
    } else if (dt == kBF16) {
        __m256i u = _mm256_castps_si256(x);
        __m256i lsb = _mm256_and_si256(_mm256_srli_epi32(u, 16), _mm256_set1_epi32(1));
        u = _mm256_add_epi32(u, _mm256_add_epi32(lsb, _mm256_set1_epi32(0x7fff)));
        u = _mm256_srli_epi32(u, 16);
        __m128i lo = _mm256_castsi256_si128(u), hi = _mm256_extracti128_si256(u, 1);
        _mm_storeu_si128(reinterpret_cast<__m128i*>(static_cast<uint16_t*>(o) + i), _mm_packus_epi32(lo, hi));
    } else {
        _mm_storeu_si128(reinterpret_cast<__m128i*>(static_cast<uint16_t*>(o) + i),
                         _mm256_cvtps_ph(x, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
    }
}

__attribute__((target("avx2,fma,f16c"))) static void adam_avx2(float* p, const void* g, float* m, float* v, void* out,
                                                               int64_t lo, int64_t hi, int gdt, int odt,
                                                               const AdamH& h)
{
    const __m256 b1 = _mm256_set1_ps(h.b1), b1m = _mm256_set1_ps(1.f - h.b1);
    const __m256 b2 = _mm256_set1_ps(h.b2), b2m = _mm256_set1_ps(1.f - h.b2);
    const __m256 eps = _mm256_set1_ps(h.eps), wd = _mm256_set1_ps(h.wd), gs = _mm256_set1_ps(h.gs);
    const __m256 ibc1 = _mm256_set1_ps(1.f / h.bc1), ibc2 = _mm256_set1_ps(1.f / h.bc2);
    const __m256 nlr = _mm256_set1_ps(-h.lr);
    int64_t i = lo;
    for (; i + 8 <= hi; i += 8) {
        __m256 gi = _mm256_mul_ps(ld_g256(g, gdt, i), gs);
        __m256 pi = _mm256_loadu_ps(p + i);
        if (!h.adamw) gi = _mm256_fmadd_ps(wd, pi, gi);
        __m256 mi = _mm256_fmadd_ps(b1, _mm256_loadu_ps(m + i), _mm256_mul_ps(b1m, gi));
        __m256 vi = _mm256_fmadd_ps(b2, _mm256_loadu_ps(v + i), _mm256_mul_ps(b2m, _mm256_mul_ps(gi, gi)));
        __m256 den = _mm256_add_ps(_mm256_sqrt_ps(_mm256_mul_ps(vi, ibc2)), eps);
        __m256 upd = _mm256_div_ps(_mm256_mul_ps(mi, ibc1), den);
        if (h.adamw) upd = _mm256_fmadd_ps(wd, pi, upd);
        pi = _mm256_fmadd_ps(nlr, upd, pi);
        _mm256_storeu_ps(p + i, pi);
        _mm256_storeu_ps(m + i, mi);
        _mm256_storeu_ps(v + i, vi);
        st_o256(out, odt, i, pi);
    }
    if (i < hi) adam_scalar(p, g, m, v, out, i, hi, gdt, odt, h);
}

static int simd_level()
{
    static int level = -1;
    if (level >= 0) return level;
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
        __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq"))
        level = 2;
    else if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma") && __builtin_cpu_supports("f16c"))
        level = 1;
    else
        level = 0;
    return level;
}

DSB_EXPORT int dsb_cpu_simd_level() { return simd_level(); }

// Threads of the host optimizers.  Launchers (torchrun) export OMP_NUM_THREADS=1 for every rank, which would leave the
// offload tier's Adam single-threaded (measured: 0.76 G parameters/s per rank); the runtime instead gives each rank its
// share of the cores (``ops/adam/cpu_adam.py: configure_threads``): all `parallel for` regions below use this count.
static int g_threads = 0;
DSB_EXPORT void dsb_cpu_set_threads(int n) { g_threads = n > 0 ? n : 0; }
DSB_EXPORT int dsb_cpu_get_threads() { return g_threads > 0 ? g_threads : omp_get_max_threads(); }
static inline int nthreads() { return g_threads > 0 ? g_threads : omp_get_max_threads(); }

// Tile so that each OpenMP task streams a cache-friendly block.
static const int64_t kTile = 1 << 16;

DSB_EXPORT int dsb_cpu_adam(float* p, const void* g, float* m, float* v, void* out, int64_t n, int gdt, int odt,
                            float lr, float b1, float b2, float eps, float wd, float bc1, float bc2, int adamw,
                            float grad_scale)
{
    if (n <= 0) return 0;
    if (gdt < 0 || gdt > 2 || odt < 0 || odt > 2) return -1;
    const AdamH h{lr, b1, b2, eps, wd, bc1, bc2, grad_scale, adamw};
    const int lvl = simd_level();
    const int64_t ntiles = (n + kTile - 1) / kTile;
#pragma omp parallel for schedule(static) num_threads(nthreads())
    for (int64_t t = 0; t < ntiles; ++t) {
        const int64_t lo = t * kTile, hi = (lo + kTile < n) ? lo + kTile : n;
        if (lvl == 2)
            adam_avx512(p, g, m, v, out, lo, hi, gdt, odt, h);
        else if (lvl == 1)
            adam_avx2(p, g, m, v, out, lo, hi, gdt, odt, h);
        else
            adam_scalar(p, g, m, v, out, lo, hi, gdt, odt, h);
    }
    return 0;
}

DSB_EXPORT int dsb_cpu_lion(float* p, const void* g, float* m, void* out, int64_t n, int gdt, int odt, float lr,
                            float b1, float b2, float wd, float grad_scale)
{
    if (n <= 0) return 0;
    const float decay = 1.f - lr * wd;
#pragma omp parallel for schedule(static) num_threads(nthreads())
    for (int64_t t = 0; t < (n + kTile - 1) / kTile; ++t) {
        const int64_t lo = t * kTile, hi = (lo + kTile < n) ? lo + kTile : n;
#pragma omp simd
        for (int64_t i = lo; i < hi; ++i) {
            const float gi = load_g(g, gdt, i) * grad_scale;
            const float c = b1 * m[i] + (1.f - b1) * gi;
            const float sgn = (c > 0.f) ? 1.f : ((c < 0.f) ? -1.f : 0.f);
            const float pi = p[i] * decay - lr * sgn;
            m[i] = b2 * m[i] + (1.f - b2) * gi;
            p[i] = pi;
            store_o(out, odt, i, pi);
        }
    }
    return 0;
}

DSB_EXPORT int dsb_cpu_adagrad(float* p, const void* g, float* hsum, void* out, int64_t n, int gdt, int odt, float lr,
                               float eps, float wd, float grad_scale)
{
    if (n <= 0) return 0;
#pragma omp parallel for schedule(static) num_threads(nthreads())
    for (int64_t t = 0; t < (n + kTile - 1) / kTile; ++t) {
        const int64_t lo = t * kTile, hi = (lo + kTile < n) ? lo + kTile : n;
#pragma omp simd
        for (int64_t i = lo; i < hi; ++i) {
            float gi = load_g(g, gdt, i) * grad_scale;
            float pi = p[i];
            gi += wd * pi;
            const float hv = hsum[i] + gi * gi;
            pi -= lr * gi / (sqrtf(hv) + eps);
            hsum[i] = hv;
            p[i] = pi;
            store_o(out, odt, i, pi);
        }
    }
    return 0;
}

// fp32 -> bf16/fp16 bulk conversion (used to stage updated params for the H2D copy).
DSB_EXPORT int dsb_cpu_cast(const float* src, void* dst, int64_t n, int odt)
{
    if (odt != kBF16 && odt != kF16) return -1;
#pragma omp parallel for schedule(static) num_threads(nthreads())
    for (int64_t t = 0; t < (n + kTile - 1) / kTile; ++t) {
        const int64_t lo = t * kTile, hi = (lo + kTile < n) ? lo + kTile : n;
        for (int64_t i = lo; i < hi; ++i) store_o(dst, odt, i, src[i]);
    }
    return 0;
}

// sum of squares of a flat fp32 buffer (host-side grad norm under offload).
DSB_EXPORT double dsb_cpu_sumsq(const float* x, int64_t n)
{
    double acc = 0.0;
#pragma omp parallel for reduction(+ : acc) schedule(static) num_threads(nthreads())
    for (int64_t i = 0; i < n; ++i) acc += static_cast<double>(x[i]) * x[i];
    return acc;
}
