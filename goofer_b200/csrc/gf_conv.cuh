// gf_conv.cuh -- N-point complex Stockham FFT (radix 8, plus one radix-2 pass when log2 N = 1 mod 3) for the
// overlap-save form of gaussian_filter1d (/root/reference/GOOFER.py:241-261) when the kernel is long: sigma 441
// (3,529 taps, SillySampler.py:865-880) and the jitter curves' sigma 73.5 / 49 (589 / 393 taps, GOOFER.py:654, 667).
//
// One CTA of N / 8 threads transforms N points in shared memory; thread j owns one radix-8 butterfly per pass.
// The Gaussian is real and even, so its spectrum H is real: two real signal blocks ride in one complex transform
// (z = x1 + i x2; FFT; times H; inverse FFT; Re -> block 1, Im -> block 2), and the inverse transform is the forward
// one on the conjugate.  Everything is __host__ __device__ so that tests/cpu_emul can run the same index arithmetic
// serially (test-only).
#pragma once
#include "gf_hd.h"

template <typename T> struct alignas(2 * sizeof(T)) GfC { T x, y; };
template <typename T> GF_HD GfC<T> gf_c(T x, T y) { GfC<T> r; r.x = x; r.y = y; return r; }
template <typename T> GF_HD GfC<T> gf_cadd(GfC<T> a, GfC<T> b) { return gf_c<T>(a.x + b.x, a.y + b.y); }
template <typename T> GF_HD GfC<T> gf_csub(GfC<T> a, GfC<T> b) { return gf_c<T>(a.x - b.x, a.y - b.y); }
template <typename T> GF_HD GfC<T> gf_cmul(GfC<T> a, GfC<T> b) { return gf_c<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <typename T> GF_HD GfC<T> gf_cnegi(GfC<T> a) { return gf_c<T>(a.y, -a.x); }        // times -i

// padded index of the exchange buffer: one slot skipped per 128 bytes, which makes the stride-8 and stride-64
// stores of the first passes conflict free (a wavefront serves 16 float2 / 8 double2 slots)
template <typename T> GF_HD int gf_cpad(int i) { return i + (i >> (sizeof(T) == 4 ? 4 : 3)); }
template <typename T, int N> struct GfConvBuf { enum { LEN = N + (N >> (sizeof(T) == 4 ? 4 : 3)) }; };

// forward 8-point DFT in registers, natural order in and out
template <typename T> GF_HD void gf_cdft8(GfC<T> *v)
{
    const T h = (T)0.70710678118654752440;
    GfC<T> a0 = gf_cadd(v[0], v[4]), a4 = gf_csub(v[0], v[4]);
    GfC<T> a1 = gf_cadd(v[1], v[5]), a5 = gf_csub(v[1], v[5]);
    GfC<T> a2 = gf_cadd(v[2], v[6]), a6 = gf_csub(v[2], v[6]);
    GfC<T> a3 = gf_cadd(v[3], v[7]), a7 = gf_csub(v[3], v[7]);
    a5 = gf_c<T>((a5.x + a5.y) * h, (a5.y - a5.x) * h);          // times exp(-i pi / 4)
    a6 = gf_cnegi(a6);                                           // times -i
    a7 = gf_c<T>((a7.y - a7.x) * h, -(a7.x + a7.y) * h);         // times exp(-3 i pi / 4)
    // two 4-point DFTs: even outputs from a0..a3, odd outputs from a4..a7
    {
        GfC<T> c0 = gf_cadd(a0, a2), c2 = gf_csub(a0, a2), c1 = gf_cadd(a1, a3), c3 = gf_cnegi(gf_csub(a1, a3));
        v[0] = gf_cadd(c0, c1); v[4] = gf_csub(c0, c1); v[2] = gf_cadd(c2, c3); v[6] = gf_csub(c2, c3);
    }
    {
        GfC<T> c0 = gf_cadd(a4, a6), c2 = gf_csub(a4, a6), c1 = gf_cadd(a5, a7), c3 = gf_cnegi(gf_csub(a5, a7));
        v[1] = gf_cadd(c0, c1); v[5] = gf_csub(c0, c1); v[3] = gf_cadd(c2, c3); v[7] = gf_csub(c2, c3);
    }
}

// log2 N mod 3 must be 0 (4096) or 1 (8192: one radix-2 pass first, the radix-8 passes then start at NS = 2)
template <int N> struct GfConvShape {
    enum { R2 = (N == 8192 || N == 1024 || N == 128) ? 1 : 0, NS0 = R2 ? 2 : 1, NSF = R2 ? 2 : 8 /* first pass with twiddles */ };
};

// Twiddles, laid out per pass and per thread: pass NS needs exp(-2 pi i r k / (8 NS)) for r = 1..7, k = j mod NS.
// Rows [r - 1][k] of pass NS start at NS - NSF (7 (sum of the earlier NS) = NS - NSF), so a warp reads consecutive
// entries for each r -- read from one natural table exp(-2 pi i m / N) the same factors are strided gathers
// (m = r k N / (8 NS)): 16 to 28 cache lines per request, which kept the L1 tag stage busy 60 % of the time and the
// fp64 pipe 11 % (ncu, c3, first version).  7 (N - NSF) / 7 < N entries in all.
template <typename T, int N> static inline void gf_conv_tw_fill(GfC<T> *tw)
{
    const double PI = 3.141592653589793238462643383279502884;
    for (int m = 0; m < N; ++m) tw[m] = gf_c<T>((T)1, (T)0);
    for (int NS = GfConvShape<N>::NSF; NS < N; NS *= 8)
        for (int r = 1; r < 8; ++r)
            for (int k = 0; k < NS; ++k) {
                const double a = -2.0 * PI * (double)(r * k) / (double)(8 * NS);
                tw[(NS - GfConvShape<N>::NSF) + (r - 1) * NS + k] = gf_c<T>((T)cos(a), (T)sin(a));
            }
}

// One radix-8 Stockham pass for thread j in [0, N / 8).  NS = product of the radices of the earlier passes.
// In place: every thread loads, the CTA synchronises, every thread stores.
template <typename T, int N, int NS> GF_HD void gf_conv_pass_load(int j, const GfC<T> *buf, const GfC<T> *tw, GfC<T> *v)
{
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = buf[gf_cpad<T>(j + r * (N / 8))];
    if (NS > 1) {
        const GfC<T> *t = tw + (NS - GfConvShape<N>::NSF) + (j & (NS - 1));
#pragma unroll
        for (int r = 1; r < 8; ++r) v[r] = gf_cmul(v[r], t[(r - 1) * NS]);
    }
    gf_cdft8(v);
}

template <typename T, int N, int NS> GF_HD void gf_conv_pass_store(int j, GfC<T> *buf, const GfC<T> *v)
{
    const int k = j & (NS - 1);
    const int base = (j - k) * 8 + k;
#pragma unroll
    for (int r = 0; r < 8; ++r) buf[gf_cpad<T>(base + r * NS)] = v[r];
}

// The radix-2 pass that comes first when N = 2 * 8^m (no twiddles at NS = 1).  Thread j holds x[j + r N / 8], r = 0..7,
// in registers -- exactly the N / 2 pairs (q, q + N / 2), q = j + t N / 8, t = 0..3 -- and stores the butterflies.
template <typename T, int N> GF_HD void gf_conv_r2_store(int j, GfC<T> *buf, const GfC<T> *v)
{
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int q = j + t * (N / 8);
        buf[gf_cpad<T>(2 * q)] = gf_cadd(v[t], v[t + 4]);
        buf[gf_cpad<T>(2 * q + 1)] = gf_csub(v[t], v[t + 4]);
    }
}

// A whole transform is: thread j starts with x[j + r N / 8] in v (straight from global memory, or from the previous
// transform: its last pass leaves X[j + r N / 8] in the same registers, which is what the spectrum product and the
// first pass of the inverse transform need -- no exchange in between), runs
//     [r2_store, sync, pass_load<NS0>] or [dft8]; pass_store<NS0>; sync; pass_load<8 NS0>; sync; pass_store<8 NS0>; sync; ...
// and ends after pass_load<N / 8> with X[j + r N / 8] in v.  k_conv.cu and tests/cpu_emul/emul.cpp spell that out.

// which jobs take the overlap-save path: long kernels only (the direct sliding window wins below ~300 taps), and
// the block must keep at least half of its points as valid outputs
#define GF_CONV_N32 8192
#define GF_CONV_N64 4096
#ifndef GF_CONV_MIN_RADIUS
#define GF_CONV_MIN_RADIUS 150
#endif
GF_HD int gf_fir_radius(double sigma) { return (int)(4.0 * sigma + 0.5); }
GF_HD bool gf_fir_f32(int in_f64, const void *maxabs, int in_cast_f32) { return !in_f64 && !maxabs && !in_cast_f32; }
GF_HD bool gf_fir_wants_fft(double sigma, bool f32)
{
#if defined(GF_CONV_OFF)
    return false;
#else
    const int radius = gf_fir_radius(sigma);
    return radius >= GF_CONV_MIN_RADIUS && 4 * radius <= (f32 ? GF_CONV_N32 : GF_CONV_N64);
#endif
}
