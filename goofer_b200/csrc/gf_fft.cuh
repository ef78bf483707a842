// gf_fft.cuh -- 1024-point real FFT / inverse real FFT as a 512-point complex radix-8 Stockham
// transform plus the even/odd split, written per *thread* so that 64 threads cooperate on one
// transform through shared memory (three radix-8 passes, two exchanges).
//
// Replaces numpy's pocketfft calls in /root/reference/GOOFER.py:370 (np.fft.rfft, float32 in =>
// complex64 out on numpy >= 2) and GOOFER.py:400 (np.fft.irfft, n = 1024).
//
// Every function is __host__ __device__ so that tests/cpu_emul can drive the very same index
// arithmetic with a serial loop over "threads" (test-only; the product never runs it on the CPU).
#pragma once
#include "gf_hd.h"

#define GF_FFT_N 512            // complex points
#define GF_FFT_THREADS 64       // threads per transform (one radix-8 butterfly each per pass)

// swizzled index for the 512-float2 exchange buffer.  A shared-memory wavefront serves 16 float2 slots; the
// Stockham passes touch the buffer with stride 1 (loads, last store), stride 8 (first store) and 8-runs 64 apart
// (second store).  XOR-ing the low four index bits with bits 3..6 makes all three patterns conflict free:
//   stride 1 : 16 consecutive i -> {0..7} ^ c and {8..15} ^ (c + 1), disjoint halves
//   stride 8 : i = 8 j + r     -> low bits r ^ (j & 7), bit 3 = (j & 1) ^ (j >> 3 & 1): a bijection of j & 15
//   8-runs   : i = 64 g + k + 8 r -> low bits k ^ r, bit 3 = (r ^ g) & 1: the two runs of a half-warp differ in bit 3
GF_HD int gf_fpad(int i) { return i ^ ((i >> 3) & 15); }
#define GF_FFT_BUF (512 + 20)   // float2 elements per padded transform buffer; 532 mod 16 == 4 puts consecutive
                                // transforms 8 banks apart, so (bin, frame)-interleaved accesses do not collide

GF_HD float2 gf_cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
GF_HD float2 gf_cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
GF_HD float2 gf_csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
GF_HD float2 gf_conj(float2 a) { return make_float2(a.x, -a.y); }

// multiply by -i (forward) or +i (inverse)
template <bool INV> GF_HD float2 gf_rot(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

template <bool INV> GF_HD void gf_dft4(float2 b0, float2 b1, float2 b2, float2 b3, float2 &y0, float2 &y1, float2 &y2, float2 &y3)
{
    float2 c0 = gf_cadd(b0, b2), c2 = gf_csub(b0, b2), c1 = gf_cadd(b1, b3), c3 = gf_rot<INV>(gf_csub(b1, b3));
    y0 = gf_cadd(c0, c1); y2 = gf_csub(c0, c1); y1 = gf_cadd(c2, c3); y3 = gf_csub(c2, c3);
}

// in-register 8-point DFT, natural order in and out
template <bool INV> GF_HD void gf_dft8(float2 *v)
{
    const float h = 0.70710678118654752440f;
    float2 a0 = gf_cadd(v[0], v[4]), a4 = gf_csub(v[0], v[4]);
    float2 a1 = gf_cadd(v[1], v[5]), a5 = gf_csub(v[1], v[5]);
    float2 a2 = gf_cadd(v[2], v[6]), a6 = gf_csub(v[2], v[6]);
    float2 a3 = gf_cadd(v[3], v[7]), a7 = gf_csub(v[3], v[7]);
    if (INV) {
        a5 = make_float2((a5.x - a5.y) * h, (a5.x + a5.y) * h);
        a7 = make_float2(-(a7.x + a7.y) * h, (a7.x - a7.y) * h);
    } else {
        a5 = make_float2((a5.x + a5.y) * h, (a5.y - a5.x) * h);
        a7 = make_float2((a7.y - a7.x) * h, -(a7.x + a7.y) * h);
    }
    a6 = gf_rot<INV>(a6);
    gf_dft4<INV>(a0, a1, a2, a3, v[0], v[2], v[4], v[6]);
    gf_dft4<INV>(a4, a5, a6, a7, v[1], v[3], v[5], v[7]);
}

// ---- twiddles ---------------------------------------------------------------------------------
// Thread j of a transform needs, for every radix-8 pass after the first, the seven factors w^r (r = 1..7) of ITS OWN
// butterfly: w = exp(-2 pi i k / 64), k = j & 7 (pass NS = 8) and w = exp(-2 pi i j / 512) (pass NS = 64).  Read from
// the natural table exp(-2 pi i m / 512) those are strided gathers (m = 8 k r / j r): on the frame kernel they were
// 32 % of all shared-memory wavefronts and 91 % of its bank conflicts (ncu source page, round 1).  The table below
// is laid out per thread instead: pairs (w^2q, w^(2q+1)) of thread j sit in one 16-byte slot, consecutive threads
// in consecutive slots, so a warp reads a contiguous 512-byte run per pair (16-byte loads, no conflicts):
//   NS = 8 : twl[(q * 8 + k) * 2 + (r & 1)]            k = j & 7, q = r >> 1           (64 entries)
//   NS = 64: twl[64 + (q * 64 + j) * 2 + (r & 1)]                                       (512 entries)
#define GF_TWL_N (64 + 512)

static inline void gf_twl_fill(float2 *twl)     // host: fp64 cos / sin rounded once, like the natural table it replaces
{
    const double PI = 3.141592653589793238462643383279502884;
    for (int r = 0; r < 8; ++r) {
        for (int k = 0; k < 8; ++k) {
            const double a = -2.0 * PI * (double)(8 * k * r) / 512.0;
            twl[((r >> 1) * 8 + k) * 2 + (r & 1)] = make_float2((float)cos(a), (float)sin(a));
        }
        for (int j = 0; j < 64; ++j) {
            const double a = -2.0 * PI * (double)(j * r) / 512.0;
            twl[64 + ((r >> 1) * 64 + j) * 2 + (r & 1)] = make_float2((float)cos(a), (float)sin(a));
        }
    }
}

// the pair (w^2q, w^(2q+1)) of one thread: one 16-byte load on the device
GF_HD void gf_tw_pair(const float2 *p, float2 &a, float2 &b)
{
#if defined(__CUDA_ARCH__)
    const float4 q = *reinterpret_cast<const float4 *>(p);
    a = make_float2(q.x, q.y); b = make_float2(q.z, q.w);
#else
    a = p[0]; b = p[1];
#endif
}

// One Stockham radix-8 pass of the 512-point transform for thread j (0..63).
//   NS = 1, 8, 64 for the three passes.  twl: the per-thread twiddle table above (16-byte aligned).
// Reads buf[j + 64 r]; the caller must barrier between gf_fft_pass_load and gf_fft_pass_store when
// running in place.
template <bool INV, int NS> GF_HD void gf_fft_pass_load(int j, const float2 *buf, const float2 *twl, float2 *v)
{
    // gf_fpad(j + 64 r) == 64 r + (gf_fpad(j) ^ (8 (r & 1))) for j < 64: two bases, immediate offsets
    const float2 *src0 = buf + gf_fpad(j), *src1 = buf + (gf_fpad(j) ^ 8);
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = (r & 1) ? src1[64 * r] : src0[64 * r];
    if (NS > 1) {
        const float2 *tw = (NS == 8) ? twl + 2 * (j & 7) : twl + 64 + 2 * j;
        const int qs = (NS == 8) ? 16 : 128;      // float2 elements between the pairs of a thread
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float2 w0, w1;
            gf_tw_pair(tw + q * qs, w0, w1);
            if (INV) { w0.y = -w0.y; w1.y = -w1.y; }
            if (q > 0) v[2 * q] = gf_cmul(v[2 * q], w0);
            v[2 * q + 1] = gf_cmul(v[2 * q + 1], w1);
        }
    }
    gf_dft8<INV>(v);
}

template <int NS> GF_HD void gf_fft_pass_store(int j, float2 *buf, const float2 *v)
{
    // swizzled index of j0 + r NS with j0 = 8 (j - k) + k, k = j mod NS
    if (NS == 1) {
        // i = 8 j + r: low three bits r ^ (j & 7), bit 3 of 8 j flipped by bit 3 of j
        float2 *dst = buf + ((8 * j) ^ (j & 8));
        const int x = j & 7;
#pragma unroll
        for (int r = 0; r < 8; ++r) dst[r ^ x] = v[r];
    } else if (NS == 8) {
        // i = 64 g + k + 8 r (g = j >> 3, k = j & 7): low three bits k ^ r, bit 3 = (r ^ g) & 1, bits 4..5 = r >> 1
        const int g = j >> 3, k = j & 7;
        float2 *dst = buf + 64 * g;
        const int gx = (g & 1) << 3;
#pragma unroll
        for (int r = 0; r < 8; ++r) dst[(((k ^ r) + ((r >> 1) << 4)) ^ ((r & 1) << 3)) ^ gx] = v[r];
    } else {
        float2 *dst0 = buf + gf_fpad(j), *dst1 = buf + (gf_fpad(j) ^ 8);
#pragma unroll
        for (int r = 0; r < 8; ++r) { if (r & 1) dst1[64 * r] = v[r]; else dst0[64 * r] = v[r]; }
    }
}

// ---- even/odd split --------------------------------------------------------------------------
// tw1024[k] = exp(-2 pi i k / 1024), k <= 512.
//
// Forward: z[n] = x[2n] + i x[2n+1], Z = FFT512(z).  For the bin pair (k, 512 - k), 0 <= k <= 256:
GF_HD void gf_rfft_split(float2 Zk, float2 Zm /* Z[(512-k) & 511] */, float2 w /* tw1024[k] */, float2 &Xk, float2 &Xm)
{
    // E = (Zk + conj Zm)/2, O = -i (Zk - conj Zm)/2 ; X[k] = E + w O ; X[512-k] = conj(E - w O)
    float2 E = make_float2(0.5f * (Zk.x + Zm.x), 0.5f * (Zk.y - Zm.y));
    float2 D = make_float2(0.5f * (Zk.x - Zm.x), 0.5f * (Zk.y + Zm.y));
    float2 O = make_float2(D.y, -D.x);
    float2 wO = gf_cmul(w, O);
    Xk = gf_cadd(E, wO);
    Xm = gf_conj(gf_csub(E, wO));
}

// Inverse: from the Hermitian half-spectrum pair (X[k], X[512-k]) build Z[k], Z[512-k] such that
// IFFT512(Z) (unnormalised) * (1/512) = x[2n] + i x[2n+1] with numpy irfft scaling (1/1024 overall).
// The imaginary parts of X[0] and X[512] must be zeroed by the caller (pocketfft c2r ignores them).
GF_HD void gf_irfft_merge(float2 Xk, float2 Xm, float2 w /* tw1024[k] */, float2 &Zk, float2 &Zm)
{
    // E = (Xk + conj Xm)/2 ; O = conj(w) (Xk - conj Xm)/2 ; Z[k] = E + i O ; Z[512-k] = conj(E - i O) ... (k != 512-k)
    float2 E = make_float2(0.5f * (Xk.x + Xm.x), 0.5f * (Xk.y - Xm.y));
    float2 D = make_float2(0.5f * (Xk.x - Xm.x), 0.5f * (Xk.y + Xm.y));
    float2 O = gf_cmul(gf_conj(w), D);
    float2 iO = make_float2(-O.y, O.x);
    Zk = gf_cadd(E, iO);
    Zm = gf_conj(gf_csub(E, iO));
}
