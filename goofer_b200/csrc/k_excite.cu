// k_excite.cu -- excitation: target f0 curve, bit-faithful glottal pulse onsets, LF pulse rendering.
//
//   gf_f0_kernel      target pitch + vocal-fry f0 + per-pass f0      SillySampler.py:836-855, 884-934,
//                     1038-1041 (su), 1062-1067 (sj); GOOFER.py:989, 1069-1071 (sh jitter)
//   gf_walk_kernel    pulse_train_numba's sequential fp64 phase walk GOOFER.py:479-493, 534-554
//   gf_pulse_kernel   LF pulse tables + overlap-add as a gather       GOOFER.py:495-552
//
// The onset walk is the one place where rounding order is part of the result (SURVEY.md section 0
// fact 4): the phase is accumulated sample by sample in fp64, left to right, exactly like the
// reference; only the crossing *detection* is done in parallel (floor of the running totals).
#include "gf_device.cuh"
#include "gf_maps.cuh"
#include <climits>
#include <cstdlib>

#define GF_PI_D 3.141592653589793
#define GF_RCP12 0.083333333333333329          // RN(1 / 12)

// a / b, correctly rounded, from the correctly rounded reciprocal y = RN(1 / b) (Markstein): q = RN(a y),
// r = a - b q (exact in an FMA), RN(q + r y).  Exact whenever b's significand is not all ones (b is a sample rate
// or a small constant here) and nothing over- or underflows (a is a float widened to double); three instructions
// instead of the ~35 of the general fp64 division.  Checked against exact rational arithmetic on 5.8e5 values.
__device__ __forceinline__ double gf_div_by(double a, double b, double y)
{
    const double q = __dmul_rn(a, y);
    const double r = __fma_rn(-q, b, a);
    return __fma_rn(r, y, q);
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double gf_midi_at(const GfNotePlan &pl, const float *__restrict__ bend, int i)
{
    // SillySampler.py:836-853; interp1d == np.interp here because the query is clipped to the knots
    const double add = (double)pl.pitch_midi;
    const double tadd = pl.t_cents ? ((double)pl.t_cents / 100.0) : 0.0;
    const double rcp100 = 0.01;                                    // RN(1 / 100)
    auto semi = [&](int k) {
        double s = gf_div_by((double)bend[k], 100.0, rcp100) + add;
        if (pl.t_cents) s = s + tadd;
        return s;
    };
    if (pl.bend_len == 1) return semi(0);
    const double dt = 60.0 / (pl.tempo * 96.0);
    const int last = pl.bend_len - 1;
    double x = gf_div_by((double)i, (double)pl.sr, __drcp_rn((double)pl.sr));
    const double xl = (double)last * dt;
    x = fmin(fmax(x, 0.0), xl);
    if (x >= xl) return semi(last);
    int j = (int)(x * __drcp_rn(dt));                              // bracket estimate, fixed up below
    if (j > last - 1) j = last - 1;
    while (j > 0 && (double)j * dt > x) --j;
    while (j < last - 1 && (double)(j + 1) * dt <= x) ++j;
    const double x0 = (double)j * dt, x1 = (double)(j + 1) * dt;
    const double y0 = semi(j), y1 = semi(j + 1);
    if (x0 == x) return y0;
    const double slope = __ddiv_rn(y1 - y0, x1 - x0);
    return __dadd_rn(__dmul_rn(slope, x - x0), y0);
}

// gf_midi_at with its operands passed by value (nothing is re-read from the plan record inside the sample loop)
__device__ __forceinline__ double gf_midi_at_fast(const float *__restrict__ bend, int bend_len, double add, int t_cents,
                                                  double tempo, int sr, int i)
{
    const double tadd = t_cents ? ((double)t_cents / 100.0) : 0.0;
    auto semi = [&](int k) {
        double s = gf_div_by((double)bend[k], 100.0, 0.01) + add;
        if (t_cents) s = s + tadd;
        return s;
    };
    const double dt = 60.0 / (tempo * 96.0);
    const int last = bend_len - 1;
    double x = gf_div_by((double)i, (double)sr, __drcp_rn((double)sr));
    const double xl = (double)last * dt;
    x = fmin(fmax(x, 0.0), xl);
    if (x >= xl) return semi(last);
    int j = (int)(x * __drcp_rn(dt));                              // bracket estimate, fixed up below
    if (j > last - 1) j = last - 1;
    while (j > 0 && (double)j * dt > x) --j;
    while (j < last - 1 && (double)(j + 1) * dt <= x) ++j;
    const double x0 = (double)j * dt, x1 = (double)(j + 1) * dt;
    const double y0 = semi(j), y1 = semi(j + 1);
    if (x0 == x) return y0;
    const double slope = __ddiv_rn(y1 - y0, x1 - x0);
    return __dadd_rn(__dmul_rn(slope, x - x0), y0);
}

// smooth_mask_ds at sample i.  The lerp reads ms_short[j], ms_short[j + 1] with i / 4 - 1.75 <= j <= i / 4 (the ratio
// (M - 1) / (N - 1) is at most 1 / 4, the f32 abscissae move the bracket by at most one): when the four values
// ms_short[i / 4 - 2 .. i / 4 + 1] are equal -- everywhere except near voicing edges -- the result is that value
// (np.interp with slope 0), without the fp64 abscissae and the division of the general path.
__device__ __forceinline__ float gf_ms_at_fast(const float *__restrict__ s, int M, int i, int N)
{
    const int q = i >> 2;
    if (q >= 2 && q + 1 < M) {
        const float a = s[q - 2], b = s[q - 1], c = s[q], d = s[q + 1];
        if (a == b && b == c && c == d) return c;
    }
    return gf_ms_at(s, M, i, N);
}

// smooth_mask_ds (GOOFER.py:564-569) for every sample of a note + per 256-sample block: is the smoothed mask exactly 1
// everywhere?  (Where it is, aper_uv * (1 - mask) vanishes identically: the block is flagged, not stored, and the frame
// kernel skips the unvoiced stream.)  Four consecutive samples per thread -- they share the four decimated values of the
// fast path: 4 loads per 4 samples instead of 16 -- one flag per two warps.  Called by every thread of a 256-thread CTA.
__device__ __forceinline__ void gf_ms_loop(const float *__restrict__ mshort, int M, int n, unsigned char *__restrict__ ms_one, float *__restrict__ out_ms)
{
    __shared__ int s_one[2][8];
    const int warp = threadIdx.x >> 5;
    int it = 0;
    for (int base = blockIdx.x * 1024; base < n; base += gridDim.x * 1024, it ^= 1) {
        const int i = base + 4 * threadIdx.x;
        float msv[4] = {1.0f, 1.0f, 1.0f, 1.0f};
        if (i < n) {
            const int q = i >> 2;                      // i is a multiple of 4
            bool fast = false;
            if (q >= 2 && q + 1 < M) {
                const float a = mshort[q - 2], b = mshort[q - 1], c = mshort[q], d = mshort[q + 1];
                if (a == b && b == c && c == d) { fast = true; msv[0] = msv[1] = msv[2] = msv[3] = c; }
            }
            if (!fast) {
#pragma unroll
                for (int k = 0; k < 4; ++k) if (i + k < n) msv[k] = gf_ms_at(mshort, M, i + k, n);
            }
        }
        const bool one = msv[0] == 1.0f && msv[1] == 1.0f && msv[2] == 1.0f && msv[3] == 1.0f;
        const int w_one = __all_sync(0xffffffffu, one);
        if ((threadIdx.x & 31) == 0) s_one[it][warp] = w_one;
        __syncthreads();                               // s_one[it] is rewritten two iterations later: one barrier per round
        const int all_one = s_one[it][warp & ~1] & s_one[it][warp | 1];
        if ((threadIdx.x & 63) == 0 && i < n) ms_one[i >> 8] = (unsigned char)all_one;
        if (i < n && !all_one) {
            if (i + 4 <= n) *reinterpret_cast<float4 *>(out_ms + i) = make_float4(msv[0], msv[1], msv[2], msv[3]);
            else for (int k = 0; k < 4; ++k) if (i + k < n) out_ms[i + k] = msv[k];
        }
    }
}

#ifndef GF_F0_TICK_MAJOR
#define GF_F0_TICK_MAJOR 1
#endif
#ifndef GF_F0_CTAS
#define GF_F0_CTAS 8                // register cap 32 (with spills): the kernel waits on loads (long scoreboard 8.9 per issue), resident warps pay:
                                    // 0.75 ms uncapped (80 registers) -> 0.60 (cap 64) -> 0.47 (cap 40) -> 0.41 ms (cap 32)
#endif
__global__ void __launch_bounds__(256, GF_F0_CTAS)
gf_f0_kernel(const GfNotePlan *__restrict__ plans, const GfNoteDev *__restrict__ notes, const GfPassDev *__restrict__ passes,
             const GfSourceDev *__restrict__ srcs, const float *__restrict__ bend_all, const double *__restrict__ normals,
             const float *__restrict__ f0_curves)
{
    const GfNotePlan &pl = plans[blockIdx.y];
    const GfNoteDev nd = notes[blockIdx.y];
    const float *mask_src = srcs[pl.src].mask;
    const float *bend = bend_all + pl.bend_off;
    const int n = pl.n_total;
    double sh_scale = 0.0;
    if (pl.f0_jitter) sh_scale = nd.noteScal[GF_NS_SHMAX];
    // 440 * 2 ** ((midi - 69) / 12): exp2 instead of pow (exact for the integer exponents of the A notes, where
    // the pulse onsets sit on rounding ties; elsewhere the two differ by at most an ulp of the fp64 value)
    const bool flat = pl.bend_len == 1;
    const double midi_flat = flat ? gf_midi_at(pl, bend, 0) : 0.0;
    const double hz_flat = flat ? 440.0 * exp2((midi_flat - 69.0) / 12.0) : 0.0;
    const int M = (n + 3) / 4;
    // per-pass outputs, loaded once (they used to be re-read from the pass records for every sample)
    const int npass = pl.n_passes;
    float *pf0[GF_MAX_PASSES];
    int pkind[GF_MAX_PASSES];
#pragma unroll
    for (int p = 0; p < GF_MAX_PASSES; ++p) {
        pf0[p] = (p < npass) ? passes[nd.pass0 + p].f0 : nullptr;
        pkind[p] = (p < npass) ? passes[nd.pass0 + p].kind : -1;
    }
    // ---- common case (c1 / c2 / c5 notes): one pass, no velocity stretch, fry, jitter or pitch dynamics ----
    const float *__restrict__ f0_direct = pl.f0_off >= 0 ? f0_curves + pl.f0_off : nullptr;      // direct gf.synthesize call: f0_interp as given
    const bool simple = npass == 1 && !pl.vel_active && pl.fry_L <= 0 && !pl.f0_jitter && !nd.f0n && !nd.pd_in && !f0_direct;
    if (simple) {
        float *__restrict__ out_f0 = pf0[0];
        float *__restrict__ out_ms = nd.ms;
        const float *__restrict__ vm = nd.vm;
        const float *__restrict__ mshort = nd.ms_short;
        unsigned char *__restrict__ ms_one = nd.ms_one;
        const int sr_i = pl.sr;
#if GF_F0_TICK_MAJOR
        gf_ms_loop(mshort, M, n, ms_one, out_ms);
        // ---- f0 = mask * 440 * 2^((midi - 69) / 12) ----
        if (flat) {
            for (int i = 4 * (blockIdx.x * blockDim.x + threadIdx.x); i < n; i += 4 * gridDim.x * blockDim.x) {
                if (i + 4 <= n) {
                    const float4 m4 = *reinterpret_cast<const float4 *>(vm + i);
                    *reinterpret_cast<float4 *>(out_f0 + i) = make_float4((float)((double)m4.x * hz_flat), (float)((double)m4.y * hz_flat),
                                                                          (float)((double)m4.z * hz_flat), (float)((double)m4.w * hz_flat));
                } else for (int k = i; k < n; ++k) out_f0[k] = (float)((double)vm[k] * hz_flat);
            }
            return;
        }
        // Tick-major: the pitch curve is piece-wise linear over the pitch-bend ticks (SillySampler.py:836-853; 5.2 ms =
        // 230 samples at 120 bpm), and everything but the final lerp and the exp2 depends on the tick alone -- the bracket
        // search, two knot values (a division by 100 each) and the slope (a division).  A warp takes a tick: every lane
        // computes the tick's constants once and then walks the tick's samples, 32 at a time (the sample-major loop
        // recomputed them for every sample; tables of them in shared memory or shuffled across a warp cost more than
        // they saved).  Which samples belong to tick j is decided by the reference's own
        // comparisons on x = i / sr -- j dt <= x < (j + 1) dt, the same fp64 products for the shared boundary of two
        // ticks -- so the ticks partition the samples exactly as np.interp's bracket does; samples at or beyond the last
        // knot (x >= last dt) take its value and are spread over all threads.
        {
            const int bend_len = pl.bend_len, last = bend_len - 1;
            const double add = (double)pl.pitch_midi;
            const double tadd = pl.t_cents ? ((double)pl.t_cents / 100.0) : 0.0;
            const bool has_t = pl.t_cents != 0;
            auto semi = [&](int k) {
                double v = gf_div_by((double)bend[k], 100.0, 0.01) + add;
                if (has_t) v = v + tadd;
                return v;
            };
            const double dt = 60.0 / (pl.tempo * 96.0);
            const double srd = (double)sr_i, rcp_sr = __drcp_rn((double)sr_i);
            const double xl = (double)last * dt;
            const int lane = threadIdx.x & 31;
            const int warps = (int)(blockDim.x >> 5);
            for (int j = blockIdx.x * warps + (threadIdx.x >> 5); j < last; j += gridDim.x * warps) {
                const double xj = (double)j * dt, xj1 = (double)(j + 1) * dt;
                const double y0 = semi(j), y1 = semi(j + 1);
                const double slope = __ddiv_rn(y1 - y0, xj1 - xj);
                // candidates: two samples of slack on either side of [xj sr, xj1 sr); the exact test below decides
                const int lo = max(0, (int)(xj * srd) - 2), hi = min(n, (int)(xj1 * srd) + 3);
                if (lo >= n) break;                                // the bend string runs past the end of the note
                for (int i = lo + lane; i < hi; i += 32) {
                    const double x = gf_div_by((double)i, srd, rcp_sr);
                    if (!(xj <= x && x < xj1)) continue;
                    const double midi = (xj == x) ? y0 : __dadd_rn(__dmul_rn(slope, x - xj), y0);
                    const double hz = 440.0 * exp2(gf_div_by(midi - 69.0, 12.0, GF_RCP12));
                    out_f0[i] = (float)((double)vm[i] * hz);
                }
            }
            // at or beyond the last knot (x clipped to xl: SillySampler.py:846)
            const double ytail = semi(last);
            const double hz_tail = 440.0 * exp2(gf_div_by(ytail - 69.0, 12.0, GF_RCP12));
            for (int i = max(0, (int)(xl * srd) - 2) + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
                const double x = gf_div_by((double)i, srd, rcp_sr);
                if (x >= xl) out_f0[i] = (float)((double)vm[i] * hz_tail);
            }
        }
        return;
#else
        for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
            const int i = base + threadIdx.x;
            float msv = 1.0f;
            if (i < n) msv = gf_ms_at_fast(mshort, M, i, n);
            const int all_one = __syncthreads_and(msv == 1.0f);
            if (threadIdx.x == 0) ms_one[base >> 8] = (unsigned char)all_one;
            if (i < n && !all_one) out_ms[i] = msv;           // blocks where the smoothed mask is 1 throughout are flagged, not stored
            if (i >= n) continue;
            const double m = (double)vm[i];
            double hz = hz_flat;
            if (!flat) {
                const double midi = gf_midi_at_fast(bend, pl.bend_len, (double)pl.pitch_midi, pl.t_cents, pl.tempo, sr_i, i);
                hz = 440.0 * exp2(gf_div_by(midi - 69.0, 12.0, GF_RCP12));
            }
            out_f0[i] = (float)(m * hz);
        }
        return;
#endif
    }
    gf_ms_loop(nd.ms_short, M, n, nd.ms_one, nd.ms);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // without the velocity stretch mask_new is a plain copy of source samples: vm (f32) holds it exactly
        const double m = pl.vel_active ? gf_mask_new(pl, mask_src, i) : (double)nd.vm[i];
        const double midi = flat ? midi_flat : gf_midi_at(pl, bend, i);
        const double hz = flat ? hz_flat : 440.0 * exp2(gf_div_by(midi - 69.0, 12.0, GF_RCP12));
        double f0 = f0_direct ? (double)f0_direct[i] : m * hz;
        // ---- vocal fry f0 override (SillySampler.py:890-934) ----
        if (pl.fry_L > 0) {
            const double base = pl.vh * (m > 0.0 ? 1.0 : 0.0);
            const int L = pl.fry_L, glide = pl.fry_glide, cst = pl.fry_const;
            if (pl.vf > 0) {
                if (i < cst) f0 = base;
                else if (i < L) { const double w = gf_lin01(i - cst, glide); f0 = (1.0 - w) * base + w * f0; }
            } else {
                const int s = n - L;
                if (i >= s + glide) f0 = base;
                else if (i >= s) { const double w = gf_lin10(i - s, glide); f0 = (1.0 - w) * base + w * f0; }
            }
        }
        if (nd.f0n) nd.f0n[i] = (float)f0;
        if (nd.pd_in) {
            // SillySampler.py:860-865: bend relative to the note (+ t), as float32
            const double base = (double)pl.pitch_midi + ((double)pl.t_cents / 100.0);
            nd.pd_in[i] = (float)(midi - base);
        }
#pragma unroll
        for (int p = 0; p < GF_MAX_PASSES; ++p) {
            if (p >= npass) break;
            float v;
            switch (pkind[p]) {
            case GF_PASS_MAIN: {
                v = (float)f0;
                if (pl.f0_jitter) {
                    // GOOFER.py:666-669, 1070-1071: f0 (f32) *= 1 + (jitter - 1) * mask, evaluated in fp64
                    const double z = nd.z_sh[i] / sh_scale;
                    const double jit = 1.0 + z * pl.f0_jitter_strength;
                    v = (float)((double)v * (1.0 + ((jit - 1.0) * (double)(float)m)));
                }
                break;
            }
            case GF_PASS_SU: v = (float)(f0 * 0.5); break;
            case GF_PASS_SJ: {
                const double z = 0.0 + (pl.sj * pl.sj) * normals[pl.nrm_off[3] + i];     // SillySampler.py:1064
                v = (float)(f0 * (0.5 * pow(2.0, z)));
                break;
            }
            default: v = (float)f0; break;
            }
            pf0[p][i] = v;
        }
    }
}

void gf_launch_f0(const GfNotePlan *plans, const GfNoteDev *notes, const GfPassDev *passes, const GfSourceDev *srcs,
                  const float *bend, const double *normals, const float *f0_curves, int n_notes, int max_n, cudaStream_t st)
{
    if (n_notes <= 0) return;
    // few fat CTAs per note (the per-CTA prologue reads three records) unless the batch is too small to fill the GPU
#ifndef GF_F0_GX
#define GF_F0_GX 8
#endif
    const int gx = max(GF_F0_GX, min(64, (1184 + n_notes - 1) / n_notes));
    dim3 grid(min(gx, (max_n + 255) / 256), n_notes);
    gf_f0_kernel<<<grid, 256, 0, st>>>(plans, notes, passes, srcs, bend, normals, f0_curves);
}

// ------------------------------------------------------------------------------------------------
// LF pulse sample (GOOFER.py:507-522), before the per-table max normalisation
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gf_lf_value(int j, double T, int T0)
{
    const double Ra = 0.02, Rg = 1.7, Rk = 0.8;
    const double Tp = Ra * T;
    const double Tc = __dadd_rn(Tp, __dmul_rn(Rk, T - Tp));
    const double ti = __ddiv_rn(__dmul_rn((double)j, T), (double)T0);
    double v;
    if (ti < Tp) {
        const double s = sin(__ddiv_rn(__dmul_rn(GF_PI_D, ti), __dadd_rn(__dmul_rn(2.0, Tp), 1e-12)));
        v = s * s;
    } else if (ti < Tc) {
        const double tau = __ddiv_rn(ti - Tp, __dadd_rn(Tc - Tp, 1e-12));
        v = exp(-Rg * tau) * cos(GF_PI_D * tau / 2.0);
    } else v = 0.0;
    return (float)v;
}

// max |table| : the rise is increasing and the fall decreasing, so the peak sits at the transition
__device__ __forceinline__ float gf_lf_table_max(double T, int T0)
{
    int jc = (int)(0.02 * (double)T0);
    float m = 0.0f;
    for (int j = jc - 1; j <= jc + 2; ++j)
        if (j >= 0 && j < T0) m = fmaxf(m, fabsf(gf_lf_value(j, T, T0)));
    return m;
}

#define GF_WALK_S 8                   // samples per lane (a warp advances 256 samples per attempt)

// ------------------------------------------------------------------------------------------------
// Phase walk of pulse_train_numba (GOOFER.py:479-493): total_phase += f0[i] / sr in fp64, sample by sample; a pulse
// fires whenever the running total passes the next integer.  The sequential rounding is part of the result (rounding
// ties on every A note), but the chain does not have to be EXECUTED serially
// (44,100 dependent DADDs per second of audio, ~36 cycles each):
//
//   While the exponent e of the running total is fixed, total = M * 2^(e-52) with an integer mantissa M in
//   [2^52, 2^53), and adding the increment +-m_k * 2^(e_k-52) in round-to-nearest-even is an INTEGER step.  With
//   s = e - e_k, q = m_k >> s, r = the s bits shifted out, half = 2^(s-1):
//       x > 0:  M' = M + q + [r > half] + [r == half] * ((M + q) & 1)
//       x < 0:  M' = M - q                                         (r == 0)
//               M' = M - q - 1 + [r < half] + [r == half] * ((M - q - 1) & 1)     (r > 0)
//   i.e. a map  M -> M + delta[M & 1]  with a pair (delta[0], delta[1]) that depends only on the increment.  Such
//   maps are closed under composition, (a then b)[p] = a[p] + b[(p + a[p]) & 1], so the running mantissas come
//   out of a scan of integer pairs -- bit for bit what the scalar fp64 loop produces, ties included.
//   A step leaves this regime ("event") when the result changes binade (M' >= 2^53, or the exact difference drops
//   below 2^52), when there is no positive normal total yet, or when the increment is above the total / subnormal:
//   about 16 per note.  Events are executed as one real fp64 addition.
//   (The delta pair itself is obtained from two real fp64 additions on representative totals of the binade, see
//   gf_walk_delta: 0.64 -> 0.54 ms against evaluating the integer rule with 64-bit shifts and masks.)
//
// Per block of 256 x WARPS samples: ATTEMPT (8 samples per lane: local composition, warp scan, cross-warp hop, walk),
// find the first event k*, COMMIT the samples before it (onsets = increments of the running max of floor(total)),
// execute k* in fp64, continue after it.  The same arithmetic in Python integers, checked against the scalar
// loop on ties, gaps, negative and tiny increments: tests/test_walk_arith_cpu.py.
// ------------------------------------------------------------------------------------------------
struct GfDelta { long long d0, d1; };                  // M -> M + (M & 1 ? d1 : d0)


__device__ __forceinline__ GfDelta gf_delta_then(const GfDelta &a, const GfDelta &b)
{
    GfDelta r;
    r.d0 = a.d0 + ((a.d0 & 1ll) ? b.d1 : b.d0);                    // parity of 0 + a.d0
    r.d1 = a.d1 + (((1ll + a.d1) & 1ll) ? b.d1 : b.d0);            // parity of 1 + a.d1
    return r;
}

// delta pair of one increment under the exponent e of a positive normal running total.  Returns true when the step
// needs a real fp64 addition instead (subnormal increment, increment above the total).  guard = q + [r > 0] for a
// subtraction: the exact difference stays in the binade iff M - guard >= 2^52.
__device__ __forceinline__ bool gf_walk_delta(double inc, int e, GfDelta &d, long long &guard)
{
    d.d0 = 0ll; d.d1 = 0ll; guard = 0ll;
    const long long ib = __double_as_longlong(inc);
    if ((ib << 1) == 0) return false;                              // +-0: identity
    const int efield = (int)((ib >> 52) & 0x7ff);
#ifndef GF_WALK_INT_DELTA
    // The pair is read off two REAL additions on representative totals of the binade: 1.5 * 2^e (even mantissa) and
    // its successor (odd).  The delta depends on the total only through the parity of its mantissa as long as the
    // sum stays in the binade, which the representatives do for |inc| < 2^(e-1); larger increments (the first
    // samples of a note) are events.  t - base is exact, scaling by 2^(52-e) is exact.  Same pairs as the integer
    // rule below (tests/test_walk_arith_cpu.py), a dozen fp64 instructions instead of ~90 integer ones.
    if (efield == 0 || (efield - 1023) >= e - 1) return true;
    const long long eb = (long long)(e + 1023) << 52;
    const double base0 = __longlong_as_double(eb | (1ll << 51)), base1 = __longlong_as_double(eb | (1ll << 51) | 1ll);
    const double scale = __longlong_as_double((long long)(1023 + 52 - e) << 52);
    d.d0 = __double2ll_rn(__dmul_rn(__dadd_rn(__dadd_rn(base0, inc), -base0), scale));
    d.d1 = __double2ll_rn(__dmul_rn(__dadd_rn(__dadd_rn(base1, inc), -base1), scale));
    if (ib < 0) guard = __double2ll_ru(__dmul_rn(-inc, scale));
    return false;
#else
    // the same pair by the integer rounding rule stated above (kept as the readable definition; -DGF_WALK_INT_DELTA)
    const int s = e - (efield - 1023);
    if (efield == 0 || s < 0) return true;
    if (s >= 64) return false;                                     // far below half an ulp: no change
    const unsigned long long mk = ((unsigned long long)ib & ((1ull << 52) - 1ull)) | (1ull << 52);
    const long long q = (long long)(mk >> s);
    const unsigned long long r = s > 0 ? (mk & ((1ull << s) - 1ull)) : 0ull, half = s > 0 ? (1ull << (s - 1)) : 0ull;
    const long long tie = (s > 0 && r == half) ? 1ll : 0ll;
    if (ib > 0) {
        const long long c = (s > 0 && r > half) ? 1ll : 0ll;
        d.d0 = q + c + (tie & q);                                  // (0 + q) & 1
        d.d1 = q + c + (tie & (q + 1ll));                          // (1 + q) & 1
    } else {
        guard = q + (r > 0 ? 1ll : 0ll);
        if (r == 0) { d.d0 = -q; d.d1 = -q; }
        else {
            const long long c = r < half ? 1ll : 0ll;
            d.d0 = -q - 1ll + c + (tie & (q + 1ll));               // (0 - q - 1) & 1
            d.d1 = -q - 1ll + c + (tie & q);                       // (1 - q - 1) & 1
        }
    }
    return false;
#endif
}

#ifndef GF_WALK_MINB
#define GF_WALK_MINB 1
#endif
template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS, GF_WALK_MINB)
gf_walk_kernel(const GfPassDev *__restrict__ passes, GfPassScal *scal, int n_pass, int sr_i, const int *__restrict__ flag_count, int cnt_lo, int cnt_hi)
{
    // one CTA of WARPS warps per (note, pass); every warp keeps an identical copy of the walk state
    __shared__ GfDelta s_tot[WARPS];
    __shared__ long long s_M;
    __shared__ int s_kbad[WARPS], s_wmax[WARPS], s_hasv[WARPS];
    __shared__ float s_lastv[WARPS];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pi = blockIdx.x;
    if (pi >= n_pass) return;
#if !defined(GF_WALK_SEQ_ONLY)
    if (!scal[pi].walk_seq) return;                        // gf_walk_scan_kernel placed this pass's onsets
    if (flag_count) {
        // the launcher issues this kernel in two shapes (many warps per pass for a few flagged passes, few warps for many):
        // the number of flagged passes, counted by the scan, decides on the device which of the two does the work
        const int c = *flag_count;
        if (c < cnt_lo || c > cnt_hi) return;
    }
#endif
    const GfPassDev ps = passes[pi];
    const int n = ps.n_total;
    const double sr = (double)sr_i;
    const float *__restrict__ f0 = ps.f0;
    const long long MANT = (1ll << 52) - 1ll, ONE52 = 1ll << 52, ONE53 = 1ll << 53;
    // running total = started ? M * 2^(e - 52) : (raw_mode ? raw : 0).  raw_mode: the total is negative or subnormal
    // (only possible with f0 jitter beyond 100 % at the very start of a note): every non-zero sample is an event
    bool started = false, raw_mode = false;
    double raw = 0.0;
    long long M = 0ll;
    int e = 0;
    int fired = 0;                      // next_k - 1
    float lv_carry = 160.0f;            // last_valid_f0 (GOOFER.py:477)
    int count = 0;
    const bool aligned16 = (reinterpret_cast<size_t>(f0) & 15) == 0;
    const double rcp_sr = __drcp_rn(sr);
    const int BLK = 32 * GF_WALK_S * WARPS;

    for (int blk = 0; blk < n; blk += BLK) {
        const int blk_end = min(n, blk + BLK);
        int lo = blk;
        while (lo < blk_end) {
            // ================= attempt: samples [lo, blk_end) under the current (M, e) =================
            const int i0 = blk + 32 * GF_WALK_S * w + GF_WALK_S * lane;
            float f[GF_WALK_S];
            GfDelta d[GF_WALK_S];
            long long guard[GF_WALK_S];
            unsigned evbits = 0u;                                  // bit j: sample j needs a real fp64 addition
            const bool whole = aligned16 && lo <= i0 && i0 + GF_WALK_S <= blk_end;
            if (whole) {
                const float4 a = *reinterpret_cast<const float4 *>(f0 + i0), b4 = *reinterpret_cast<const float4 *>(f0 + i0 + 4);
                f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b4.x; f[5] = b4.y; f[6] = b4.z; f[7] = b4.w;
            }
#pragma unroll
            for (int j = 0; j < GF_WALK_S; ++j) {
                const int i = i0 + j;
                const bool act = i >= lo && i < blk_end;
                if (!whole) f[j] = act ? f0[i] : 0.0f;
                d[j].d0 = 0ll; d[j].d1 = 0ll; guard[j] = 0ll;
                if (act && f[j] != 0.0f) {
                    if (!started || raw_mode) evbits |= 1u << j;
                    else if (gf_walk_delta(gf_div_by((double)f[j], sr, rcp_sr), e, d[j], guard[j])) evbits |= 1u << j;
                }
            }
            // A delta pair differs in its two entries only when the increment sits on a rounding TIE of the binade (the
            // bits shifted out are exactly half an ulp): rare outside a few special pitches.  A warp whose 256 samples hold
            // no tie composes its maps as plain 64-bit prefix sums -- two instructions per composition instead of fourteen
            // (parity test, two 64-bit selects, two 64-bit adds), which were 41 % of the kernel's instructions (ncu, round 2).
            bool tie = false;
#pragma unroll
            for (int j = 0; j < GF_WALK_S; ++j) tie = tie || (d[j].d0 != d[j].d1);
#ifdef GF_WALK_ASSUME_NO_TIE
            const bool warp_tie = false;                           // timing experiment only: wrong on ties
#else
            const bool warp_tie = __any_sync(0xffffffffu, tie);
#endif
            GfDelta F;                                             // inclusive scan of the lane maps inside the warp
            if (!warp_tie) {
                long long g = d[0].d0;
#pragma unroll
                for (int j = 1; j < GF_WALK_S; ++j) g += d[j].d0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const long long a = __shfl_up_sync(0xffffffffu, g, o);
                    if (lane >= o) g += a;
                }
                F.d0 = g; F.d1 = g;
            } else {
                GfDelta G = d[0];
#pragma unroll
                for (int j = 1; j < GF_WALK_S; ++j) G = gf_delta_then(G, d[j]);
                F = G;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    GfDelta a;
                    a.d0 = __shfl_up_sync(0xffffffffu, F.d0, o);
                    a.d1 = __shfl_up_sync(0xffffffffu, F.d1, o);
                    if (lane >= o) F = gf_delta_then(a, F);
                }
            }
            if (WARPS > 1) {
                if (lane == 31) s_tot[w] = F;
                __syncthreads();
            }
            GfDelta E; E.d0 = 0ll; E.d1 = 0ll;                     // map from the block start to this lane's first sample
            if (WARPS > 1) for (int q = 0; q < w; ++q) E = gf_delta_then(E, s_tot[q]);
            {
                GfDelta p;
                p.d0 = __shfl_up_sync(0xffffffffu, F.d0, 1);
                p.d1 = __shfl_up_sync(0xffffffffu, F.d1, 1);
                if (lane > 0) E = gf_delta_then(E, p);
            }
            // walk the lane's samples with concrete mantissas; stop being meaningful at the first bad sample
            long long Mv[GF_WALK_S];
            long long Mj = M + ((M & 1ll) ? E.d1 : E.d0);
            int kbad = INT_MAX;
#pragma unroll
            for (int j = 0; j < GF_WALK_S; ++j) {
                bool bad = (evbits >> j) & 1u;
                if (started && !raw_mode) {
                    bad |= (Mj - guard[j]) < ONE52;                // the exact difference leaves the binade
                    Mj += (Mj & 1ll) ? d[j].d1 : d[j].d0;
                    bad |= Mj >= ONE53 || Mj < ONE52;
                }
                Mv[j] = Mj;
                if (bad && kbad == INT_MAX && (i0 + j) >= lo && (i0 + j) < blk_end) kbad = i0 + j;
            }
            // a bad sample poisons every later mantissa: the first one over the whole block decides
            int kstar = __reduce_min_sync(0xffffffffu, kbad);
            if (WARPS > 1) {
                if (lane == 0) s_kbad[w] = kstar;
                __syncthreads();
                kstar = INT_MAX;
                for (int q = 0; q < WARPS; ++q) kstar = min(kstar, s_kbad[q]);
            }
            if (kstar == INT_MAX) kstar = blk_end;
            // ================= commit samples [lo, kstar) =================
            const int sh = 52 - e;
            int m[GF_WALK_S];
            int lmax = INT_MIN;
            float lastv = 0.0f; bool hasv = false;
#pragma unroll
            for (int j = 0; j < GF_WALK_S; ++j) {
                const int i = i0 + j;
                const bool com = i >= lo && i < kstar;
                m[j] = com ? ((started && !raw_mode && e >= 0) ? (int)(Mv[j] >> sh) : 0) : INT_MIN;
                lmax = max(lmax, m[j]);
                if (com && (f[j] > 1e-6f)) { lastv = f[j]; hasv = true; }          // (double) f > 1e-6  <=>  f > RN_f32(1e-6): RN_f32(1e-6) < 1e-6
            }
            // running max of floor(total) before this lane: earlier lanes, earlier warps, `fired`
            int pmax = lmax;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, pmax, o);
                if (lane >= o) pmax = max(pmax, v);
            }
            const int wmax = __shfl_sync(0xffffffffu, pmax, 31);
            int before = __shfl_up_sync(0xffffffffu, pmax, 1);
            if (lane == 0) before = INT_MIN;
            const unsigned hv = __ballot_sync(0xffffffffu, hasv);
            const int owner = kstar - 1;                           // last committed sample (if kstar > lo)
            if (WARPS > 1) {
                if (lane == 0) { s_wmax[w] = wmax; s_hasv[w] = hv != 0u; }
                if (hv && lane == 31 - __clz(hv)) s_lastv[w] = lastv;
            }
            if (kstar > lo && owner >= i0 && owner < i0 + GF_WALK_S) s_M = Mv[owner - i0];
            __syncthreads();
            float lv_prior = lv_carry, lv_after = lv_carry;
            int rm_all = fired;
            if (WARPS > 1) {
                for (int q = 0; q < WARPS; ++q) {
                    if (q < w) before = max(before, s_wmax[q]);
                    rm_all = max(rm_all, s_wmax[q]);
                    if (s_hasv[q]) { if (q < w) lv_prior = s_lastv[q]; lv_after = s_lastv[q]; }
                }
            } else {
                rm_all = max(rm_all, wmax);
                if (hv) lv_after = __shfl_sync(0xffffffffu, lastv, 31 - __clz(hv));
            }
            before = max(before, fired);
            if (rm_all > fired) {
                const unsigned below = hv & ((1u << lane) - 1u);
                const float prior = __shfl_sync(0xffffffffu, lastv, below ? (31 - __clz(below)) : 0);
                float lv = below ? prior : lv_prior;
                int rm = before;
#pragma unroll
                for (int j = 0; j < GF_WALK_S; ++j) {
                    const int i = i0 + j;
                    if (i >= lo && i < kstar) {
                        if (f[j] > 1e-6f) lv = f[j];
                        // onsets number rm+1 .. m[j] (1-based since the note start) sit in slots rm .. m[j]-1: `count`
                        // onsets were written when `fired` pulses had fired
                        for (int c = rm; c < m[j]; ++c) {
                            const int slot = count + (c - fired);
                            if (slot < ps.onset_cap) ps.onsets[slot] = make_int4(i, 0, __float_as_int(lv), 0);   // T0 / table max: gf_onset_kernel
                        }
                        rm = max(rm, m[j]);
                    }
                }
            } else {
                // keep the shuffle above convergent for all lanes
            }
            count += rm_all - fired;
            fired = rm_all;
            lv_carry = lv_after;
            if (kstar > lo && started && !raw_mode) M = s_M;
            __syncthreads();                                       // shared slots are rewritten by the next attempt
            // ================= the event sample: one real fp64 addition =================
            if (kstar < blk_end) {
                const float fe = f0[kstar];
                const double prev = raw_mode ? raw
                                  : (started ? __longlong_as_double((long long)(((unsigned long long)(e + 1023) << 52) | (unsigned long long)(M & MANT))) : 0.0);
                const double tot = __dadd_rn(prev, gf_div_by((double)fe, sr, rcp_sr));
                const long long tb = __double_as_longlong(tot);
                const int tf = (int)((tb >> 52) & 0x7ff);
                if (tb > 0 && tf != 0) { started = true; raw_mode = false; M = (tb & MANT) | ONE52; e = tf - 1023; }
                else if ((tb << 1) == 0) { started = false; raw_mode = false; M = 0ll; e = 0; }
                else { started = false; raw_mode = true; raw = tot; M = 0ll; e = 0; }
                if (fe > 1e-6f) lv_carry = fe;
                const int me = (int)fmin(fmax(floor(tot), -2.0e9), 2.0e9);
                if (me > fired) {
                    if (threadIdx.x == 0)
                        for (int c = fired; c < me; ++c) {
                            const int slot = count + (c - fired);
                            if (slot < ps.onset_cap) ps.onsets[slot] = make_int4(kstar, 0, __float_as_int(lv_carry), 0);
                        }
                    count += me - fired;
                    fired = me;
                }
                lo = kstar + 1;
            } else lo = blk_end;
        }
    }
    if (threadIdx.x == 0) {
        scal[pi].n_onsets = min(count, ps.onset_cap);
        if (count > ps.onset_cap) scal[pi].err = 1;
    }
}

// ------------------------------------------------------------------------------------------------
// The same onsets WITHOUT the rounding chain, for the passes where rounding cannot matter.  Only the onset
// positions (and last_valid_f0 there) leave pulse_train_numba's loop: pulse number c fires at the first sample
// where the running total reaches c, i.e. where the running maximum of floor(total) grows.  The fp64 total after k
// additions differs from the exact sum of the increments by at most k half-ulps of the largest total so far; the
// exact sum exceeds a fixed-point sum of the increments truncated (floor) to multiples of 2^-44 by less than
// k 2^-44.  So if, at every sample that adds something, the interval
//     [A_k - m_k, A_k + m_k + (k + 2) 2^-44],   A_k = fixed-point sum,  m_k = (k + 2) max(2^-44, half an ulp)
// either lies below the next firing level or has the same floor at both ends, the sequential loop and the integer
// scan fire at the same samples, and the pass is done after one block scan of 64-bit sums and maxima (44,100
// samples: 22 tiles of 2,048) instead of the ~86 attempt / commit rounds of the bit-exact walk below (0.5 ms
// however many notes there are: every pass is resident at once and waits on its own chain).  Otherwise -- a flat A
// note, whose total sits on an integer every 2,205 samples up to rounding; ~2e-4 of the other 1 s notes by
// chance -- the pass is flagged and gf_walk_kernel renders it.  Negative increments (f0 jitter beyond 100 %) are
// part of the scan; a negative TOTAL, a non-finite or absurd f0 and an onset on a sample with f0 <= 1e-6 are flagged.
// Passes longer than 4 s go straight to the walk: the bound grows with k^2 (a 16 s note at 440 Hz would be flagged one
// time in four) and one CTA scanning 345 tiles in a row is no faster than 16 warps walking them (c4: 5.7 -> 7.1 ms with it).
// The rule in Python integers against the scalar loop: tests/test_walk_arith_cpu.py.
// ------------------------------------------------------------------------------------------------
#define GF_WSC_THREADS 256
#define GF_WSC_PER 8
#define GF_WSC_UNIT 44
#define GF_WSC_MAX_N (4 * 44100)
__global__ void __launch_bounds__(GF_WSC_THREADS)
gf_walk_scan_kernel(const GfPassDev *__restrict__ passes, GfPassScal *scal, int n_pass, int sr_i, int *flag_count)
{
    __shared__ long long s_warp[GF_WSC_THREADS / 32];
    __shared__ int s_wmax[GF_WSC_THREADS / 32];
    const int pi = blockIdx.x;
    if (pi >= n_pass) return;
    const GfPassDev ps = passes[pi];
    const int n = ps.n_total;
    if (n > GF_WSC_MAX_N) {
        if (threadIdx.x == 0) { scal[pi].walk_seq = 1; if (flag_count) atomicAdd(flag_count, 1); }
        return;
    }
    const double sr = (double)sr_i, rcp_sr = __drcp_rn((double)sr_i);
    const float fmax_ok = 0.25f * (float)sr_i;
    const float *__restrict__ f0 = ps.f0;
    const bool aligned16 = (reinterpret_cast<size_t>(f0) & 15) == 0;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    long long carry = 0ll;              // fixed-point total before this tile
    int fired = 0;                      // running maximum of floor(total) = pulses fired before this tile
    int bad = 0;
    // a sample can only be borderline if the fraction of the total is within the LARGEST margin of the pass of 0 or 1:
    // one comparison screens (nearly) every sample before the exact test below
    const int sh_max = max(0, (31 - __clz(n + 1)) - 9);
    const long long quick = ((long long)n + 2ll) * 2ll + (((long long)n + 2ll) << sh_max);
    const long long ONE = 1ll << GF_WSC_UNIT;
    auto load_tile = [&](int i0, float *f) {
        if (aligned16 && i0 + GF_WSC_PER <= n) {
            const float4 a = *reinterpret_cast<const float4 *>(f0 + i0), b = *reinterpret_cast<const float4 *>(f0 + i0 + 4);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        } else {
#pragma unroll
            for (int e = 0; e < GF_WSC_PER; ++e) f[e] = (i0 + e < n) ? f0[i0 + e] : 0.0f;
        }
    };
    float fnext[GF_WSC_PER];
    load_tile(tid * GF_WSC_PER, fnext);
    for (int base = 0; base < n; base += GF_WSC_THREADS * GF_WSC_PER) {
        const int i0 = base + tid * GF_WSC_PER;
        float f[GF_WSC_PER];
#pragma unroll
        for (int e = 0; e < GF_WSC_PER; ++e) f[e] = fnext[e];
        load_tile(i0 + GF_WSC_THREADS * GF_WSC_PER, fnext);      // the next tile's samples travel while this one is scanned
        long long loc[GF_WSC_PER];
        long long run = 0ll;
#pragma unroll
        for (int e = 0; e < GF_WSC_PER; ++e) {
            long long u = 0ll;
            if (f[e] != 0.0f) {
                if (!(fabsf(f[e]) < fmax_ok)) bad = 1;                                                 // NaN, inf, absurd
                else u = __double2ll_rd(gf_div_by((double)f[e], sr, rcp_sr) * 17592186044416.0);      // 2^44: exact scaling
            }
            run += u;
            loc[e] = run;
        }
        long long incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        long long pre = carry + (incl - run), total = carry;
#pragma unroll
        for (int q = 0; q < GF_WSC_THREADS / 32; ++q) {
            const long long t = s_warp[q];
            if (q < w) pre += t;
            total += t;
        }
        carry = total;
        // floor(total) after each sample, and its running maximum
        int F[GF_WSC_PER];
        int lmax = INT_MIN;
#pragma unroll
        for (int e = 0; e < GF_WSC_PER; ++e) {
            const long long A = pre + loc[e];
            if (A < 0ll || (A >> 62)) bad = 1;
            F[e] = (int)(A >> GF_WSC_UNIT);
            lmax = max(lmax, F[e]);
        }
        int pmax = lmax;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, pmax, o);
            if (lane >= o) pmax = max(pmax, t);
        }
        if (lane == 31) s_wmax[w] = pmax;
        int before = __shfl_up_sync(0xffffffffu, pmax, 1);
        if (lane == 0) before = INT_MIN;
        __syncthreads();
        int tile_max = fired;
#pragma unroll
        for (int q = 0; q < GF_WSC_THREADS / 32; ++q) {
            const int t = s_wmax[q];
            if (q < w) before = max(before, t);
            tile_max = max(tile_max, t);
        }
        int R = max(before, fired);                       // pulses fired before this thread's first sample
#pragma unroll
        for (int e = 0; e < GF_WSC_PER; ++e) {
            if (f[e] != 0.0f) {
                const long long A = pre + loc[e];
                const int Rn = max(R, F[e]);
                const long long fr = A & (ONE - 1ll);
                if (fr <= quick || fr >= ONE - quick) {
                    const long long k2 = (long long)(i0 + e) + 2ll;
                    const int sh = max(0, 10 - __clzll((long long)(Rn + 1) << GF_WSC_UNIT));  // half an ulp of the largest total so far (< Rn + 1), units of 2^-44
                    const long long m = k2 << sh;
                    const long long lo = max(A - m, 0ll), hi = A + m + k2;
                    if ((lo >> GF_WSC_UNIT) != (hi >> GF_WSC_UNIT) && (hi >> GF_WSC_UNIT) > (long long)R) bad = 1;
                }
                if (Rn > R) {
                    if (!(f[e] > 1e-6f)) bad = 1;                                             // last_valid_f0 would come from an earlier sample
                    for (int c = R; c < Rn; ++c)
                        if (c < ps.onset_cap) ps.onsets[c] = make_int4(i0 + e, 0, __float_as_int(f[e]), 0);
                    R = Rn;
                }
            }
        }
        fired = tile_max;
        __syncthreads();
    }
    bad = __syncthreads_or(bad);
    if (tid == 0) {
        if (bad) { scal[pi].walk_seq = 1; if (flag_count) atomicAdd(flag_count, 1); }
        else {
            scal[pi].n_onsets = min(fired, ps.onset_cap);
            if (fired > ps.onset_cap) scal[pi].err = 1;
        }
    }
}

// per onset: period length T0 = round(sr / last_valid_f0) clipped to [3, 8192] (GOOFER.py:495-499, Python
// round = half to even), the peak of its LF table (GOOFER.py:524-528) and the two branch points of the table
// (first j with ti >= Tp, first j with ti >= Tc -- decided with the reference's own fp64 expressions, because
// 0.02 * T0 is an integer for T0 = 50 k and the comparison then sits on a rounding tie); thread per onset
__global__ void __launch_bounds__(128)
gf_onset_kernel(const GfPassDev *__restrict__ passes, GfPassScal *scal, int sr_i)
{
    const GfPassDev ps = passes[blockIdx.y];
    const int count = scal[blockIdx.y].n_onsets;
    const double sr = (double)sr_i;
    int mx = 0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        int4 o = ps.onsets[e];
        double lv = (double)__int_as_float(o.z);
        if (!(lv > 1e-6)) lv = 1e-6;
        const double T = __ddiv_rn(1.0, lv);
        int T0 = (int)rint(__dmul_rn(sr, T));
        T0 = T0 < 3 ? 3 : (T0 > 8192 ? 8192 : T0);
        const double Tp = 0.02 * T;
        const double Tc = __dadd_rn(Tp, __dmul_rn(0.8, T - Tp));
        auto ti = [&](int j) { return __ddiv_rn(__dmul_rn((double)j, T), (double)T0); };
        int jp = max(0, (int)(0.02 * (double)T0) - 1);
        while (jp < T0 && ti(jp) < Tp) ++jp;
        while (jp > 0 && !(ti(jp - 1) < Tp)) --jp;
        int jc = max(jp, (int)(0.804 * (double)T0) - 1);
        while (jc < T0 && ti(jc) < Tc) ++jc;
        while (jc > jp && !(ti(jc - 1) < Tc)) --jc;
        o.y = T0;
        o.z = jp | (jc << 16);
        o.w = __float_as_int(gf_lf_table_max(T, T0));
        ps.onsets[e] = o;
        mx = max(mx, T0);
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(&scal[blockIdx.y].max_T0, mx);
}

int gf_launch_walk(const GfPassDev *passes, GfPassScal *scal, int n_pass, int max_n, int sr, cudaStream_t st, int *flag_count)
{
    if (n_pass <= 0) return 0;
    // One warp per note costs the fewest issue slots: right when there are enough notes to fill the GPU.  More warps per
    // note cut the latency of a note's chain: right for few and / or long notes.  Measured on B200 (ms, 1 s notes, every
    // pass on the bit-exact walk; warps per note 1 / 2 / 4 / 8 / 16): 128 notes .65 .38 .24 .18 .19 | 256: .68 .40 .26 .34 .36 |
    // 512: .69 .43 .50 .65 .69 | 1,024: .74 .55 .76 1.13 1.20 | 2,048: .95 1.00 1.29 2.2 2.4; 96 notes of 16 s: 9.9 5.4 2.9 1.8 1.4.
    static int force = -1;
    if (force < 0) { const char *e = getenv("GOOFER_WALK_WARPS"); force = e ? atoi(e) : 0; }
    auto launch = [&](int nw, const int *fc, int lo, int hi) {
        if (nw == 16) gf_walk_kernel<16><<<n_pass, 512, 0, st>>>(passes, scal, n_pass, sr, fc, lo, hi);
        else if (nw == 8) gf_walk_kernel<8><<<n_pass, 256, 0, st>>>(passes, scal, n_pass, sr, fc, lo, hi);
        else if (nw == 4) gf_walk_kernel<4><<<n_pass, 128, 0, st>>>(passes, scal, n_pass, sr, fc, lo, hi);
        else if (nw == 2) gf_walk_kernel<2><<<n_pass, 64, 0, st>>>(passes, scal, n_pass, sr, fc, lo, hi);
        else gf_walk_kernel<1><<<n_pass, 32, 0, st>>>(passes, scal, n_pass, sr, fc, lo, hi);
    };
    int launches = 0;
#if !defined(GF_WALK_SEQ_ONLY)
    // The scan places the onsets of every pass it can decide; the bit-exact walk then runs for the flagged ones only (its
    // other CTAs leave at once).  How many those are is known on the device only (the scan counts them in *flag_count).
    // -DGF_WALK_FEW=n issues the walk twice -- GF_WALK_FEW_WARPS warps per pass for up to n flagged passes, the batch-size
    // heuristic beyond -- and lets the count decide on the device.  Measured on B200 (walk bucket / whole step, ms): c2 (43
    // flagged) off 0.413 / 4.074, 256 x 8 warps 0.322 / 4.080, 96 x 16 warps 0.329 / 4.101; c3 (150 tie-heavy flagged passes,
    // 256 notes) off 0.883 / 7.913, 256 x 8 0.941 / 7.982, 192 x 16 1.43 / 8.46.  The walk's latency is hidden behind the
    // envelope kernel either way, so the step does not move: off by default.
    gf_walk_scan_kernel<<<n_pass, GF_WSC_THREADS, 0, st>>>(passes, scal, n_pass, sr, flag_count);
    ++launches;
    int nw = n_pass <= 640 ? 8 : (n_pass <= 6144 ? 4 : 2);
#else
    int nw = n_pass <= 160 ? 8 : (n_pass <= 320 ? 4 : (n_pass <= 1536 ? 2 : 1));
    flag_count = nullptr;
#endif
    if (max_n >= 4 * 44100) nw = min(16, 2 * nw);
    if (force == 1 || force == 2 || force == 4 || force == 8 || force == 16) { nw = force; flag_count = nullptr; }
#ifndef GF_WALK_FEW
#define GF_WALK_FEW 0           // tuning knob, off: up to this many flagged passes get GF_WALK_FEW_WARPS warps each (decided on the device)
#endif
#ifndef GF_WALK_FEW_WARPS
#define GF_WALK_FEW_WARPS 8
#endif
    if (flag_count && nw < GF_WALK_FEW_WARPS && GF_WALK_FEW > 0) {
        launch(GF_WALK_FEW_WARPS, flag_count, 1, GF_WALK_FEW);
        launch(nw, flag_count, GF_WALK_FEW + 1, INT_MAX);
        launches += 2;
    } else { launch(nw, nullptr, 0, 0); ++launches; }
    gf_onset_kernel<<<dim3(4, n_pass), 128, 0, st>>>(passes, scal, sr);
    return launches + 1;
}


// ------------------------------------------------------------------------------------------------
// pulse[i] = sum over the onsets whose table covers i, in onset order (f32 adds)  GOOFER.py:542-552
// One CTA per 256-sample tile: two binary searches find the onsets that can reach the tile, their records are
// staged in shared memory, every thread sums the tables covering its sample.  The LF table value is
// evaluated in f32 from (j, T0) alone: T cancels in ti / Tp and (ti - Tp) / (Tc - Tp) up to the 1e-12
// guards (< 1e-8 relative, like the reference's own 5-slot table cache, which reuses one T per T0).
// ------------------------------------------------------------------------------------------------
#define GF_PULSE_CAP 512
#ifndef GF_PULSE_SPAN
#define GF_PULSE_SPAN 4096            // samples per CTA: a contiguous run of 256-sample tiles whose onsets are staged once
#endif
__device__ __forceinline__ float gf_lf_value_f32(int d, int jp, int jc, float r_rise, float r_fall, float inv_max)
{
    float v = 0.0f;
    // arguments in [0, pi / 2] and [-1.7, 0]: the SFU approximations are good to ~5e-7 absolute there (the table is
    // normalised to peak 1; the stage test holds the pulse train to 2e-6 of the oracle); 0.29 -> 0.26 ms
    if (d < jp) { const float s = __sinf((float)d * r_rise); v = s * s; }
    else if (d < jc) {
        const float tau = fmaf((float)d, r_fall, -(0.02f / 0.784f));
        v = __expf(-1.7f * tau) * __cosf(1.57079637f * tau);
    }
    return v * inv_max;
}

__global__ void __launch_bounds__(256)
gf_pulse_kernel(const GfPassDev *__restrict__ passes, const GfPassScal *__restrict__ scal)
{
    // staged onset records: (x, T0, jp | jc << 16, bits of 1 / table max) in one 16-byte slot, the two slopes in an 8-byte one
    __shared__ __align__(16) int4 s_a[GF_PULSE_CAP];
    __shared__ __align__(8) float2 s_r[GF_PULSE_CAP];
    __shared__ int s_rng[2];
    const GfPassDev ps = passes[blockIdx.y];
    const int n = ps.n_total;
    const int count = scal[blockIdx.y].n_onsets;
    const int max_T0 = scal[blockIdx.y].max_T0;
    const int4 *__restrict__ on = ps.onsets;
    const int tiles = (n + 255) / 256;
    const int per = (tiles + gridDim.x - 1) / gridDim.x;
    const int c_begin = blockIdx.x * per * 256, c_end = min(n, c_begin + per * 256);
    if (c_begin >= n) return;
    // onsets that can reach this CTA's run of samples: two binary searches in global memory, once
    if (threadIdx.x < 2) {
        const int key = threadIdx.x == 0 ? c_begin - max_T0 : c_end - 1;
        int lo = 0, hi = count;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (on[mid].x <= key) lo = mid + 1; else hi = mid; }
        s_rng[threadIdx.x] = lo;
    }
    __syncthreads();
    const int e0 = s_rng[0], e1 = s_rng[1];
    const int ne = e1 - e0;
    const bool staged = ne <= GF_PULSE_CAP;
    if (staged) {
        for (int q = threadIdx.x; q < ne; q += blockDim.x) {
            const int4 o = on[e0 + q];
            const float mxv = __int_as_float(o.w);
            s_a[q] = make_int4(o.x, o.y, o.z, __float_as_int(mxv > 0.0f ? 1.0f / mxv : 1.0f));
            s_r[q] = make_float2(1.57079637f / (0.02f * (float)o.y), 1.0f / (0.784f * (float)o.y));
        }
    }
    __syncthreads();
    int q_lo = 0, q_hi = 0;                   // staged onsets [q_lo, q_hi) can reach this WARP's 32 samples of the current tile (both only
                                              // move right).  Per 256-sample tile the window held (256 + T0) / T0 onsets -- 7 at 1 kHz -- of
                                              // which a sample is covered by one or two: 167 instructions per sample; per warp it holds ~1.5
    for (int s0 = c_begin; s0 < c_end; s0 += 256) {
        const int i = s0 + threadIdx.x;
        const int w0 = s0 + (threadIdx.x & ~31);
        float acc = 0.0f;
        if (staged) {
            while (q_hi < ne && s_a[q_hi].x <= w0 + 31) ++q_hi;
            while (q_lo < q_hi && s_a[q_lo].x + max_T0 <= w0) ++q_lo;
            if (i < n) {
                for (int q = q_lo; q < q_hi; ++q) {
                    const int4 a = s_a[q];
                    const int d = i - a.x;
                    if (d >= 0 && d < a.y) {
                        const float2 r = s_r[q];
                        acc += gf_lf_value_f32(d, a.z & 0xffff, a.z >> 16, r.x, r.y, __int_as_float(a.w));
                    }
                }
            }
        } else if (i < n) {
            for (int e = e0; e < e1; ++e) {
                const int4 o = on[e];
                const int d = i - o.x;
                if (d >= 0 && d < o.y) {
                    const float mxv = __int_as_float(o.w);
                    acc += gf_lf_value_f32(d, o.z & 0xffff, o.z >> 16, 1.57079637f / (0.02f * (float)o.y),
                                           1.0f / (0.784f * (float)o.y), mxv > 0.0f ? 1.0f / mxv : 1.0f);
                }
            }
        }
        if (i < n) ps.pulse[i] = acc;
    }
}

void gf_launch_pulse(const GfPassDev *passes, const GfPassScal *scal, int n_pass, int max_n, cudaStream_t st)
{
    if (n_pass <= 0) return;
    dim3 grid((max_n + GF_PULSE_SPAN - 1) / GF_PULSE_SPAN, n_pass);
    gf_pulse_kernel<<<grid, 256, 0, st>>>(passes, scal);
}

// ------------------------------------------------------------------------------------------------
// Noise phases on the device: the stream numpy's Generator(PCG64).uniform(0, 2 pi, (513, T)).astype(float32) draws
// (GOOFER.py:1151-1152), bit for bit.  PCG64 = 128-bit LCG (multiplier 0x2360ED051FC65DA44385DF649FCCF645, per-stream
// odd increment), output XSL-RR 128 -> 64 of the state AFTER the step; next_double = (u64 >> 11) * 2^-53; uniform =
// low + (high - low) * next_double with low = 0.  Element e = k T + t of the row-major (513, T) array is the (e + 1)-th
// draw.  One thread per bin row k: it jumps the LCG ahead to the row's first element (O(log e) 128-bit multiply-adds,
// once) and then steps through the row's T frames, one plain LCG step each.  The output is FRAME-MAJOR (T, GF_ENVS_LD)
// -- a warp stores 32 neighbouring bins of one frame, 128 bytes -- which is how the frame kernel reads it.
// ------------------------------------------------------------------------------------------------
struct GfPhiJob { float *dst; int T; int pad; unsigned long long s_hi, s_lo, i_hi, i_lo; };

typedef unsigned __int128 gf_u128;

// LCG jump: (mult, plus) of `delta` steps of x -> x * m + c
__device__ __forceinline__ void gf_lcg_jump(gf_u128 m, gf_u128 c, unsigned long long delta, gf_u128 &acc_mult, gf_u128 &acc_plus)
{
    acc_mult = 1; acc_plus = 0;
    while (delta > 0) {
        if (delta & 1ull) { acc_mult *= m; acc_plus = acc_plus * m + c; }
        c = (m + 1) * c;
        m *= m;
        delta >>= 1;
    }
}

#define GF_PHI_THREADS 128
__global__ void __launch_bounds__(GF_PHI_THREADS) gf_phi_kernel(const GfPhiJob *__restrict__ jobs)
{
    const GfPhiJob jb = jobs[blockIdx.y];
    const int k = blockIdx.x * GF_PHI_THREADS + threadIdx.x;             // bin row
    if (k >= GF_NBINS) return;
    const gf_u128 MULT = ((gf_u128)0x2360ED051FC65DA4ull << 64) | (gf_u128)0x4385DF649FCCF645ull;
    const gf_u128 inc = ((gf_u128)jb.i_hi << 64) | (gf_u128)jb.i_lo;
    gf_u128 state = ((gf_u128)jb.s_hi << 64) | (gf_u128)jb.s_lo;
    gf_u128 am, ap;
    gf_lcg_jump(MULT, inc, (unsigned long long)k * (unsigned long long)jb.T + 1ull, am, ap);   // state after the step of the row's first element
    state = am * state + ap;
    float *dst = jb.dst + k;
    for (int t = 0; t < jb.T; ++t) {
        const unsigned long long hi = (unsigned long long)(state >> 64), lo = (unsigned long long)state;
        const unsigned rot = (unsigned)(hi >> 58);                         // state >> 122
        const unsigned long long x = hi ^ lo;
        const unsigned long long r = (x >> rot) | (x << ((64u - rot) & 63u));
        const double d = __dmul_rn((double)(r >> 11), 1.0 / 9007199254740992.0);
        dst[(size_t)t * GF_ENVS_LD] = (float)__dmul_rn(6.283185307179586, d);
        state = state * MULT + inc;
    }
}

void gf_launch_phi(const GfPhiJob *jobs, int n_jobs, int max_T, cudaStream_t st)
{
    if (n_jobs <= 0 || max_T <= 0) return;
    dim3 grid((GF_NBINS + GF_PHI_THREADS - 1) / GF_PHI_THREADS, n_jobs);
    gf_phi_kernel<<<grid, GF_PHI_THREADS, 0, st>>>(jobs);
}

// ------------------------------------------------------------------------------------------------
// Host-supplied noise phases arrive as numpy draws them, (513, T) row-major (GOOFER.py:1151-1152); the frame kernel
// reads a frame's 513 bins at a time.  32 x 32 tiles through shared memory: coalesced on both sides, the
// (T, GF_ENVS_LD) result sits next to the envelope rows of the same frame.
// ------------------------------------------------------------------------------------------------
// One CTA per (pass, 32-frame tile): it walks the 17 bin tiles two at a time (eight loads in flight per thread).
__global__ void __launch_bounds__(256) gf_phi_fm_kernel(const GfPassDev *__restrict__ passes)
{
    __shared__ float tile[2][32][33];
    const GfPassDev ps = passes[blockIdx.y];
    if (!ps.phi_src) return;                                   // drawn on the device: already frame-major
    const int T = ps.T_out;
    const int t0 = blockIdx.x * 32;
    if (t0 >= T) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float *__restrict__ src = ps.phi_src;
    float *__restrict__ dst = ps.phi;
    for (int k0 = 0; k0 < GF_NBINS; k0 += 64) {
        float v[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = k0 + 32 * h + ty + 8 * i, t = t0 + tx;
                v[h][i] = (k < GF_NBINS && t < T) ? src[(size_t)k * T + t] : 0.0f;
            }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) tile[h][ty + 8 * i][tx] = v[h][i];
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int t = t0 + ty + 8 * i, k = k0 + 32 * h + tx;
                if (t < T && k < GF_NBINS) dst[(size_t)t * GF_ENVS_LD + k] = tile[h][tx][ty + 8 * i];
            }
        __syncthreads();
    }
}

void gf_launch_phi_fm(const GfPassDev *passes, int n_pass, int max_T, cudaStream_t st)
{
    if (n_pass <= 0 || max_T <= 0) return;
    dim3 grid((max_T + 31) / 32, n_pass);
    gf_phi_fm_kernel<<<grid, 256, 0, st>>>(passes);
}
