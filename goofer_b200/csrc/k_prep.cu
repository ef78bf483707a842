// k_prep.cu -- feature preparation kernels (everything GooferResampler.resample does to the envelope,
// mask and formant tracks before it calls gf.synthesize, plus synthesize's own envelope warps).
//
//   gf_src_env_kernel   gf.decode_env_from_knots                       GOOFER.py:149-168 (84-95)
//   gf_tracks_kernel    formant tracks: loop / stretch / canon / sanitise+smooth
//                                                                       SillySampler.py:714-763, 776-792, 242-283
//   gf_env_kernel       br tilt :503-515, es :518-551, fw :554-574, loop/velocity maps :625-788,
//                       fst bells :808-832, fry warp :967-995; then synthesize's env4breath blur
//                       GOOFER.py:993, F1-F4 warp :1004-1014 (:805-875), g shift :1016-1017 (:618-627)
//   gf_mask_kernel      mask_new (tile + velocity stretch)             SillySampler.py:699-712, 788
//   gf_fir_kernel       gaussian_filter1d along time (numpy-'reflect')  GOOFER.py:241-261
#include <cuda_fp16.h>
#include <cuda_pipeline.h>
#include "gf_device.cuh"
#include "gf_kernels.h"
#include "gf_maps.cuh"

// ------------------------------------------------------------------------------------------------
// source envelope: knots (K, T) f16 log-values -> dense (T, 520) f32, frame-major
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gf_half_bits_to_float(uint16_t h)
{
    return __half2float(__ushort_as_half(h));
}

__global__ void __launch_bounds__(256) gf_src_env_kernel(const GfSourceDev *__restrict__ srcs)
{
    __shared__ float tile[32][33];
    const GfSourceDev s = srcs[blockIdx.y];
    const int t0 = blockIdx.x * 32;
    if (t0 >= s.T) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const double fstep = 1024.0 * (1.0 / (double)d_tab.sr);     // rfftfreq: k / (n * d)
    // one 32-frame x 32-bin tile per CTA (blockIdx.z = bin chunk): 17 times the CTAs of a loop over the chunks; the
    // binary search is a chain of dependent loads and the kernel is latency bound
    {
        const int bc = blockIdx.z * 32;
        for (int i = ty; i < 32; i += 8) {
            const int b = bc + i, t = t0 + tx;
            float v = 0.0f;
            if (b < GF_NBINS && t < s.T) {
                if (s.knots) {
                    // precompute_interp_matrix (GOOFER.py:84-95): 2-tap lerp between mel knots, all f32
                    const float f = (float)((double)b / fstep);
                    int lo = 0, hi = s.K;                       // searchsorted(hz, f, side='right')
                    while (lo < hi) { int mid = (lo + hi) >> 1; if (s.hz_knots[mid] <= f) lo = mid + 1; else hi = mid; }
                    int idx = lo - 1;
                    idx = idx < 0 ? 0 : (idx > s.K - 2 ? s.K - 2 : idx);
                    const float x0 = s.hz_knots[idx], x1 = s.hz_knots[idx + 1];
                    const float w1 = (f - x0) / fmaxf(x1 - x0, 1e-12f);
                    const float w0 = 1.0f - w1;
                    const float k0 = gf_half_bits_to_float(s.knots[(size_t)idx * s.T + t]);
                    const float k1 = gf_half_bits_to_float(s.knots[(size_t)(idx + 1) * s.T + t]);
                    v = expf(fmaf(w1, k1, w0 * k0));
                } else {
                    v = s.dense[(size_t)b * s.T + t];
                }
            }
            tile[i][tx] = v;
        }
        __syncthreads();
        for (int i = ty; i < 32; i += 8) {
            const int t = t0 + i, b = bc + tx;
            if (t < s.T && b < GF_NBINS) s.envS[(size_t)t * GF_ENVS_LD + b] = tile[tx][i];
        }
    }
}

void gf_launch_src_env(const GfSourceDev *srcs, int n_src, int max_T, cudaStream_t st)
{
    if (n_src <= 0 || max_T <= 0) return;
    dim3 grid((max_T + 31) / 32, n_src, (GF_NBINS + 31) / 32);
    gf_src_env_kernel<<<grid, 256, 0, st>>>(srcs);
}

#define GF_MAX_ES_TAPS 64   // es radius <= 28 -> 57 taps, zero-padded to a multiple of 8
// layout of GfNoteDev.env_aux (f32): per-note tables shared by all the envelope-kernel CTAs of the note
#define GF_AUX_TILT 0
#define GF_AUX_FREQ 520
#define GF_AUX_TAPS 1040
#define GF_AUX_LEN (1040 + GF_MAX_ES_TAPS)

// ------------------------------------------------------------------------------------------------
// formant tracks
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double gf_gauss_tap(int j, int radius, double sigma, double norm)
{
    const double t = (double)(j - radius) / sigma;
    return exp(-0.5 * t * t) / norm;
}

__global__ void __launch_bounds__(128)
gf_tracks_kernel(const GfNotePlan *__restrict__ plans, const GfNoteDev *__restrict__ notes, const GfSourceDev *__restrict__ srcs)
{
    const GfNotePlan &pl = plans[blockIdx.x];
    const GfNoteDev nd = notes[blockIdx.x];
    const GfSourceDev &sc = srcs[pl.src];
    const int T = pl.T_env;
    __shared__ int any_bad, any_good;
    __shared__ double taps[33];
    __shared__ double ev4[33], eves[GF_MAX_ES_TAPS];          // un-normalised taps: one exp per thread, summed in index order by all
    double es_sigma = 0.0;
    int es_radius = 0;
    if (pl.es != 0.0) {
        const double s = fabs(pl.es);
        es_sigma = pl.es < 0.0 ? (1.0 + 6.0 * s) : (0.8 + 4.0 * s);
        es_radius = (int)(4.0 * es_sigma + 0.5);
    }
    if (threadIdx.x < 33) { const double t = (double)((int)threadIdx.x - 16) / 4.0; ev4[threadIdx.x] = exp(-0.5 * t * t); }
    else if (threadIdx.x >= 64 && threadIdx.x - 64 <= 2 * es_radius && pl.es != 0.0) {
        const double q = (double)((int)threadIdx.x - 64 - es_radius) / es_sigma;
        eves[threadIdx.x - 64] = exp(-0.5 * q * q);
    }
    __syncthreads();
    if (threadIdx.x < 33) {
        double norm = 0.0;
        for (int j = 0; j < 33; ++j) norm += ev4[j];
        taps[threadIdx.x] = gf_gauss_tap(threadIdx.x, 16, 4.0, norm);
    }
    // ---- per-note tables of the envelope kernel: br tilt, es taps, f32 bin frequencies ----
    {
        __shared__ double red[4];
        float *aux = nd.env_aux;
        const int sr = pl.sr;
        const double nyq = (double)sr / 2.0, step = nyq / 512.0;
        for (int b = threadIdx.x; b < GF_NBINS; b += blockDim.x)
            aux[GF_AUX_FREQ + b] = (b == 512) ? (float)nyq : (float)((double)b * step);      // linspace(0, sr/2, 513) f32
        if (pl.brightness_env != 1.0) {
            // SillySampler.py:506-510: f32 linspace(1e-6, nyq), clip(f / nyq, .02, 1) ** alpha, / (mean + 1e-12)
            const float alpha = (float)fmin(fmax(pl.brightness_env - 1.0, -0.9), 1.0);
            const float nyqf = (float)((double)sr * 0.5);
            double part = 0.0;
            float tv[5];
            int cnt = 0;
            for (int b = threadIdx.x; b < GF_NBINS; b += blockDim.x, ++cnt) {
                const double fv = (b == GF_NBINS - 1) ? (double)sr * 0.5 : (double)b * (((double)sr * 0.5 - 1e-6) / 512.0) + 1e-6;
                float nf = (float)fv / nyqf;
                nf = fminf(fmaxf(nf, 0.02f), 1.0f);
                tv[cnt] = powf(nf, alpha);
                part += (double)tv[cnt];
            }
            part = gf_warp_sum(part);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
            __syncthreads();
            const float mean = (float)((red[0] + red[1] + red[2] + red[3]) / (double)GF_NBINS);
            cnt = 0;
            for (int b = threadIdx.x; b < GF_NBINS; b += blockDim.x, ++cnt) aux[GF_AUX_TILT + b] = tv[cnt] / (mean + 1e-12f);
        }
        if (pl.es != 0.0 && threadIdx.x < GF_MAX_ES_TAPS) {
            const double sigma = es_sigma;
            const int radius = es_radius;
            float t = 0.0f;
            if (threadIdx.x < 2 * radius + 1) {
                double norm = 0.0;
                for (int j = 0; j <= 2 * radius; ++j) norm += eves[j];
                t = (float)gf_gauss_tap(threadIdx.x, radius, sigma, norm);
            }
            aux[GF_AUX_TAPS + threadIdx.x] = t;
        }
    }
    const float min_hz[4] = {120.0f, 300.0f, 1500.0f, 2000.0f};
    const float max_hz = (float)((double)pl.sr * 0.48);
    for (int k = 0; k < 4; ++k) {
        float *canon = nd.trk_canon + (size_t)k * T;
        const GfTrackSlices sl = gf_track_slices(pl, k);
        for (int t = threadIdx.x; t < T; t += blockDim.x)
            canon[t] = gf_track_canon(pl, sl, sc.formants[k], k, t);
        if (!nd.trk_clean) continue;
        // sanitize_smooth_formant (SillySampler.py:264-283)
        float *tmp = nd.trk_clean + (size_t)(4 + k) * T;     // scratch rows 4..7
        float *clean = nd.trk_clean + (size_t)k * T;
        if (threadIdx.x == 0) { any_bad = 0; any_good = 0; }
        __syncthreads();
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            const float x = canon[t];
            const bool bad = !isfinite(x) || x < min_hz[k] || x > max_hz;
            if (bad) any_bad = 1; else any_good = 1;
        }
        __syncthreads();
        const bool hb = any_bad != 0, hg = any_good != 0;
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            float x = canon[t];
            if (hb) {
                if (!hg) x = 300.0f;
                else {
                    auto is_bad = [&](int q) { const float y = canon[q]; return !isfinite(y) || y < min_hz[k] || y > max_hz; };
                    if (is_bad(t)) {
                        // interp1d(good_idx, x[good], 'linear', extrapolate) at t   (GOOFER.py:173-239)
                        int gl = t - 1; while (gl >= 0 && is_bad(gl)) --gl;
                        int gr = t + 1; while (gr < T && is_bad(gr)) ++gr;
                        double r;
                        if (gl >= 0 && gr < T) {
                            const double y0 = canon[gl], y1 = canon[gr];
                            r = ((y1 - y0) / ((double)(float)gr - (double)(float)gl)) * ((double)(float)t - (double)(float)gl) + y0;
                        } else if (gl < 0) {
                            int g2 = gr + 1; while (g2 < T && is_bad(g2)) ++g2;
                            if (g2 >= T) r = canon[gr];
                            else {
                                const double sl_ = ((double)canon[g2] - (double)canon[gr]) / ((double)(float)g2 - (double)(float)gr + 1e-10);
                                r = (double)canon[gr] + sl_ * ((double)(float)t - (double)(float)gr);
                            }
                        } else {
                            int g2 = gl - 1; while (g2 >= 0 && is_bad(g2)) --g2;
                            if (g2 < 0) r = canon[gl];
                            else {
                                const double sr_ = ((double)canon[gl] - (double)canon[g2]) / ((double)(float)gl - (double)(float)g2 + 1e-10);
                                r = (double)canon[gl] + sr_ * ((double)(float)t - (double)(float)gl);
                            }
                        }
                        x = (float)r;
                    }
                }
            }
            tmp[t] = x;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            double a = 0.0;
            if (t >= 16 && t + 16 < T) {                      // interior: no reflection
                const float *q = tmp + (t - 16);
#pragma unroll
                for (int j = 0; j < 33; ++j) a += taps[j] * (double)q[j];
            } else {
                for (int j = 0; j < 33; ++j) a += taps[j] * (double)tmp[gf_reflect(t + j - 16, T)];
            }
            clean[t] = (float)a;
        }
        __syncthreads();
    }
}

void gf_launch_tracks(const GfNotePlan *plans, const GfNoteDev *notes, const GfSourceDev *srcs, int n_notes, cudaStream_t st)
{
    if (n_notes > 0) gf_tracks_kernel<<<n_notes, 128, 0, st>>>(plans, notes, srcs);
}

// ------------------------------------------------------------------------------------------------
// envelope: one warp per output frame, eight frames per CTA.  Lane L owns the 17 contiguous bins
// [17 L, 17 L + 17) (513 = 30 * 17 + 3: lane 30 owns three bins, lane 31 none), which makes every FIR a
// register-tiled sliding window over a reflect-padded shared-memory row (stride 17 is conflict free).
// The GATHER stages (fw / fry / F1-F4 / g resampling: reads at b * ratio) use the other mapping, lane L owns bins
// L, L + 32, .. : neighbouring lanes then read neighbouring words, whereas 17 L * ratio collides whenever the
// scaled stride shares a factor with 32 (ncu, round 1: 31 M bank conflicts, 23 % of the kernel's shared-memory
// wavefronts, all on the gather lines).  Stages hand rows over through shared memory, so each picks its own mapping.
// Values are f32 like the arrays the reference stores; positions / abscissae stay fp64.
// ------------------------------------------------------------------------------------------------
#define GF_ENV_WARPS GF_FT
#define GF_EPL 17           // bins per lane
#define GF_ROW_L 32         // left padding of a row (reflect halo, radius <= 28)
#define GF_ROW_LEN 600      // 32 + 513 + 55: the FIR windows of the last lanes read up to bin index 561

#ifndef GF_ENV_FST_SKIP
#define GF_ENV_FST_SKIP 0
#endif
// unroll factor of the gather loops that write their row straight back to shared memory (fry, F1-F4, g): full unrolling
// (17 copies each) is what made the kernel 150 KB of code
#ifndef GF_ENV_GU
#define GF_ENV_GU 17
#endif
#define GF_STR_(x) #x
#define GF_UNROLL_(n) _Pragma(GF_STR_(unroll n))
#define GF_ENV_GATHER_UNROLL GF_UNROLL_(GF_ENV_GU)

// gf_env_mix (gf_maps.cuh: loop / cross-fade / stretch / velocity maps, ~1,000 instructions inlined) runs once per frame
// and warp: called out of line so that it does not sit in the instruction stream of the per-bin hot loop (the kernel's
// 150 KB of code overflowed the instruction cache: 0.98 "no instruction" stalls per issue in ncu, round 2)
__device__ __noinline__ void gf_env_mix_ool(const GfNotePlan &p, int t, GfMix &m) { gf_env_mix(p, t, m); }

#ifndef GF_ENV_TPC
#define GF_ENV_TPC 1                // tiles (of GF_FT frames) per CTA.  With 2..4 the CTA prologue (a chain of dependent record loads,
                                    // the per-note tables, a barrier) is paid once per GF_ENV_TPC frames of a warp and, from the second
                                    // frame on, the source row is already in shared memory (cp.async, see `nxt`).  Measured on B200
                                    // (round 2, c2): 1.251 / 1.279 / 1.343 / 1.340 ms at 1 / 2 / 3 / 4 tiles -- one tile per CTA stays
#endif
struct GfEnvSmem {
    float rows[GF_ENV_WARPS][2][GF_ROW_LEN];
    float tilt[GF_ENVS_LD];
    float freq[GF_ENVS_LD];
    float es_taps[GF_MAX_ES_TAPS];
    double knots[GF_ENV_WARPS][18];                       // F1..F4 warp of the warp's frame: xd[0..5], xs[6..11], slopes [12..16]
#if GF_ENV_TPC > 1
    __align__(16) float nxt[GF_ENV_WARPS][GF_ENVS_LD];    // the source row of the warp's NEXT frame, fetched by cp.async while this one is shaped
#endif
};

// fill the reflect halo of a row whose bins 0..512 are valid (numpy 'reflect': edge sample not repeated)
__device__ __forceinline__ void gf_row_halo(float *row /* -> bin 0 */, int lane, int radius)
{
    for (int q = lane + 1; q <= radius; q += 32) { row[-q] = row[q]; row[512 + q] = row[512 - q]; }
    __syncwarp();
}

// sliding-window FIR with runtime radius: out[e] = sum_j taps[j] * row[b0 + e + j - radius], taps zero-padded to x8
template <typename T>
__device__ __forceinline__ void gf_fir_rt(const T *row, int b0, const T *taps, int radius, T *out)
{
#pragma unroll
    for (int e = 0; e < GF_EPL; ++e) out[e] = (T)0;
    const int ntap = 2 * radius + 1;
    for (int j0 = 0; j0 < ntap; j0 += 8) {
        T win[GF_EPL + 7];
        const T *p = row + b0 + j0 - radius;
#pragma unroll
        for (int q = 0; q < GF_EPL + 7; ++q) win[q] = p[q];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const T w = taps[j0 + jj];
#pragma unroll
            for (int e = 0; e < GF_EPL; ++e) out[e] = fma(w, win[e + jj], out[e]);
        }
    }
}

// linear interpolation of a 513-bin row at fractional bin position u in [0, 512]: the position is fp64, the lerp f32
// (the reference evaluates slope * (x - x0) + y0 in fp64 and stores f32: at most one f32 ulp apart)
__device__ __forceinline__ float gf_grid_lerp(const float *row, double u)
{
    int j = (int)u;
    if (j > 511) j = 511;
    const float t = (float)(u - (double)j);
    const float y0 = row[j];
    return fmaf(t, row[j + 1] - y0, y0);
}

// np.interp on the uniform grid freqs[i] = i * step (freqs[512] = nyq) with the linear extrapolation of
// GOOFER.py:173-239 outside [0, nyq]
__device__ __forceinline__ float gf_grid_interp(const float *row, double x, double step, double inv_step, double nyq)
{
    if (x < 0.0) {
        const double sl = ((double)row[1] - (double)row[0]) / (step - 0.0 + 1e-10);
        return (float)((double)row[0] + sl * (x - 0.0));
    }
    if (x > nyq) {
        const double x1 = 511.0 * step;
        const double sr_ = ((double)row[512] - (double)row[511]) / (nyq - x1 + 1e-10);
        return (float)((double)row[512] + sr_ * (x - nyq));
    }
    // piece-wise linear interpolation is continuous, so a bracket that is off by one at a grid point (x * inv_step
    // rounds across an integer) yields the same value: no fix-up of j is needed, and t comes from the same product
    return gf_grid_lerp(row, x * inv_step);
}


#ifndef GF_ENV_CTAS
#define GF_ENV_CTAS 3               // 80 registers (64 B of spills), 44 KB of shared memory per CTA: 1.94 -> 1.69 ms against two CTAs at 127
#endif
#ifdef GF_ENV_MAXNREG
__global__ void __maxnreg__(GF_ENV_MAXNREG)
#else
__global__ void __launch_bounds__(32 * GF_ENV_WARPS, GF_ENV_CTAS)
#endif
gf_env_kernel(const int2 *__restrict__ work, const GfNotePlan *__restrict__ plans, const GfNoteDev *__restrict__ notes,
              const GfSourceDev *__restrict__ srcs)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GfEnvSmem &sm = *reinterpret_cast<GfEnvSmem *>(smem_raw);
    const int2 wk = work[blockIdx.x];                 // x = note (in wave), y = tile index
    const GfNotePlan &pl = plans[wk.x];
    const GfNoteDev nd = notes[wk.x];
    const GfSourceDev &sc = srcs[pl.src];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sr = pl.sr;
    const double nyq = (double)sr / 2.0;
    const double step = nyq / 512.0;
    const double inv_step = 1.0 / step;

    const bool do_tilt = pl.brightness_env != 1.0;
    const bool do_es = pl.es != 0.0;
    const bool do_fw = pl.fw != 0.0;
    int es_radius = 0;
    if (do_es) {
        const double s = fabs(pl.es);
        const double sigma = pl.es < 0.0 ? (1.0 + 6.0 * s) : (0.8 + 4.0 * s);
        es_radius = (int)(4.0 * sigma + 0.5);
    }
    // A CTA renders GF_ENV_TPC consecutive tiles of its note: the chain of dependent record loads, the table loads and the
    // barrier of the prologue are paid once.  The first source frame of the warp's first output frame is requested before
    // anything else: its latency hides behind the table loads and the barrier.
    const int tile0 = wk.y * GF_ENV_TPC;
    GfMix mix;
    mix.n = 0;
    float pre[GF_EPL];
#pragma unroll
    for (int e = 0; e < GF_EPL; ++e) pre[e] = 0.0f;
    {
        const int t = tile0 * GF_FT + warp;
        if (t < pl.T_out) {
            gf_env_mix_ool(pl, min(t, pl.T_env - 1), mix);
            const float *src0 = sc.envS + (size_t)gf_src_frame(pl, mix.f[0]) * GF_ENVS_LD;
#pragma unroll
            for (int e = 0; e < GF_EPL; ++e) { const int b = lane + 32 * e; if (b < GF_NBINS) pre[e] = src0[b]; }
        }
    }
#if GF_ENV_TPC > 1
    // source row of frame t -> the warp's `nxt` buffer, asynchronously (16-byte cp.async, no registers held)
    auto fetch_next = [&](int t) {
        if (t >= pl.T_out) return;
        GfMix mx;
        gf_env_mix_ool(pl, min(t, pl.T_env - 1), mx);
        const float *src = sc.envS + (size_t)gf_src_frame(pl, mx.f[0]) * GF_ENVS_LD;
        for (int c = lane; c < GF_ENVS_LD / 4; c += 32) __pipeline_memcpy_async(&sm.nxt[warp][4 * c], src + 4 * c, 16);
        __pipeline_commit();
    };
    fetch_next((tile0 + 1) * GF_FT + warp);
#endif
    // per-note tables prepared by gf_tracks_kernel
    for (int b = threadIdx.x; b < GF_NBINS; b += blockDim.x) {
        sm.freq[b] = nd.env_aux[GF_AUX_FREQ + b];
        if (do_tilt) sm.tilt[b] = nd.env_aux[GF_AUX_TILT + b];
    }
    if (do_es && threadIdx.x < GF_MAX_ES_TAPS) sm.es_taps[threadIdx.x] = nd.env_aux[GF_AUX_TAPS + threadIdx.x];
    __syncthreads();

    float *rA = sm.rows[warp][0] + GF_ROW_L, *rB = sm.rows[warp][1] + GF_ROW_L;
    // the FIR windows also touch cells outside [-radius, 512 + radius] (zero-padded taps, idle lanes): they
    // must hold finite values, 0 * NaN left over from an earlier kernel would poison the sums
    // (only the pads: the 513 bins of a row are always written before they are read)
    for (int q = lane; q < GF_ROW_LEN - GF_NBINS; q += 32) {
        const int c = q < GF_ROW_L ? q : q + GF_NBINS;                 // [0, 32) and [545, 600) of a row
        sm.rows[warp][0][c] = 0.0f;
        sm.rows[warp][1][c] = 0.0f;
    }
    __syncwarp();
    const int b0 = min(GF_EPL * lane, 510);               // lane 31 owns nothing: it shadows lane 30 (reads stay inside the row)
    const int nown = (lane == 31) ? 0 : min(GF_EPL, GF_NBINS - b0);   // bins this lane owns

    for (int it = 0; it < GF_ENV_TPC; ++it) {
    const int t = (tile0 + it) * GF_FT + warp;
    if (t >= pl.T_out) break;                             // warp-uniform; no CTA-wide barrier below
    const int te = min(t, pl.T_env - 1);                  // GOOFER.py:1115-1119 trim / edge-pad to the STFT grid
    // the frame's formant tracks (fst bells: lanes 0-3, F1-F4 warp: lanes 4-7) are requested now and handed round by
    // shuffle where they are used: one register instead of eight dependent loads in the middle of the frame
    float trk = 0.0f;
    if (lane < 4) { if (pl.any_fst && !(fabs(pl.fst[lane]) < 1e-6)) trk = nd.trk_clean[(size_t)lane * pl.T_env + te]; }
    else if (lane < 8) { if (pl.any_F_shift) trk = nd.trk_canon[(size_t)(lane - 4) * pl.T_env + te]; }
#if GF_ENV_TPC > 1
    if (it > 0) {
        gf_env_mix_ool(pl, te, mix);
        __pipeline_wait_prior(0);                         // this frame's source row, requested while the previous frame was shaped
        __syncwarp();                                     // (also: the previous frame's row reads are done before the rows are rewritten)
#pragma unroll
        for (int e = 0; e < GF_EPL; ++e) { const int b = lane + 32 * e; pre[e] = (b < GF_NBINS) ? sm.nxt[warp][b] : 0.0f; }
        __syncwarp();
        if (it + 1 < GF_ENV_TPC) fetch_next((tile0 + it + 1) * GF_FT + warp);
    }
#endif

    float acc[GF_EPL];
#pragma unroll
    for (int e = 0; e < GF_EPL; ++e) acc[e] = 0.0f;
    for (int m = 0; m < mix.n; ++m) {
        const float *src = sc.envS + (size_t)gf_src_frame(pl, mix.f[m]) * GF_ENVS_LD;
        float *cur = rA, *oth = rB;
        if (m == 0) {
#pragma unroll
            for (int e = 0; e < GF_EPL; ++e) { const int b = lane + 32 * e; if (b < GF_NBINS) cur[b] = do_tilt ? pre[e] * sm.tilt[b] : pre[e]; }
        } else {
            for (int b = lane; b < GF_NBINS; b += 32) cur[b] = do_tilt ? src[b] * sm.tilt[b] : src[b];
        }
        __syncwarp();
        if (do_es) {
            // SillySampler.py:518-551: blur (es<0) or unsharp mask (es>0) along frequency, then per-frame mean match
            gf_row_halo(cur, lane, es_radius);
            float mod[GF_EPL];
            gf_fir_rt<float>(cur, b0, sm.es_taps, es_radius, mod);
            const float s5 = (float)(5.0 * fabs(pl.es));
            float sum0 = 0.0f, sum1 = 0.0f;
#pragma unroll
            for (int e = 0; e < GF_EPL; ++e) {
                if (e < nown) {
                    const float x = cur[b0 + e];
                    if (pl.es > 0.0) mod[e] = fmaxf(0.0f, fmaf(s5, x - mod[e], x));
                    sum0 += x;
                    sum1 += mod[e];
                }
            }
            const double t0 = gf_warp_sum((double)sum0), t1 = gf_warp_sum((double)sum1);
            const float m0 = (float)(t0 / (double)GF_NBINS);
            const double m1 = t1 / (double)GF_NBINS;
            const float ratio = (float)((double)m0 / (m1 + 1e-12));
#pragma unroll
            for (int e = 0; e < GF_EPL; ++e)
                if (e < nown) { const float y = mod[e] * ratio; oth[b0 + e] = pl.es < 0.0 ? fmaxf(0.0f, y) : y; }
            __syncwarp();
            float *sw = cur; cur = oth; oth = sw;
        }
        const float wm = (float)mix.w[m];
        if (do_fw) {
            // SillySampler.py:555-569: resample at (b - 256.5) (1 + fw) + 256.5, clipped, 2-tap lerp
#pragma unroll
            for (int e = 0; e < GF_EPL; ++e) {
                const int b = lane + 32 * e;
                if (b < GF_NBINS) {
                    // np.clip(pos, 0, 512): the position is increasing in b, so the clip is a pair of integer selects on the
                    // RESULT (row[0] below, row[512] above) instead of two fp64 min / max on the position
                    const double pos = ((double)b - 513.0 / 2.0) * (1.0 + pl.fw) + 513.0 / 2.0;
                    const int lo = min(max((int)pos, 0), 511);
                    const float fr = (float)(pos - (double)lo);
                    float y = fmaf(fr, cur[lo + 1] - cur[lo], cur[lo]);
                    y = (pos <= 0.0) ? cur[0] : y;
                    y = (pos >= 512.0) ? cur[512] : y;
                    acc[e] = fmaf(wm, y, acc[e]);
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < GF_EPL; ++e) {
                const int b = lane + 32 * e;
                if (b < GF_NBINS) acc[e] = fmaf(wm, cur[b], acc[e]);
            }
        }
        __syncwarp();
    }
    // ---- fst bells (SillySampler.py:808-832), f32 ----
    if (pl.any_fst) {
        float Fk[4], isg[4], sv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float sig = (k == 0) ? 100.0f : (k == 1) ? 200.0f : (k == 2) ? 350.0f : 500.0f;
            const double sk = pl.fst[k];
            float fk = 0.0f;
            bool on = !(fabs(sk) < 1e-6);
            const float fk_pf = __shfl_sync(0xffffffffu, trk, k);   // requested at the top of the frame
            if (on) {
                fk = fk_pf;
                on = isfinite(fk) && (fk > 50.0f) && ((double)fk < (double)sr * 0.5);
            }
            Fk[k] = fk;
            isg[k] = 1.0f / sig;
            sv[k] = on ? (float)((1.0 + sk) - 1.0) : 0.0f;      // an inactive formant multiplies by exactly 1
        }
        // A bell further than 6 sigma away multiplies by EXACTLY 1.0f (|s| e^-18 = 1.5e-8 is below half an ulp of 1), so it is
        // skipped -- bit for bit the same gain.  With the interleaved mapping the 32 lanes hold 32 NEIGHBOURING bins at every
        // step e (1,378 Hz), so "is any of them within 6 sigma of F_k" is a warp-uniform test on the block's edge
        // frequencies: on average five of six bell evaluations (sub, 3 mul, ex2, fma) disappear (the bells span 600 / 1,200 /
        // 2,100 / 3,000 Hz of the 22 kHz).  (Round 1 tried this with the contiguous mapping, where the test is per lane and
        // cost what it saved.)  Measured on B200 (round 2, c2): 1.20 ms with the skip against 1.165 ms without -- 68 uniform
        // branches per frame break the interleaving of the independent bell evaluations; off by default (-DGF_ENV_FST_SKIP=1).
        float blo[4], bhi[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float sig = (k == 0) ? 100.0f : (k == 1) ? 200.0f : (k == 2) ? 350.0f : 500.0f;
            blo[k] = (sv[k] != 0.0f) ? Fk[k] - 6.0f * sig : 3.0e38f;          // inactive: never in range
            bhi[k] = (sv[k] != 0.0f) ? Fk[k] + 6.0f * sig : -3.0e38f;
        }
#pragma unroll
        for (int e = 0; e < GF_EPL; ++e) {
            const float fb = sm.freq[min(lane + 32 * e, 512)];
            const float f_lo = sm.freq[32 * e], f_hi = sm.freq[min(32 * e + 31, 512)];      // uniform: the block's edge bins
            float gain = 1.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#if GF_ENV_FST_SKIP
                if (bhi[k] > f_lo && blo[k] < f_hi)                            // warp-uniform
#endif
                {
                    const float d = (fb - Fk[k]) * isg[k];
                    gain *= fmaf(sv[k], __expf(-0.5f * (d * d)), 1.0f);
                }
            }
            acc[e] *= gain;
        }
    }
    float *cur = rA, *oth = rB;
#pragma unroll
    for (int e = 0; e < GF_EPL; ++e) { const int b = lane + 32 * e; if (b < GF_NBINS) cur[b] = acc[e]; }
    __syncwarp();
    // ---- vocal-fry envelope compression (SillySampler.py:967-995) ----
    if (pl.fry_mask_on) {
        const int c = min(pl.n_total - 1, te * GF_HOP + GF_HOP / 2);
        const float wfr = gf_fry_at(pl, c);
        if (wfr > 1e-6f) {
            const double s = 1.0 - (double)wfr * (1.0 - 0.92);
            if (!(fabs(s - 1.0) < 1e-6)) {
                const double inv_s = 1.0 / s;
GF_ENV_GATHER_UNROLL
                for (int e = 0; e < GF_EPL; ++e) {
                    const int b = lane + 32 * e;
                    if (b < GF_NBINS) {
                        const double sp = (double)b * inv_s;                 // >= 0; clipped to 512 on the result
                        const int lo = min((int)sp, 511);
                        const float fr = (float)(sp - (double)lo);
                        const float y = fmaf(fr, cur[lo + 1] - cur[lo], cur[lo]);
                        oth[b] = (sp >= 512.0) ? cur[512] : y;
                    }
                }
                __syncwarp();
                float *sw = cur; cur = oth; oth = sw;
            }
        }
    }
    // ---- env4breath: Gaussian sigma 1.75 of env_new (GOOFER.py:993), before the formant warps ----
    {
        gf_row_halo(cur, lane, 7);
        float g[15];
#pragma unroll
        for (int j = 0; j < 15; ++j) g[j] = (float)d_tab.g175[j];
        float win[GF_EPL + 14];
#pragma unroll
        for (int q = 0; q < GF_EPL + 14; ++q) win[q] = cur[b0 - 7 + q];
#pragma unroll
        for (int e = 0; e < GF_EPL; ++e) {
            float a = 0.0f;
#pragma unroll
            for (int j = 0; j < 15; ++j) a = fmaf(g[j], win[e + j], a);
            if (e < nown) oth[b0 + e] = a;
        }
        __syncwarp();
        float *dstN = nd.envN + (size_t)t * GF_ENVS_LD;
        for (int b = lane; b < GF_NBINS; b += 32) dstN[b] = oth[b];
        __syncwarp();
    }
    // ---- F1..F4 warp (GOOFER.py:840-875) ----
    // The knots (0,0), (F_k r_k -> F_k) for the valid formants, (nyq, nyq) belong to the frame, i.e. to the whole warp:
    // they live in a per-warp shared-memory table, lane j computes the slope, the order check and the first bin of
    // segment j + 1, and every bin finds its segment by counting integer thresholds (no per-bin search, no local arrays).
    if (pl.any_F_shift) {
        double *kx = sm.knots[warp];
        int nk = 1;
        if (lane == 0) { kx[0] = 0.0; kx[6] = 0.0; }
        for (int k = 0; k < 4; ++k) {
            const double fo = (double)__shfl_sync(0xffffffffu, trk, 4 + k);
            const double fs = fo * pl.F_shift[k];
            if (fo > 50.0 && fo < nyq && fs > 50.0) {         // warp-uniform
                if (lane == 0) { kx[nk] = fs; kx[6 + nk] = fo; }
                ++nk;
            }
        }
        if (lane == 0) { kx[nk] = nyq; kx[6 + nk] = nyq; }
        ++nk;
        __syncwarp();
        bool ok = true;
        int th = 1 << 30;                                     // first bin b with b * step >= xd[lane + 1]
        if (lane < nk - 1) {
            const double d0 = kx[lane], d1 = kx[lane + 1];
            kx[12 + lane] = (kx[6 + lane + 1] - kx[6 + lane]) / (d1 - d0);
            ok = d0 <= d1;
            int bt = (int)fmin(fmax(ceil(d1 * inv_step), 0.0), 513.0);
            while (bt > 0 && (double)(bt - 1) * step >= d1) --bt;
            while (bt <= 512 && (double)bt * step < d1) ++bt;
            th = bt;
        }
        const bool mono = __all_sync(0xffffffffu, ok);
        __syncwarp();
        const int th1 = __shfl_sync(0xffffffffu, th, 0), th2 = __shfl_sync(0xffffffffu, th, 1), th3 = __shfl_sync(0xffffffffu, th, 2),
                  th4 = __shfl_sync(0xffffffffu, th, 3), th5 = __shfl_sync(0xffffffffu, th, 4);
        if (mono) {
GF_ENV_GATHER_UNROLL
            for (int e = 0; e < GF_EPL; ++e) {
                const int b = lane + 32 * e;
                if (b < GF_NBINS) {
                    const double x = (double)b * step;       // 512 * step == nyq exactly
                    // np.interp(x, xd, xs): segment j = number of knots 1 .. nk-1 at or below x
                    const int j = (b >= th1) + (b >= th2) + (b >= th3) + (b >= th4) + (b >= th5);
                    double wf;
                    if (j >= nk - 1) wf = kx[6 + nk - 1];
                    else {
                        const double xdj = kx[j], xsj = kx[6 + j];
                        wf = (xdj == x) ? xsj : kx[12 + j] * (x - xdj) + xsj;
                    }
                    // ascending knots: wf lies in [0, nyq] up to an ulp, where extrapolation and the end lerp coincide to 1e-12
                    oth[b] = gf_grid_lerp(cur, wf * inv_step);
                }
            }
        } else {
            // shifted formants out of order: numpy's search on unsorted knots, reproduced as a linear scan from the left
            for (int b = lane; b < GF_NBINS; b += 32) {
                const double x = (b == 512) ? nyq : (double)b * step;
                double wf;
                if (x > kx[nk - 1]) wf = kx[6 + nk - 1];
                else if (x < kx[0]) wf = kx[6];
                else {
                    int i = 0;
                    while (i < nk && x >= kx[i]) ++i;
                    const int j = i - 1;
                    if (j >= nk - 1) wf = kx[6 + nk - 1];
                    else if (kx[j] == x) wf = kx[6 + j];
                    else wf = kx[12 + j] * (x - kx[j]) + kx[6 + j];
                }
                oth[b] = gf_grid_interp(cur, wf, step, inv_step, nyq);
            }
        }
        __syncwarp();
        float *sw = cur; cur = oth; oth = sw;
    }
    // ---- g: shift all formants (GOOFER.py:618-627) ----
    if (pl.formant_shift != 1.0) {
        const double inv_r = 1.0 / pl.formant_shift;
GF_ENV_GATHER_UNROLL
        for (int e = 0; e < GF_EPL; ++e) {
            // freqs / ratio on the freqs grid: in bins that is b / ratio, clipped to [0, 512]
            const int b = lane + 32 * e;
            if (b < GF_NBINS) {
                const double u = (double)b * inv_r;                          // >= 0; np.clip(.., 0, nyq) on the result
                const float y = gf_grid_lerp(cur, u);
                oth[b] = (u >= 512.0) ? cur[512] : y;
            }
        }
        __syncwarp();
        float *sw = cur; cur = oth; oth = sw;
    }
    float *dstF = nd.envF + (size_t)t * GF_ENVS_LD;
    for (int b = lane; b < GF_NBINS; b += 32) dstF[b] = cur[b];
    }   // tiles of this CTA
}

void gf_launch_env(const int2 *work, int n_work, const GfNotePlan *plans, const GfNoteDev *notes, const GfSourceDev *srcs,
                   cudaStream_t st)
{
    if (n_work <= 0) return;
    static GfSmemLimit memo;
    gf_smem_limit(gf_env_kernel, sizeof(GfEnvSmem), memo);
    gf_env_kernel<<<n_work, 32 * GF_ENV_WARPS, sizeof(GfEnvSmem), st>>>(work, plans, notes, srcs);
}

// ------------------------------------------------------------------------------------------------
// mask_new -> vm (f32), and the decimated + smoothed mask of smooth_mask_ds (GOOFER.py:556-563)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gf_mask_kernel(const GfNotePlan *__restrict__ plans, const GfNoteDev *__restrict__ notes, const GfSourceDev *__restrict__ srcs)
{
    const GfNotePlan &pl = plans[blockIdx.y];
    const GfNoteDev nd = notes[blockIdx.y];
    const float *mask_src = srcs[pl.src].mask;
    if (!pl.vel_active) {
        // without the velocity stretch mask_new is a gather of source samples (gf_mask_prevel_src, gf_maps.cuh): the plan's
        // fields in registers, four consecutive samples per thread, one 16-byte store (35 -> ~10 instructions per sample)
        const int n = pl.n_total;
        const int pre_n = pl.pre_s_n, pre_a = pl.pre_s_a, tail_n = pl.tail_s_n, tail_a = pl.tail_s_a;
        const bool wrap = pl.tail_s_n < pl.want_samples, rev = pl.reverse != 0, fv = pl.FV != 0;
        const int n_src = pl.N_src;
        float *__restrict__ vm = nd.vm;
        for (int i = 4 * (blockIdx.x * blockDim.x + threadIdx.x); i < n; i += 4 * gridDim.x * blockDim.x) {
            float v[4] = {1.0f, 1.0f, 1.0f, 1.0f};
            if (!fv) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int u = i + k;
                    if (u < n) {
                        int sidx;
                        if (u < pre_n) sidx = pre_a + u;
                        else {
                            int w = u - pre_n;
                            if (wrap) w %= tail_n;
                            sidx = tail_a + w;
                        }
                        v[k] = mask_src[rev ? (n_src - 1 - sidx) : sidx];
                    }
                }
            }
            if (i + 4 <= n) *reinterpret_cast<float4 *>(vm + i) = make_float4(v[0], v[1], v[2], v[3]);
            else for (int k = 0; k < 4; ++k) if (i + k < n) vm[i + k] = v[k];
            nd.vm4[i >> 2] = v[0];                             // mask[::4] for the sigma-25 smoothing (a strided read of vm pulled all of it through DRAM)
        }
        return;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pl.n_total; i += gridDim.x * blockDim.x) {
        const float v = (float)gf_mask_new(pl, mask_src, i);
        nd.vm[i] = v;
        if ((i & 3) == 0) nd.vm4[i >> 2] = v;
    }
}

void gf_launch_mask(const GfNotePlan *plans, const GfNoteDev *notes, const GfSourceDev *srcs, int n_notes, int max_n, cudaStream_t st)
{
    if (n_notes <= 0) return;
#ifndef GF_MASK_GX
#define GF_MASK_GX 16
#endif
    dim3 grid(min(GF_MASK_GX, (max_n + 255) / 256), n_notes);
    gf_mask_kernel<<<grid, 256, 0, st>>>(plans, notes, srcs);
}

// ------------------------------------------------------------------------------------------------
// generic Gaussian FIR along time with numpy-'reflect' padding, fp64 accumulation
//   job: in (f32 or f64, strided), out (f32 or f64), n, sigma ; optional max(|y| + 1e-6) reduction
// ------------------------------------------------------------------------------------------------

#define GF_FIR_TILE 1024

__device__ __forceinline__ bool gf_fir_is_f32(const GfFirJob &jb)
{
    // f32 input without a max-abs reduction: f32 taps / accumulation (the output may still be stored as fp64: the
    // pitch-dynamics deviation curve, whose 3,529-tap smoothing only feeds a clipped gain in dB)
    return gf_fir_f32(jb.in_f64, jb.maxabs, jb.in_cast_f32);
}
// long kernels are rendered by the overlap-save kernels of k_conv.cu
__device__ __forceinline__ bool gf_fir_is_fft(const GfFirJob &jb) { return gf_fir_wants_fft(jb.sigma, gf_fir_is_f32(jb)); }

// f32 -> f32 jobs (voicing-mask smoothing: sigma 25 on the decimated mask, sigma 20, sigma 441): the same
// register-tiled sliding window as the envelope kernel, 17 outputs per thread, f32 taps and accumulation
#define GF_F32_THREADS 128
#define GF_F32_TILE (GF_EPL * GF_F32_THREADS)
__global__ void __launch_bounds__(GF_F32_THREADS) gf_fir32_kernel(const GfFirJob *__restrict__ jobs)
{
    extern __shared__ __align__(16) float f32sm[];
    __shared__ double red[GF_F32_THREADS / 32];
    const GfFirJob jb = jobs[blockIdx.y];
    if (!gf_fir_is_f32(jb) || gf_fir_is_fft(jb)) return;
    const int n = jb.n;
    const int start = blockIdx.x * GF_F32_TILE;
    if (start >= n) return;
    const int radius = (int)(4.0 * jb.sigma + 0.5);
    const int ntap8 = ((2 * radius + 1 + 7) & ~7) + 8;
    float *taps = f32sm;                                  // ntap8
    float *row = taps + ntap8;                            // radius + GF_F32_TILE + radius + 16
    float *outb = row + (2 * radius + GF_F32_TILE + 16);  // GF_F32_TILE
    double part = 0.0;
    for (int j = threadIdx.x; j <= 2 * radius; j += blockDim.x) { const double t = (double)(j - radius) / jb.sigma; part += exp(-0.5 * t * t); }
    part = gf_warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    double norm = 0.0;
    for (int w = 0; w < GF_F32_THREADS / 32; ++w) norm += red[w];
    for (int j = threadIdx.x; j < ntap8; j += blockDim.x) taps[j] = (j <= 2 * radius) ? (float)gf_gauss_tap(j, radius, jb.sigma, norm) : 0.0f;
    const float *in = (const float *)jb.in;
    const int span = 2 * radius + GF_F32_TILE + 16;
    const float first = in[(size_t)gf_reflect(start - radius, n) * jb.in_stride];
    bool same = true;
    for (int i = threadIdx.x; i < span; i += blockDim.x) {
        const int p = start + i - radius;
        float v = 0.0f;
        if (p < n + radius) {
            const int q = (p >= 0 && p < n) ? p : gf_reflect(p, n);
            v = in[(size_t)q * jb.in_stride];
            same = same && (v == first);
        }
        row[i] = v;
    }
    // a normalised kernel maps a constant stretch to that constant: exactly so in the reference's fp64
    // (sum of taps = 1 within 1e-16, rounded to f32), within 1e-7 in f32 -- return the exact value
    const int constant = __syncthreads_and(same);
    float out[GF_EPL];
    if (constant) {                                           // CTA-uniform: a voiced stretch (mask == 1) or silence
#pragma unroll
        for (int e = 0; e < GF_EPL; ++e) out[e] = first;
    } else gf_fir_rt<float>(row + radius, GF_EPL * threadIdx.x, taps, radius, out);
#pragma unroll
    for (int e = 0; e < GF_EPL; ++e) outb[GF_EPL * threadIdx.x + e] = out[e];
    __syncthreads();
    const int cnt = min(GF_F32_TILE, n - start);
    if (jb.out_f64) {
        double *dst = (double *)jb.out;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) dst[start + i] = (double)outb[i];
    } else {
        float *dst = (float *)jb.out;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) dst[start + i] = outb[i];
    }
}

// fp64 jobs (sh / sr jitter curves: the result scales f0, so the taps and the accumulation stay fp64): the same
// register-tiled sliding window, 17 outputs per thread
__global__ void __launch_bounds__(GF_F32_THREADS) gf_fir_kernel(const GfFirJob *__restrict__ jobs)
{
    extern __shared__ double fsm[];
    __shared__ double red[GF_F32_THREADS / 32];
    const GfFirJob jb = jobs[blockIdx.y];
    if (gf_fir_is_f32(jb) || gf_fir_is_fft(jb)) return;       // handled by gf_fir32_kernel / gf_fftconv_kernel
    const int n = jb.n;
    const int start = blockIdx.x * GF_F32_TILE;
    if (start >= n) return;
    const int radius = (int)(4.0 * jb.sigma + 0.5);
    const int ntap8 = ((2 * radius + 1 + 7) & ~7) + 8;
    double *taps = fsm;                                   // ntap8
    double *row = taps + ntap8;                           // radius + tile + radius + 16
    double part = 0.0;
    for (int j = threadIdx.x; j <= 2 * radius; j += blockDim.x) { const double t = (double)(j - radius) / jb.sigma; part += exp(-0.5 * t * t); }
    part = gf_warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    double norm = 0.0;
    for (int w = 0; w < GF_F32_THREADS / 32; ++w) norm += red[w];
    for (int j = threadIdx.x; j < ntap8; j += blockDim.x) taps[j] = (j <= 2 * radius) ? gf_gauss_tap(j, radius, jb.sigma, norm) : 0.0;
    const int span = 2 * radius + GF_F32_TILE + 16;
    for (int i = threadIdx.x; i < span; i += blockDim.x) {
        const int p = start + i - radius;
        double v = 0.0;
        if (p < n + radius) {
            const int q = (p >= 0 && p < n) ? p : gf_reflect(p, n);
            v = jb.in_f64 ? ((const double *)jb.in)[(size_t)q * jb.in_stride] : (double)((const float *)jb.in)[(size_t)q * jb.in_stride];
            if (jb.in_cast_f32) v = (double)(float)v;
        }
        row[i] = v;
    }
    __syncthreads();
    double out[GF_EPL];
    gf_fir_rt<double>(row + radius, GF_EPL * threadIdx.x, taps, radius, out);
    double mx = 0.0;
#pragma unroll
    for (int e = 0; e < GF_EPL; ++e) {
        const int i = start + GF_EPL * threadIdx.x + e;
        if (i < n) {
            if (jb.out_f64) ((double *)jb.out)[i] = out[e]; else ((float *)jb.out)[i] = (float)out[e];
            mx = fmax(mx, fabs(out[e]) + 1e-6);
        }
    }
    if (jb.maxabs) {
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0)
            atomicMax((unsigned long long *)jb.maxabs, (unsigned long long)__double_as_longlong(mx));
    }
}

int gf_launch_fir(const GfFirJob *h_jobs, const GfFirJob *d_jobs, int n_jobs, cudaStream_t st)
{
    if (n_jobs <= 0) return 0;
    int launches = gf_launch_fftconv(h_jobs, d_jobs, n_jobs, st);
    // the direct kernels: shared memory and grid sized by the jobs they keep
    int max_n = 0, radius = 0;
    bool any_f64 = false, any_f32 = false;
    for (int k = 0; k < n_jobs; ++k) {
        const GfFirJob &j = h_jobs[k];
        const bool f32 = gf_fir_f32(j.in_f64, j.maxabs, j.in_cast_f32);
        if (j.n <= 0 || gf_fir_wants_fft(j.sigma, f32)) continue;
        max_n = std::max(max_n, j.n);
        radius = std::max(radius, gf_fir_radius(j.sigma));
        (f32 ? any_f32 : any_f64) = true;
    }
    if (max_n <= 0) return launches;
    const dim3 grid((max_n + GF_F32_TILE - 1) / GF_F32_TILE, n_jobs);
    if (any_f64) {
        const size_t smem = sizeof(double) * (size_t)((((2 * radius + 1 + 7) & ~7) + 8) + (2 * radius + GF_F32_TILE + 16));
        static GfSmemLimit memo64;
        gf_smem_limit(gf_fir_kernel, smem, memo64);
        gf_fir_kernel<<<grid, GF_F32_THREADS, smem, st>>>(d_jobs);
        ++launches;
    }
    if (any_f32) {
        const size_t smem32 = sizeof(float) * (size_t)((((2 * radius + 1 + 7) & ~7) + 8) + (2 * radius + GF_F32_TILE + 16) + GF_F32_TILE);
        static GfSmemLimit memo32;
        gf_smem_limit(gf_fir32_kernel, smem32, memo32);
        gf_fir32_kernel<<<grid, GF_F32_THREADS, smem32, st>>>(d_jobs);
        ++launches;
    }
    return launches;
}
