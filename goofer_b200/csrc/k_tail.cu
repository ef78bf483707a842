// k_tail.cu -- time-domain tail of gf.synthesize and of GooferResampler.resample.
//
//   gf_peak_kernel     GOOFER.py:1179-1193, 1210: mask cross-fade gains, sr volume jitter, peak
//   gf_mix_kernel      GOOFER.py:1208-1218 normalise; SillySampler.py:1143-1182 V/B/U mix, sa blend, pd
//   gf_onepole_kernel  SillySampler.py:95-174 dynamic one-pole cascades (su, sj, fry, st)
#include "gf_device.cuh"
#include "gf_maps.cuh"

struct GfStreams { float h, b, u; };

// the three streams of one synthesize pass before the peak normalisation
__device__ __forceinline__ GfStreams gf_pass_streams(const GfNotePlan &pl, const GfNoteDev &nd, const GfPassDev &ps,
                                                     const GfPassScal &sc, int i)
{
    GfStreams r;
    const float mag = __uint_as_float(sc.mag_bits);
    r.h = ps.harm[i] / mag;                                   // S / max|S| (GOOFER.py:1121-1128), applied after the iSTFT
    if (ps.mask_ones) {                                       // sa pass: mask == 1, strengths 1 (SillySampler.py:1156-1170)
        r.b = ps.bre[i];
        r.u = 0.0f * ps.uv[i];
    } else {
        const float ms = nd.ms[i];                            // smooth_mask_ds, expanded once by gf_f0_kernel
        r.b = ps.bre[i] * ms * 0.1f;
        r.u = ps.uv[i] * (1.0f - ms) * 0.75f;
    }
    if (ps.kind == GF_PASS_MAIN && pl.vol_jitter) {
        // GOOFER.py:1185-1191 (create_volume_jitter :638-659 without vibrato)
        const double zh = nd.z_srh[i] / nd.noteScal[GF_NS_SRHMAX];
        const double zb = nd.z_srb[i] / nd.noteScal[GF_NS_SRBMAX];
        const double hj = 1.0 + zh * pl.vol_jitter_strength;
        const double bj = 1.0 + zb * (pl.vol_jitter_strength * 2);
        const double vj = (double)nd.vjm[i];
        r.h = (float)((double)r.h * (1.0 + (hj - 1.0) * vj));
        r.b = (float)((double)r.b * (1.0 + (bj - 1.0) * vj));
    }
    return r;
}

__global__ void __launch_bounds__(256)
gf_peak_kernel(const GfNotePlan *__restrict__ plans, const GfNoteDev *__restrict__ notes, const GfPassDev *__restrict__ passes,
               GfPassScal *scal, int pass0)
{
    const int pi = pass0 + blockIdx.y;
    const GfPassDev ps = passes[pi];
    const GfNotePlan &pl = plans[ps.note];
    const GfNoteDev nd = notes[ps.note];
    const GfPassScal sc = scal[pi];
    float mx = 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ps.n_total; i += gridDim.x * blockDim.x) {
        const GfStreams s = gf_pass_streams(pl, nd, ps, sc, i);
        mx = fmaxf(mx, fabsf((s.h + s.u) + s.b));
    }
    mx = gf_warp_max(mx);
    if ((threadIdx.x & 31) == 0 && mx > 0.0f) gf_atomic_max_pos(&scal[pi].peak_bits, mx);
}

void gf_launch_peak(const GfNotePlan *plans, const GfNoteDev *notes, const GfPassDev *passes, GfPassScal *scal, int pass0,
                    int n_pass, int max_n, cudaStream_t st)
{
    if (n_pass <= 0) return;
    dim3 grid(min(8, (max_n + 255) / 256), n_pass);         // few fat CTAs: the per-thread record loads amortise over ~20 samples
    gf_peak_kernel<<<grid, 256, 0, st>>>(plans, notes, passes, scal, pass0);
}

__device__ __forceinline__ float gf_pass_gain(const GfNotePlan &pl, const GfPassScal &sc)
{
    // GOOFER.py:1208-1213
    const float pk = __uint_as_float(sc.peak_bits) + 1e-12f;
    const double nrm = fmin(fmax(pl.normalize, 0.0), 1.0);
    return (float)pow(1.0 / (double)pk, nrm);
}

// stage 1 of the tail: normalised streams of every pass -> fx scratch (only for notes that need the
// sequential filters); stage 2: mix.  Notes without filters go straight through gf_mix_kernel.
__global__ void __launch_bounds__(256)
gf_mix_kernel(const GfNotePlan *__restrict__ plans, const GfNoteDev *__restrict__ notes, const GfPassDev *__restrict__ passes,
              const GfPassScal *__restrict__ scal, int note0)
{
    const GfNotePlan &pl = plans[note0 + blockIdx.y];
    const GfNoteDev nd = notes[note0 + blockIdx.y];
    const int n = pl.n_total;
    const GfPassDev &p0 = passes[nd.pass0];
    const GfPassScal &s0 = scal[nd.pass0];
    const float g0 = gf_pass_gain(pl, s0);
    int p_sa = -1;
    for (int p = 1; p < pl.n_passes; ++p)
        if (pl.pass_kind[p] == GF_PASS_SA) p_sa = nd.pass0 + p;
    float g_sa = 0.0f;
    if (p_sa >= 0) g_sa = gf_pass_gain(pl, scal[p_sa]);
    const bool fx = nd.fx[0] != nullptr;                  // harm / bre already post-processed into fx[0] / fx[1]
    double fx_scale = 1.0;
    if (fx && pl.tension != 0.0) {
        // SillySampler.py:1136-1140 (gf.rms GOOFER.py:170-171)
        const double r0 = sqrt(nd.noteScal[GF_NS_R0] / (double)n + 1e-12), r1 = sqrt(nd.noteScal[GF_NS_R1] / (double)n + 1e-12);
        if (r1 > 0.0) fx_scale = r0 / r1;
    }
    const double pd_ref = nd.pd_dev ? nd.noteScal[GF_NS_PDREF] : 1.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const GfStreams s = gf_pass_streams(pl, nd, p0, s0, i);
        float h = s.h * g0, b = s.b * g0;
        const float u = s.u * g0;
        if (nd.tap_harm) { nd.tap_harm[i] = h; nd.tap_uv[i] = u; nd.tap_bre[i] = b; }
        if (fx) { h = (float)((double)nd.fx[0][i] * fx_scale); b = (float)((double)nd.fx[1][i] * fx_scale); }
        // SillySampler.py:1143-1151: harm * V (np.float64) + bre * B (f32) + uv * U (f32), * volume
        double out = (((double)h * pl.V + (double)(b * (float)pl.B)) + (double)(u * (float)pl.U)) * pl.volume;
        if (p_sa >= 0) {
            // SillySampler.py:1153-1172
            const GfStreams a = gf_pass_streams(pl, nd, passes[p_sa], scal[p_sa], i);
            const float au = a.u * g_sa, ab = a.b * g_sa;
            out = out * (1.0 - pl.sa) + ((double)((au + ab) * (float)pl.volume)) * pl.sa;
        }
        if (nd.pd_dev) {
            // SillySampler.py:869-881, 1174-1182
            const double v = fmin(fmax(nd.pd_dev[i] / pd_ref, -1.0), 1.0);
            const double db = (12.0 * fabs(pl.pd)) * (pl.pd > 0.0 ? v : -v);
            float dynf = (float)pow(10.0, db / 20.0);
            dynf = fminf(fmaxf(dynf, 1e-3f), 1e3f);
            const double dyn = 1.0 + (double)(dynf - 1.0f) * (double)nd.pd_gm[i];
            out = out * dyn;
        }
        nd.out[i] = (float)out;
    }
}

void gf_launch_mix(const GfNotePlan *plans, const GfNoteDev *notes, const GfPassDev *passes, const GfPassScal *scal,
                   int note0, int n_notes, int max_n, cudaStream_t st)
{
    if (n_notes <= 0) return;
    dim3 grid(min(8, (max_n + 255) / 256), n_notes);
    gf_mix_kernel<<<grid, 256, 0, st>>>(plans, notes, passes, scal, note0);
}

// ------------------------------------------------------------------------------------------------
// dynamic_butter_filter (SillySampler.py:95-174): `order` cascaded one-pole sections whose
// coefficient follows f0 per sample.  One CTA per signal; each thread owns a contiguous chunk.  Per
// section: (1) every thread composes the affine map of its chunk, (2) the 256 maps are scanned,
// (3) every thread replays its chunk from the right initial state.  f32 throughout like the reference.
// ------------------------------------------------------------------------------------------------

// walk a lane's chunk eight samples at a time: the sixteen loads of a batch are in flight together (the
// recurrence itself is serial, the memory latency no longer is); the body may store to position i (y may alias x:
// the batch was loaded before its first store)
template <typename Body>
__device__ __forceinline__ void gf_chunk8(int c0, int c1, const float *__restrict__ alpha, const float *src, Body body)
{
    int i = c0;
    for (; i + 8 <= c1; i += 8) {
        float a[8], x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { a[k] = alpha[i + k]; x[k] = src[i + k]; }
#pragma unroll
        for (int k = 0; k < 8; ++k) body(a[k], x[k], i + k);
    }
    for (; i < c1; ++i) body(alpha[i], src[i], i);
}

#define GF_OP_THREADS 256
__global__ void __launch_bounds__(GF_OP_THREADS) gf_onepole_kernel(const GfOnepoleJob *__restrict__ jobs)
{
    __shared__ float wA[GF_OP_THREADS / 32], wB[GF_OP_THREADS / 32];
    const GfOnepoleJob jb = jobs[blockIdx.x];
    const int n = jb.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n <= 0) return;
    const double srd = (double)jb.sr;
    // ---- per-sample coefficient (SillySampler.py:128-152), f32 stores like the numba kernel ----
    bool any_pos_l = false;
    for (int i = tid; i < n; i += GF_OP_THREADS) {
        float f = jb.f0 ? jb.f0[i] : (float)jb.f0_const;
        if (jb.f0_floor > 0.0) f = fmaxf(f, (float)jb.f0_floor);
        any_pos_l |= (f > 0.0f);
    }
    const bool any_pos = __syncthreads_or(any_pos_l) != 0;
    auto drv = [&](int i) {
        i = i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
        float f = jb.f0 ? jb.f0[i] : (float)jb.f0_const;
        if (jb.f0_floor > 0.0) f = fmaxf(f, (float)jb.f0_floor);
        return f;
    };
    for (int i = tid; i < n; i += GF_OP_THREADS) {
        float f0s;
        if (jb.smooth_f0 && any_pos) {
            // np.convolve(pad(f0, 2, 'edge'), ones(5)/5, 'valid') in f32
            const float k = 1.0f / 5.0f;
            float a = 0.0f;
            // np.convolve accumulates sum_j k[j] * x[i + 4 - j]: order of terms j = 0..4
            for (int j = 0; j < 5; ++j) a = fmaf(k, drv(i + 2 - j), a);
            f0s = a;
        } else f0s = drv(i);
        float fc = (f0s > 0.0f) ? (float)((double)f0s * jb.cutoff_factor) : (float)jb.cutoff_factor;
        fc = fmaxf(fc, jb.highpass ? 20.0f : 60.0f);
        fc = (float)fmin((double)fc, 0.45 * srd);
        const double w = (2.0 * 3.141592653589793) * (double)fc;
        jb.alpha[i] = (float)(jb.highpass ? srd / (w + srd) : w / (w + srd));
    }
    __syncthreads();
    // thread t owns samples [c0, c1); per section: (1) compose the affine map of the chunk, (2) scan the 256 maps
    // (warp shuffles + one shared-memory hop), (3) replay the chunk from the right initial state
    const int chunk = (n + GF_OP_THREADS - 1) / GF_OP_THREADS;
    const int c0 = min(n, tid * chunk), c1 = min(n, c0 + chunk);
    for (int pass = 0; pass < max(1, jb.order); ++pass) {
        const float *src = (pass == 0) ? jb.x : jb.y;
        float A = 1.0f, B = 0.0f;
        // x[c0 - 1] is read before any thread stores this section's output (y may alias the input): barrier below
        const float xp_first = (c0 > 0 && c0 < n) ? src[c0 - 1] : src[0];
        if (!jb.highpass) {
            // y = y + a (x - y) = (1 - a) y + a x
            gf_chunk8(c0, c1, jb.alpha, src, [&](float a, float x, int) { A = (1.0f - a) * A; B = fmaf(a, x - B, B); });
        } else {
            float xp = xp_first;
            // y = a (y + x - xp)
            gf_chunk8(c0, c1, jb.alpha, src, [&](float a, float x, int) { A = a * A; B = a * ((B - xp) + x); xp = x; });
        }
        // inclusive scan inside the warp: (sA, sB) = map of chunks [warp start .. this thread]
        float sA = A, sB = B;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float pA = __shfl_up_sync(0xffffffffu, sA, o), pB = __shfl_up_sync(0xffffffffu, sB, o);
            if (lane >= o) { sB = fmaf(sA, pB, sB); sA = sA * pA; }
        }
        if (lane == 31) { wA[warp] = sA; wB[warp] = sB; }
        __syncthreads();                                  // also orders the xp_first reads before the stores below
        // state entering this warp: zero initial state pushed through the maps of the earlier warps
        float y = 0.0f;
        for (int w = 0; w < warp; ++w) y = fmaf(wA[w], y, wB[w]);
        // ... and through the earlier chunks of this warp
        const float eA = __shfl_up_sync(0xffffffffu, sA, 1), eB = __shfl_up_sync(0xffffffffu, sB, 1);
        if (lane > 0) y = fmaf(eA, y, eB);
        float *dst = jb.y;
        if (!jb.highpass) {
            gf_chunk8(c0, c1, jb.alpha, src, [&](float a, float x, int i) { y = fmaf(a, x - y, y); dst[i] = y; });
        } else {
            float xp = xp_first;
            gf_chunk8(c0, c1, jb.alpha, src, [&](float a, float x, int i) { y = a * ((y - xp) + x); xp = x; dst[i] = y; });
        }
        __syncthreads();
    }
}

void gf_launch_onepole(const GfOnepoleJob *jobs, int n_jobs, cudaStream_t st)
{
    if (n_jobs > 0) gf_onepole_kernel<<<n_jobs, GF_OP_THREADS, 0, st>>>(jobs);
}
