// k_tail.cu -- time-domain tail of gf.synthesize and of GooferResampler.resample.
//
//   gf_peak_kernel     GOOFER.py:1179-1193, 1210: mask cross-fade gains, sr volume jitter, peak
//   gf_mix_kernel      GOOFER.py:1208-1218 normalise; SillySampler.py:1143-1182 V/B/U mix, sa blend, pd
//   gf_onepole_kernel  SillySampler.py:95-174 dynamic one-pole cascades (su, sj, fry, st)
#include "gf_device.cuh"
#include "gf_maps.cuh"

struct GfStreams { float h, b, u; };

// gains of one sample of one synthesize pass before the peak normalisation (GOOFER.py:1121-1128, 1179-1191).
// `ms` is the smoothed mask at the sample (ignored by the sa pass); the unvoiced stream is only read where its
// gain (1 - ms) * 0.75 is not exactly zero: the frame kernel does not store it elsewhere.
// SIMPLE: a note with one synthesize pass and no volume jitter (see gf_note_tail_simple): those branches compile out.
template <bool SIMPLE = false>
__device__ __forceinline__ GfStreams gf_pass_gains(const GfNotePlan &pl, const GfNoteDev &nd, const GfPassDev &ps, float mag,
                                                   float hraw, float braw, float uraw, float ms, int i)
{
    GfStreams r;
    r.h = hraw / mag;                                         // S / max|S| (GOOFER.py:1121-1128), applied after the iSTFT
    if (!SIMPLE && ps.mask_ones) {                                       // sa pass: mask == 1, strengths 1 (SillySampler.py:1156-1170)
        r.b = braw;
        r.u = 0.0f;
    } else {
        r.b = braw * ms * pl.breath_strength;                          // 0.1 / 0.75 unless a direct gf.synthesize call overrides them
        r.u = (ms == 1.0f) ? 0.0f : uraw * (1.0f - ms) * pl.uv_strength;
    }
    if (!SIMPLE && ps.kind == GF_PASS_MAIN && pl.vol_jitter) {
        // GOOFER.py:1185-1191 (create_volume_jitter :638-659 without vibrato)
        const double zh = nd.z_srh[i] / nd.noteScal[GF_NS_SRHMAX];
        const double zb = nd.z_srb[i] / nd.noteScal[GF_NS_SRBMAX];
        const double hj = 1.0 + zh * pl.vol_jitter_strength;
        const double bj = 1.0 + zb * pl.vol_jitter_strength_breath;
        const double vj = (double)nd.vjm[i];
        r.h = (float)((double)r.h * (1.0 + (hj - 1.0) * vj));
        r.b = (float)((double)r.b * (1.0 + (bj - 1.0) * vj));
    }
    return r;
}

// the three streams of one synthesize pass before the peak normalisation, one sample
__device__ __forceinline__ GfStreams gf_pass_streams(const GfNotePlan &pl, const GfNoteDev &nd, const GfPassDev &ps,
                                                     const GfPassScal &sc, int i)
{
    // smooth_mask_ds, expanded once by gf_f0_kernel; hop blocks where it is 1 throughout are only flagged (ms_one), not stored
    const float ms = (ps.mask_ones || nd.ms_one[i >> 8]) ? 1.0f : nd.ms[i];
    const float uraw = (ps.mask_ones || ms == 1.0f) ? 0.0f : ps.uv[i];
    return gf_pass_gains(pl, nd, ps, __uint_as_float(sc.mag_bits), ps.harm[i], ps.bre[i], uraw, ms, i);
}

// four consecutive samples starting at i (a multiple of 4; the workspace arrays are 256-byte aligned): one 16-byte
// load per array when all four exist, element loads at the ragged end
__device__ __forceinline__ void gf_ld4(const float *__restrict__ p, int i, int cnt, float (&v)[4])
{
    if (cnt == 4) {
        const float4 q = *reinterpret_cast<const float4 *>(p + i);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (k < cnt) ? p[i + k] : 0.0f;
    }
}

template <bool SIMPLE>
__device__ __forceinline__ void gf_pass_streams4(const GfNotePlan &pl, const GfNoteDev &nd, const GfPassDev &ps, float mag,
                                                 int i, int cnt, GfStreams (&out)[4])
{
    float h[4], b[4], u[4] = {0.f, 0.f, 0.f, 0.f}, ms[4] = {1.f, 1.f, 1.f, 1.f};
    gf_ld4(ps.harm, i, cnt, h);
    gf_ld4(ps.bre, i, cnt, b);
    if (SIMPLE || !ps.mask_ones) {
        // i is a multiple of 4, so the four samples share a 256-sample hop block: inside a voiced stretch (block flagged
        // all-one by gf_f0_kernel) neither the smoothed mask nor the unvoiced stream is read
        if (!nd.ms_one[i >> 8]) gf_ld4(nd.ms, i, cnt, ms);
        if (cnt < 4 || !(ms[0] == 1.0f && ms[1] == 1.0f && ms[2] == 1.0f && ms[3] == 1.0f)) gf_ld4(ps.uv, i, cnt, u);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (k < cnt) out[k] = gf_pass_gains<SIMPLE>(pl, nd, ps, mag, h[k], b[k], u[k], ms[k], i + k);
}

// The common note (c1 / c2 / c5: one pass, no sr jitter, no post-FX, no pitch dynamics, no stage taps) takes lean
// instantiations of the peak and mix kernels (fewer registers, no dead branches); everything else the general ones.
// The host evaluates the same predicate on the plans (gf_note_tail_simple in api.cu) to decide which to launch.
__device__ __forceinline__ bool gf_tail_simple(const GfNotePlan &pl, const GfNoteDev &nd)
{
    return pl.n_passes == 1 && !pl.vol_jitter && nd.fx[0] == nullptr && nd.pd_dev == nullptr && nd.tap_harm == nullptr;
}

#ifndef GF_TAIL_CTAS
#define GF_TAIL_CTAS 6                // register cap 40: both kernels wait on HBM loads
#endif
template <bool SIMPLE>
__global__ void __launch_bounds__(256, GF_TAIL_CTAS)
gf_peak_kernel(const GfNotePlan *__restrict__ plans, const GfNoteDev *__restrict__ notes, const GfPassDev *__restrict__ passes,
               GfPassScal *scal, int pass0)
{
    const int pi = pass0 + blockIdx.y;
    const GfPassDev ps = passes[pi];
    const GfNotePlan &pl = plans[ps.note];
    const GfNoteDev nd = notes[ps.note];
    if (gf_tail_simple(pl, nd) != SIMPLE) return;
    const float mag = __uint_as_float(scal[pi].mag_bits);
    const int n = ps.n_total;
    float mx = 0.0f;
    for (int i = 4 * (blockIdx.x * blockDim.x + threadIdx.x); i < n; i += 4 * gridDim.x * blockDim.x) {
        const int cnt = min(4, n - i);
        GfStreams s[4];
        gf_pass_streams4<SIMPLE>(pl, nd, ps, mag, i, cnt, s);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < cnt) mx = fmaxf(mx, fabsf((s[k].h + s[k].u) + s[k].b));
    }
    mx = gf_warp_max(mx);
    if ((threadIdx.x & 31) == 0 && mx > 0.0f) gf_atomic_max_pos(&scal[pi].peak_bits, mx);
}

void gf_launch_peak(const GfNotePlan *plans, const GfNoteDev *notes, const GfPassDev *passes, GfPassScal *scal, int pass0,
                    int n_pass, int max_n, bool any_simple, bool any_general, cudaStream_t st)
{
    if (n_pass <= 0) return;
#ifndef GF_TAIL_GX
#define GF_TAIL_GX 8
#endif
    dim3 grid(min(GF_TAIL_GX, (max_n + 1023) / 1024), n_pass);       // few fat CTAs: the per-thread record loads amortise over ~20 samples
    if (any_simple) gf_peak_kernel<true><<<grid, 256, 0, st>>>(plans, notes, passes, scal, pass0);
    if (any_general) gf_peak_kernel<false><<<grid, 256, 0, st>>>(plans, notes, passes, scal, pass0);
}

__device__ __forceinline__ float gf_pass_gain(const GfNotePlan &pl, const GfPassScal &sc)
{
    // GOOFER.py:1208-1213
    const float pk = __uint_as_float(sc.peak_bits) + 1e-12f;
    const double nrm = fmin(fmax(pl.normalize, 0.0), 1.0);
    if (nrm == 1.0) return (float)(1.0 / (double)pk);       // pow(x, 1.0) == x exactly (IEEE pow)
    return (float)pow(1.0 / (double)pk, nrm);
}

// 16-bit PCM of one output sample the way SillySampler.py:1185 stores it: sf.write(.wav) -> libsndfile PCM_16 with
// clipping enabled by python-soundfile (SFC_SET_CLIPPING): pcm.c d2s_clip_array scales by 2^31, saturates, rounds to
// nearest (lrint, ties to even) and keeps the high 16 bits.  x * 2^31 is exact in fp64 for an f32 x.
__device__ __forceinline__ short gf_pcm16(float x)
{
    const double s = (double)x * 2147483648.0;
    if (s >= 2147483647.0) return (short)0x7fff;
    if (s <= -2147483648.0) return (short)-0x8000;
    return (short)(__double2ll_rn(s) >> 16);              // NaN -> 0
}

// stage 1 of the tail: normalised streams of every pass -> fx scratch (only for notes that need the
// sequential filters); stage 2: mix.  Notes without filters go straight through gf_mix_kernel.
template <bool SIMPLE>
__global__ void __launch_bounds__(256, GF_TAIL_CTAS)
gf_mix_kernel(const GfNotePlan *__restrict__ plans, const GfNoteDev *__restrict__ notes, const GfPassDev *__restrict__ passes,
              const GfPassScal *__restrict__ scal, int note0)
{
    const GfNotePlan &pl = plans[note0 + blockIdx.y];
    const GfNoteDev nd = notes[note0 + blockIdx.y];
    if (gf_tail_simple(pl, nd) != SIMPLE) return;
    const int n = pl.n_total;
    const GfPassDev &p0 = passes[nd.pass0];
    const float mag0 = __uint_as_float(scal[nd.pass0].mag_bits);
    const float g0 = gf_pass_gain(pl, scal[nd.pass0]);
    int p_sa = -1;
    if (!SIMPLE)
        for (int p = 1; p < pl.n_passes; ++p)
            if (pl.pass_kind[p] == GF_PASS_SA) p_sa = nd.pass0 + p;
    float g_sa = 0.0f, mag_sa = 1.0f;
    if (p_sa >= 0) { g_sa = gf_pass_gain(pl, scal[p_sa]); mag_sa = __uint_as_float(scal[p_sa].mag_bits); }
    const bool fx = !SIMPLE && nd.fx[0] != nullptr;       // harm / bre already post-processed into fx[0] / fx[1]
    double fx_scale = 1.0;
    if (fx && pl.tension != 0.0) {
        // SillySampler.py:1136-1140 (gf.rms GOOFER.py:170-171)
        const double r0 = sqrt(nd.noteScal[GF_NS_R0] / (double)n + 1e-12), r1 = sqrt(nd.noteScal[GF_NS_R1] / (double)n + 1e-12);
        if (r1 > 0.0) fx_scale = r0 / r1;
    }
    const bool pd_on = !SIMPLE && nd.pd_dev != nullptr;
    const bool taps = !SIMPLE && nd.tap_harm != nullptr;
    const double pd_ref = pd_on ? nd.noteScal[GF_NS_PDREF] : 1.0;
    // scalars of the note, read once (the stores below may alias the plan record as far as the compiler knows)
    const double V = pl.V, volume = pl.volume, sa = pl.sa, pdv = pl.pd;
    const float Bf = (float)pl.B, Uf = (float)pl.U, volf = (float)pl.volume;
    // the caller's output (and tap) arrays start at any element offset: 16-byte stores only when aligned
    const bool out_vec = (((uintptr_t)nd.out) & 15) == 0;
    const bool pcm_vec = (((uintptr_t)nd.pcm) & 7) == 0;
    const bool tap_vec = taps && ((((uintptr_t)nd.tap_harm) | ((uintptr_t)nd.tap_uv) | ((uintptr_t)nd.tap_bre)) & 15) == 0;
    for (int i = 4 * (blockIdx.x * blockDim.x + threadIdx.x); i < n; i += 4 * gridDim.x * blockDim.x) {
        const int cnt = min(4, n - i);
        GfStreams s[4], a[4];
        gf_pass_streams4<SIMPLE>(pl, nd, p0, mag0, i, cnt, s);
        if (p_sa >= 0) gf_pass_streams4<false>(pl, nd, passes[p_sa], mag_sa, i, cnt, a);
        float o[4], th[4], tu[4], tb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= cnt) { o[k] = th[k] = tu[k] = tb[k] = 0.0f; continue; }
            float h = s[k].h * g0, b = s[k].b * g0;
            const float u = s[k].u * g0;
            th[k] = h; tu[k] = u; tb[k] = b;
            if (fx) { h = (float)((double)nd.fx[0][i + k] * fx_scale); b = (float)((double)nd.fx[1][i + k] * fx_scale); }
            // SillySampler.py:1143-1151: harm * V (np.float64) + bre * B (f32) + uv * U (f32), * volume
            double out = (((double)h * V + (double)(b * Bf)) + (double)(u * Uf)) * volume;
            if (p_sa >= 0) {
                // SillySampler.py:1153-1172
                const float au = a[k].u * g_sa, ab = a[k].b * g_sa;
                out = out * (1.0 - sa) + ((double)((au + ab) * volf)) * sa;
            }
            if (pd_on) {
                // SillySampler.py:869-881, 1174-1182
                const double v = fmin(fmax(nd.pd_dev[i + k] / pd_ref, -1.0), 1.0);
                const double db = (12.0 * fabs(pdv)) * (pdv > 0.0 ? v : -v);
                float dynf = (float)pow(10.0, db / 20.0);
                dynf = fminf(fmaxf(dynf, 1e-3f), 1e3f);
                const double dyn = 1.0 + (double)(dynf - 1.0f) * (double)nd.pd_gm[i + k];
                out = out * dyn;
            }
            o[k] = (float)out;
        }
        if (taps) {
            if (cnt == 4 && tap_vec) {
                *reinterpret_cast<float4 *>(nd.tap_harm + i) = make_float4(th[0], th[1], th[2], th[3]);
                *reinterpret_cast<float4 *>(nd.tap_uv + i) = make_float4(tu[0], tu[1], tu[2], tu[3]);
                *reinterpret_cast<float4 *>(nd.tap_bre + i) = make_float4(tb[0], tb[1], tb[2], tb[3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < cnt) { nd.tap_harm[i + k] = th[k]; nd.tap_uv[i + k] = tu[k]; nd.tap_bre[i + k] = tb[k]; }
            }
        }
        if (nd.out) {
            if (cnt == 4 && out_vec) *reinterpret_cast<float4 *>(nd.out + i) = make_float4(o[0], o[1], o[2], o[3]);
            else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < cnt) nd.out[i + k] = o[k];
            }
        }
        if (nd.pcm) {
            short q[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) q[k] = gf_pcm16(o[k]);
            if (cnt == 4 && pcm_vec) *reinterpret_cast<short4 *>(nd.pcm + i) = make_short4(q[0], q[1], q[2], q[3]);
            else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < cnt) nd.pcm[i + k] = q[k];
            }
        }
    }
}

void gf_launch_mix(const GfNotePlan *plans, const GfNoteDev *notes, const GfPassDev *passes, const GfPassScal *scal,
                   int note0, int n_notes, int max_n, bool any_simple, bool any_general, cudaStream_t st)
{
    if (n_notes <= 0) return;
    dim3 grid(min(GF_TAIL_GX, (max_n + 1023) / 1024), n_notes);
    if (any_simple) gf_mix_kernel<true><<<grid, 256, 0, st>>>(plans, notes, passes, scal, note0);
    if (any_general) gf_mix_kernel<false><<<grid, 256, 0, st>>>(plans, notes, passes, scal, note0);
}

// ------------------------------------------------------------------------------------------------
// dynamic_butter_filter (SillySampler.py:95-174): `order` cascaded one-pole sections whose coefficient follows f0 per
// sample.  One CTA per signal, which it walks in tiles of 256 threads x 16 samples.  A tile is read ONCE (16 contiguous
// samples per thread, 16-byte loads), taken through ALL sections in registers, and written once: per section every
// thread composes the affine map y_out = A y_in + B of its 16 samples, the 256 maps are scanned (warp shuffles + one
// shared-memory hop), every thread replays its samples from the right incoming state, and the outputs are the next
// section's inputs.  Section states and (high-pass) previous inputs are carried from tile to tile.  f32 throughout like
// the numba kernel.  (The first version gave each thread one 172-sample chunk and re-read coefficient and signal from
// global memory with a 172-element lane stride, twice per section: 32 sectors per warp load -- L2-transaction bound.)
// ------------------------------------------------------------------------------------------------
#define GF_OP_THREADS 256
#define GF_OP_K 16
#define GF_OP_TILE (GF_OP_THREADS * GF_OP_K)
#define GF_OP_MAX_ORDER 12

__global__ void __launch_bounds__(GF_OP_THREADS) gf_onepole_kernel(const GfOnepoleJob *__restrict__ jobs)
{
    __shared__ float wA[GF_OP_THREADS / 32], wB[GF_OP_THREADS / 32], sEdge[GF_OP_THREADS / 32];
    __shared__ float sY[GF_OP_MAX_ORDER], sXp[GF_OP_MAX_ORDER];     // tile-to-tile carries: section state, last input of the section
    const GfOnepoleJob jb = jobs[blockIdx.x];
    const int n = jb.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n <= 0) return;
    const int order = min(max(1, jb.order), GF_OP_MAX_ORDER);
    const double srd = (double)jb.sr;
    // np.any(f0 > 0) decides whether the 5-tap smoothing runs (SillySampler.py:107-113)
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
    if (tid < GF_OP_MAX_ORDER) { sY[tid] = 0.0f; sXp[tid] = 0.0f; }
    __syncthreads();
    const bool vec = ((((uintptr_t)jb.x) | ((uintptr_t)jb.y)) & 15) == 0;
    const bool hp = jb.highpass != 0;

    for (int tile0 = 0; tile0 < n; tile0 += GF_OP_TILE) {
        const int c0 = tile0 + GF_OP_K * tid;
        const int cnt = max(0, min(GF_OP_K, n - c0));
        const int tile_last = (min(n, tile0 + GF_OP_TILE) - 1 - tile0) / GF_OP_K;       // thread that owns the tile's last sample
        float xin[GF_OP_K], al[GF_OP_K];
        if (cnt == GF_OP_K && vec) {
#pragma unroll
            for (int q = 0; q < GF_OP_K / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4 *>(jb.x + c0 + 4 * q);
                xin[4 * q] = v.x; xin[4 * q + 1] = v.y; xin[4 * q + 2] = v.z; xin[4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < GF_OP_K; ++k) xin[k] = (k < cnt) ? jb.x[c0 + k] : 0.0f;
        }
        // ---- per-sample coefficient (SillySampler.py:128-152), f32 stores like the numba kernel ----
#pragma unroll
        for (int k = 0; k < GF_OP_K; ++k) {
            al[k] = 0.0f;
            if (k < cnt) {
                const int i = c0 + k;
                float f0s;
                if (jb.smooth_f0 && any_pos) {
                    // np.convolve(pad(f0, 2, 'edge'), ones(5)/5, 'valid') in f32: sum_j k[j] * x[i + 4 - j], j = 0..4
                    const float kk = 1.0f / 5.0f;
                    float a = 0.0f;
                    for (int j = 0; j < 5; ++j) a = fmaf(kk, drv(i + 2 - j), a);
                    f0s = a;
                } else f0s = drv(i);
                float fc = (f0s > 0.0f) ? (float)((double)f0s * jb.cutoff_factor) : (float)jb.cutoff_factor;
                fc = fmaxf(fc, hp ? 20.0f : 60.0f);
                fc = (float)fmin((double)fc, 0.45 * srd);
                const double w = (2.0 * 3.141592653589793) * (double)fc;
                al[k] = (float)(hp ? srd / (w + srd) : w / (w + srd));
            }
        }
        for (int pass = 0; pass < order; ++pass) {
            // carries of this section from the previous tile: read before the barriers below, rewritten after them
            const float y_carry = sY[pass], xp_carry = sXp[pass];
            float mylast = xin[0];
#pragma unroll
            for (int k = 1; k < GF_OP_K; ++k) if (k < cnt) mylast = xin[k];
            // ---- high-pass: the section's input just before this thread's first sample ----
            float xp_first = xin[0];                               // very first sample of the signal: x[-1] := x[0]
            if (hp) {
                if (lane == 31) sEdge[warp] = mylast;
                __syncthreads();
                const float up = __shfl_up_sync(0xffffffffu, mylast, 1);
                if (lane > 0) xp_first = up;
                else if (warp > 0) xp_first = sEdge[warp - 1];
                else if (tile0 > 0) xp_first = xp_carry;
            }
            // ---- (1) affine map of the thread's samples from a zero state ----
            float A = 1.0f, B = 0.0f;
            if (!hp) {
#pragma unroll
                for (int k = 0; k < GF_OP_K; ++k)
                    if (k < cnt) { A = (1.0f - al[k]) * A; B = fmaf(al[k], xin[k] - B, B); }       // y = y + a (x - y)
            } else {
                float xp = xp_first;
#pragma unroll
                for (int k = 0; k < GF_OP_K; ++k)
                    if (k < cnt) { A = al[k] * A; B = al[k] * ((B - xp) + xin[k]); xp = xin[k]; }  // y = a (y + x - xp)
            }
            // ---- (2) scan of the 256 maps ----
            float sA = A, sB = B;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float pA = __shfl_up_sync(0xffffffffu, sA, o), pB = __shfl_up_sync(0xffffffffu, sB, o);
                if (lane >= o) { sB = fmaf(sA, pB, sB); sA = sA * pA; }
            }
            if (lane == 31) { wA[warp] = sA; wB[warp] = sB; }
            __syncthreads();
            float y = (tile0 > 0) ? y_carry : 0.0f;
            for (int w = 0; w < warp; ++w) y = fmaf(wA[w], y, wB[w]);
            const float eA = __shfl_up_sync(0xffffffffu, sA, 1), eB = __shfl_up_sync(0xffffffffu, sB, 1);
            if (lane > 0) y = fmaf(eA, y, eB);
            // ---- (3) replay; the outputs are the next section's inputs ----
            if (!hp) {
#pragma unroll
                for (int k = 0; k < GF_OP_K; ++k)
                    if (k < cnt) { y = fmaf(al[k], xin[k] - y, y); xin[k] = y; }
            } else {
                float xp = xp_first;
#pragma unroll
                for (int k = 0; k < GF_OP_K; ++k)
                    if (k < cnt) { const float x = xin[k]; y = al[k] * ((y - xp) + x); xp = x; xin[k] = y; }
            }
            if (tid == tile_last) { sY[pass] = y; sXp[pass] = mylast; }
            __syncthreads();                                       // wA / wB / sEdge and the carries are rewritten by the next section
        }
        if (cnt == GF_OP_K && vec) {
#pragma unroll
            for (int q = 0; q < GF_OP_K / 4; ++q)
                *reinterpret_cast<float4 *>(jb.y + c0 + 4 * q) = make_float4(xin[4 * q], xin[4 * q + 1], xin[4 * q + 2], xin[4 * q + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < GF_OP_K; ++k) if (k < cnt) jb.y[c0 + k] = xin[k];
        }
    }
}

void gf_launch_onepole(const GfOnepoleJob *jobs, int n_jobs, cudaStream_t st)
{
    if (n_jobs > 0) gf_onepole_kernel<<<n_jobs, GF_OP_THREADS, 0, st>>>(jobs);
}
