// k_fx.cu -- growl (sg), pitch dynamics (pd) and the sequential post-FX of GooferResampler.resample
// (su / sj layers, vocal-fry high-pass blend, sd tremolo, st tension).
//
//   sg   gf_sg_f0_kernel      apply_subharm_vibrato                 GOOFER.py:748-766
//        gf_sg_walk_kernel    _detect_pulse_events (one ratio)      GOOFER.py:672-698
//        gf_sg_bank_kernel    pulse bank keyed by f'{sub_f0:.2f}'   GOOFER.py:717-723 (first occurrence wins)
//        gf_sg_render_kernel  lf_model_pulse + overlap-add          GOOFER.py:437-471, 724-729
//   pd   gf_pd_ref_kernel     np.percentile(|dev|, 95)              SillySampler.py:868
//   fx   gf_fx_stage_kernel   SillySampler.py:1038-1140 element-wise steps between the one-pole cascades
#include "gf_device.cuh"
#include "gf_maps.cuh"

// ------------------------------------------------------------------------------------------------
// sg: modulated f0
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gf_sg_f0_kernel(const int *__restrict__ list, const GfNotePlan *__restrict__ plans, const GfNoteDev *__restrict__ notes,
                const GfPassDev *__restrict__ passes)
{
    const int ni = list[blockIdx.y];
    const GfNotePlan &pl = plans[ni];
    const GfNoteDev nd = notes[ni];
    const float *f0 = passes[nd.pass0].f0;
    const int n = pl.n_total;
    const double sr = (double)pl.sr;
    const int nf = (int)(0.01 * sr);                       // subharm_vibrato_delay = 0.01   SillySampler.py:1031
    const double w = (2.0 * 3.141592653589793) * 75.0;     // subharm_vibrato_rate = 75
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float f = f0[i];
        float o = f;
        if (f > 0.0f) {
            double vib = sin(w * ((double)i / sr) + 0.0);
            if (nf < n && i < nf) vib *= gf_dlin01(i, nf);
            o = (float)((double)f * (1.0 + vib * 3.0));    // depth 3
        }
        nd.sg_f0[i] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// sg: event detection as a scan.  _detect_pulse_events (GOOFER.py:672-698) walks  phase += sub_f0 / sr;
// if phase >= 1: event, phase -= 1  in fp64, sample by sample.  Only the event POSITIONS leave the loop (the phase
// itself is never used), and they are the integer crossings of the running sum S_k of the increments -- unless
// rounding decides a borderline case.  Every addition lands below 2 and rounds by at most 2^-53, the subtraction
// of 1 is exact, so after k steps the fp64 phase is within k 2^-53 of S_k - (events so far).  The kernel therefore
//   * sums the increments EXACTLY (each is a 53-bit integer times a power of two: fixed point in units of 2^-88,
//     128-bit adds) with a block-wide scan,
//   * places event number F at the sample where floor(S_k) reaches F (slot F - 1: no compaction pass),
//   * and checks that no S_k lies within (n + 2) 2^-52 of an integer.  If one does (about 1e-11 per sample with
//     the 75 Hz vibrato on the increments), or an increment is outside [2^-24, 1), the note is flagged and the
//     sequential walk below renders it, bit for bit like the reference either way.
// 44,100 dependent DADD / compare / DSUB steps per second of audio (1.3 ms whatever the batch size) become one
// pass over the samples.
// ------------------------------------------------------------------------------------------------
struct GfU128 { unsigned long long lo, hi; };
__device__ __forceinline__ GfU128 gf_u128_add(GfU128 a, GfU128 b)
{
    GfU128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
    return r;
}
#define GF_SGS_THREADS 256
#define GF_SGS_PER 8
#define GF_SGS_UNIT 88          // S in units of 2^-88

__global__ void __launch_bounds__(GF_SGS_THREADS)
gf_sg_scan_kernel(const int *__restrict__ list, const GfNotePlan *__restrict__ plans, const GfNoteDev *__restrict__ notes, GfPassScal *scal)
{
    __shared__ GfU128 s_warp[GF_SGS_THREADS / 32];
    const int ni = list[blockIdx.x];
    const GfNotePlan &pl = plans[ni];
    const GfNoteDev nd = notes[ni];
    const int n = pl.n_total;
    const double sr = (double)pl.sr;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // |fp64 phase - exact| <= k 2^-53: twice that, in units of 2^-64 (the top 64 bits of the fraction)
    const unsigned long long margin = ((unsigned long long)n + 2ull) << 12;
    GfU128 carry; carry.lo = 0ull; carry.hi = 0ull;
    int bad = 0;
    for (int base = 0; base < n; base += GF_SGS_THREADS * GF_SGS_PER) {
        const int i0 = base + tid * GF_SGS_PER;
        float fv[GF_SGS_PER], mv[GF_SGS_PER];
        if (i0 + GF_SGS_PER <= n) {
            const float4 a = *reinterpret_cast<const float4 *>(nd.sg_f0 + i0), b = *reinterpret_cast<const float4 *>(nd.sg_f0 + i0 + 4);
            const float4 c = *reinterpret_cast<const float4 *>(nd.vm + i0), d = *reinterpret_cast<const float4 *>(nd.vm + i0 + 4);
            fv[0] = a.x; fv[1] = a.y; fv[2] = a.z; fv[3] = a.w; fv[4] = b.x; fv[5] = b.y; fv[6] = b.z; fv[7] = b.w;
            mv[0] = c.x; mv[1] = c.y; mv[2] = c.z; mv[3] = c.w; mv[4] = d.x; mv[5] = d.y; mv[6] = d.z; mv[7] = d.w;
        } else {
#pragma unroll
            for (int e = 0; e < GF_SGS_PER; ++e) {
                const bool in = i0 + e < n;
                fv[e] = in ? nd.sg_f0[i0 + e] : 0.0f;
                mv[e] = in ? nd.vm[i0 + e] : 0.0f;
            }
        }
        GfU128 loc[GF_SGS_PER];
        GfU128 run; run.lo = 0ull; run.hi = 0ull;
        unsigned actm = 0u;
#pragma unroll
        for (int e = 0; e < GF_SGS_PER; ++e) {
            const double sub = (double)fv[e] * 2.0;                 // ratio = 2 ** (12 / 12)
            const bool act = (mv[e] > 0.0f) && (fv[e] > 0.0f) && !(sub < 1e-2);
            GfU128 v; v.lo = 0ull; v.hi = 0ull;
            if (act) {
                const unsigned long long bits = (unsigned long long)__double_as_longlong(__ddiv_rn(sub, sr));
                const int ex = (int)(bits >> 52) & 0x7ff;
                if (ex < 1023 - 24 || ex > 1022) bad = 1;
                else {
                    const unsigned long long mant = (bits & 0xfffffffffffffull) | 0x10000000000000ull;
                    const int sh = ex - 1075 + GF_SGS_UNIT;         // 12 .. 35
                    v.lo = mant << sh;
                    v.hi = mant >> (64 - sh);
                    actm |= 1u << e;
                }
            }
            run = gf_u128_add(run, v);
            loc[e] = run;
        }
        // block-wide exclusive prefix of the per-thread totals
        GfU128 incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            GfU128 t;
            t.lo = __shfl_up_sync(0xffffffffu, incl.lo, o);
            t.hi = __shfl_up_sync(0xffffffffu, incl.hi, o);
            if (lane >= o) incl = gf_u128_add(incl, t);
        }
        if (lane == 31) s_warp[w] = incl;
        GfU128 pre;
        pre.lo = __shfl_up_sync(0xffffffffu, incl.lo, 1);
        pre.hi = __shfl_up_sync(0xffffffffu, incl.hi, 1);
        if (lane == 0) { pre.lo = 0ull; pre.hi = 0ull; }
        __syncthreads();
        GfU128 total = carry;
        pre = gf_u128_add(pre, carry);
#pragma unroll
        for (int q = 0; q < GF_SGS_THREADS / 32; ++q) {
            const GfU128 t = s_warp[q];
            if (q < w) pre = gf_u128_add(pre, t);
            total = gf_u128_add(total, t);
        }
        carry = total;
        // crossings of this thread's samples
        unsigned long long Fp = pre.hi >> (GF_SGS_UNIT - 64);
#pragma unroll
        for (int e = 0; e < GF_SGS_PER; ++e) {
            const GfU128 S = gf_u128_add(pre, loc[e]);
            const unsigned long long F = S.hi >> (GF_SGS_UNIT - 64);
            if ((actm >> e) & 1u) {
                const unsigned long long frac = (S.hi << (128 - GF_SGS_UNIT)) | (S.lo >> (GF_SGS_UNIT - 64));     // top 64 bits of the fraction
                if (frac <= margin || frac >= ~margin) bad = 1;
                if (F != Fp) {
                    if (F != Fp + 1ull) bad = 1;
                    const unsigned long long slot = F - 1ull;
                    if (slot < (unsigned long long)nd.sg_cap) { nd.sg_ev_i[slot] = i0 + e; nd.sg_ev_f[slot] = (double)fv[e] * 2.0; }
                }
            }
            Fp = F;
        }
        __syncthreads();
    }
    bad = __syncthreads_or(bad);
    if (tid == 0) {
        if (bad) scal[nd.pass0].sg_seq = 1;
        else {
            const unsigned long long count = carry.hi >> (GF_SGS_UNIT - 64);
            scal[nd.pass0].n_sub_events = (int)min(count, (unsigned long long)nd.sg_cap);
            if (count > (unsigned long long)nd.sg_cap) scal[nd.pass0].err = 2;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// sg: sequential event walk for the notes the scan flagged (one warp per note; the fp64 phase chain runs left to
// right like the reference, only the per-sample increments are prepared in parallel)
// ------------------------------------------------------------------------------------------------
#define GF_SG_WARPS 4
__global__ void __launch_bounds__(32 * GF_SG_WARPS)
gf_sg_walk_kernel(const int *__restrict__ list, int n_list, const GfNotePlan *__restrict__ plans,
                  const GfNoteDev *__restrict__ notes, GfPassScal *scal)
{
    __shared__ double s_inc[GF_SG_WARPS][32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int li = blockIdx.x * GF_SG_WARPS + w;
    if (li >= n_list) return;
    const int ni = list[li];
    const GfNotePlan &pl = plans[ni];
    const GfNoteDev nd = notes[ni];
#if !defined(GF_SG_SEQ_ONLY)
    if (!scal[nd.pass0].sg_seq) return;                     // the scan placed this note's events
#endif
    const int n = pl.n_total;
    const double sr = (double)pl.sr;
    double phase = 0.0;
    int count = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const double f = (i < n) ? (double)nd.sg_f0[i] : 0.0;
        const double m = (i < n) ? (double)nd.vm[i] : 0.0;
        const double sub = f * 2.0;                       // ratio = 2 ** (12 / 12)
        const bool act = (i < n) && (m > 0.0) && (f > 0.0) && !(sub < 1e-2);
        s_inc[w][lane] = act ? __ddiv_rn(sub, sr) : 0.0;
        const unsigned am = __ballot_sync(0xffffffffu, act);
        __syncwarp();
        unsigned fire = 0u;
        if (am) {
            // straight-line chain: an inactive sample adds +0.0, which leaves the (non-negative) phase and the
            // comparison unchanged, so no branch sits between the dependent additions
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                phase = __dadd_rn(phase, s_inc[w][k]);
                const bool f = phase >= 1.0;
                fire |= f ? (1u << k) : 0u;
                phase = f ? __dsub_rn(phase, 1.0) : phase;
            }
        }
        __syncwarp();
        if ((fire >> lane) & 1u) {
            const int slot = count + __popc(fire & ((1u << lane) - 1u));
            if (slot < nd.sg_cap) { nd.sg_ev_i[slot] = i; nd.sg_ev_f[slot] = sub; }
        }
        count += __popc(fire);
    }
    if (lane == 0) {
        scal[nd.pass0].n_sub_events = min(count, nd.sg_cap);
        if (count > nd.sg_cap) scal[nd.pass0].err = 2;
    }
}

// length of lf_model_pulse for period T = 1 / sub_f0   (GOOFER.py:441-442: int(round(sr * T)), min 3)
__device__ __forceinline__ int gf_sub_len(double sf0, double sr)
{
    const double T = __ddiv_rn(1.0, sf0);
    double r = rint(__dmul_rn(sr, T));
    if (r > 1.0e9) r = 1.0e9;
    int n = (int)r;
    return n <= 3 ? 3 : n;
}

// ------------------------------------------------------------------------------------------------
// sg: pulse bank.  The reference caches one pulse per f'{sub_f0:.2f}' key: the FIRST event with a
// given two-decimal frequency fixes the pulse (and its length) of all later ones.
// One CTA per note: clear the table, insert every event (atomicCAS on the key, atomicMin on the
// event index), then resolve.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gf_sub_lf_max(double T, int n);

__global__ void __launch_bounds__(512)
gf_sg_bank_kernel(const int *__restrict__ list, const GfNotePlan *__restrict__ plans, const GfNoteDev *__restrict__ notes,
                  GfPassScal *scal)
{
    const int ni = list[blockIdx.x];
    const GfNotePlan &pl = plans[ni];
    const GfNoteDev nd = notes[ni];
    const int E = scal[nd.pass0].n_sub_events;
    const int M = nd.sg_tab_n;
    int *tab = reinterpret_cast<int *>(nd.sg_tab);
    for (int s = threadIdx.x; s < M; s += blockDim.x) { tab[2 * s] = -1; tab[2 * s + 1] = 0x7fffffff; }
    __syncthreads();
    const unsigned shift = 32u - (unsigned)(31 - __clz(M));
    auto slot_of = [&](int key) { return (int)(((unsigned)key * 2654435761u) >> shift) & (M - 1); };
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        const int key = (int)llrint(nd.sg_ev_f[e] * 100.0);     // '%.2f' bucket
        int s = slot_of(key);
        for (;;) {
            const int prev = atomicCAS(&tab[2 * s], -1, key);
            if (prev == -1 || prev == key) { atomicMin(&tab[2 * s + 1], e); break; }
            s = (s + 1) & (M - 1);
        }
    }
    __syncthreads();
    int mx = 0;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        const int key = (int)llrint(nd.sg_ev_f[e] * 100.0);
        int s = slot_of(key);
        while (tab[2 * s] != key) s = (s + 1) & (M - 1);
        const int r = tab[2 * s + 1];
        nd.sg_rep[e] = r;
        // the pulse of the event is its representative's: length and table peak once per event, not per rendered sample
        const double sf0 = nd.sg_ev_f[r];
        const int len = gf_sub_len(sf0, (double)pl.sr);
        nd.sg_len[e] = len;
        nd.sg_m[e] = gf_sub_lf_max(__ddiv_rn(1.0, sf0), len);
        mx = max(mx, len);
    }
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(&scal[nd.pass0].sub_max_len, mx);
}

// lf_model_pulse(T, sr, Ra=.02, Rg=1.7, Rk=1) sample j before the max normalisation (GOOFER.py:437-471):
// float32 time axis, fp64 comparisons and divisions (T is an np.float64 scalar), only pi * t in f32
__device__ __forceinline__ float gf_sub_lf_value(int j, double T, int n)
{
    const double step = __ddiv_rn(T, (double)n);
    const float t32 = (float)__dmul_rn((double)j, step);
    const double t = (double)t32;
    const double Tp = __dmul_rn(0.02, T);
    const double Tc = __dadd_rn(Tp, __dmul_rn(1.0, T - Tp));
    double v = 0.0;
    if (t < Tp) {
        const double a = (double)__fmul_rn(3.14159274101257324f, t32);
        const double s = sin(__ddiv_rn(a, __dmul_rn(2.0, Tp)));
        v = s * s;
    } else if (t < Tc) {
        const double tau = __ddiv_rn(t - Tp, Tc - Tp);
        v = exp(-1.7 * tau) * cos(3.141592653589793 * tau / 2.0);
    }
    return (float)v;
}

__device__ __forceinline__ float gf_sub_lf_max(double T, int n)
{
    const int jc = (int)(0.02 * (double)n);
    float m = 0.0f;
    for (int j = jc - 2; j <= jc + 3; ++j)
        if (j >= 0 && j < n) m = fmaxf(m, fabsf(gf_sub_lf_value(j, T, n)));
    return m;
}

// sub[i] = sum over the events covering i (event order) of bank[rep][i - onset]; also max |sub * mask|
__global__ void __launch_bounds__(256)
gf_sg_render_kernel(const int *__restrict__ list, const GfNotePlan *__restrict__ plans, const GfNoteDev *__restrict__ notes,
                    const GfPassDev *__restrict__ passes, GfPassScal *scal)
{
    const int ni = list[blockIdx.y];
    const GfNotePlan &pl = plans[ni];
    const GfNoteDev nd = notes[ni];
    const GfPassDev &ps = passes[nd.pass0];
    const int n = pl.n_total;
    const int E = scal[nd.pass0].n_sub_events;
    const int max_len = scal[nd.pass0].sub_max_len;
    float mx = 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int lo = 0, hi = E;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (nd.sg_ev_i[mid] <= i) lo = mid + 1; else hi = mid; }
        const int last = lo - 1;
        double acc = 0.0;
        if (last >= 0) {
            int e0 = last;
            while (e0 > 0 && i - nd.sg_ev_i[e0 - 1] < max_len) --e0;
            for (int e = e0; e <= last; ++e) {
                const int d = i - nd.sg_ev_i[e];
                const int len = nd.sg_len[e];
                if (d < len) {
                    const double T = __ddiv_rn(1.0, nd.sg_ev_f[nd.sg_rep[e]]);
                    const float raw = gf_sub_lf_value(d, T, len);
                    const float m = nd.sg_m[e];
                    acc += (double)((m > 0.0f) ? __fdiv_rn(raw, m) : raw);
                }
            }
        }
        const float s = (float)acc;
        ps.sub[i] = s;
        mx = fmaxf(mx, fabsf(s * nd.vm[i]));
    }
    mx = gf_warp_max(mx);
    if ((threadIdx.x & 31) == 0 && mx > 0.0f) gf_atomic_max_pos(&scal[nd.pass0].submax_bits, mx);
}

int gf_growl(const WaveHost &wh, const GfNotePlan *d_plans, const GfNoteDev *d_notes, const GfPassDev *d_passes,
             GfPassScal *d_scal, Bump &bp, int max_n, cudaStream_t st, int64_t *launches)
{
    std::vector<int> list;
    for (size_t i = 0; i < wh.plans.size(); ++i)
        if (wh.plans[i].add_subharm) list.push_back((int)i);
    if (list.empty()) return GOOFER_OK;
    int *d_list;
    int rc = gf_upload(bp, list, &d_list, st);
    if (rc != GOOFER_OK) return rc;
    if (bp.off > bp.cap) { gf_set_error("internal: growl list overflows the workspace"); return GOOFER_ERR_WORKSPACE; }
    const int nl = (int)list.size();
    dim3 g1(std::min(64, (max_n + 255) / 256), nl);
    gf_sg_f0_kernel<<<g1, 256, 0, st>>>(d_list, d_plans, d_notes, d_passes); ++*launches; GF_STEP("sg_f0");
#if !defined(GF_SG_SEQ_ONLY)
    gf_sg_scan_kernel<<<nl, GF_SGS_THREADS, 0, st>>>(d_list, d_plans, d_notes, d_scal); ++*launches; GF_STEP("sg_scan");
#endif
    gf_sg_walk_kernel<<<(nl + GF_SG_WARPS - 1) / GF_SG_WARPS, 32 * GF_SG_WARPS, 0, st>>>(d_list, nl, d_plans, d_notes, d_scal);
    ++*launches; GF_STEP("sg_walk");
    gf_sg_bank_kernel<<<nl, 512, 0, st>>>(d_list, d_plans, d_notes, d_scal); ++*launches; GF_STEP("sg_bank");
    gf_sg_render_kernel<<<g1, 256, 0, st>>>(d_list, d_plans, d_notes, d_passes, d_scal); ++*launches; GF_STEP("sg_render");
    return GOOFER_OK;
}

// ------------------------------------------------------------------------------------------------
// pd: 95th percentile (numpy 'linear' method) of |dev| by an 8 x 8-bit radix select on the fp64 bit
// patterns (non-negative doubles order like unsigned integers).  One CTA per note.
// ------------------------------------------------------------------------------------------------
#define GF_SEL_THREADS 1024
__device__ unsigned long long gf_select_rank(const double *__restrict__ x, int n, int rank, unsigned *hist, unsigned *bcast)
{
    // returns the bit pattern of the rank-th smallest |x| (0-based)
    unsigned long long prefix = 0ull;
    int remaining = rank;
    for (int pass = 7; pass >= 0; --pass) {
        const int sh = pass * 8;
        for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0u;
        __syncthreads();
        const unsigned long long himask = (pass == 7) ? 0ull : (~0ull << (sh + 8));
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long u = (unsigned long long)__double_as_longlong(fabs(x[i]));
            if ((u & himask) == prefix) atomicAdd(&hist[(unsigned)(u >> sh) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0, b = 0;
            for (; b < 256; ++b) {
                if (acc + (int)hist[b] > remaining) break;
                acc += (int)hist[b];
            }
            bcast[0] = (unsigned)b;
            bcast[1] = (unsigned)acc;
        }
        __syncthreads();
        prefix |= ((unsigned long long)bcast[0]) << sh;
        remaining -= (int)bcast[1];
        __syncthreads();
    }
    return prefix;
}

__global__ void __launch_bounds__(GF_SEL_THREADS)
gf_pd_ref_kernel(const int *__restrict__ list, const GfNotePlan *__restrict__ plans, const GfNoteDev *__restrict__ notes)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned bcast[2];
    __shared__ unsigned long long s_min;
    __shared__ int s_cnt;
    const int ni = list[blockIdx.x];
    const GfNotePlan &pl = plans[ni];
    const GfNoteDev nd = notes[ni];
    const int n = pl.n_total;
    const double *dev = nd.pd_dev;
    // numpy: virtual index (n - 1) * 0.95, gamma = frac, _lerp(a[lo], a[lo + 1], gamma)
    const double virt = (double)(n - 1) * (95.0 / 100.0);
    int lo = (int)floor(virt);
    if (lo > n - 1) lo = n - 1;
    const double gamma = virt - (double)lo;
    const unsigned long long ulo = gf_select_rank(dev, n, lo, hist, bcast);
    // the next order statistic: equal to a[lo] when more than lo + 1 elements are <= a[lo], else the smallest larger one
    if (threadIdx.x == 0) { s_min = ~0ull; s_cnt = 0; }
    __syncthreads();
    int cnt = 0;
    unsigned long long mn = ~0ull;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long u = (unsigned long long)__double_as_longlong(fabs(dev[i]));
        if (u <= ulo) ++cnt; else mn = min(mn, u);
    }
    atomicAdd(&s_cnt, cnt);
    atomicMin(&s_min, mn);
    __syncthreads();
    if (threadIdx.x == 0) {
        const double a = __longlong_as_double((long long)ulo);
        double b = a;
        if (lo + 1 <= n - 1 && s_cnt < lo + 2) b = __longlong_as_double((long long)s_min);
        const double diff = b - a;
        double r = (gamma >= 0.5) ? (b - diff * (1.0 - gamma)) : (a + diff * gamma);
        nd.noteScal[GF_NS_PDREF] = r + 1e-8;
    }
}

int gf_pitch_dyn(const WaveHost &wh, int n0, int n1, const GfNotePlan *d_plans, const GfNoteDev *d_notes, Bump &bp, int sr,
                 int max_n, cudaStream_t st, int64_t *launches)
{
    std::vector<int> list;
    std::vector<GfFirJob> jobs;
    for (size_t i = (size_t)n0; i < (size_t)n1; ++i) {
        const GfNotePlan &p = wh.plans[i];
        if (p.pd == 0.0) continue;
        list.push_back((int)i);
        const GfNoteDev &nd = wh.notes[i];
        GfFirJob j;
        std::memset(&j, 0, sizeof(j));
        j.in = nd.pd_in; j.in_stride = 1; j.n = p.n_total; j.out = nd.pd_dev; j.out_f64 = 1;
        j.sigma = (double)std::max(1, (int)(0.010 * sr));                 // SillySampler.py:865-866
        jobs.push_back(j);
        std::memset(&j, 0, sizeof(j));
        j.in = nd.vm; j.in_stride = 1; j.n = p.n_total; j.out = nd.pd_gm; j.out_f64 = 0;
        j.sigma = (double)(int)(0.01 * sr);                               // SillySampler.py:880
        jobs.push_back(j);
    }
    if (list.empty()) return GOOFER_OK;
    int *d_list; GfFirJob *d_jobs;
    int rc;
    if ((rc = gf_upload(bp, list, &d_list, st)) != GOOFER_OK) return rc;
    if ((rc = gf_upload(bp, jobs, &d_jobs, st)) != GOOFER_OK) return rc;
    if (bp.off > bp.cap) { gf_set_error("internal: pd jobs overflow the workspace"); return GOOFER_ERR_WORKSPACE; }
    *launches += gf_launch_fir(jobs.data(), d_jobs, (int)jobs.size(), st); GF_STEP("pd_fir");
    gf_pd_ref_kernel<<<(int)list.size(), GF_SEL_THREADS, 0, st>>>(d_list, d_plans, d_notes); ++*launches; GF_STEP("pd_ref");
    return GOOFER_OK;
}

// ------------------------------------------------------------------------------------------------
// post-FX element-wise stages (one CTA per note: the rms reductions stay deterministic)
// ------------------------------------------------------------------------------------------------
#define GF_FX_THREADS 1024

__device__ __forceinline__ double gf_block_sum(double v, double *red)
{
    v = gf_warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 32) {
        t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        t = gf_warp_sum(t);
    }
    return t;       // valid in warp 0
}

__global__ void __launch_bounds__(GF_FX_THREADS)
gf_fx_stage_kernel(int stage, const int *__restrict__ list, const GfNotePlan *__restrict__ plans,
                   const GfNoteDev *__restrict__ notes, const GfPassDev *__restrict__ passes, const GfPassScal *__restrict__ scal)
{
    __shared__ double red[32];
    const int ni = list[blockIdx.x];
    const GfNotePlan &pl = plans[ni];
    const GfNoteDev nd = notes[ni];
    const int n = pl.n_total;
    float *harm = nd.fx[0], *bre = nd.fx[1];
    int p_su = -1, p_sj = -1;
    for (int p = 1; p < pl.n_passes; ++p) {
        if (pl.pass_kind[p] == GF_PASS_SU) p_su = nd.pass0 + p;
        if (pl.pass_kind[p] == GF_PASS_SJ) p_sj = nd.pass0 + p;
    }
    const double a = fabs(pl.tension);
    if (stage == 0) {
        // normalised streams of the main pass and the harmonic layers (GOOFER.py:1208-1218)
        const GfPassDev &p0 = passes[nd.pass0];
        const GfPassScal &s0 = scal[nd.pass0];
        const float g0 = gf_pass_gain(pl, s0);
        const float gu = p_su >= 0 ? gf_pass_gain(pl, scal[p_su]) : 0.f, gj = p_sj >= 0 ? gf_pass_gain(pl, scal[p_sj]) : 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const GfStreams s = gf_pass_streams(pl, nd, p0, s0, i);
            harm[i] = s.h * g0;
            bre[i] = s.b * g0;
            if (p_su >= 0) nd.fx[2][i] = gf_pass_streams(pl, nd, passes[p_su], scal[p_su], i).h * gu;
            if (p_sj >= 0) nd.fx[3][i] = gf_pass_streams(pl, nd, passes[p_sj], scal[p_sj], i).h * gj;
        }
    } else if (stage == 1) {
        // SillySampler.py:1059 harm += hp * su ; :1081 harm = (1 - sj) * harm + sj * hp
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            double h = (double)harm[i];
            if (p_su >= 0) h = (double)(float)(h + (double)nd.fx[2][i] * pl.su);
            if (p_sj >= 0) h = (1.0 - pl.sj) * h + pl.sj * (double)nd.fx[3][i];
            harm[i] = (float)h;
        }
    } else if (stage == 2) {
        // fry blend :1096-1097, sd tremolo :1102-1112, rms before tension :1116
        const double sr = (double)pl.sr;
        const int fade = (int)(0.1 * sr);
        double part = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            float h = harm[i], b = bre[i];
            if (pl.fry_mask_on) {
                const float fm = gf_fry_at(pl, i);
                h = h * (1.0f - fm) + nd.fx[2][i] * fm;
                b = b * (1.0f - fm) + nd.fx[3][i] * fm;
            }
            if (pl.sd > 0) {
                // create_volume_jitter(vibrato=True, speed=150, strength=sd/200, seed=None)  GOOFER.py:642-659
                double s = sin((2.0 * 3.141592653589793) * 150.0 * ((double)i / sr) + 0.0);
                if (fade < n && i < fade) s *= gf_dlin01(i, fade);
                const double ej = fmin(fmax(1.0 + s * (pl.sd / 200.0), 0.5), 1.5);
                b = (float)((double)b * (1.0 + (ej - 1.0) * (double)nd.sdm[i]));
                b = b * (float)(1.0 + (pl.sd / 100.0) * 10.0);
            }
            harm[i] = h; bre[i] = b;
            const double x = (double)(h + b);
            part += x * x;
        }
        if (pl.tension != 0.0) {
            const double tot = gf_block_sum(part, red);
            if (threadIdx.x == 0) nd.noteScal[GF_NS_R0] = tot;
        }
    } else if (stage == 3) {
        // tension combine :1126-1134 and rms after :1136
        if (pl.tension == 0.0) return;
        double part = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            float h = harm[i], b = bre[i];
            if (pl.tension > 0.0) {
                h = h + nd.fx[2][i] * (float)(1.0 + a * 20.0);
                b = b * (float)(1.0 - a);
                harm[i] = h; bre[i] = b;
            }
            const double x = (double)(h + b);
            part += x * x;
        }
        const double tot = gf_block_sum(part, red);
        if (threadIdx.x == 0) nd.noteScal[GF_NS_R1] = tot;
    }
}

static GfOnepoleJob gf_job(const float *x, const float *f0, float *y, float *alpha, int n, int order, int highpass,
                           double cutoff, double f0_const, double f0_floor, int sr)
{
    GfOnepoleJob j;
    std::memset(&j, 0, sizeof(j));
    j.x = x; j.f0 = f0; j.y = y; j.alpha = alpha; j.n = n; j.order = order; j.highpass = highpass; j.smooth_f0 = 1;
    j.cutoff_factor = cutoff; j.f0_const = f0_const; j.f0_floor = f0_floor; j.sr = sr;
    return j;
}

int gf_post_fx(const WaveHost &wh, int n0, int n1, const GfNotePlan *d_plans, const GfNoteDev *d_notes, const GfPassDev *d_passes,
               GfPassScal *d_scal, Bump &bp, int sr, int max_n, cudaStream_t st, int64_t *launches)
{
    (void)max_n;
    std::vector<int> list;
    std::vector<GfOnepoleJob> jA, jB, jC;
    for (size_t i = (size_t)n0; i < (size_t)n1; ++i) {
        const GfNotePlan &p = wh.plans[i];
        const GfNoteDev &nd = wh.notes[i];
        if (!nd.fx[0]) continue;
        list.push_back((int)i);
        const int n = p.n_total;
        // su / sj: two order-6 high-pass calls on the same driver == one order-12 cascade  SillySampler.py:1052-1058, 1078-1080
        if (p.su > 0.0) jA.push_back(gf_job(nd.fx[2], nd.f0n, nd.fx[2], nd.alpha[0], n, 12, 1, 1.0, 0.0, 120.0, sr));
        if (p.sj > 0.0) jA.push_back(gf_job(nd.fx[3], nd.f0n, nd.fx[3], nd.alpha[1], n, 12, 1, 1.0, 0.0, 120.0, sr));
        if (p.fry_mask_on) {                                                        // :1090-1095
            jB.push_back(gf_job(nd.fx[0], nullptr, nd.fx[2], nd.alpha[0], n, 6, 1, 200.0, 1.0, 0.0, sr));
            jB.push_back(gf_job(nd.fx[1], nullptr, nd.fx[3], nd.alpha[1], n, 6, 1, 200.0, 1.0, 0.0, sr));
        }
        if (p.tension != 0.0) {                                                     // :1115-1134
            const double a = std::fabs(p.tension);
            if (p.tension < 0.0) {
                int order = (int)std::nearbyint(1.0 + a * 4.0);
                order = std::min(std::max(order, 1), 6);
                jC.push_back(gf_job(nd.fx[0], nd.f0n, nd.fx[0], nd.alpha[0], n, order, 0, 2.0 - a * 0.75, 0.0, 0.0, sr));
                jC.push_back(gf_job(nd.fx[1], nd.f0n, nd.fx[1], nd.alpha[1], n, 4, 1, a, 0.0, 0.0, sr));
            } else {
                jC.push_back(gf_job(nd.fx[0], nd.f0n, nd.fx[2], nd.alpha[0], n, 4, 1, a * 4.0, 0.0, 0.0, sr));
                jC.push_back(gf_job(nd.fx[1], nd.f0n, nd.fx[1], nd.alpha[1], n, 6, 0, (2.0 - a) / 0.5, 0.0, 0.0, sr));
            }
        }
    }
    if (list.empty()) return GOOFER_OK;
    int *d_list; GfOnepoleJob *dA, *dB, *dC;
    int rc;
    if ((rc = gf_upload(bp, list, &d_list, st)) != GOOFER_OK) return rc;
    if ((rc = gf_upload(bp, jA, &dA, st)) != GOOFER_OK) return rc;
    if ((rc = gf_upload(bp, jB, &dB, st)) != GOOFER_OK) return rc;
    if ((rc = gf_upload(bp, jC, &dC, st)) != GOOFER_OK) return rc;
    if (bp.off > bp.cap) { gf_set_error("internal: post-FX jobs overflow the workspace"); return GOOFER_ERR_WORKSPACE; }
    const int nl = (int)list.size();
    auto stage = [&](int s) { gf_fx_stage_kernel<<<nl, GF_FX_THREADS, 0, st>>>(s, d_list, d_plans, d_notes, d_passes, d_scal); ++*launches; };
    stage(0); GF_STEP("fx_prepare");
    if (!jA.empty()) { gf_launch_onepole(dA, (int)jA.size(), st); ++*launches; GF_STEP("fx_hp12"); stage(1); GF_STEP("fx_layers"); }
    if (!jB.empty()) { gf_launch_onepole(dB, (int)jB.size(), st); ++*launches; GF_STEP("fx_fry_hp"); }
    stage(2); GF_STEP("fx_fry_sd");
    if (!jC.empty()) { gf_launch_onepole(dC, (int)jC.size(), st); ++*launches; GF_STEP("fx_tension"); stage(3); GF_STEP("fx_rms"); }
    return GOOFER_OK;
}
