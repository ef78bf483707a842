// gf_maps.cuh -- index maps: which source frames / samples feed output frame t / sample i of a note.
//
// GooferResampler builds env_new / mask_new / formant tracks by slicing, tiling, cross-fading,
// averaging and linearly stretching *source* features (/root/reference/SillySampler.py:494-500,
// :625-763, :766-788; GOOFER.py:597-616).  All of those are linear maps with <= 4 taps, so instead of
// materialising the intermediate arrays the kernels evaluate the map per output element.
// __host__ __device__ so that tests/cpu_emul can compare the maps with the oracle (test-only).
#pragma once
#include "gf_hd.h"
#include "gf_plan.h"

struct GfMix {              // sum_k w[k] * src_frame[f[k]]   (f indexes the source in *stored* order)
    int n;
    int f[4];
    double w[4];
};

GF_HD void gf_py_slice(long long start, long long stop, long long n, int *a, int *cnt)
{
    long long s = start, e = stop;
    if (s < 0) { s += n; if (s < 0) s = 0; } else if (s > n) s = n;
    if (e < 0) { e += n; if (e < 0) e = 0; } else if (e > n) e = n;
    *a = (int)s;
    *cnt = (int)(e > s ? e - s : 0);
}

// np.linspace(0, 1, num)[i] in fp64 (numpy: i * step, endpoint stored exactly)
GF_HD double gf_lin01(int i, int num)
{
    if (num <= 1) return 0.0;
    if (i == num - 1) return 1.0;
    return (double)i * (1.0 / (double)(num - 1));
}
// np.linspace(1, 0, num)[i]
GF_HD double gf_lin10(int i, int num)
{
    if (num <= 1) return 1.0;
    if (i == num - 1) return 0.0;
    return (double)i * (-1.0 / (double)(num - 1)) + 1.0;
}

// position of x in xo = linspace(0, 1, m): largest j with xo[j] <= x (np.interp's bracket), j <= m-2
GF_HD int gf_lin01_bracket(double x, int m)
{
    int j = (int)(x * (double)(m - 1));
    if (j > m - 2) j = m - 2;
    if (j < 0) j = 0;
    while (j > 0 && gf_lin01(j, m) > x) --j;
    while (j < m - 2 && gf_lin01(j + 1, m) <= x) ++j;
    return j;
}

GF_HD void gf_mix_add(GfMix &m, int f, double w)
{
    if (w == 0.0) return;
    for (int k = 0; k < m.n; ++k)
        if (m.f[k] == f) { m.w[k] += w; return; }
    if (m.n < 4) { m.f[m.n] = f; m.w[m.n] = w; ++m.n; }
}

// stored index of (possibly reversed) source frame / sample
GF_HD int gf_src_frame(const GfNotePlan &p, int f) { return p.reverse ? (p.T_src - 1 - f) : f; }
GF_HD int gf_src_sample(const GfNotePlan &p, int s) { return p.reverse ? (p.N_src - 1 - s) : s; }

// env_tail_looped[:, v] as a mix of tail frames (SillySampler.py:628-696), scaled by `scale`
GF_HD void gf_env_loop_mix(const GfNotePlan &p, int v, double scale, GfMix &m)
{
    const int have = p.tail_f_n, a = p.tail_f_a;
    if (p.env_direct) { gf_mix_add(m, a + v, scale); return; }
    if (p.loop_mode == GF_LOOP_STRETCH) {                       // GOOFER.py:597-616
        if (have == 1) { gf_mix_add(m, a, scale); return; }
        const double x = gf_lin01(v, p.stretch_target);
        if (x >= 1.0) { gf_mix_add(m, a + have - 1, scale); return; }
        const int j = gf_lin01_bracket(x, have);
        const double x0 = gf_lin01(j, have), x1 = gf_lin01(j + 1, have);
        const double t = (x - x0) / (x1 - x0);
        gf_mix_add(m, a + j, scale * (1.0 - t));
        gf_mix_add(m, a + j + 1, scale * t);
        return;
    }
    if (p.loop_mode == GF_LOOP_AVG) {                           // SillySampler.py:647-652
        const int q = v % have;
        gf_mix_add(m, a + q, 0.5 * scale);
        gf_mix_add(m, a + have - 1 - q, 0.5 * scale);
        return;
    }
    // concat with 8-frame cross-fades inside each pair (SillySampler.py:654-696)
    const int body = (p.reps - 1) * p.unit_len;
    if (v < body) {
        const int q = v % p.unit_len;
        const int fade = p.fade;
        if (q < have - fade) { gf_mix_add(m, a + q, scale); return; }
        if (q < have) {
            const int j = q - (have - fade);
            gf_mix_add(m, a + q, scale * gf_lin10(j, fade));
            gf_mix_add(m, a + j, scale * gf_lin01(j, fade));
            return;
        }
        gf_mix_add(m, a + fade + (q - have), scale);
        return;
    }
    const int w = v - body;
    const int fr = p.fade_r;
    if (w < have - fr) { gf_mix_add(m, a + w, scale); return; }
    if (w < have) {
        const int j = w - (have - fr);
        gf_mix_add(m, a + w, scale * gf_lin10(j, fr));
        gf_mix_add(m, a + j, scale * gf_lin01(j, fr));
        return;
    }
    gf_mix_add(m, a + fr + (w - have), scale);
}

// env_new[:, u] before the velocity stretch
GF_HD void gf_env_prevel_mix(const GfNotePlan &p, int u, double scale, GfMix &m)
{
    if (u < p.pre_f_n) { gf_mix_add(m, p.pre_f_a + u, scale); return; }
    gf_env_loop_mix(p, u - p.pre_f_n, scale, m);
}

// stretch_prefix_* source position of output index idx (SillySampler.py:176-204)
GF_HD double gf_prefix_pos(int idx, int pre_new, int pre_len, double factor)
{
    return (idx < pre_new) ? ((double)idx / factor) : ((double)(idx - pre_new) + (double)pre_len);
}

// env_new[:, t] (after the velocity stretch), t < T_env
GF_HD void gf_env_mix(const GfNotePlan &p, int t, GfMix &m)
{
    m.n = 0;
    if (!p.vel_active) { gf_env_prevel_mix(p, t, 1.0, m); return; }
    const double pos = gf_prefix_pos(t, p.pre_new_f, p.pre_f_n, p.vel);
    const int n = p.T0_frames;
    int j = (int)pos;
    if (j >= n - 1) { gf_env_prevel_mix(p, n - 1, 1.0, m); return; }
    const double a = pos - (double)j;
    gf_env_prevel_mix(p, j, 1.0 - a, m);
    if (a != 0.0) gf_env_prevel_mix(p, j + 1, a, m);
}

// mask_new before the velocity stretch: source sample index (stored order), or -1 for "1.0" (FV)
GF_HD int gf_mask_prevel_src(const GfNotePlan &p, int u)
{
    int s;
    if (u < p.pre_s_n) s = p.pre_s_a + u;
    else {
        int w = u - p.pre_s_n;
        if (p.tail_s_n < p.want_samples) w %= p.tail_s_n;
        s = p.tail_s_a + w;
    }
    return gf_src_sample(p, s);
}

GF_HD double gf_mask_prevel(const GfNotePlan &p, const float *mask_src, int u)
{
    if (p.FV) return 1.0;
    return (double)mask_src[gf_mask_prevel_src(p, u)];
}

// mask_new[i] (fp64 when the velocity stretch ran, else the f32 source value)   SillySampler.py:766-788
GF_HD double gf_mask_new(const GfNotePlan &p, const float *mask_src, int i)
{
    if (!p.vel_active) return gf_mask_prevel(p, mask_src, i);
    const double pos = gf_prefix_pos(i, p.pre_new_s, p.pre_s_n, p.vel);
    const int n = p.n0_total;
    int j = (int)pos;
    if (j >= n - 1) return gf_mask_prevel(p, mask_src, n - 1);
    const double y0 = gf_mask_prevel(p, mask_src, j), y1 = gf_mask_prevel(p, mask_src, j + 1);
    const double slope = (y1 - y0) / 1.0;
    double d = pos - (double)j;
#if defined(__CUDA_ARCH__)
    return __dadd_rn(__dmul_rn(slope, d), y0);
#else
    volatile double prod = slope * d;
    return prod + y0;
#endif
}

// ---- formant tracks (SillySampler.py:714-763, :776-788, :242-262) ------------------------------
struct GfTrackSlices { int pre_a, pre_n, tail_a, tail_n, loop_len; };

GF_HD GfTrackSlices gf_track_slices(const GfNotePlan &p, int k)
{
    GfTrackSlices s;
    const int L = p.F_len[k];
    gf_py_slice(p.fr0, p.fr1, L, &s.pre_a, &s.pre_n);
    gf_py_slice(p.fr1, p.fr2, L, &s.tail_a, &s.tail_n);
    const int want = p.want_frames, size = s.tail_n;
    if (size == 0) s.loop_len = want;
    else if (p.loop_mode == GF_LOOP_STRETCH) {
        const double st = (double)want / (double)size;
        s.loop_len = (st == 1.0) ? size : (int)((double)size * st);
    } else s.loop_len = want;
    return s;
}

GF_HD double gf_trk_src(const GfNotePlan &p, const double *trk, int k, int f)
{
    return trk[p.reverse ? (p.F_len[k] - 1 - f) : f];
}

// _loop_track output element v (f32)
GF_HD float gf_track_loop(const GfNotePlan &p, const GfTrackSlices &s, const double *trk, int k, int v)
{
    const int size = s.tail_n, a = s.tail_a;
    if (size == 0) return 0.0f;
    if (p.loop_mode == GF_LOOP_STRETCH) {
        const double st = (double)p.want_frames / (double)size;
        if (st == 1.0) return (float)gf_trk_src(p, trk, k, a + v);
        if (size == 1) return (float)gf_trk_src(p, trk, k, a);
        const double x = gf_lin01(v, s.loop_len);
        if (x >= 1.0) return (float)gf_trk_src(p, trk, k, a + size - 1);
        const int j = gf_lin01_bracket(x, size);
        const double x0 = gf_lin01(j, size), x1 = gf_lin01(j + 1, size);
        const double y0 = (double)(float)gf_trk_src(p, trk, k, a + j), y1 = (double)(float)gf_trk_src(p, trk, k, a + j + 1);
        const double slope = (y1 - y0) / (x1 - x0);
        return (float)(slope * (x - x0) + y0);
    }
    const int q = v % size;
    if (p.loop_mode == GF_LOOP_AVG) {
        const float t0 = (float)gf_trk_src(p, trk, k, a + q), t1 = (float)gf_trk_src(p, trk, k, a + size - 1 - q);
        return (t0 + t1) * 0.5f;
    }
    return (float)gf_trk_src(p, trk, k, a + q);
}

// track before the velocity stretch, padded (edge) / trimmed to T0_frames; u < T0_frames
GF_HD double gf_track_prevel(const GfNotePlan &p, const GfTrackSlices &s, const double *trk, int k, int u)
{
    const int len = s.pre_n + s.loop_len;
    if (len <= 0) return 0.0;
    if (u > len - 1) u = len - 1;
    if (u < s.pre_n) return gf_trk_src(p, trk, k, s.pre_a + u);
    return (double)gf_track_loop(p, s, trk, k, u - s.pre_n);
}

// canon["Fk"] evaluated on the env_new frame grid: t < T_env   (f32 like canon_formants)
GF_HD float gf_track_canon(const GfNotePlan &p, const GfTrackSlices &s, const double *trk, int k, int t)
{
    if (p.F_len[k] <= 0) return 0.0f;
    if (t > p.T0_frames - 1) t = p.T0_frames - 1;
    if (!p.vel_active) return (float)gf_track_prevel(p, s, trk, k, t);
    const double pos = gf_prefix_pos(t, p.pre_new_f, p.pre_f_n, p.vel);
    const int n = p.T0_frames;
    int j = (int)pos;
    if (j >= n - 1) return (float)gf_track_prevel(p, s, trk, k, n - 1);
    const double y0 = gf_track_prevel(p, s, trk, k, j), y1 = gf_track_prevel(p, s, trk, k, j + 1);
    return (float)((y1 - y0) * (pos - (double)j) + y0);
}
