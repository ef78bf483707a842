// plan.cpp -- host planner: the integer / length bookkeeping of GooferResampler, no sample data.
//
// Mirrors, expression by expression (Python float semantics: int() truncates, // floors, round()
// is half-to-even), /root/reference/SillySampler.py:
//   :286-410  flag scalars          :453-500  slice bounds        :625-696  loop lengths
//   :699-712  mask tiling           :766-788  velocity stretch    :884-965  vocal-fry lengths
// Nothing here touches audio or envelope samples -- every numeric array is processed on the GPU.
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>

#include "../../include/goofer_b200.h"
#include "gf_plan.h"
#include "gf_internal.h"

namespace {

inline long long py_int(double x) { return (long long)std::trunc(x); }
inline long long py_floordiv(long long a, long long b) {
    long long q = a / b, r = a % b;
    if (r != 0 && ((r < 0) != (b < 0))) --q;
    return q;
}
inline long long py_round(double x) { return (long long)std::nearbyint(x); }   // FE_TONEAREST: half-even
inline double clipd(double x, double lo, double hi) { return std::min(std::max(x, lo), hi); }

// python slice a[start:stop] on a length-n sequence -> (first index, count)
inline void py_slice(long long start, long long stop, long long n, int32_t *a, int32_t *cnt) {
    auto norm = [n](long long i) { if (i < 0) { i += n; if (i < 0) i = 0; } else if (i > n) i = n; return i; };
    long long s = norm(start), e = norm(stop);
    *a = (int32_t)s;
    *cnt = (int32_t)std::max(0LL, e - s);
}

inline bool has(const GooferNote &n, int f) { return (n.present >> f) & 1ULL; }
inline int fl(const GooferNote &n, int f, int dflt) { return has(n, f) ? n.flag[f] : dflt; }

}  // namespace

int gf_plan_note(const GooferBatch *b, int idx, GfNotePlan *pl)
{
    const GooferNote &nt = b->notes[idx];
    std::memset(pl, 0, sizeof(*pl));
    pl->status = GOOFER_NOTE_OK;
    if (nt.source < 0 || nt.source >= b->n_sources) { pl->status = GOOFER_NOTE_BAD_SOURCE; return 0; }
    const GooferSource &sc = b->sources[nt.source];
    if (sc.T <= 0 || sc.N <= 0 || sc.sr <= 0 || sc.ylen <= 0) { pl->status = GOOFER_NOTE_BAD_SOURCE; return 0; }
    if (fl(nt, GF_SE, 0) == 1) { pl->status = GOOFER_NOTE_EDITOR; return 0; }      // SillySampler.py:309-310

    const int sr = sc.sr;
    pl->src = nt.source;
    pl->sr = sr;
    pl->T_src = sc.T;
    pl->N_src = sc.N;
    for (int k = 0; k < 4; ++k) pl->F_len[k] = sc.formants[k] ? sc.formant_len[k] : 0;

    // ---- flag scalars (SillySampler.py:313-410) ----
    pl->formant_shift = 1.0 + (fl(nt, GF_g, 0) / 200.0);
    pl->brightness_env = (fl(nt, GF_br, 0) + 100) / 100.0;
    const int fk[4] = {GF_fa, GF_fb, GF_fc, GF_fd};
    for (int k = 0; k < 4; ++k) {
        pl->F_shift[k] = 1.0 + (fl(nt, fk[k], 0) / 100.0);
        if (pl->F_shift[k] != 1.0) pl->any_F_shift = 1;
    }
    pl->f0_jitter = has(nt, GF_sh) && nt.flag[GF_sh] > 0;
    pl->f0_jitter_strength = fl(nt, GF_sh, 0) / 50.0;
    pl->vol_jitter = has(nt, GF_sr) && nt.flag[GF_sr] > 0;
    pl->vol_jitter_strength = fl(nt, GF_sr, 0) / 50.0;
    pl->sd = (double)fl(nt, GF_sd, 0);
    pl->B = (fl(nt, GF_B, 0) + 100) / 100.0;
    pl->U = (fl(nt, GF_U, 0) + 100) / 100.0;
    pl->V = clipd(fl(nt, GF_V, 100), 0, 100) / 100.0;
    {
        int L = fl(nt, GF_L, 0);
        pl->loop_mode = has(nt, GF_L) ? (L == 1 ? GF_LOOP_AVG : (L == 2 ? GF_LOOP_STRETCH : GF_LOOP_CONCAT)) : GF_LOOP_CONCAT;
    }
    pl->tension = fl(nt, GF_st, 0) / 100.0;
    {
        int sg = fl(nt, GF_sg, 0);
        pl->subharm_weight = (sg / 100.0) * 1.5;
        pl->add_subharm = sg > 0;
    }
    pl->reverse = fl(nt, GF_R, 0) == 1;
    pl->sj = clipd(fl(nt, GF_sj, 0), 0, 100) / 100.0;
    pl->sa = clipd(fl(nt, GF_sa, 0), 0, 100) / 100.0;
    pl->su = clipd(fl(nt, GF_su, 0), 0, 100) / 100.0;
    pl->normalize = has(nt, GF_P) ? clipd(nt.flag[GF_P], 0, 100) / 100.0 : 1.0;
    pl->es = clipd(fl(nt, GF_es, 0), -100, 100) / 100.0;
    pl->FV = fl(nt, GF_FV, 0) == 1;
    pl->pd = (double)(long long)clipd(fl(nt, GF_pd, 0), -100, 100) / 100.0;
    pl->fw = (fl(nt, GF_fw, 0) / 100.0) * 0.1;
    {
        double fst = clipd(fl(nt, GF_fst, 0), -100, 100) / 100.0;
        const int sk[4] = {GF_fsta, GF_fstb, GF_fstc, GF_fstd};
        for (int k = 0; k < 4; ++k) {
            pl->fst[k] = clipd(fst + (fl(nt, sk[k], 0) / 100.0), -1.0, 1.0);
            if (std::fabs(pl->fst[k]) >= 1e-6) pl->any_fst = 1;
        }
    }
    pl->vol_jitter_strength_breath = pl->vol_jitter_strength * 2;
    pl->breath_strength = 0.1f;
    pl->uv_strength = 0.75f;
    // continuous overrides of the direct gf.synthesize seam (GooferNote.override_val, GOOFER.py:971-983)
    {
        auto ovr = [&](int k) { return (nt.override_mask >> k) & 1u; };
        if (ovr(GF_OVR_FORMANT_SHIFT)) pl->formant_shift = nt.override_val[GF_OVR_FORMANT_SHIFT];
        for (int k = 0; k < 4; ++k)
            if (ovr(GF_OVR_F1_SHIFT + k)) pl->F_shift[k] = nt.override_val[GF_OVR_F1_SHIFT + k];
        pl->any_F_shift = 0;
        for (int k = 0; k < 4; ++k) if (pl->F_shift[k] != 1.0) pl->any_F_shift = 1;
        if (ovr(GF_OVR_F0_JITTER_STRENGTH)) pl->f0_jitter_strength = nt.override_val[GF_OVR_F0_JITTER_STRENGTH];
        if (ovr(GF_OVR_VOL_JITTER_HARM)) { pl->vol_jitter_strength = nt.override_val[GF_OVR_VOL_JITTER_HARM]; pl->vol_jitter_strength_breath = pl->vol_jitter_strength * 2; }
        if (ovr(GF_OVR_VOL_JITTER_BREATH)) pl->vol_jitter_strength_breath = nt.override_val[GF_OVR_VOL_JITTER_BREATH];
        if (ovr(GF_OVR_NORMALIZE)) pl->normalize = clipd(nt.override_val[GF_OVR_NORMALIZE], 0.0, 1.0);      // np.clip(normalize, 0, 1)  GOOFER.py:1211
        if (ovr(GF_OVR_BREATH_STRENGTH)) pl->breath_strength = (float)nt.override_val[GF_OVR_BREATH_STRENGTH];
        if (ovr(GF_OVR_UV_STRENGTH)) pl->uv_strength = (float)nt.override_val[GF_OVR_UV_STRENGTH];
    }
    pl->volume = nt.volume;
    pl->pitch_midi = nt.pitch_midi;
    pl->t_cents = fl(nt, GF_t, 0);
    pl->tempo = nt.tempo;
    pl->bend_off = nt.bend_off;
    pl->bend_len = nt.bend_len;
    if (nt.bend_len < 1 || nt.bend_off < 0 || nt.bend_off + nt.bend_len > b->bend_total) return -1;

    // ---- slice bounds (SillySampler.py:453-500) ----
    const double dur = (double)sc.ylen / (double)sr;
    const double end_base = (nt.cutoff_s < 0) ? (nt.offset_s - nt.cutoff_s) : (dur - nt.cutoff_s);
    double off_u, cut_u;
    if (pl->reverse) {
        double Lsec = end_base - nt.offset_s;
        off_u = dur - end_base;
        cut_u = dur - (off_u + Lsec);
    } else {
        off_u = nt.offset_s;
        cut_u = nt.cutoff_s;
    }
    long long s0 = py_int(off_u * sr);
    long long s1 = s0 + py_int(nt.consonant_s * sr);
    long long s2 = py_int(((cut_u < 0) ? (off_u - cut_u) : (dur - cut_u)) * sr);
    long long f0_ = py_floordiv(s0, GF_HOP), f1_ = py_floordiv(s1, GF_HOP), f2_ = py_floordiv(s2, GF_HOP);
    pl->f0_off = nt.f0_off;
    const bool direct = nt.f0_off >= 0;
    if (direct) {
        // direct gf.synthesize call: env_spec, voicing_mask and the formant tracks are the whole source
        // (GOOFER.py:986-1002), len(y) = N; nothing is sliced, looped or stretched
        if (pl->reverse) { pl->status = GOOFER_NOTE_BAD_SOURCE; return 0; }
        s0 = 0; s1 = 0; s2 = sc.N;
        f0_ = 0; f1_ = 0; f2_ = sc.T;
    }
    pl->fr0 = (int32_t)f0_; pl->fr1 = (int32_t)f1_; pl->fr2 = (int32_t)f2_;
    py_slice(f0_, f1_, sc.T, &pl->pre_f_a, &pl->pre_f_n);
    py_slice(f1_, f2_, sc.T, &pl->tail_f_a, &pl->tail_f_n);
    py_slice(s0, s1, sc.N, &pl->pre_s_a, &pl->pre_s_n);
    py_slice(s1, s2, sc.N, &pl->tail_s_a, &pl->tail_s_n);

    // ---- loop lengths (SillySampler.py:625-712) ----
    const long long want_samples = direct ? (long long)sc.N : py_int(nt.length_s * sr);
    const long long want_frames = direct ? (long long)sc.T : py_int(std::ceil(nt.length_s * sr / GF_HOP));
    pl->want_samples = (int32_t)want_samples;
    pl->want_frames = (int32_t)want_frames;
    const int have = pl->tail_f_n;
    if (want_samples < 0 || want_frames < 0) { pl->status = GOOFER_NOTE_TOO_SHORT; return 0; }
    if (have >= want_frames) {
        pl->env_direct = 1;
        pl->T_loop = (int32_t)want_frames;
    } else {
        if (have == 0) { pl->status = GOOFER_NOTE_EMPTY_TAIL; return 0; }     // ZeroDivisionError at :634
        pl->reps = (int32_t)(want_frames / have);
        pl->rem = (int32_t)(want_frames % have);
        if (pl->loop_mode == GF_LOOP_STRETCH) {
            pl->stretch_target = (int32_t)py_int(have * ((double)want_frames / (double)have));  // GOOFER.py:601
            pl->T_loop = pl->stretch_target;
        } else if (pl->loop_mode == GF_LOOP_AVG) {
            pl->T_loop = (int32_t)want_frames;
        } else {
            pl->fade = std::min(8, have / 2);
            // have == 1 => fade == 0: prev[:, :-0] is empty in the reference, each round leaves a bare tail
            pl->unit_len = pl->fade ? 2 * have - pl->fade : have;
            pl->fade_r = pl->rem ? std::min(8, pl->rem / 2) : 0;
            pl->T_loop = (pl->reps - 1) * pl->unit_len + have + pl->rem - pl->fade_r;
        }
    }
    if (pl->tail_s_n < want_samples && pl->tail_s_n == 0) { pl->status = GOOFER_NOTE_EMPTY_TAIL; return 0; }  // :704
    pl->T0_frames = pl->pre_f_n + pl->T_loop;
    pl->n0_total = pl->pre_s_n + (int32_t)want_samples;

    // ---- velocity (SillySampler.py:766-788, :176-204) ----
    pl->vel = std::pow(2.0, 1.0 - (nt.velocity / 100.0));
    pl->T_env = pl->T0_frames;
    pl->n_total = pl->n0_total;
    pl->pre_new_f = pl->pre_f_n;
    pl->pre_new_s = pl->pre_s_n;
    if (!direct && std::fabs(pl->vel - 1.0) > 1e-6 && pl->pre_f_n > 1 && pl->pre_s_n > 1) {
        pl->vel_active = 1;
        // _prefix_positions also needs n > 1, true here because pre_len > 1
        pl->pre_new_f = (int32_t)std::max(1LL, py_round(pl->pre_f_n * pl->vel));
        pl->pre_new_s = (int32_t)std::max(1LL, py_round(pl->pre_s_n * pl->vel));
        pl->T_env = pl->pre_new_f + (pl->T0_frames - pl->pre_f_n);
        pl->n_total = pl->pre_new_s + (pl->n0_total - pl->pre_s_n);
    }
    if (pl->n_total < 2 || pl->T_env < 1) { pl->status = GOOFER_NOTE_TOO_SHORT; return 0; }
    pl->T_out = 1 + pl->n_total / GF_HOP;

    // ---- vocal fry lengths (SillySampler.py:884-965) ----
    {
        double vf = (double)fl(nt, GF_vf, 0);
        pl->vh = std::max(1.0, (double)fl(nt, GF_vh, 50));
        pl->vl = clipd((double)fl(nt, GF_vl, 15), 0.0, 100.0);
        if (vf != 0) {
            vf = clipd(vf, -100.0, 100.0);
            pl->fry_on = 1;
            const int n = pl->n_total;
            long long L = py_int((double)py_round(n * (std::fabs(vf) / 100.0)));
            if (L > 0) {
                long long glide = py_round(L * (pl->vl / 100.0));
                glide = std::min(std::max(glide, 0LL), L);
                pl->fry_L = (int32_t)L;
                pl->fry_glide = (int32_t)glide;
                pl->fry_const = (int32_t)(L - glide);
            }
            const int mid = n / 2;
            long long a, e;
            if (vf > 0) {
                long long Lm = py_round(mid * (vf / 100.0));
                a = 0; e = std::max(0LL, std::min((long long)n, Lm));
            } else {
                long long Lm = py_round((n - mid) * (std::fabs(vf) / 100.0));
                a = std::max(0LL, n - Lm); e = n;
            }
            if (e > a) {
                pl->fry_mask_on = 1;
                pl->fry_a = (int32_t)a; pl->fry_b = (int32_t)e;
                pl->fry_fade = (int32_t)py_int(0.01 * sr);
            }
        }
        pl->vf = vf;
    }

    // ---- passes (SillySampler.py:1006, 1038, 1062, 1153) ----
    pl->n_passes = 0;
    pl->pass_kind[pl->n_passes++] = GF_PASS_MAIN;
    if (pl->su > 0.0) pl->pass_kind[pl->n_passes++] = GF_PASS_SU;
    if (pl->sj > 0.0) pl->pass_kind[pl->n_passes++] = GF_PASS_SJ;
    if (pl->sa > 0.0) pl->pass_kind[pl->n_passes++] = GF_PASS_SA;
    for (int k = 0; k < 4; ++k) { pl->phi_off[k] = nt.phi_off[k]; pl->nrm_off[k] = nt.nrm_off[k]; }
    pl->out_off = nt.out_off;
    pl->phi_rng_mask = nt.phi_rng_mask & 15u;
    for (int k = 0; k < 4; ++k)
        for (int q = 0; q < 4; ++q) pl->phi_rng[k][q] = nt.phi_rng[k][q];
    return 0;
}

void gf_plan_info(const GfNotePlan *pl, GooferNotePlanInfo *info)
{
    std::memset(info, 0, sizeof(*info));
    info->status = pl->status;
    if (pl->status != GOOFER_NOTE_OK) return;
    info->n_total = pl->n_total;
    info->t_out = pl->T_out;
    info->t_env = pl->T_env;
    info->n_passes = pl->n_passes;
    info->need_phi[0] = 1;
    info->need_phi[1] = pl->su > 0.0;
    info->need_phi[2] = pl->sj > 0.0;
    info->need_phi[3] = pl->sa > 0.0;
    info->need_nrm[0] = pl->f0_jitter;
    info->need_nrm[1] = pl->vol_jitter;
    info->need_nrm[2] = pl->vol_jitter;
    info->need_nrm[3] = pl->sj > 0.0;
}

extern "C" int goofer_plan_batch(const GooferBatch *b, GooferNotePlanInfo *info)
{
    if (!b || !info || b->n_notes < 0 || (b->n_notes > 0 && !b->notes) || (b->n_sources > 0 && !b->sources)) {
        gf_set_error("goofer_plan_batch: invalid descriptor");
        return GOOFER_ERR_INVALID;
    }
    int bad = 0;
    for (int i = 0; i < b->n_notes; ++i) {
        GfNotePlan pl;
        if (gf_plan_note(b, i, &pl) != 0) {
            gf_set_error("goofer_plan_batch: note %d has an invalid pitch-bend range", i);
            return GOOFER_ERR_INVALID;
        }
        gf_plan_info(&pl, &info[i]);
        if (pl.status != GOOFER_NOTE_OK) ++bad;
    }
    if (bad) {
        gf_set_error("goofer_plan_batch: %d note(s) cannot be rendered (see GooferNotePlanInfo.status)", bad);
        return GOOFER_ERR_NOTE;
    }
    return GOOFER_OK;
}

extern "C" int goofer_debug_plan(const GooferBatch *b, int32_t idx, void *out, size_t bytes)
{
    if (!b || !out || idx < 0 || idx >= b->n_notes) { gf_set_error("goofer_debug_plan: invalid arguments"); return GOOFER_ERR_INVALID; }
    GfNotePlan pl;
    if (gf_plan_note(b, idx, &pl) != 0) { gf_set_error("goofer_debug_plan: bad pitch-bend range"); return GOOFER_ERR_INVALID; }
    std::memcpy(out, &pl, std::min(bytes, sizeof(pl)));
    return (int)sizeof(pl);
}
