// api.cu -- the C ABI of libgoofer_b200.so (include/goofer_b200.h): batch orchestration, workspace
// carving, wave scheduling, the host-buffer convenience entry point and the stage-level calls.
//
// There is no CPU fallback anywhere in this file: every numeric array is produced by a kernel.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include <mutex>
#include <algorithm>
#include <functional>
#include <ctime>
#include <utility>

#include "gf_internal.h"
#include "gf_kernels.h"

// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static thread_local GooferStats g_stats = {0, 0, 0, 0};

void gf_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *goofer_last_error(void) { return g_err; }
extern "C" int goofer_version(void) { return GOOFER_ABI_VERSION; }
extern "C" void goofer_last_stats(GooferStats *s) { if (s) *s = g_stats; }
extern "C" size_t goofer_struct_size(int which)
{
    switch (which) {
    case 0: return sizeof(GooferSource);
    case 1: return sizeof(GooferNote);
    case 2: return sizeof(GooferNotePlanInfo);
    case 3: return sizeof(GooferBatch);
    case 4: return sizeof(GooferStats);
    default: return 0;
    }
}

#define GF_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            gf_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return GOOFER_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

// GOOFER_DEBUG_SYNC=1: synchronise after every launch and name the kernel that failed (bring-up aid)
static bool gf_debug_sync()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("GOOFER_DEBUG_SYNC"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

// per-kernel device timing (goofer_profile): one CUDA event after every launch, on the launching stream
struct GfProf {
    bool on = false;
    bool serial = false;                                   // goofer_profile(2): both preparation chains on the caller's stream
    std::vector<cudaEvent_t> pool;
    std::vector<const char *> names;
    std::vector<cudaStream_t> streams;                     // the stream each mark was recorded on
    size_t used = 0;
    std::string summary;
};
static thread_local GfProf g_prof;

static void gf_prof_mark(const char *name, cudaStream_t st)
{
    if (!g_prof.on) return;
    if (g_prof.used == g_prof.pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        g_prof.pool.push_back(e);
    }
    cudaEventRecord(g_prof.pool[g_prof.used], st);
    if (g_prof.names.size() <= g_prof.used) { g_prof.names.push_back(name); g_prof.streams.push_back(st); }
    else { g_prof.names[g_prof.used] = name; g_prof.streams[g_prof.used] = st; }
    ++g_prof.used;
}

extern "C" void goofer_profile(int enable)
{
    g_prof.on = enable != 0;
    g_prof.serial = enable == 2;
    g_prof.used = 0;
}

// "name:launches:total_ms;..." over every render call since goofer_profile(1); synchronises on every event.
// A kernel's time is the span between its mark and the previous mark ON THE SAME STREAM ("begin" / "fork" /
// "join" marks only open a span).  Kernels of the two preparation chains run concurrently on two streams, so
// their spans overlap in time and do not add up to the step; the frame kernel runs alone after the join.
extern "C" const char *goofer_profile_summary(void)
{
    g_prof.summary.clear();
    if (g_prof.used < 2) return g_prof.summary.c_str();
    for (size_t i = 0; i < g_prof.used; ++i) cudaEventSynchronize(g_prof.pool[i]);
    std::vector<const char *> order;
    std::vector<double> tot;
    std::vector<long> cnt;
    for (size_t i = 1; i < g_prof.used; ++i) {
        const char *nm = g_prof.names[i];
        if (std::strcmp(nm, "begin") == 0 || std::strcmp(nm, "fork") == 0 || std::strcmp(nm, "join") == 0) continue;
        size_t j = i;
        while (j > 0 && g_prof.streams[j - 1] != g_prof.streams[i]) --j;
        if (j == 0) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_prof.pool[j - 1], g_prof.pool[i]) != cudaSuccess) continue;
        size_t k = 0;
        while (k < order.size() && std::strcmp(order[k], nm) != 0) ++k;
        if (k == order.size()) { order.push_back(nm); tot.push_back(0.0); cnt.push_back(0); }
        tot[k] += ms; ++cnt[k];
    }
    char buf[128];
    for (size_t k = 0; k < order.size(); ++k) {
        snprintf(buf, sizeof(buf), "%s:%ld:%.6f;", order[k], cnt[k], tot[k]);
        g_prof.summary += buf;
    }
    return g_prof.summary.c_str();
}

// GOOFER_HOST_TRACE: host-side timestamps of the enqueue path (printed by goofer_render_batch_host)
struct GfHostTrace { bool on = false; std::vector<std::pair<const char *, double>> marks; };
static thread_local GfHostTrace g_htrace;
static inline void gf_htrace(const char *name)
{
    if (!g_htrace.on) return;
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    g_htrace.marks.push_back({name, 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec});
}

#define GF_STEP(name)                                                                              \
    do {                                                                                           \
        gf_prof_mark(name, st);                                                                    \
        if (gf_debug_sync()) {                                                                     \
            cudaError_t e_ = cudaStreamSynchronize(st);                                            \
            if (e_ == cudaSuccess) e_ = cudaGetLastError();                                        \
            if (e_ != cudaSuccess) { gf_set_error("kernel %s failed: %s", name, cudaGetErrorString(e_)); return GOOFER_ERR_CUDA; } \
        }                                                                                          \
    } while (0)

// ------------------------------------------------------------------------------------------------
// constant tables, computed in fp64 with the reference's own formulas
// ------------------------------------------------------------------------------------------------
static void gf_brightness(float *dst, int sr, double start_hz, double end_hz, double gain_db)
{
    // GOOFER.py:585-595
    const int n = GF_NBINS;
    const double nyq = sr / 2.0;
    std::vector<double> f(n);
    for (int i = 0; i < n; ++i) f[i] = (i == n - 1) ? nyq : i * (nyq / (n - 1));
    int a = (int)(std::lower_bound(f.begin(), f.end(), start_hz) - f.begin());
    int b = (int)(std::lower_bound(f.begin(), f.end(), end_hz) - f.begin());
    const double top = std::pow(10.0, gain_db / 20.0);
    for (int i = 0; i < n; ++i) {
        double g = 1.0;
        if (i >= a && i < b) {
            const int m = b - a;
            const double lin = (m <= 1) ? 0.0 : ((i - a == m - 1) ? 1.0 : (i - a) * (1.0 / (m - 1)));
            g = 1.0 + lin * (top - 1.0);
        } else if (i >= b) g = top;
        dst[i] = (float)g;
    }
}

static void gf_gauss_taps(double *dst, double sigma)
{
    const int r = (int)(4.0 * sigma + 0.5);
    double norm = 0.0;
    for (int j = 0; j <= 2 * r; ++j) { const double t = (j - r) / sigma; dst[j] = std::exp(-0.5 * t * t); norm += dst[j]; }
    for (int j = 0; j <= 2 * r; ++j) dst[j] /= norm;
}

static std::mutex g_tab_mutex;
static int g_tab_sr[64] = {0};

// sr <= 0: any table will do (the STFT / iSTFT stage calls use only the sr-independent members); 44,100 is loaded when
// the device has none yet.  Replacing a table of another rate first waits for the device: kernels of an earlier
// asynchronous call on a non-blocking stream may still be reading d_tab (the copy below is synchronous only with
// respect to the null stream).
int gf_tables_init(int sr)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { gf_set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError())); return GOOFER_ERR_CUDA; }
    std::lock_guard<std::mutex> lk(g_tab_mutex);
    if (dev < 64 && g_tab_sr[dev] != 0 && (sr <= 0 || g_tab_sr[dev] == sr)) return 0;
    if (sr <= 0) sr = 44100;
    if (dev < 64 && g_tab_sr[dev] != 0) cudaDeviceSynchronize();
    GfTables *t = new GfTables();
    const double PI = 3.141592653589793238462643383279502884;
    gf_twl_fill(t->twl);
    for (int k = 0; k <= 512; ++k) t->tw1024[k] = make_float2((float)std::cos(-2.0 * PI * k / 1024.0), (float)std::sin(-2.0 * PI * k / 1024.0));
    for (int i = 0; i < 1024; ++i) {
        // np.hanning(M): 0.5 + 0.5 cos(pi n / (M - 1)), n = 1 - M, 3 - M, ... ; cast f32, then sqrt in f32
        const double nn = (double)(1 - 1024 + 2 * i);
        const float h = (float)(0.5 + 0.5 * std::cos(PI * nn / 1023.0));
        t->win[i] = std::sqrt(h);
        t->win2[i] = t->win[i] * t->win[i];
    }
    for (int i = 0; i < GF_NBINS; ++i) {
        t->boost[i] = (float)((i == GF_NBINS - 1) ? 100.0 : i * (99.0 / 512.0) + 1.0);
        t->freq32[i] = (float)((double)i / (1024.0 * (1.0 / (double)sr)));
    }
    gf_brightness(t->bright_h, sr, 2000, 3500, 3.0);
    gf_brightness(t->bright_b, sr, 3500, 5000, 20.0);
    gf_gauss_taps(t->g175, 1.75);
    gf_gauss_taps(t->g05, 0.5);
    {
        // time-domain image of the 5-tap blur and the taps of its inverse (k_frame.cu gf_blur_edges)
        std::vector<double> G(1024), ginv(1024);
        for (int n = 0; n < 1024; ++n)
            G[n] = t->g05[2] + 2.0 * t->g05[1] * std::cos(2.0 * PI * n / 1024.0) + 2.0 * t->g05[0] * std::cos(4.0 * PI * n / 1024.0);
        for (int j = 0; j < 32; ++j) {
            double a = 0.0;
            for (int n = 0; n < 1024; ++n) a += std::cos(2.0 * PI * (double)j * n / 1024.0) / G[n];
            ginv[j] = a / 1024.0;
        }
        for (int n = 0; n < 1024; ++n) t->winG[n] = (float)((double)t->win[n] * G[n]);
        for (int k = 0; k < 16; ++k) {
            // ginv is even: ginv[-j] = ginv[j]
            t->bq1[k] = (float)(ginv[std::abs(k - 1)] - ginv[k + 1]);
            t->bq2[k] = (float)(ginv[std::abs(k - 2)] - ginv[k + 2]);
        }
    }
    t->sr = sr;
    cudaError_t e = cudaMemcpyToSymbol(d_tab, t, sizeof(GfTables));
    delete t;
    if (e == cudaSuccess && (dev >= 64 || g_tab_sr[dev] == 0)) {      // independent of the sample rate: once per device
        GfConvTables *c = new GfConvTables();
        gf_conv_tw_fill<float, GF_CONV_N32>(c->tw32);
        gf_conv_tw_fill<double, GF_CONV_N64>(c->tw64);
        e = cudaMemcpyToSymbol(d_conv, c, sizeof(GfConvTables));
        delete c;
    }
    if (e != cudaSuccess) { gf_set_error("table upload failed: %s", cudaGetErrorString(e)); return GOOFER_ERR_CUDA; }
    if (dev < 64) g_tab_sr[dev] = sr;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// workspace carving
// ------------------------------------------------------------------------------------------------
struct Bump {
    char *base; size_t cap; size_t off;
    void *take(size_t bytes) {
        // align the absolute address (base itself may sit at any offset inside the caller's workspace)
        const size_t mis = base ? (size_t)((uintptr_t)base & 255) : 0;
        off = ((off + mis + 255) & ~(size_t)255) - mis;
        void *p = base ? base + off : nullptr;
        off += bytes;
        return p;
    }
    template <typename T> T *arr(size_t n) { return (T *)take(n * sizeof(T)); }
};

static inline bool note_needs_fx(const GfNotePlan &p)
{
    return p.su > 0.0 || p.sj > 0.0 || p.fry_mask_on || p.sd > 0 || p.tension != 0.0;
}

struct NoteLayout {      // what one note takes from the wave region (computed identically for sizing and rendering)
    size_t bytes;
};

// carve (or just size, when bp.base == NULL) the buffers of one note
static void gf_carve_note(const GfNotePlan &p, Bump &bp, GfNoteDev *nd, GfPassDev *pd /* n_passes entries */, bool taps)
{
    const size_t n = (size_t)p.n_total;
    GfNoteDev d;
    std::memset(&d, 0, sizeof(d));
    d.trk_canon = bp.arr<float>(4 * (size_t)p.T_env);
    d.trk_clean = p.any_fst ? bp.arr<float>(8 * (size_t)p.T_env) : nullptr;
    d.env_aux = bp.arr<float>(1040 + 64);
    d.envF = bp.arr<float>((size_t)p.T_out * GF_ENVS_LD);
    d.envN = bp.arr<float>((size_t)p.T_out * GF_ENVS_LD);
    d.vm = bp.arr<float>(n);
    d.vm4 = bp.arr<float>((n + 3) / 4);
    d.ms_short = bp.arr<float>((n + 3) / 4);
    d.ms = bp.arr<float>(n);
    d.ms_one = bp.arr<unsigned char>((n + 255) / 256 + 4);
    if (p.f0_jitter) d.z_sh = bp.arr<double>(n);
    if (p.vol_jitter) { d.z_srh = bp.arr<double>(n); d.z_srb = bp.arr<double>(n); d.vjm = bp.arr<float>(n); }
    if (p.sd > 0) d.sdm = bp.arr<float>(n);
    if (p.pd != 0.0) { d.pd_in = bp.arr<float>(n); d.pd_dev = bp.arr<double>(n); d.pd_gm = bp.arr<float>(n); }
    if (note_needs_fx(p)) {
        d.f0n = bp.arr<float>(n);
        for (int k = 0; k < 4; ++k) d.fx[k] = bp.arr<float>(n);
        for (int k = 0; k < 2; ++k) d.alpha[k] = bp.arr<float>(n);
    }
    if (p.add_subharm) {
        d.sg_f0 = bp.arr<float>(n);
        d.sg_cap = p.n_total / 4 + 64;
        d.sg_ev_i = bp.arr<int>((size_t)d.sg_cap);
        d.sg_ev_f = bp.arr<double>((size_t)d.sg_cap);
        d.sg_rep = bp.arr<int>((size_t)d.sg_cap);
        d.sg_len = bp.arr<int>((size_t)d.sg_cap);
        d.sg_m = bp.arr<float>((size_t)d.sg_cap);
        int m = 128;
        while (m < 2 * d.sg_cap) m <<= 1;
        d.sg_tab_n = m;
        d.sg_tab = bp.arr<int2>((size_t)m);
    }
    (void)taps;
    for (int k = 0; k < p.n_passes; ++k) {
        GfPassDev q;
        std::memset(&q, 0, sizeof(q));
        q.kind = p.pass_kind[k];
        q.n_total = p.n_total;
        q.T_out = p.T_out;
        q.f0 = bp.arr<float>(n);
        q.pulse = bp.arr<float>(n);
        q.harm = bp.arr<float>(n);
        q.bre = bp.arr<float>(n);
        q.uv = bp.arr<float>(n);
        q.onset_cap = p.n_total / 8 + 64;
        q.onsets = bp.arr<int4>((size_t)q.onset_cap);
        q.phi = bp.arr<float>((size_t)GF_ENVS_LD * p.T_out);                                             // frame-major; drawn or transposed
        if (k == 0 && p.add_subharm) q.sub = bp.arr<float>(n);
        q.mask_ones = (q.kind == GF_PASS_SA);
        if (pd) pd[k] = q;
    }
    if (nd) *nd = d;
}

// What gf_carve_note takes from the wave region depends on a handful of plan fields only; sizing 1,024 notes by dry
// carving cost 0.15 ms per pass over the batch (twice per host-entry call), so the result is memoised per key.
static size_t gf_note_bytes(const GfNotePlan &p)
{
    struct Key {
        int n_total, T_env, T_out, n_passes, kinds, bits; unsigned rng;
        bool operator==(const Key &o) const { return n_total == o.n_total && T_env == o.T_env && T_out == o.T_out && n_passes == o.n_passes && kinds == o.kinds && bits == o.bits && rng == o.rng; }
    };
    static thread_local Key last_key = {-1, 0, 0, 0, 0, 0, 0};
    static thread_local size_t last_val = 0;
    Key k;
    k.n_total = p.n_total; k.T_env = p.T_env; k.T_out = p.T_out; k.n_passes = p.n_passes;
    k.kinds = 0;
    for (int q = 0; q < p.n_passes; ++q) k.kinds = k.kinds * 8 + p.pass_kind[q] + 1;
    k.bits = (p.any_fst ? 1 : 0) | (p.f0_jitter ? 2 : 0) | (p.vol_jitter ? 4 : 0) | (p.sd > 0 ? 8 : 0) | (p.pd != 0.0 ? 16 : 0) |
             (note_needs_fx(p) ? 32 : 0) | (p.add_subharm ? 64 : 0);
    k.rng = p.phi_rng_mask;
    if (k == last_key) return last_val;
    Bump bp{nullptr, 0, 0};
    gf_carve_note(p, bp, nullptr, nullptr, false);
    last_key = k;
    last_val = bp.off + 256;
    return last_val;
}

static size_t gf_sources_bytes(const GooferBatch *b)
{
    size_t tot = 256;
    for (int s = 0; s < b->n_sources; ++s) tot += (((size_t)std::max(b->sources[s].T, 0) * GF_ENVS_LD * sizeof(float)) + 255) & ~(size_t)255;
    tot += (size_t)b->n_sources * sizeof(GfSourceDev) + 256;
    return tot;
}

// per-wave bookkeeping arrays (plans, note / pass records, scalars, work lists, job lists)
static size_t gf_wave_meta_bytes(size_t n_notes, size_t n_pass, size_t n_env_work, size_t n_frame_work, size_t n_fir)
{
    return n_notes * (sizeof(GfNotePlan) + sizeof(GfNoteDev) + GF_NS_COUNT * sizeof(double)) + n_pass * (sizeof(GfPassDev) + sizeof(GfPassScal)) +
           n_env_work * sizeof(int2) + n_frame_work * sizeof(int4) + n_fir * sizeof(GfFirJob) +
           n_pass * (16 * sizeof(GfOnepoleJob) + sizeof(GfPhiJob) + 256) + n_notes * (3 * sizeof(int) + 2 * sizeof(GfFirJob)) + 32 * 256;
}

#ifndef GF_BLOCKS_PER_CTA
#define GF_BLOCKS_PER_CTA 90      // at most this many output hop blocks per frame-kernel CTA (3 frames of halo each); a 1 s note
                                  // (172 blocks) splits in 2 x 86: 1.97 -> 1.89 ms against 3 x 58 (cap 64)
#endif

static_assert(GF_BLOCKS_PER_CTA + 3 <= GF_FRAME_META, "per-frame tables of the frame kernel (k_frame.cu)");

// balanced split of a note's hop blocks: as few CTAs as the cap allows, equal shares
static inline int gf_blocks_per_cta(int n_blocks)
{
    const int parts = (n_blocks + GF_BLOCKS_PER_CTA - 1) / GF_BLOCKS_PER_CTA;
    return (n_blocks + parts - 1) / std::max(parts, 1);
}

static void gf_note_work_counts(const GfNotePlan &p, size_t *n_env, size_t *n_frame, size_t *n_fir)
{
    *n_env = (size_t)(p.T_out + GF_FT - 1) / GF_FT;
    const int n_blocks = std::max(p.T_out, 2) - 2 + 1;
    const int bpc = gf_blocks_per_cta(n_blocks);
    *n_frame = (size_t)p.n_passes * (2 * ((n_blocks + bpc - 1) / bpc) + 2);      // upper bound: gf_balance_frame_group may split finer
    *n_fir = 8;
}

// CTA slots of the frame kernel on this device (resident CTAs per SM x SMs)
static int gf_frame_slots()
{
    static int slots[GF_MAX_DEVICES] = {0};                  // per device: one process may drive several GPUs
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= GF_MAX_DEVICES) return 148 * GF_FRAME_CTAS;
    if (slots[dev] == 0)
        slots[dev] = (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0 ? sms : 148) * GF_FRAME_CTAS;
    return slots[dev];
}

// Tail of a frame-kernel launch.  The notes [a, e) of one launch are cut into `fparts[i]` CTAs per pass; the CTAs run
// in waves of `slots`.  When the last wave is far from full, the notes whose CTAs make it up are cut finer, so that
// their (shorter) CTAs fill the slots of that wave once: 1,024 one-second notes = 2,048 CTAs of 86 hop blocks on 444
// slots = 4.6 waves; with the last 136 notes in three CTAs each the tail wave takes 61 / 89 of a full one.
static void gf_balance_frame_group(const std::vector<GfNotePlan> &plans, int a, int e, int slots, std::vector<int> &fparts)
{
    long items = 0;
    for (int i = a; i < e; ++i) items += (long)fparts[i] * plans[i].n_passes;
    const long rem = items % slots;
    if (rem == 0 || rem * 10 >= (long)slots * 9) return;
    long acc = 0;
    int t0 = e;
    while (t0 > a && acc < rem) { --t0; acc += (long)fparts[t0] * plans[t0].n_passes; }
    if (acc <= 0 || acc > slots) return;
    for (int i = t0; i < e; ++i) {
        const int n_blocks = std::max(plans[i].T_out, 2) - 2 + 1;
        int want = (int)((long)fparts[i] * slots / acc);
        want = std::min(want, std::min(2 * fparts[i] + 2, std::max(1, n_blocks / 16)));     // keep the 3-frame halo small
        if (want > fparts[i]) fparts[i] = want;
    }
}

static int gf_validate(const GooferBatch *b)
{
    if (!b || b->n_notes < 0 || b->n_sources < 0 || (b->n_notes && !b->notes) || (b->n_sources && !b->sources) ||
        b->bend_total < 0 || b->phi_total < 0 || b->nrm_total < 0 || b->out_total < 0 || b->f0_total < 0) {
        gf_set_error("invalid batch descriptor (NULL or negative field)");
        return GOOFER_ERR_INVALID;
    }
    for (int s = 0; s < b->n_sources; ++s) {
        const GooferSource &g = b->sources[s];
        bool bad = g.K < 0 || g.T < 0 || g.N < 0;
        for (int k = 0; k < 4; ++k) bad = bad || g.formant_len[k] < 0;
        if (bad) { gf_set_error("source %d: negative size", s); return GOOFER_ERR_INVALID; }
    }
    return GOOFER_OK;
}

static int gf_make_plans(const GooferBatch *b, std::vector<GfNotePlan> &plans)
{
    plans.resize(b->n_notes);
    int bad = 0, first_bad = -1, sr = 0;
    for (int i = 0; i < b->n_notes; ++i) {
        if (gf_plan_note(b, i, &plans[i]) != 0) { gf_set_error("note %d: pitch-bend range outside the bend buffer", i); return GOOFER_ERR_INVALID; }
        if (plans[i].status != GOOFER_NOTE_OK) { ++bad; if (first_bad < 0) first_bad = i; continue; }
        if (sr == 0) sr = plans[i].sr;
        if (plans[i].sr != sr) { gf_set_error("note %d: mixed sample rates in one batch (%d vs %d)", i, plans[i].sr, sr); return GOOFER_ERR_INVALID; }
    }
    if (bad) {
        gf_set_error("%d note(s) cannot be rendered; first: note %d status %d (see goofer_plan_batch)", bad, first_bad, plans[first_bad].status);
        return GOOFER_ERR_NOTE;
    }
    return GOOFER_OK;
}

static size_t gf_workspace_bytes_planned(const GooferBatch *b, const std::vector<GfNotePlan> &plans, int32_t notes_per_wave);

extern "C" size_t goofer_workspace_bytes(const GooferBatch *b, int32_t notes_per_wave)
{
    if (gf_validate(b) != GOOFER_OK) return 0;
    std::vector<GfNotePlan> plans;
    if (gf_make_plans(b, plans) != GOOFER_OK) return 0;
    return gf_workspace_bytes_planned(b, plans, notes_per_wave);
}

static size_t gf_workspace_bytes_planned(const GooferBatch *b, const std::vector<GfNotePlan> &plans, int32_t notes_per_wave)
{
    if (notes_per_wave <= 0) notes_per_wave = 2048;
    size_t best = 0;
    for (int i0 = 0; i0 < b->n_notes; i0 += notes_per_wave) {
        const int i1 = std::min(b->n_notes, i0 + notes_per_wave);
        size_t tot = 0, np = 0, ne = 0, nf = 0, nfir = 0;
        for (int i = i0; i < i1; ++i) {
            tot += gf_note_bytes(plans[i]);
            size_t a, c, d;
            gf_note_work_counts(plans[i], &a, &c, &d);
            ne += a; nf += c; nfir += d; np += plans[i].n_passes;
        }
        tot += gf_wave_meta_bytes(i1 - i0, np, ne, nf, nfir);
        best = std::max(best, tot);
    }
    return gf_sources_bytes(b) + best + 4096 + 512;          // + the status word at the head of the workspace
}

// ------------------------------------------------------------------------------------------------
// one wave
// ------------------------------------------------------------------------------------------------
struct WaveHost {
    std::vector<GfNotePlan> plans;
    std::vector<GfNoteDev> notes;
    std::vector<GfPassDev> passes;
    std::vector<int2> env_work;
    std::vector<int4> frame_work;
    std::vector<GfFirJob> fir;
    std::vector<GfOnepoleJob> op_jobs;
};

// Metadata (plans, records, work lists, job lists) is staged in a thread-local page-locked arena so that
// the uploads are truly asynchronous: a pageable source would make cudaMemcpyAsync wait for the stream.
struct GfPinned { char *base = nullptr; size_t cap = 0, off = 0; int dev = -1; };
static thread_local GfPinned g_pin;

static void *gf_pin_take(size_t bytes)
{
    bytes = (bytes + 255) & ~(size_t)255;
    int dev = -1;
    cudaGetDevice(&dev);
    if (g_pin.dev != dev) {
        // the thread moved to another GPU: what the previous device still reads from the arena must complete first
        if (g_pin.dev >= 0 && g_pin.off > 0) { cudaSetDevice(g_pin.dev); cudaDeviceSynchronize(); cudaSetDevice(dev); }
        g_pin.off = 0;
        g_pin.dev = dev;
    }
    if (g_pin.off + bytes > g_pin.cap) {
        // wrap (or grow): every copy that read the arena must have completed
        cudaDeviceSynchronize();
        if (bytes > g_pin.cap) {
            if (g_pin.base) cudaFreeHost(g_pin.base);
            g_pin.base = nullptr;
            g_pin.cap = std::max(bytes * 2, (size_t)64 << 20);
            if (cudaMallocHost((void **)&g_pin.base, g_pin.cap) != cudaSuccess) { g_pin.base = nullptr; g_pin.cap = 0; return nullptr; }
        }
        g_pin.off = 0;
    }
    void *p = g_pin.base + g_pin.off;
    g_pin.off += bytes;
    return p;
}

// The staged bytes are pulled in by a kernel that reads the (device-mapped) pinned arena directly, NOT by
// cudaMemcpyAsync: a DMA copy would queue behind whatever bulk transfer occupies the H2D copy engine (the
// host entry point streams hundreds of MB of noise phases on another stream) and stall the whole wave.
__global__ void __launch_bounds__(256) gf_meta_copy_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

static int gf_meta_copy(void *dst, const void *stage, size_t bytes, cudaStream_t st)
{
    const size_t n16 = (bytes + 15) / 16;            // both sides are 256-byte aligned and padded
    const int blocks = (int)std::min<size_t>((n16 + 255) / 256, 148);
    gf_meta_copy_kernel<<<blocks, 256, 0, st>>>((uint4 *)dst, (const uint4 *)stage, n16);
    ++g_stats.kernel_launches;
    GF_CUDA(cudaGetLastError());
    return GOOFER_OK;
}

template <typename T>
static int gf_upload(Bump &bp, const std::vector<T> &v, T **dptr, cudaStream_t st)
{
    *dptr = bp.arr<T>(std::max<size_t>(v.size(), 1));
    if (v.empty() || !bp.base) return GOOFER_OK;
    if (bp.off > bp.cap) { gf_set_error("internal: metadata overflows the workspace (%zu > %zu)", bp.off, bp.cap); return GOOFER_ERR_WORKSPACE; }
    const size_t bytes = v.size() * sizeof(T);
    void *stage = gf_pin_take(bytes);
    if (!stage) { gf_set_error("cudaMallocHost failed for the metadata staging arena"); return GOOFER_ERR_CUDA; }
    std::memcpy(stage, v.data(), bytes);
    return gf_meta_copy(*dptr, stage, bytes, st);
}

// Several host vectors -> consecutive workspace arrays with ONE staging block and ONE copy kernel (the arrays are
// carved back to back, each 256-byte aligned, and the staging block mirrors their relative offsets; the alignment
// gaps carry don't-care bytes).  Six separate uploads cost six launches of ~12 us each per wave.
struct GfUploadBatch {
    struct Item { const void *src; size_t bytes; char *dst; };
    std::vector<Item> items;
    template <typename T> void add(Bump &bp, const std::vector<T> &v, T **dptr)
    {
        *dptr = bp.arr<T>(std::max<size_t>(v.size(), 1));
        if (!v.empty() && bp.base) items.push_back({v.data(), v.size() * sizeof(T), (char *)*dptr});
    }
    int flush(const Bump &bp, cudaStream_t st)
    {
        if (items.empty()) return GOOFER_OK;
        if (bp.off > bp.cap) { gf_set_error("internal: metadata overflows the workspace (%zu > %zu)", bp.off, bp.cap); return GOOFER_ERR_WORKSPACE; }
        char *d0 = items.front().dst;
        const size_t span = (size_t)(items.back().dst - d0) + items.back().bytes;
        char *stage = (char *)gf_pin_take(span);
        if (!stage) { gf_set_error("cudaMallocHost failed for the metadata staging arena"); return GOOFER_ERR_CUDA; }
        for (const Item &it : items) std::memcpy(stage + (it.dst - d0), it.src, it.bytes);
        items.clear();
        return gf_meta_copy(d0, stage, span, st);
    }
};

int gf_post_fx(const WaveHost &wh, int n0, int n1, const GfNotePlan *d_plans, const GfNoteDev *d_notes, const GfPassDev *d_passes,
               GfPassScal *d_scal, Bump &bp, int sr, int max_n, cudaStream_t st, int64_t *launches);
int gf_growl(const WaveHost &wh, const GfNotePlan *d_plans, const GfNoteDev *d_notes, const GfPassDev *d_passes,
             GfPassScal *d_scal, Bump &bp, int max_n, cudaStream_t st, int64_t *launches);
int gf_pitch_dyn(const WaveHost &wh, int n0, int n1, const GfNotePlan *d_plans, const GfNoteDev *d_notes, Bump &bp, int sr,
                 int max_n, cudaStream_t st, int64_t *launches);

// The excitation chain (mask -> fir -> f0 -> walk -> pulse [-> growl]) and the envelope chain (tracks -> env) of a
// wave are independent until the frame kernel.  The excitation chain runs on a library-owned high-priority side
// stream, forked from and joined back into the caller's stream with events, beside the envelope kernel
// (GOOFER_OVERLAP=0: everything on the caller's stream).  The two chains mostly take turns -- the envelope kernel
// fills the register file at three CTAs per SM -- but the tails of the low-occupancy kernels (walk: two warps per
// note) no longer leave SMs idle.  Measured on B200, forked against one stream: c1 3.74 / 3.89 ms, c2 4.80 / 4.90 ms,
// c3 11.43 / 11.52 ms, c4 (96 x 16 s, long walk chains) 6.05 / 6.68 ms.  (With the first versions of the kernels the
// fork lost, 7.94 against 7.52 ms; it was re-measured after they had been tightened.)
struct GfSide {
    cudaStream_t sx = nullptr;                      // excitation chain (high priority)
    cudaStream_t alt = nullptr;                     // every other part of the frame / peak / mix tail (default priority)
    cudaEvent_t fork = nullptr, join = nullptr, prep = nullptr, alt_done = nullptr;
    int dev = -1;
};
static thread_local GfSide g_side;

static bool gf_overlap_on()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("GOOFER_OVERLAP"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1 && !gf_debug_sync() && !(g_prof.on && g_prof.serial);
}

static cudaStream_t gf_side_stream()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    if (g_side.sx && g_side.dev == dev) return g_side.sx;
    if (g_side.sx) {
        cudaStreamDestroy(g_side.sx); cudaEventDestroy(g_side.fork); cudaEventDestroy(g_side.join);
        if (g_side.alt) { cudaStreamDestroy(g_side.alt); cudaEventDestroy(g_side.prep); cudaEventDestroy(g_side.alt_done); }
        g_side = GfSide();
    }
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);            // hi = greatest priority (numerically lowest)
    const char *pe = getenv("GOOFER_SIDE_PRIORITY");          // "low": the side stream yields to the caller's stream (tuning knob)
    const int prio = (pe && pe[0] == 'l') ? lo : hi;
    if (cudaStreamCreateWithPriority(&g_side.sx, cudaStreamNonBlocking, prio) != cudaSuccess) { g_side.sx = nullptr; return nullptr; }
    if (cudaEventCreateWithFlags(&g_side.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&g_side.join, cudaEventDisableTiming) != cudaSuccess) {
        cudaStreamDestroy(g_side.sx); g_side = GfSide(); return nullptr;
    }
    g_side.dev = dev;
    return g_side.sx;
}

// Tuning knob, off by default: GOOFER_PART_STREAMS=2 alternates the parts of a wave's tail (host entry point) between the
// caller's stream and this one, so that the next part's CTAs fill the half-empty last round of a part's frame launch
// (a part of 30 % of 1,024 notes covers the GPU 1.4 times).  Measured on B200 (c2, production transfer set, same box):
// 5.44 / 5.47 ms per call against 5.25 / 5.23 on one stream -- the parts then finish together instead of one after the
// other, and the downloads that were hidden behind the later parts' kernels are not.
static cudaStream_t gf_alt_stream()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("GOOFER_PART_STREAMS"); v = (e && e[0] == '2') ? 1 : 0; }
    if (!v || !gf_overlap_on() || !gf_side_stream()) return nullptr;
    if (g_side.alt) return g_side.alt;
    if (cudaStreamCreateWithFlags(&g_side.alt, cudaStreamNonBlocking) != cudaSuccess) { g_side.alt = nullptr; return nullptr; }
    if (cudaEventCreateWithFlags(&g_side.prep, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&g_side.alt_done, cudaEventDisableTiming) != cudaSuccess) {
        cudaStreamDestroy(g_side.alt); g_side.alt = nullptr; return nullptr;
    }
    return g_side.alt;
}

// A part = a run of consecutive notes of the batch whose noise phases arrive together (host entry point):
// the frame kernel of the part waits for `phi_ready`, `done` is recorded after the part's mix kernel.
struct GfPart { int note_end; cudaEvent_t phi_ready; cudaEvent_t done; };

// Render status word at the head of the workspace: {code, first offending note}.  A pulse-onset or growl-event list that
// overflows its capacity (mean f0 above sr / 8: beyond any MIDI pitch, but a caller-supplied f0 curve can do it) is
// flagged per pass by the walk kernels (GfPassScal.err); one tiny kernel per wave folds those flags into the status
// word, which goofer_render_batch_host checks after its final synchronisation and goofer_render_status() exposes to
// callers of the asynchronous device entry point.
__global__ void __launch_bounds__(256) gf_err_scan_kernel(const GfPassScal *__restrict__ scal, const GfPassDev *__restrict__ passes, int n_pass,
                                                           int note0, int *status)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pass) return;
    const int e = scal[i].err;
    if (e) { atomicOr(&status[0], e); atomicMin(&status[1], note0 + passes[i].note); }
}

static int gf_render_wave(const GooferBatch *b, const std::vector<GfNotePlan> &all, int i0, int i1, const GfSourceDev *d_srcs,
                          Bump bp /* by value: wave region restarts every wave */, cudaStream_t st, const GfPart *parts, int n_parts,
                          int *d_status, std::function<int()> *sources_first /* run once, after the first wave's phase generator */,
                          std::function<int()> *uploads /* run once, right after the phase generator launch (host entry point) */)
{
    WaveHost wh;
    const int nn = i1 - i0;
    wh.plans.assign(all.begin() + i0, all.begin() + i1);
    wh.notes.resize(nn);
    int sr = wh.plans[0].sr;
    int max_n = 0;
    size_t n_pass = 0;
    for (int i = 0; i < nn; ++i) n_pass += wh.plans[i].n_passes;
    wh.passes.resize(n_pass);
    // frame-kernel CTAs per pass of every note: as few as GF_BLOCKS_PER_CTA allows, finer at the tail of each launch
    std::vector<int> fparts(nn);
    for (int i = 0; i < nn; ++i) {
        const int n_blocks = std::max(wh.plans[i].T_out, 2) - 2 + 1;
        fparts[i] = std::max(1, (n_blocks + GF_BLOCKS_PER_CTA - 1) / GF_BLOCKS_PER_CTA);
    }
    if (!getenv("GOOFER_NO_TAIL_BALANCE")) {
        const int slots = gf_frame_slots();
        if (n_parts > 0) {
            int prev_end = 0;
            for (int k = 0; k < n_parts; ++k) {
                const int a = std::max(prev_end, i0) - i0, e = std::min(parts[k].note_end, i1) - i0;
                prev_end = parts[k].note_end;
                if (e > a) gf_balance_frame_group(wh.plans, a, e, slots, fparts);
            }
        } else gf_balance_frame_group(wh.plans, 0, nn, slots, fparts);
    }
    size_t pi = 0;
    // per-note double scalars (maxima, rms sums, percentile): one block, one memset
    // per-note double scalars and per-pass device-written scalars are carved back to back: ONE memset zeroes both
    double *d_nscal = bp.arr<double>((size_t)nn * GF_NS_COUNT);
    GfPassScal *d_scal = bp.arr<GfPassScal>(n_pass + 1);       // one spare record: its walk_seq counts the passes the onset scan flags
    const size_t scal_span = (size_t)((char *)(d_scal + n_pass + 1) - (char *)d_nscal);
    for (int i = 0; i < nn; ++i) {
        const GfNotePlan &p = wh.plans[i];
        max_n = std::max(max_n, p.n_total);
        gf_carve_note(p, bp, &wh.notes[i], &wh.passes[pi], false);
        GfNoteDev &nd = wh.notes[i];
        nd.noteScal = d_nscal + (size_t)i * GF_NS_COUNT;
        nd.pass0 = (int)pi;
        nd.out = b->out ? b->out + p.out_off : nullptr;
        nd.pcm = b->out_pcm16 ? b->out_pcm16 + p.out_off : nullptr;
        if (b->tap_harm && b->tap_uv && b->tap_bre) {
            nd.tap_harm = b->tap_harm + p.out_off; nd.tap_uv = b->tap_uv + p.out_off; nd.tap_bre = b->tap_bre + p.out_off;
        }
        for (int k = 0; k < p.n_passes; ++k) {
            GfPassDev &q = wh.passes[pi + k];
            q.note = i;
            const int slot = q.kind;                      // phi slot == pass kind (0 main, 1 su, 2 sj, 3 sa)
            q.phi_src = ((p.phi_rng_mask >> slot) & 1u) ? nullptr : b->phi + p.phi_off[slot];
        }
        pi += p.n_passes;
    }
    int64_t &L = g_stats.kernel_launches;
    int rc;
    gf_htrace("wave: notes carved");
    // ---- noise phases drawn on the device (GooferNote.phi_rng): needs nothing but its own job list, so it is uploaded
    // and launched before the rest of the wave's metadata is even built: the generator runs (0.3 ms per 1,024 notes)
    // while the host assembles the work lists, and overlaps the arrival of the source arrays over PCIe ----
    {
        std::vector<GfPhiJob> pj;
        int max_T = 0;
        for (size_t q = 0; q < wh.passes.size(); ++q) {
            const GfPassDev &pd = wh.passes[q];
            if (pd.phi_src) continue;
            const GfNotePlan &p = wh.plans[pd.note];
            GfPhiJob j;
            j.dst = pd.phi; j.T = p.T_out; j.pad = 0;
            j.s_hi = p.phi_rng[pd.kind][0]; j.s_lo = p.phi_rng[pd.kind][1]; j.i_hi = p.phi_rng[pd.kind][2]; j.i_lo = p.phi_rng[pd.kind][3];
            max_T = std::max(max_T, j.T);
            pj.push_back(j);
        }
        if (!pj.empty()) {
            // carved from the END of the wave region so that the arrays built below keep their places
            Bump tail{bp.base + bp.cap - ((pj.size() * sizeof(GfPhiJob) + 511) & ~(size_t)255), (pj.size() * sizeof(GfPhiJob) + 511) & ~(size_t)255, 0};
            GfPhiJob *d_pj;
            if ((rc = gf_upload(tail, pj, &d_pj, st)) != GOOFER_OK) return rc;
            gf_launch_phi(d_pj, (int)pj.size(), max_T, st); ++L; GF_STEP("phi");
        }
    }
    gf_htrace("wave: phase generator launched");
    {
        size_t ne = 0, nf = 0;
        for (int i = 0; i < nn; ++i) {
            size_t a, c, d;
            gf_note_work_counts(wh.plans[i], &a, &c, &d);
            ne += a; nf += c;
        }
        wh.env_work.reserve(ne + 8); wh.frame_work.reserve(nf + 8); wh.fir.reserve((size_t)nn * 6 + 8);
    }
    if (uploads && *uploads) {
        // the host entry point's H2D copies: issued while the generator runs and before the host builds the work lists,
        // so that the sources are in HBM by the time the first kernel that needs them is enqueued
        const int rcu = (*uploads)();
        *uploads = nullptr;
        if (rcu != GOOFER_OK) return rcu;
        gf_htrace("wave: uploads issued");
    }
    pi = 0;
    for (int i = 0; i < nn; ++i) {
        const GfNotePlan &p = wh.plans[i];
        GfNoteDev &nd = wh.notes[i];
        // work lists
        const int tiles = (p.T_out + GF_FT - 1) / GF_FT;
        for (int t = 0; t * GF_ENV_TPC < tiles; ++t) wh.env_work.push_back(make_int2(i, t));      // GF_ENV_TPC tiles per CTA
        const int n_blocks = std::max(p.T_out, 2) - 2 + 1;
        const int bpc = (n_blocks + fparts[i] - 1) / fparts[i];
        for (int k = 0; k < p.n_passes; ++k)
            for (int bb = 0; bb < n_blocks; bb += bpc)
                wh.frame_work.push_back(make_int4((int)(pi + k), 2 + bb, std::min(bpc, n_blocks - bb), 0));
        // FIR jobs (Gaussian smoothing along time)
        {
            GfFirJob j;
            std::memset(&j, 0, sizeof(j));
            j.in = nd.vm4; j.in_f64 = 0; j.in_stride = 1; j.n = (p.n_total + 3) / 4; j.out = nd.ms_short; j.out_f64 = 0;    // mask[::4], compact
            j.sigma = 25.0;                                // max(1, 100 / 4)   GOOFER.py:562
            wh.fir.push_back(j);
            if (p.vol_jitter || p.sd > 0) {
                std::memset(&j, 0, sizeof(j));
                j.in = nd.vm; j.in_stride = 1; j.n = p.n_total; j.sigma = 20.0;
                if (p.vol_jitter) { j.out = nd.vjm; wh.fir.push_back(j); }          // GOOFER.py:1189
                if (p.sd > 0) { j.out = nd.sdm; wh.fir.push_back(j); }               // SillySampler.py:1109
            }
            if (p.f0_jitter) {
                std::memset(&j, 0, sizeof(j));
                j.in = b->normals + p.nrm_off[0]; j.in_f64 = 1; j.in_stride = 1; j.n = p.n_total; j.out = nd.z_sh; j.out_f64 = 1;
                j.sigma = (double)sr / (100.0 * 6);        // f0_jitter_speed = 100   GOOFER.py:667
                j.maxabs = nd.noteScal + GF_NS_SHMAX;
                wh.fir.push_back(j);
            }
            if (p.vol_jitter) {
                for (int s = 0; s < 2; ++s) {
                    std::memset(&j, 0, sizeof(j));
                    j.in = b->normals + p.nrm_off[1 + s]; j.in_f64 = 1; j.in_stride = 1; j.n = p.n_total;
                    j.out = s ? nd.z_srb : nd.z_srh; j.out_f64 = 1;
                    j.sigma = (double)sr / (150.0 * 6);    // volume_jitter_speed = 150   GOOFER.py:654
                    j.maxabs = nd.noteScal + (s ? GF_NS_SRBMAX : GF_NS_SRHMAX);
                    wh.fir.push_back(j);
                }
            }
        }
        pi += p.n_passes;
    }
    GfNotePlan *d_plans; GfNoteDev *d_notes; GfPassDev *d_passes; int2 *d_envw; int4 *d_framew; GfFirJob *d_fir;
    gf_htrace("wave: work lists built");
    {
        GfUploadBatch up;
        up.add(bp, wh.plans, &d_plans);
        up.add(bp, wh.notes, &d_notes);
        up.add(bp, wh.passes, &d_passes);
        up.add(bp, wh.env_work, &d_envw);
        up.add(bp, wh.frame_work, &d_framew);
        up.add(bp, wh.fir, &d_fir);
        if ((rc = up.flush(bp, st)) != GOOFER_OK) return rc;
    }
    if (bp.off > bp.cap) { gf_set_error("internal: wave overflows the workspace (%zu > %zu)", bp.off, bp.cap); return GOOFER_ERR_WORKSPACE; }
    GF_CUDA(cudaMemsetAsync(d_nscal, 0, scal_span, st));

    gf_htrace("wave: metadata uploaded");
    if (sources_first && *sources_first) {
        if ((rc = (*sources_first)()) != GOOFER_OK) return rc;
        *sources_first = nullptr;
    }
    // ---- excitation chain on the side stream (sx == st with GOOFER_OVERLAP=0) ----
    cudaStream_t sx = gf_overlap_on() ? gf_side_stream() : nullptr;
    const bool forked = sx != nullptr;
    if (forked) {
        GF_CUDA(cudaEventRecord(g_side.fork, st));
        GF_CUDA(cudaStreamWaitEvent(sx, g_side.fork, 0));
        gf_prof_mark("fork", sx);
    } else sx = st;
    {
        cudaStream_t st = sx;                                 // shadows: GF_STEP marks / syncs the stream the kernel went to
        gf_launch_mask(d_plans, d_notes, d_srcs, nn, max_n, st); ++L; GF_STEP("mask");
        {
            L += gf_launch_fir(wh.fir.data(), d_fir, (int)wh.fir.size(), st); GF_STEP("fir");
        }
        gf_launch_f0(d_plans, d_notes, d_passes, d_srcs, b->bend_cents, b->normals, b->f0_curves, nn, max_n, st); ++L; GF_STEP("f0");
        L += gf_launch_walk(d_passes, d_scal, (int)n_pass, max_n, sr, st, &d_scal[n_pass].walk_seq); GF_STEP("walk");     // onset scan + bit-exact walk + onset kernels
        gf_launch_pulse(d_passes, d_scal, (int)n_pass, max_n, st); ++L; GF_STEP("pulse");
        if ((rc = gf_growl(wh, d_plans, d_notes, d_passes, d_scal, bp, max_n, st, &L)) != GOOFER_OK) return rc;
    }
    // ---- host-supplied phases that are already resident (device entry point: no phi_ready events): transposed to the
    // frame-major layout here, on the caller's stream -- a DRAM-bound copy beside the issue-bound excitation chain ----
    bool phi_fm_done = false;
    {
        bool waits = false, any_src = false;
        for (int k = 0; k < n_parts; ++k) waits = waits || (parts[k].phi_ready != nullptr);
        int max_T = 0;
        for (const GfPassDev &q : wh.passes) if (q.phi_src) { any_src = true; max_T = std::max(max_T, q.T_out); }
        if (any_src && !waits) { gf_launch_phi_fm(d_passes, (int)n_pass, max_T, st); ++L; GF_STEP("phi_fm"); phi_fm_done = true; }
    }
    // ---- envelope chain on the caller's stream ----
    gf_launch_tracks(d_plans, d_notes, d_srcs, nn, st); ++L; GF_STEP("tracks");
    gf_launch_env(d_envw, (int)wh.env_work.size(), d_plans, d_notes, d_srcs, st); ++L; GF_STEP("env");
    if (forked) {
        GF_CUDA(cudaEventRecord(g_side.join, sx));
        GF_CUDA(cudaStreamWaitEvent(st, g_side.join, 0));
        gf_prof_mark("join", st);
    }
    gf_htrace("wave: preparation kernels enqueued");
    // ---- phase-dependent tail, part by part: frame -> peak -> pd / post-FX -> mix ----
    {
        std::vector<int> first_work(nn + 1, 0), first_pass(nn + 1, 0);     // per note: first frame-work item / first pass
        {
            size_t w = 0, q = 0;
            for (int i = 0; i < nn; ++i) {
                first_work[i] = (int)w; first_pass[i] = (int)q;
                while (w < wh.frame_work.size() && wh.passes[wh.frame_work[w].x].note == i) ++w;
                q += wh.plans[i].n_passes;
            }
            first_work[nn] = (int)w; first_pass[nn] = (int)q;
        }
        const GfPart whole = {i1, nullptr, nullptr};
        const GfPart *pp = n_parts > 0 ? parts : &whole;
        const int np = n_parts > 0 ? n_parts : 1;
        int prev_end = 0;
        cudaStream_t alt = np >= 2 ? gf_alt_stream() : nullptr;
        cudaStream_t const st_main = st;
        bool alt_used = false;
        if (alt) {
            GF_CUDA(cudaEventRecord(g_side.prep, st_main));
            GF_CUDA(cudaStreamWaitEvent(alt, g_side.prep, 0));
        }
        int part_no = 0;
        for (int k = 0; k < np; ++k) {
            const int a = std::max(prev_end, i0) - i0, e = std::min(pp[k].note_end, i1) - i0;     // notes [a, e) of this wave
            prev_end = pp[k].note_end;
            if (e <= a) continue;
            cudaStream_t st = (alt && (part_no++ & 1)) ? alt : st_main;                            // shadows: this part's stream
            alt_used = alt_used || (alt && st == alt);
            if (pp[k].phi_ready) GF_CUDA(cudaStreamWaitEvent(st, pp[k].phi_ready, 0));             // first kernel that reads the noise phases
            {
                // host-supplied phases, (513, T) as numpy draws them: transposed to the frame-major layout the frame kernel reads
                int max_T = 0;
                bool any = false;
                for (int q = first_pass[a]; q < first_pass[e]; ++q)
                    if (wh.passes[q].phi_src) { any = true; max_T = std::max(max_T, wh.passes[q].T_out); }
                if (any && !phi_fm_done) { gf_launch_phi_fm(d_passes + first_pass[a], first_pass[e] - first_pass[a], max_T, st); ++L; GF_STEP("phi_fm"); }
            }
            gf_launch_frame(d_framew + first_work[a], first_work[e] - first_work[a], d_passes, d_scal, d_notes, d_plans, st); ++L; GF_STEP("frame");
            // lean peak / mix instantiations for the common note, the general ones for the rest (gf_tail_simple in k_tail.cu)
            bool any_simple = false, any_general = false;
            for (int i = a; i < e; ++i) {
                const GfNotePlan &p = wh.plans[i];
                const bool simple = p.n_passes == 1 && !p.vol_jitter && !note_needs_fx(p) && p.pd == 0.0 && !wh.notes[i].tap_harm;
                (simple ? any_simple : any_general) = true;
            }
            gf_launch_peak(d_plans, d_notes, d_passes, d_scal, first_pass[a], first_pass[e] - first_pass[a], max_n, any_simple, any_general, st);
            L += (int)any_simple + (int)any_general; GF_STEP("peak");
            if ((rc = gf_pitch_dyn(wh, a, e, d_plans, d_notes, bp, sr, max_n, st, &L)) != GOOFER_OK) return rc;
            if ((rc = gf_post_fx(wh, a, e, d_plans, d_notes, d_passes, d_scal, bp, sr, max_n, st, &L)) != GOOFER_OK) return rc;
            gf_launch_mix(d_plans, d_notes, d_passes, d_scal, a, e - a, max_n, any_simple, any_general, st);
            L += (int)any_simple + (int)any_general; GF_STEP("mix");
            if (pp[k].done && pp[k].note_end <= i1) GF_CUDA(cudaEventRecord(pp[k].done, st));
        }
        if (alt_used) {
            GF_CUDA(cudaEventRecord(g_side.alt_done, alt));
            GF_CUDA(cudaStreamWaitEvent(st_main, g_side.alt_done, 0));
        }
    }
    gf_htrace("wave: everything enqueued");
    gf_err_scan_kernel<<<(int)((n_pass + 255) / 256), 256, 0, st>>>(d_scal, d_passes, (int)n_pass, i0, d_status); ++L;
    GF_CUDA(cudaGetLastError());
    ++g_stats.waves;
    return GOOFER_OK;
}

// A sub-batch = the notes [n0, n1) of the batch rendered by one call (n1 < 0: all of them).  The host entry point renders a
// large batch as two or more sub-batches back to back on one stream, so that the first one's results travel while the
// next one computes; `first` = this call decodes the sources and resets the status word (later sub-batches of the same
// batch find both in the workspace, whose head is laid out identically every time).
struct GfSubBatch { int n0 = 0, n1 = -1; bool first = true; };

static int gf_render_batch_ex(const GooferBatch *b, void *workspace, size_t workspace_bytes, void *stream, const GfPart *parts, int n_parts,
                              const std::vector<GfNotePlan> *planned = nullptr, cudaEvent_t src_ready = nullptr,
                              std::function<int()> *uploads = nullptr, GfSubBatch sb = GfSubBatch());

extern "C" int goofer_render_batch(const GooferBatch *b, void *workspace, size_t workspace_bytes, void *stream)
{
    return gf_render_batch_ex(b, workspace, workspace_bytes, stream, nullptr, 0);
}

// parts (optional): see GfPart; only the frame kernels wait for the noise phases.  planned (optional): the plans of
// b's notes when the caller made them already (the host entry point plans once for sizing and rendering: 0.17 ms per
// 1,024 notes each time).  src_ready (optional): event after which the source arrays / bends / normals are in device
// memory -- the first kernels that read them wait for it, the noise-phase generator of the first wave runs before.
// uploads (optional): issues the copies that src_ready (and the parts' phi_ready events) stand for; called once, right
// after the first wave's phase generator has been launched, and reset to empty.
static int gf_render_batch_ex(const GooferBatch *b, void *workspace, size_t workspace_bytes, void *stream, const GfPart *parts, int n_parts,
                              const std::vector<GfNotePlan> *planned, cudaEvent_t src_ready, std::function<int()> *uploads, GfSubBatch sb)
{
    int rc = gf_validate(b);
    if (rc != GOOFER_OK) return rc;
    if (sb.first) { g_stats.kernel_launches = 0; g_stats.waves = 0; }
    if (b->n_notes == 0) return GOOFER_OK;
    const int note_lo = std::max(0, sb.n0), note_hi = sb.n1 < 0 ? b->n_notes : std::min(sb.n1, b->n_notes);
    if (note_hi <= note_lo) return GOOFER_OK;
    if (!workspace || (!b->out && !b->out_pcm16) || !b->bend_cents) { gf_set_error("NULL workspace / out (and out_pcm16) / bend_cents"); return GOOFER_ERR_INVALID; }
    std::vector<GfNotePlan> own_plans;
    if (!planned && (rc = gf_make_plans(b, own_plans)) != GOOFER_OK) return rc;
    const std::vector<GfNotePlan> &plans = planned ? *planned : own_plans;
    for (int i = note_lo; i < note_hi; ++i) {
        const GfNotePlan &p = plans[i];
        for (int k = 0; k < p.n_passes; ++k) {
            const int slot = p.pass_kind[k];
            if ((p.phi_rng_mask >> slot) & 1u) continue;                      // drawn on the device
            if (!b->phi || p.phi_off[slot] < 0 || p.phi_off[slot] + (int64_t)GF_NBINS * p.T_out > b->phi_total) {
                gf_set_error("note %d: phi slot %d outside the phi buffer", i, slot);
                return GOOFER_ERR_INVALID;
            }
        }
        const int need[4] = {p.f0_jitter, p.vol_jitter, p.vol_jitter, p.sj > 0.0};
        for (int k = 0; k < 4; ++k)
            if (need[k] && (!b->normals || p.nrm_off[k] < 0 || p.nrm_off[k] + p.n_total > b->nrm_total)) {
                gf_set_error("note %d: normal slot %d missing or outside the normals buffer", i, k);
                return GOOFER_ERR_INVALID;
            }
        if (p.out_off < 0 || p.out_off + p.n_total > b->out_total) { gf_set_error("note %d: output range outside the out buffer", i); return GOOFER_ERR_INVALID; }
        if (p.f0_off >= 0 && (!b->f0_curves || p.f0_off + p.n_total > b->f0_total)) { gf_set_error("note %d: f0 curve missing or outside the f0_curves buffer", i); return GOOFER_ERR_INVALID; }
    }
    cudaStream_t st = (cudaStream_t)stream;
    gf_htrace("render: validated");
    if ((rc = gf_tables_init(plans[0].sr)) != 0) return rc;
    gf_prof_mark("begin", st);

    Bump bp{(char *)workspace, workspace_bytes, 0};
    int *d_status = bp.arr<int>(64);                          // render status word (gf_err_scan_kernel): first thing in the workspace
    if (workspace_bytes < 512) { gf_set_error("workspace too small"); return GOOFER_ERR_WORKSPACE; }
    if (sb.first) {
        static const int init[2] = {0, 0x7fffffff};
        int *stage = (int *)gf_pin_take(sizeof(init));
        if (!stage) { gf_set_error("cudaMallocHost failed for the metadata staging arena"); return GOOFER_ERR_CUDA; }
        std::memcpy(stage, init, sizeof(init));
        if ((rc = gf_meta_copy(d_status, stage, sizeof(init), st)) != GOOFER_OK) return rc;
    }
    // ---- sources: decode / transpose once per call ----
    std::vector<GfSourceDev> srcs(b->n_sources);
    int max_T = 0;
    for (int s = 0; s < b->n_sources; ++s) {
        const GooferSource &g = b->sources[s];
        GfSourceDev d;
        std::memset(&d, 0, sizeof(d));
        d.knots = g.knots_log_f16; d.hz_knots = g.hz_knots; d.dense = g.env_dense; d.mask = g.mask;
        d.K = g.K; d.T = g.T; d.N = g.N;
        for (int k = 0; k < 4; ++k) { d.formants[k] = g.formants[k]; d.formant_len[k] = g.formants[k] ? g.formant_len[k] : 0; }
        if (g.T > 0) {
            if (!d.knots && !d.dense) { gf_set_error("source %d has neither knots nor a dense envelope", s); return GOOFER_ERR_INVALID; }
            if (d.knots && (g.K < 2 || g.K > 4096 || !g.hz_knots)) { gf_set_error("source %d: knots need 2 <= K <= 4096 and hz_knots", s); return GOOFER_ERR_INVALID; }
            d.envS = bp.arr<float>((size_t)g.T * GF_ENVS_LD);
            max_T = std::max(max_T, g.T);
        }
        if (g.N < 0 || (g.N > 0 && !g.mask)) { gf_set_error("source %d: NULL voicing mask (N = %d)", s, g.N); return GOOFER_ERR_INVALID; }
        for (int k = 0; k < 4; ++k)
            if (g.formants[k] && g.formant_len[k] <= 0) { gf_set_error("source %d: formant track %d has length %d", s, k + 1, g.formant_len[k]); return GOOFER_ERR_INVALID; }
        srcs[s] = d;
    }
    for (int i = note_lo; i < note_hi; ++i) {
        const GooferSource &g = b->sources[plans[i].src];
        if (g.T <= 0 || g.N <= 0) { gf_set_error("note %d: source %d has no frames / samples", i, plans[i].src); return GOOFER_ERR_INVALID; }
    }
    GfSourceDev *d_srcs = bp.arr<GfSourceDev>(std::max(1, b->n_sources));
    if (bp.off > bp.cap) { gf_set_error("workspace too small for the source cache (%zu > %zu)", bp.off, bp.cap); return GOOFER_ERR_WORKSPACE; }
    // source table + envelope decode: issued inside the first wave, after its phase generator (gf_render_wave)
    std::function<int()> sources_first = [&]() -> int {
        if (!sb.first) return GOOFER_OK;                      // decoded by the batch's first sub-batch, same stream
        if (src_ready) GF_CUDA(cudaStreamWaitEvent(st, src_ready, 0));
        if (b->n_sources) {
            void *stage = gf_pin_take(srcs.size() * sizeof(GfSourceDev));
            if (!stage) { gf_set_error("cudaMallocHost failed for the metadata staging arena"); return GOOFER_ERR_CUDA; }
            std::memcpy(stage, srcs.data(), srcs.size() * sizeof(GfSourceDev));
            const int rcs = gf_meta_copy(d_srcs, stage, srcs.size() * sizeof(GfSourceDev), st);
            if (rcs != GOOFER_OK) return rcs;
        }
        gf_launch_src_env(d_srcs, b->n_sources, max_T, st); ++g_stats.kernel_launches; GF_STEP("src_env");
        return GOOFER_OK;
    };

    // ---- waves: greedy packing into what is left of the workspace ----
    bp.off = (bp.off + 255) & ~(size_t)255;
    if (bp.off >= workspace_bytes) { gf_set_error("workspace too small for the source cache"); return GOOFER_ERR_WORKSPACE; }
    const size_t wave_cap = workspace_bytes - bp.off;
    int i0 = note_lo;
    while (i0 < note_hi) {
        size_t tot = 0, np = 0, ne = 0, nf = 0, nfir = 0;
        int i1 = i0;
        while (i1 < note_hi) {
            size_t a, c, d;
            gf_note_work_counts(plans[i1], &a, &c, &d);
            const size_t nb = gf_note_bytes(plans[i1]);
            const size_t meta = gf_wave_meta_bytes(i1 - i0 + 1, np + plans[i1].n_passes, ne + a, nf + c, nfir + d);
            if (tot + nb + meta + 4096 > wave_cap) break;
            tot += nb; np += plans[i1].n_passes; ne += a; nf += c; nfir += d;
            ++i1;
        }
        if (i1 == i0) {
            gf_set_error("workspace too small: note %d alone needs %zu bytes, %zu available", i0, gf_note_bytes(plans[i0]), wave_cap);
            return GOOFER_ERR_WORKSPACE;
        }
        gf_htrace("render: wave packed");
        Bump wave{(char *)workspace + bp.off, wave_cap, 0};
        if ((rc = gf_render_wave(b, plans, i0, i1, d_srcs, wave, st, parts, n_parts, d_status, &sources_first, uploads)) != GOOFER_OK) return rc;
        // host-side work lists are reused by the next wave only after this one was enqueued; the
        // device regions are reused in stream order, so no extra synchronisation is needed
        i0 = i1;
    }
    return GOOFER_OK;
}

extern "C" int goofer_render_status(const void *workspace, void *stream, int32_t *first_note)
{
    if (!workspace) { gf_set_error("goofer_render_status: NULL workspace"); return GOOFER_ERR_INVALID; }
    int st[2] = {0, 0};
    workspace = (const void *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);      // where the bump allocator put the status word
    GF_CUDA(cudaMemcpyAsync(st, workspace, sizeof(st), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    GF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (first_note) *first_note = st[0] ? st[1] : -1;
    if (st[0]) {
        gf_set_error("note %d: %s list overflowed (mean f0 above sr / 8); the pulse train of that note is truncated", st[1],
                     (st[0] & 1) ? "pulse-onset" : "growl-event");
        return GOOFER_ERR_NOTE;
    }
    return GOOFER_OK;
}

// ------------------------------------------------------------------------------------------------
// stage-level entry points
// ------------------------------------------------------------------------------------------------
extern "C" int goofer_stft_batch(const float *x, int32_t n_sig, int32_t n, float *S_out, void *stream)
{
    if (!x || !S_out || n_sig < 0 || n < 2) { gf_set_error("goofer_stft_batch: invalid arguments"); return GOOFER_ERR_INVALID; }
    int rc = gf_tables_init(0);
    if (rc) return rc;
    if (n_sig == 0) return GOOFER_OK;
    gf_launch_stft(x, n_sig, n, (float2 *)S_out, (cudaStream_t)stream);
    GF_CUDA(cudaGetLastError());
    return GOOFER_OK;
}

extern "C" int goofer_istft_batch(const float *S, int32_t n_sig, int32_t T, int32_t length, float *y_out, void *stream)
{
    if (!S || !y_out || n_sig < 0 || T < 1 || length < 0) { gf_set_error("goofer_istft_batch: invalid arguments"); return GOOFER_ERR_INVALID; }
    int rc = gf_tables_init(0);
    if (rc) return rc;
    if (n_sig == 0 || length == 0) return GOOFER_OK;
    if (length < GF_HOP * (T - 1)) { gf_set_error("goofer_istft_batch: length %d shorter than hop * (T - 1) = %d is not supported", length, GF_HOP * (T - 1)); return GOOFER_ERR_INVALID; }
    gf_launch_istft((const float2 *)S, n_sig, T, length, y_out, (cudaStream_t)stream);
    GF_CUDA(cudaGetLastError());
    return GOOFER_OK;
}

extern "C" size_t goofer_pulse_work_bytes(int32_t n_sig, int32_t n)
{
    if (n_sig < 0 || n < 0) return 0;
    const size_t cap = (size_t)n / 8 + 64;
    return (size_t)n_sig * (sizeof(GfPassDev) + sizeof(GfPassScal) + cap * sizeof(int4) + 512) + sizeof(GfPassScal) + 1024;
}

extern "C" int goofer_pulse_train_batch(const float *f0, int32_t n_sig, int32_t n, int32_t sr, float *pulse_out, void *work, void *stream)
{
    if (!f0 || !pulse_out || !work || n_sig < 0 || n < 1 || sr <= 0) { gf_set_error("goofer_pulse_train_batch: invalid arguments"); return GOOFER_ERR_INVALID; }
    if (n_sig == 0) return GOOFER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    Bump bp{(char *)work, goofer_pulse_work_bytes(n_sig, n), 0};
    std::vector<GfPassDev> ps(n_sig);
    const int cap = n / 8 + 64;
    for (int s = 0; s < n_sig; ++s) {
        GfPassDev q;
        std::memset(&q, 0, sizeof(q));
        q.n_total = n; q.T_out = 1 + n / GF_HOP;
        q.f0 = const_cast<float *>(f0) + (size_t)s * n;
        q.pulse = pulse_out + (size_t)s * n;
        q.onset_cap = cap;
        q.onsets = bp.arr<int4>(cap);
        ps[s] = q;
    }
    GfPassDev *d_ps = bp.arr<GfPassDev>(n_sig);
    GfPassScal *d_sc = bp.arr<GfPassScal>(n_sig + 1);
    GF_CUDA(cudaMemcpyAsync(d_ps, ps.data(), ps.size() * sizeof(GfPassDev), cudaMemcpyHostToDevice, st));
    GF_CUDA(cudaMemsetAsync(d_sc, 0, (n_sig + 1) * sizeof(GfPassScal), st));
    gf_launch_walk(d_ps, d_sc, n_sig, n, sr, st, &d_sc[n_sig].walk_seq);
    gf_launch_pulse(d_ps, d_sc, n_sig, n, st);
    GF_CUDA(cudaGetLastError());
    std::vector<GfPassScal> sc(n_sig);
    GF_CUDA(cudaMemcpyAsync(sc.data(), d_sc, n_sig * sizeof(GfPassScal), cudaMemcpyDeviceToHost, st));
    GF_CUDA(cudaStreamSynchronize(st));
    for (int s = 0; s < n_sig; ++s)
        if (sc[s].err) { gf_set_error("goofer_pulse_train_batch: signal %d has more than n/8+64 pulse onsets (mean f0 above sr/8)", s); return GOOFER_ERR_INVALID; }
    return GOOFER_OK;
}

extern "C" int goofer_onepole_batch(const float *x, const float *f0, int32_t n_sig, int32_t n, int32_t sr, double cutoff_factor,
                                    int32_t order, int32_t btype, float *y_out, void *stream)
{
    if (!x || !f0 || !y_out || n_sig < 0 || n < 0 || sr <= 0) { gf_set_error("goofer_onepole_batch: invalid arguments"); return GOOFER_ERR_INVALID; }
    if (n_sig == 0 || n == 0) return GOOFER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    float *alpha = nullptr;
    GfOnepoleJob *d_jobs = nullptr;
    GF_CUDA(cudaMallocAsync(&alpha, (size_t)n_sig * n * sizeof(float), st));
    GF_CUDA(cudaMallocAsync(&d_jobs, (size_t)n_sig * sizeof(GfOnepoleJob), st));
    std::vector<GfOnepoleJob> jobs(n_sig);
    for (int s = 0; s < n_sig; ++s) {
        GfOnepoleJob j;
        std::memset(&j, 0, sizeof(j));
        j.x = x + (size_t)s * n; j.f0 = f0 + (size_t)s * n; j.y = y_out + (size_t)s * n; j.alpha = alpha + (size_t)s * n;
        j.n = n; j.order = order; j.highpass = btype != 0; j.smooth_f0 = 1; j.cutoff_factor = cutoff_factor; j.sr = sr;
        jobs[s] = j;
    }
    GF_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(GfOnepoleJob), cudaMemcpyHostToDevice, st));
    gf_launch_onepole(d_jobs, n_sig, st);
    GF_CUDA(cudaGetLastError());
    GF_CUDA(cudaFreeAsync(alpha, st));
    GF_CUDA(cudaFreeAsync(d_jobs, st));
    return GOOFER_OK;
}
