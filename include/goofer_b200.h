/*
 * goofer_b200 -- C ABI of the B200-native (sm_100a) GOOFER / SillySampler render path.
 *
 * Drop-in boundary: the reference has no FFI of its own -- its seams are the
 * Python calls below.  A maintainer binds this library with ctypes (INTEGRATION.md shows the stub):
 *
 *   goofer_plan_batch / goofer_render_batch[_host]
 *        replace the numeric body of SillySampler.GooferResampler.resample()
 *        (/root/reference/SillySampler.py:449-1185) incl. every gf.synthesize() call it makes
 *        (SillySampler.py:1006,1041,1067,1156 -> GOOFER.py:971-1220) and
 *        gf.decode_env_from_knots (GOOFER.py:149-168).  The 13 CLI strings are parsed by the
 *        Python host (SillySampler.py:286-410 stays Python) into GooferNote records.
 *   goofer_stft_batch / goofer_istft_batch     replace GOOFER.py:355-370 / :392-413 (+ :372-390)
 *   goofer_pulse_train_batch                   replaces GOOFER.py:473-554
 *   goofer_onepole_batch                       replaces SillySampler.py:95-174
 *   goofer_analyse_batch                       envelope half of gf.extract_features + gf.compress_env_to_knots
 *                                              (GOOFER.py:940-946, 968, 97-147) -- the step before the render path
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * GooferStatus; nothing throws; the library allocates no device memory (the caller passes a
 * workspace); there is no hidden global state besides constant tables per device, so calls are
 * re-entrant per (device, stream).  goofer_last_error() is thread-local.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with
 * GOOFER_ERR_CUDA.
 */
#ifndef GOOFER_B200_H
#define GOOFER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOOFER_ABI_VERSION 5
#define GOOFER_N_FFT 1024      /* SillySampler.py:14 */
#define GOOFER_HOP 256         /* SillySampler.py:15 */
#define GOOFER_N_BINS 513

typedef enum GooferStatus {
    GOOFER_OK = 0,
    GOOFER_ERR_INVALID = -1,      /* bad descriptor (NULL pointer, negative size, unknown enum) */
    GOOFER_ERR_WORKSPACE = -2,    /* workspace too small for even one note */
    GOOFER_ERR_CUDA = -3,         /* launch / runtime failure, or no CUDA device */
    GOOFER_ERR_NOTE = -4          /* at least one note failed planning: see GooferNotePlanInfo.status */
} GooferStatus;

/* per-note planning status (mirrors the exceptions the reference raises for the same input) */
enum {
    GOOFER_NOTE_OK = 0,
    GOOFER_NOTE_EMPTY_TAIL = 1,   /* ZeroDivisionError at SillySampler.py:634 / :704 */
    GOOFER_NOTE_TOO_SHORT = 2,    /* fewer than 2 output samples */
    GOOFER_NOTE_BAD_SOURCE = 3,
    GOOFER_NOTE_EDITOR = 4        /* SE1: Tk voicing editor is out of scope */
};

/* Flag slots of GooferNote.flag[] -- the 34-flag surface of README.md:8-41 / SillySampler.py:309-410.
 * A slot holds the integer written after the flag letters; GooferNote.present has bit i set when the
 * flag appeared in the flag string. */
enum {
    GF_t = 0, GF_g, GF_fa, GF_fb, GF_fc, GF_fd, GF_fw, GF_fst, GF_fsta, GF_fstb, GF_fstc, GF_fstd,
    GF_V, GF_B, GF_U, GF_sh, GF_sr, GF_st, GF_sg, GF_sd, GF_sj, GF_sa, GF_su, GF_br, GF_es, GF_pd,
    GF_FV, GF_L, GF_R, GF_P, GF_vf, GF_vh, GF_vl, GF_SE, GF_NFLAGS
};

/* One voicebank sample's cached features = what gf.load_features returns (GOOFER.py:319-339).
 * Pointers are DEVICE pointers for goofer_render_batch and HOST pointers for *_host. */
typedef struct GooferSource {
    const uint16_t *knots_log_f16;  /* (K, T) row-major IEEE half: 'knot_vals_log' (mode 'knots'), or NULL */
    const float *hz_knots;          /* (K,) f32 */
    int32_t K;
    const float *env_dense;         /* (513, T) row-major f32 (mode 'full', already to_compute'd), or NULL */
    int32_t T;                      /* envelope frames */
    const float *mask;              /* (N,) f32 voicing mask (fp16-quantised values) */
    int32_t N;                      /* len(voicing_mask) == len(f0_interp) */
    const double *formants[4];      /* F1..F4 tracks, fp64 (formants_to_int_keys, GOOFER.py:48-62) */
    int32_t formant_len[4];
    int32_t sr;                     /* 'sr' */
    int64_t ylen;                   /* 'y_len' */
} GooferSource;

/* One note = the 13 CLI arguments after host-side parsing (SillySampler.py:286-306). */
typedef struct GooferNote {
    int32_t source;                 /* index into the GooferSource array (in.wav -> features) */
    int32_t pitch_midi;             /* note_to_midi(pitch)            SillySampler.py:86-90 */
    double velocity;                /* float(velocity)                :297 */
    double offset_s, length_s, consonant_s, cutoff_s;   /* ms / 1000  :299-302 */
    double volume;                  /* float(volume) / 100            :303 */
    double tempo;                   /* float(tempo.lstrip('!'))       :305 */
    int64_t bend_off;               /* first element of this note's pitch bend in GooferBatch.bend_cents */
    int32_t bend_len;               /* >= 1 (pitch_string_to_cents, :72-84) */
    int32_t flag[GF_NFLAGS];
    uint64_t present;
    /* noise buffers (element offsets into GooferBatch.phi / .normals; -1 = not supplied).
     * phi slots: 0 main, 1 su pass, 2 sj pass, 3 sa pass -- each (513, T_out) row-major f32, uniform
     * [0, 2pi), exactly what rng.uniform(0, 2pi, (n_bins, T)).astype(f32) yields (GOOFER.py:1151-1152).
     * normal slots: 0 sh randn(N) (GOOFER.py:666), 1 sr harm, 2 sr breath (GOOFER.py:653),
     * 3 sj standard normal (SillySampler.py:1064 draws normal(0, sj^2) = sj^2 * z) -- fp64, (n_total,). */
    int64_t phi_off[4];
    int64_t nrm_off[4];
    int64_t out_off;                /* first element of this note's output in GooferBatch.out */
    /* Direct gf.synthesize call (GOOFER.py:971-1220 as SillyEditor.py:227, 559 and test.py:38 use it) instead of a
     * resampler note: f0_off >= 0 is the first element of this note's per-sample f0 curve (f32, `f0_interp`) in
     * GooferBatch.f0_curves.  The source is then used WHOLE -- env_spec = all T frames, voicing_mask = all N samples,
     * len(y) = N -- and offset / length / consonant / cutoff / velocity / pitch / pitch bend are ignored; flags keep
     * their meaning as synthesize keyword arguments (g formant_shift, fa-fd F1-F4_shift, sh, sr, sg, P normalize ...).
     * -1 = a resampler note. */
    int64_t f0_off;
    /* Noise phases generated on the device instead of uploaded (355 KB per note and pass): bit k of phi_rng_mask says
     * that phi slot k is the stream numpy's Generator(PCG64) would draw -- rng.uniform(0, 2 pi, (513, T_out))
     * .astype(float32), GOOFER.py:1151-1152 -- from the bit generator whose 128-bit state and increment are
     * phi_rng[k] = {state_hi, state_lo, inc_hi, inc_lo} (np.random.PCG64(seed).state["state"]).  The kernel jumps the
     * LCG ahead per thread and reproduces numpy's values bit for bit; phi_off[k] is then ignored (GooferBatch.phi may
     * be NULL when every slot in use is generated).  The reference draws this noise itself (unseeded), so this is what
     * a drop-in does in production; host-supplied buffers remain the parity-test path. */
    uint64_t phi_rng[4][4];
    uint32_t phi_rng_mask;
    /* Continuous keyword arguments of gf.synthesize (GOOFER.py:971-983) that an integer flag cannot express -- the
     * direct callers pass floats (test.py:38 formant_shift, breath_strength, uv_strength; SillyEditor.py:227, 559).
     * Bit k of override_mask: override_val[k] REPLACES the scalar the flags would give (GF_OVR_* below); the flag
     * columns keep deciding which stages run (sh > 0 turns f0 jitter on, its strength may then be overridden). */
    uint32_t override_mask;
    double override_val[12];
} GooferNote;

enum {
    GF_OVR_FORMANT_SHIFT = 0,       /* formant_shift (g: 1 + g/200) */
    GF_OVR_F1_SHIFT, GF_OVR_F2_SHIFT, GF_OVR_F3_SHIFT, GF_OVR_F4_SHIFT,     /* F1..F4_shift (fa..fd: 1 + x/100) */
    GF_OVR_F0_JITTER_STRENGTH,      /* f0_jitter_strength (sh/50) */
    GF_OVR_VOL_JITTER_HARM,         /* volume_jitter_strength_harm (sr/50) */
    GF_OVR_VOL_JITTER_BREATH,       /* volume_jitter_strength_breath (2 * sr/50) */
    GF_OVR_NORMALIZE,               /* normalize (P/100) */
    GF_OVR_BREATH_STRENGTH,         /* breath_strength (0.1) */
    GF_OVR_UV_STRENGTH,             /* uv_strength (0.75) */
    GF_N_OVERRIDES
};

/* What the planner derives for one note (lengths the host needs to size noise and output buffers). */
typedef struct GooferNotePlanInfo {
    int32_t status;                 /* GOOFER_NOTE_* */
    int32_t n_total;                /* output samples = len(f0_new)   SillySampler.py:836 */
    int32_t t_out;                  /* STFT frames = 1 + n_total // 256 */
    int32_t t_env;                  /* env_new.shape[1]               SillySampler.py:801 */
    int32_t n_passes;               /* gf.synthesize calls: 1 + [su>0] + [sj>0] + [sa>0] */
    int32_t need_phi[4];            /* which phi slots the note consumes */
    int32_t need_nrm[4];            /* which normal slots the note consumes */
} GooferNotePlanInfo;

typedef struct GooferBatch {
    int32_t n_sources;
    const GooferSource *sources;    /* HOST array of descriptors (pointers inside: device or host, see above) */
    int32_t n_notes;
    const GooferNote *notes;        /* HOST array */
    const float *bend_cents;        /* concatenated pitch bends, f32 cents (device / host like the rest) */
    int64_t bend_total;
    const float *phi;               /* concatenated noise phases */
    int64_t phi_total;
    const double *normals;          /* concatenated fp64 normals */
    int64_t nrm_total;
    float *out;                     /* concatenated outputs, f32 */
    int64_t out_total;
    /* optional stage taps of the MAIN synth pass for parity tests (device/host like out; may be NULL):
     * each (out_total,) f32 laid out like out: harmonic, aper_uv, aper_bre after normalisation
     * (the tuple gf.synthesize returns, GOOFER.py:1220) */
    float *tap_harm, *tap_uv, *tap_bre;
    /* optional 16-bit PCM copy of `out`, (out_total,) int16, laid out like out (device/host like out; may be NULL).
     * The step after the path: SillySampler.py:1185 `sf.write(out_file, out, sr)` stores PCM_16 for .wav.  Encoded on
     * the device the way python-soundfile drives libsndfile (SFC_SET_CLIPPING on, pcm.c d2s_clip_array):
     * saturate(lrint(x * 2^31)) >> 16.  When out_pcm16 is given, `out` may be NULL (the host entry point then
     * downloads half the bytes). */
    int16_t *out_pcm16;
    const float *f0_curves;         /* concatenated per-sample f0 curves of the notes with f0_off >= 0 (may be NULL) */
    int64_t f0_total;
} GooferBatch;

int goofer_version(void);
const char *goofer_last_error(void);
/* sizeof() of the ABI records as this library was compiled, for binding self-checks:
 * 0 GooferSource, 1 GooferNote, 2 GooferNotePlanInfo, 3 GooferBatch, 4 GooferStats (0 for anything else) */
size_t goofer_struct_size(int which);

/* Planning only (CPU, no CUDA): fills info[n_notes].  Mirrors the integer/length bookkeeping of
 * SillySampler.py:453-500, 625-635, 766-788. Source pointers are not dereferenced. */
int goofer_plan_batch(const GooferBatch *b, GooferNotePlanInfo *info);

/* Test hook: copies the library's internal POD plan of note `idx` (struct GfNotePlan, csrc/gf_plan.h)
 * into `out` (at most `bytes`); returns sizeof(GfNotePlan) or a negative GooferStatus.  CPU only. */
int goofer_debug_plan(const GooferBatch *b, int32_t idx, void *out, size_t bytes);

/* Workspace needed to render `b` `notes_per_wave` notes at a time (0 = library default). */
size_t goofer_workspace_bytes(const GooferBatch *b, int32_t notes_per_wave);

/* Render with every array already resident on the current CUDA device. `stream` is a cudaStream_t.  The call only
 * enqueues work; results are complete when `stream` is. */
int goofer_render_batch(const GooferBatch *b, void *workspace, size_t workspace_bytes, void *stream);

/* Outcome of the render(s) enqueued on `stream` with `workspace`: waits for the stream, then returns GOOFER_OK, or
 * GOOFER_ERR_NOTE when a note's pulse-onset / growl-event list overflowed its capacity of n/8+64 (n/4+64) entries --
 * mean f0 above sr/8, beyond every MIDI pitch but reachable with a caller-supplied f0 curve (GooferNote.f0_off); the
 * reference would render such a note (GOOFER.py:473-554), this library truncates its pulse train and says so here.
 * *first_note (may be NULL) receives the lowest offending note index or -1.  goofer_render_batch_host performs this
 * check itself. */
int goofer_render_status(const void *workspace, void *stream, int32_t *first_note);

/* Same call with HOST buffers: copies sources/noise in and results out (pinned staging, chunked and
 * overlapped with compute); allocates and caches its own device buffers per thread. */
int goofer_render_batch_host(const GooferBatch *b);
/* Frees what goofer_render_batch_host cached for the calling thread (device buffers, streams, pinned staging).  The
 * caches are per (thread, device) -- a thread that moves to another GPU gets a fresh one -- and are not freed at
 * thread exit: a host thread that used the library calls this before it ends (one thread per GPU is the intended use). */
void goofer_host_release(void);

/* counters of the last render call on this thread */
typedef struct GooferStats {
    int64_t kernel_launches;        /* CUDA kernels launched by the library */
    int64_t h2d_bytes, d2h_bytes;   /* bytes moved by goofer_render_batch_host */
    int32_t waves;
} GooferStats;
void goofer_last_stats(GooferStats *s);

/* Per-kernel device timing for bench.py's roofline line: after goofer_profile(1) every kernel launch of
 * goofer_render_batch on this thread is followed by a CUDA event on the launching stream;
 * goofer_profile_summary() waits for the last one and returns "kernel:launches:total_ms;..." accumulated
 * since the enable call (thread-local storage, valid until the next call).  The excitation chain normally runs on a
 * side stream beside the envelope kernel, so the spans of those kernels overlap in time (the frame kernel and what
 * follows it run alone); goofer_profile(2) keeps everything on the caller's stream while it is on, for per-kernel
 * durations that add up.  GOOFER_OVERLAP=0 in the environment does the same for the whole process. */
void goofer_profile(int enable);
const char *goofer_profile_summary(void);

/* ---- stage-level entry points (device pointers; used by the Python mirror of gf.* and by tests) ---- */

/* gf.stft (GOOFER.py:355-370): x (n_sig, n) f32 rows -> S (n_sig, 513, T) complex64 interleaved,
 * T = 1 + n // 256, sqrt-Hann window, reflect padding. */
int goofer_stft_batch(const float *x, int32_t n_sig, int32_t n, float *S_out, void *stream);
/* gf.istft (GOOFER.py:392-413): S (n_sig, 513, T) complex64 -> y (n_sig, length) f32. */
int goofer_istft_batch(const float *S, int32_t n_sig, int32_t T, int32_t length, float *y_out, void *stream);
/* gf.pulse_train_numba (GOOFER.py:473-554, Ra .02 Rg 1.7 Rk .8): f0 (n_sig, n) f32 -> pulse (n_sig, n) f32.
 * work: >= goofer_pulse_work_bytes(n_sig, n) bytes of device scratch. */
size_t goofer_pulse_work_bytes(int32_t n_sig, int32_t n);
int goofer_pulse_train_batch(const float *f0, int32_t n_sig, int32_t n, int32_t sr, float *pulse_out,
                             void *work, void *stream);
/* dynamic_butter_filter (SillySampler.py:95-174): x, f0 (n_sig, n) f32; btype 0 lowpass / 1 highpass. */
int goofer_onepole_batch(const float *x, const float *f0, int32_t n_sig, int32_t n, int32_t sr,
                         double cutoff_factor, int32_t order, int32_t btype, float *y_out, void *stream);

/* ---- analysis front-end: the spectral-envelope half of gf.extract_features (GOOFER.py:940-946, 968) and
 * compress_env_to_knots (GOOFER.py:97-147).  y (n_sig, n) f32 waveforms -> per signal the arrays gf.save_features
 * stores: K_out[s] in {32, 48, .. 192}, hz_out (n_sig, 192) f32 (first K valid), knots_out (n_sig, 192, T) IEEE half
 * log-envelope knots (first K rows valid), T = 1 + n // 256.  f0 / voicing / formants come from Praat in the
 * reference and stay with the caller.  Device pointers; work: >= goofer_analyse_work_bytes(n_sig, n) bytes. */
size_t goofer_analyse_work_bytes(int32_t n_sig, int32_t n);
int goofer_analyse_batch(const float *y, int32_t n_sig, int32_t n, int32_t sr, uint16_t *knots_out, float *hz_out,
                         int32_t *K_out, void *work, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GOOFER_B200_H */
