// gf_device.cuh -- device-side records, constant tables and small helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gf_hd.h"
#include "gf_plan.h"
#include "gf_fft.cuh"
#include "gf_conv.cuh"

// ---- constant tables (GOOFER.py:12-46 get_cached_window/freqs/boost/brightness; :241-261 taps) ----
struct GfTables {
    float2 twl[GF_TWL_N];   // per-thread twiddles of the radix-8 passes (gf_fft.cuh gf_twl_fill)
    float2 tw1024[513];     // exp(-2 pi i k / 1024)
    float win[1024];        // sqrt(hanning(1024)) in f32          GOOFER.py:16
    float win2[1024];       // win * win (f32)                      GOOFER.py:386
    float boost[513];       // linspace(1, 100, 513) f32            GOOFER.py:33
    float bright_h[513];    // +3 dB 2-3.5 kHz                      GOOFER.py:42
    float bright_b[513];    // +20 dB 3.5-5 kHz                     GOOFER.py:43
    float freq32[513];      // rfftfreq f32                         GOOFER.py:24
    double g175[15];        // Gaussian sigma 1.75 (env4breath)     GOOFER.py:993
    double g05[5];          // Gaussian sigma 0.5 (brightness blur) GOOFER.py:1143
    float winG[1024];       // win * G, G = time-domain image of the sigma-0.5 blur (k_frame.cu gf_blur_edges)
    float bq1[16], bq2[16]; // edge-correction taps of that blur, index = bin 1..12
    int sr;
};

// the library is built as ONE translation unit (goofer_b200.cu includes every k_*.cu)
__device__ GfTables d_tab;

// twiddles of the overlap-save transforms (gf_conv.cuh gf_conv_tw_fill: per pass, per thread), fp64 cos / sin rounded once
struct GfConvTables {
    GfC<float> tw32[GF_CONV_N32];
    GfC<double> tw64[GF_CONV_N64];
};
__device__ GfConvTables d_conv;

// one voicebank source as the kernels see it
struct GfSourceDev {
    const uint16_t *knots;  // (K, T) f16 or NULL
    const float *hz_knots;  // (K,)
    const float *dense;     // (513, T) f32 or NULL
    const float *mask;      // (N,)
    const double *formants[4];
    float *envS;            // workspace: (T, GF_ENVS_LD) frame-major decoded envelope
    int K, T, N;
    int formant_len[4];
};
#define GF_ENVS_LD 520

// per-(note, pass) buffers and device-written scalars
struct GfPassDev {
    int note;               // index of the note in the wave
    int kind;               // GF_PASS_*
    int n_total, T_out;
    float *f0;              // (n_total,) f32 excitation f0 of this pass
    float *pulse;           // (n_total,)
    float *sub;             // sg layer, raw sum (main pass, add_subharm) or NULL
    float *harm, *bre, *uv; // raw OLA streams (n_total,)
    int4 *onsets;           // (i, T0, f32 bits of last_valid_f0, f32 bits of table max)
    int onset_cap;
    float *phi;             // workspace: the pass's noise phases FRAME-MAJOR (T_out, GF_ENVS_LD) like envF / envN -- what the
                            // frame kernel reads (a frame's bins contiguously); drawn there by gf_phi_kernel (GooferNote.phi_rng)
                            // or transposed there from the caller's buffer by gf_phi_fm_kernel
    const float *phi_src;   // the caller's (513, T_out) f32 buffer as numpy draws it, or NULL when the phases are drawn on the device
    int mask_ones;          // sa pass: voicing mask == 1
};

struct GfPassScal {         // zeroed per wave, written by the device
    unsigned int mag_bits;  // max(|S * hp| + 1e-8)            GOOFER.py:1121
    unsigned int peak_bits; // max|harm + uv + bre|            GOOFER.py:1210
    unsigned int submax_bits;
    int n_onsets;
    int max_T0;
    int err;
    int n_sub_events;
    int sub_max_len;        // longest growl pulse (samples)
    int sg_seq;             // growl events: the exact-sum scan met a borderline crossing, run the sequential walk
    int walk_seq;           // pulse onsets: the fixed-point scan met a borderline crossing, run the bit-exact walk
};

struct GfNoteDev {
    float *trk_canon;       // (4, T_env) f32
    float *trk_clean;       // (4, T_env) f32 (fst) or NULL
    float *env_aux;         // per-note tables of the envelope kernel (br tilt, f32 bin frequencies, es taps)
    float *envF, *envN;     // (T_out, GF_ENVS_LD) f32 frame-major: shaped envelope / noise envelope on the STFT frame grid
    float *vm;              // (n_total,) f32(mask_new)
    float *vm4;             // ((n_total + 3) / 4,) vm[::4], written beside vm: the decimated input of smooth_mask_ds' sigma-25 smoothing
    float *f0n;             // (n_total,) f32(f0_new): cutoff driver of the post-FX filters, or NULL
    float *ms_short;        // (ceil(n/4),) f32: gaussian-smoothed decimated mask   GOOFER.py:556-563
    float *ms;              // (n_total,) f32: smooth_mask_ds result (lerp of ms_short)  GOOFER.py:564-569
    unsigned char *ms_one;  // (ceil(n/256),) 1 where ms == 1 on the whole hop block [256 b, 256 b + 256)
    double *z_sh;           // (n_total,) smoothed sh noise (f0 jitter) or NULL
    double *z_srh, *z_srb;  // smoothed sr noise
    float *vjm;             // gauss(vm, 20) (sr)
    float *sdm;             // gauss(mask_new, 20) (sd)
    float *pd_in;           // pd: f32(midi - base)                     SillySampler.py:860-866
    double *pd_dev;         // pd: gaussian-smoothed bend deviation (fp64)
    float *pd_gm;           // pd: gauss(mask_new, 441)                  SillySampler.py:880
    float *fx[4];           // post-FX streams: [0] harm, [1] bre, [2] / [3] layer or filter scratch
    float *alpha[2];        // per-sample filter coefficients of the two concurrent one-pole jobs
    // sg (growl) layer: modulated f0, event list, first-occurrence pulse bank  GOOFER.py:700-766
    float *sg_f0;           // (n_total,) f32 apply_subharm_vibrato(f0)
    int *sg_ev_i;           // (sg_cap,) event sample index
    double *sg_ev_f;        // (sg_cap,) event sub_f0
    int *sg_rep;            // (sg_cap,) index of the first event with the same '%.2f' key
    int *sg_len;            // (sg_cap,) length of the event's pulse = that of its representative (gf_sg_bank_kernel)
    float *sg_m;            // (sg_cap,) peak of that pulse table
    int2 *sg_tab;           // (sg_tab_n,) open-addressing table: x = key, y = first event index
    int sg_cap, sg_tab_n;
    double *noteScal;       // small per-note double scalars (maxima, rms)
    float *out;             // (n_total,) final output (or NULL when only PCM is wanted)
    short *pcm;             // (n_total,) 16-bit PCM of the final output or NULL
    float *tap_harm, *tap_uv, *tap_bre;
    int pass0;              // first entry of this note in the pass arrays
};

enum { GF_NS_SHMAX = 0, GF_NS_SRHMAX, GF_NS_SRBMAX, GF_NS_R0, GF_NS_R1, GF_NS_PDREF, GF_NS_COUNT = 8 };   // R0 / R1: sums of squares

// ---- job records of the generic kernels ----
struct GfFirJob {
    const void *in;  int in_f64;  int in_stride;  int n;
    void *out;       int out_f64;
    double sigma;
    double *maxabs;         // atomicMax target for max(|y| + 1e-6) (bits of a non-negative double) or NULL
    int in_cast_f32;        // round the input to f32 first
    int pad;
};

struct GfOnepoleJob {
    const float *x;         // input (n,)
    const float *f0;        // (n,) f32 cutoff driver (already 5-tap smoothed by the caller when needed) or NULL => const
    float *y;               // output (n,) (may alias x)
    float *alpha;           // scratch (n,)
    int n;
    int order;
    int highpass;
    int smooth_f0;          // apply the 5-tap edge-padded moving average first (SillySampler.py:107-113)
    double cutoff_factor;
    double f0_const;        // used when f0 == NULL
    double f0_floor;        // max(f0, floor) applied to the driver first (su/sj: 120)  SillySampler.py:1052
    int sr;
    int pad;
};

// ---- per-device launch state ---------------------------------------------------------------------
// cudaFuncSetAttribute applies to the CURRENT device only, so the "already raised" memo of a kernel's dynamic
// shared-memory limit is kept per device: one process may drive several GPUs (one host thread each).  A benign race
// between two threads on the same device sets the same value twice.
#define GF_MAX_DEVICES 64
struct GfSmemLimit { size_t bytes[GF_MAX_DEVICES]; };
template <typename K> static inline void gf_smem_limit(K kernel, size_t bytes, GfSmemLimit &memo)
{
    int dev = 0;
    const bool known = cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < GF_MAX_DEVICES;
    if (known && memo.bytes[dev] >= bytes) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (known) memo.bytes[dev] = bytes;
}

// ---- small helpers ------------------------------------------------------------------------------
__device__ __forceinline__ int gf_reflect(int q, int n)
{
    // numpy 'reflect' padding as a periodic extension (period 2(n-1)); n == 1 degenerates to 0
    if (n <= 1) return 0;
    const int per = 2 * (n - 1);
    q %= per;
    if (q < 0) q += per;
    return q < n ? q : per - q;
}

// np.linspace(0, 1, num)[i] / np.linspace(1, 0, num)[i] in fp64 (numpy: i * step + start, endpoint exact)
__device__ __forceinline__ double gf_dlin01(int i, int num)
{
    if (num <= 1) return 0.0;
    if (i == num - 1) return 1.0;
    return (double)i * (1.0 / (double)(num - 1));
}
__device__ __forceinline__ double gf_dlin10(int i, int num)
{
    if (num <= 1) return 1.0;
    if (i == num - 1) return 0.0;
    return (double)i * (-1.0 / (double)(num - 1)) + 1.0;
}

// vocal-fry mask value at sample c (SillySampler.py:937-965): ones on [fry_a, fry_b) with 10 ms linear ramps (f32)
__device__ __forceinline__ float gf_fry_at(const GfNotePlan &pl, int c)
{
    if (!pl.fry_mask_on || c < pl.fry_a || c >= pl.fry_b) return 0.0f;
    float mf = 1.0f;
    const int fade = pl.fry_fade;
    if (fade > 0) {
        const int a1 = min(pl.fry_b, pl.fry_a + fade), b0 = max(pl.fry_a, pl.fry_b - fade);
        if (c < a1) mf = (float)gf_dlin01(c - pl.fry_a, a1 - pl.fry_a);
        if (c >= b0) mf = (float)((double)mf * gf_dlin10(c - b0, pl.fry_b - b0));
    }
    return mf;
}

// smooth_mask_ds (GOOFER.py:556-569): lerp of the smoothed decimated mask back to sample rate on float32
// linspace abscissae (their f32 rounding moves the lerp weight by up to 7e-4, so it is reproduced), result f32
__device__ __forceinline__ float gf_ms_at(const float *__restrict__ s, int M, int i, int N)
{
    if (M == 1) return s[0];
    const float x = (float)gf_dlin01(i, N);
    if (x >= 1.0f) return s[M - 1];
    int j = (int)(x * (float)(M - 1));
    if (j > M - 2) j = M - 2;
    float x0 = (float)gf_dlin01(j, M);
    while (j > 0 && x0 > x) { --j; x0 = (float)gf_dlin01(j, M); }
    float x1 = (float)gf_dlin01(j + 1, M);
    while (j < M - 2 && x1 <= x) { ++j; x0 = x1; x1 = (float)gf_dlin01(j + 1, M); }
    if (x0 == x) return s[j];
    // np.interp: slope * (x - x0) + y0 in fp64; the two differences are exact, one f32 division suffices
    const float t = (x - x0) / (x1 - x0);
    return (float)fma((double)s[j + 1] - (double)s[j], (double)t, (double)s[j]);
}

__device__ __forceinline__ void gf_atomic_max_pos(unsigned int *addr, float v)
{
    atomicMax(addr, __float_as_uint(v));     // v >= 0: IEEE bit patterns order like unsigned ints
}

__device__ __forceinline__ float gf_warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double gf_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float gf_warp_sumf(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
