// gf_plan.h -- POD records shared by the host planner (plan.cpp) and the device kernels.
//
// A GfNotePlan holds every integer length / index and every flag-derived scalar that
// SillySampler.GooferResampler derives on the host before it touches sample data
// (/root/reference/SillySampler.py:286-410 flag scalars, :453-500 slice bounds, :625-763 loop
// lengths, :766-788 velocity stretch lengths).  The kernels consume it read-only.
#pragma once
#include <stdint.h>

#define GF_NFFT 1024
#define GF_HOP 256
#define GF_NBINS 513
#ifndef GF_FT
#define GF_FT 8                 // frames per envelope tile ([tile][bin][GF_FT] layout in the workspace)
#endif
#define GF_MAX_PASSES 4         // main, su, sj, sa   (SillySampler.py:1006,1041,1067,1156)

enum { GF_LOOP_CONCAT = 0, GF_LOOP_AVG = 1, GF_LOOP_STRETCH = 2 };
enum { GF_PASS_MAIN = 0, GF_PASS_SU = 1, GF_PASS_SJ = 2, GF_PASS_SA = 3 };

struct GfNotePlan {
    int32_t status;
    int32_t src;                // source index
    int32_t reverse;            // R1
    int32_t sr;
    int32_t T_src, N_src;       // source envelope frames / mask samples
    int32_t F_len[4];           // source formant track lengths
    // normalised python slices [a, b) into the (possibly reversed) source
    int32_t pre_f_a, pre_f_n;   // env_pre   = env[:, fr0:fr1]
    int32_t tail_f_a, tail_f_n; // env_tail  = env[:, fr1:fr2]
    int32_t pre_s_a, pre_s_n;   // mask_pre  = mask[s0:s1]
    int32_t tail_s_a, tail_s_n; // mask_tail = mask[s1:s2]
    int32_t fr0, fr1, fr2;      // raw frame bounds (tracks are sliced against their own length)
    int32_t want_frames, want_samples;
    int32_t loop_mode;
    int32_t env_direct;         // tail_f_n >= want_frames: plain trim
    int32_t reps, rem, fade, fade_r;   // concat bookkeeping (SillySampler.py:654-696)
    int32_t unit_len;           // 2*have - fade
    int32_t stretch_target;     // L2: int(have * (want / have))
    int32_t T_loop;             // env_tail_looped.shape[1]
    int32_t T0_frames;          // env_new.shape[1] before the velocity stretch
    int32_t n0_total;           // len(mask_new) before the velocity stretch
    int32_t vel_active;
    int32_t pre_new_f, pre_new_s;      // stretched prefix lengths
    double vel;                 // 2 ** (1 - velocity / 100)
    int32_t T_env;              // env_new.shape[1]
    int32_t n_total;            // len(f0_new)
    int32_t T_out;              // 1 + n_total // 256
    int32_t n_passes;
    int32_t pass_kind[GF_MAX_PASSES];  // GF_PASS_* of pass slot p (slot 0 is always main)
    // ---- pitch ----
    int32_t pitch_midi;
    int32_t t_cents;
    int32_t bend_len;
    int64_t bend_off;
    double tempo;
    // ---- flag scalars (SillySampler.py:313-410) ----
    double formant_shift;       // 1 + g/200
    double F_shift[4];          // 1 + fa..fd / 100
    int32_t any_F_shift;
    double brightness_env;      // (br + 100) / 100
    double es;                  // clip(es) / 100
    double fw;                  // fw / 100 * 0.1
    double fst[4];
    int32_t any_fst;
    double V, B, U, volume;
    int32_t f0_jitter;  double f0_jitter_strength;
    int32_t vol_jitter; double vol_jitter_strength;
    double vol_jitter_strength_breath;   // 2 x the harmonic strength (SillySampler.py:1024) unless overridden
    float breath_strength, uv_strength;  // gf.synthesize keyword defaults 0.1 / 0.75 (GOOFER.py:975, 1180-1181)
    double sd;
    double tension;
    int32_t add_subharm; double subharm_weight;
    double sj, sa, su;
    double normalize;
    int32_t FV;
    double pd;
    double vf, vh, vl;          // vf already clipped to [-100, 100]
    // vocal fry derived (SillySampler.py:890-965)
    int32_t fry_on;             // vf != 0
    int32_t fry_L, fry_glide, fry_const;   // f0 override lengths
    int32_t fry_mask_on, fry_a, fry_b, fry_fade;   // fry mask support [a, b), 10 ms ramps
    // ---- noise / output offsets (elements) ----
    int64_t phi_off[4];
    int64_t nrm_off[4];
    int64_t out_off;
    int64_t f0_off;             // >= 0: direct gf.synthesize call, f0 curve at GooferBatch.f0_curves + f0_off (else -1)
    uint64_t phi_rng[4][4];     // per phi slot: PCG64 state_hi, state_lo, inc_hi, inc_lo (GooferNote.phi_rng)
    uint32_t phi_rng_mask;      // slots whose phases are generated on the device
    uint32_t reserved0;
};
