"""ctypes mirror of struct GfNotePlan (goofer_b200/csrc/gf_plan.h) for the planner tests."""
import ctypes as C

i32, i64, f64 = C.c_int32, C.c_int64, C.c_double


class GfNotePlan(C.Structure):
    _fields_ = [
        ("status", i32), ("src", i32), ("reverse", i32), ("sr", i32), ("T_src", i32), ("N_src", i32),
        ("F_len", i32 * 4),
        ("pre_f_a", i32), ("pre_f_n", i32), ("tail_f_a", i32), ("tail_f_n", i32),
        ("pre_s_a", i32), ("pre_s_n", i32), ("tail_s_a", i32), ("tail_s_n", i32),
        ("fr0", i32), ("fr1", i32), ("fr2", i32),
        ("want_frames", i32), ("want_samples", i32), ("loop_mode", i32), ("env_direct", i32),
        ("reps", i32), ("rem", i32), ("fade", i32), ("fade_r", i32), ("unit_len", i32), ("stretch_target", i32),
        ("T_loop", i32), ("T0_frames", i32), ("n0_total", i32), ("vel_active", i32),
        ("pre_new_f", i32), ("pre_new_s", i32),
        ("vel", f64),
        ("T_env", i32), ("n_total", i32), ("T_out", i32), ("n_passes", i32), ("pass_kind", i32 * 4),
        ("pitch_midi", i32), ("t_cents", i32), ("bend_len", i32), ("bend_off", i64), ("tempo", f64),
        ("formant_shift", f64), ("F_shift", f64 * 4), ("any_F_shift", i32),
        ("brightness_env", f64), ("es", f64), ("fw", f64), ("fst", f64 * 4), ("any_fst", i32),
        ("V", f64), ("B", f64), ("U", f64), ("volume", f64),
        ("f0_jitter", i32), ("f0_jitter_strength", f64), ("vol_jitter", i32), ("vol_jitter_strength", f64),
        ("vol_jitter_strength_breath", f64), ("breath_strength", C.c_float), ("uv_strength", C.c_float),
        ("sd", f64), ("tension", f64), ("add_subharm", i32), ("subharm_weight", f64),
        ("sj", f64), ("sa", f64), ("su", f64), ("normalize", f64), ("FV", i32), ("pd", f64),
        ("vf", f64), ("vh", f64), ("vl", f64),
        ("fry_on", i32), ("fry_L", i32), ("fry_glide", i32), ("fry_const", i32),
        ("fry_mask_on", i32), ("fry_a", i32), ("fry_b", i32), ("fry_fade", i32),
        ("phi_off", i64 * 4), ("nrm_off", i64 * 4), ("out_off", i64), ("f0_off", i64),
        ("phi_rng", C.c_uint64 * 16), ("phi_rng_mask", C.c_uint32), ("reserved0", C.c_uint32),
    ]
