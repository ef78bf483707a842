"""ORACLE: CPU restatement of ``GOOFER.synthesize`` (GOOFER.py:971-1220).

TEST INFRASTRUCTURE ONLY (see oracle/dsp.py header).  Parity status: pinned by execution against
the unmodified reference (tests/test_oracle_vs_reference.py) and by tests/golden/.

Differences in *form* (not in result): noise is never drawn here -- the caller supplies the buffers
the reference would have drawn:
    noise['sh']      randn(N)   legacy global RNG, apply_f0_jitter      GOOFER.py:666
    noise['phi']     (513, T)   uniform[0, 2pi) f32, main noise phases  GOOFER.py:1151-1152
    noise['sr_h'], noise['sr_b']  randn(N) x2, create_volume_jitter     GOOFER.py:653
Only the keyword arguments SillySampler actually passes are supported (stretch_factor == 1,
roughness_on False, glottal_smoothing False: SURVEY.md section 2.1 'dead for the CLI surface').
"""
from __future__ import annotations

import numpy as np

from . import dsp

_TABLES = {}


def _tables(sr: int, n_fft: int):
    key = (sr, n_fft)
    t = _TABLES.get(key)
    if t is None:
        nb = n_fft // 2 + 1
        t = {
            "win": dsp.sqrt_hann(n_fft),
            "freqs": dsp.bin_freqs_f32(sr, n_fft).reshape(-1, 1),
            "boost": dsp.boost_curve(nb).reshape(-1, 1),
            "bright_harm": dsp.brightness_curve(nb, sr, 2000, 3500, 3.0).reshape(-1, 1),
            "bright_breath": dsp.brightness_curve(nb, sr, 3500, 5000, 20.0).reshape(-1, 1),
        }
        _TABLES[key] = t
    return t


def formant_rows(formants, n_frames: int) -> np.ndarray:
    """GOOFER.py:48-70, 999-1002: dict (int or 'F1'.. keys) -> (4, n_frames) f64, edge-pad / trim."""
    rows = []
    d = {}
    if isinstance(formants, dict):
        for k, v in formants.items():
            if isinstance(k, str) and k.upper().startswith("F"):
                try:
                    k = int(k[1:])
                except Exception:
                    continue
            if isinstance(k, int) and 1 <= k <= 4:
                d[k] = np.asarray(v)
    for i in (1, 2, 3, 4):
        x = np.asarray(d.get(i, np.zeros(1)), dtype=np.float64)
        if x.size < n_frames:
            x = np.zeros(n_frames) if x.size == 0 else np.pad(x, (0, n_frames - x.size), mode="edge")
        rows.append(x[:n_frames])
    return np.stack(rows, axis=0)


def warp_by_formants(env: np.ndarray, F: np.ndarray, ratios, sr: int) -> np.ndarray:
    """GOOFER.py:805-875 (fa-fd): piece-wise linear frequency warp through the (shifted -> original)
    formant knots, then resampling of each frame at the warped frequencies."""
    n_bins, T = env.shape
    nyq = sr / 2.0
    freqs = np.linspace(0.0, nyq, n_bins)
    Fs = F * np.asarray(ratios, dtype=np.float64)[:, None]
    out = np.zeros_like(env)
    for t in range(T):
        src = [0.0]
        dst = [0.0]
        for i in range(4):
            fo, fs = F[i, t], Fs[i, t]
            if fo > 50.0 and fo < nyq and fs > 50.0:
                src.append(fo)
                dst.append(fs)
        src.append(nyq)
        dst.append(nyq)
        wf = dsp.lerp_extrap(np.asarray(dst), np.asarray(src), freqs)
        out[:, t] = dsp.lerp_extrap(freqs, env[:, t], wf)
    return out


def shift_all_formants(env: np.ndarray, ratio: float, sr: int) -> np.ndarray:
    """GOOFER.py:618-627 (g): env'(f) = env(clip(f / ratio, 0, nyq))."""
    n_bins, T = env.shape
    freqs = np.linspace(0, sr / 2, n_bins)
    q = np.clip(freqs / ratio, 0, sr / 2)
    out = np.zeros_like(env)
    for t in range(T):
        out[:, t] = dsp.lerp_extrap(freqs, env[:, t], q)
    return out


def smooth_mask(mask: np.ndarray, sigma=100, ds=4) -> np.ndarray:
    """GOOFER.py:556-569: decimate by 4, Gaussian, lerp back on float32 linspace abscissae."""
    short = mask[::ds].astype(np.float32) if ds > 1 else mask.astype(np.float32)
    s = dsp.gauss1d(short, max(1.0, sigma / max(1, ds)))
    if ds <= 1:
        return s.astype(np.float32)
    xo = np.linspace(0.0, 1.0, num=s.size, dtype=np.float32)
    xn = np.linspace(0.0, 1.0, num=mask.size, dtype=np.float32)
    return dsp.lerp_extrap(xo, s, xn).astype(np.float32)


def jitter_curve(noise_randn: np.ndarray, sr: int, speed: float, strength: float) -> np.ndarray:
    """GOOFER.py:653-657 / 666-669: smoothed, max-normalised noise -> 1 + noise * strength (fp64)."""
    z = dsp.gauss1d(np.asarray(noise_randn, dtype=np.float64), sr / (speed * 6))
    z = z / np.max(np.abs(z) + 1e-6)
    return 1.0 + z * strength


def tremolo_curve(length: int, sr: int, speed: float, strength: float) -> np.ndarray:
    """GOOFER.py:642-659 with vibrato=True, seed=None (phase 0), 0.1 s linear fade-in, clip [.5, 1.5]."""
    t = np.arange(length) / sr
    s = np.sin(2 * np.pi * speed * t + 0)
    fade = int(0.1 * sr)
    if fade < length:
        s[:fade] *= np.linspace(0, 1, fade)
    return np.clip(1.0 + s * strength, 0.5, 1.5)


def growl_f0(f0_f32: np.ndarray, sr: int, rate: float, depth: float, delay: float) -> np.ndarray:
    """GOOFER.py:748-766 apply_subharm_vibrato (seed=None => phase 0); result keeps f0's f32 dtype."""
    t = np.arange(len(f0_f32)) / sr
    vib = np.sin(2 * np.pi * rate * t + 0)
    nf = int(delay * sr)
    fade = np.linspace(0, 1, nf)
    if len(fade) < len(vib):
        vib[:nf] *= fade
    v = f0_f32 > 0
    out = f0_f32.copy()
    out[v] = out[v] * (1 + vib[v] * depth)
    return out


def synthesize(env_spec, f0_interp, voicing_mask, n_out: int, sr: int, noise: dict, *,
               n_fft=dsp.N_FFT, hop=dsp.HOP, normalize=1.0, uv_strength=0.75, breath_strength=0.1,
               noise_transition_smoothness=100, formant_shift=1.0,
               f0_jitter=False, f0_jitter_speed=100, f0_jitter_strength=1.5,
               volume_jitter=False, volume_jitter_speed=150, volume_jitter_strength_harm=50,
               volume_jitter_strength_breath=100,
               add_subharm=False, subharm_semitones=-12, subharm_weight=0.5, subharm_vibrato=False,
               subharm_vibrato_rate=6.0, subharm_vibrato_depth=0.1, subharm_vibrato_delay=0.1,
               F_shifts=(1.0, 1.0, 1.0, 1.0), formants=None, taps=None):
    """GOOFER.py:971-1220.  ``n_out`` plays the role of ``len(y)``.  Returns
    (reconstruct, harmonic, aper_uv, aper_bre), all (n_out,) float32."""
    tb = _tables(sr, n_fft)
    win = tb["win"]
    if isinstance(env_spec, dict):
        env_spec = dsp.decode_knots(env_spec)                      # :986-987
    env = np.asarray(env_spec, dtype=np.float32)                   # :988
    f0 = np.array(f0_interp, dtype=np.float32)                     # :989 (+ private copy)
    vm = np.asarray(voicing_mask, dtype=np.float32)                # :990

    env_noise_src = dsp.gauss1d(env, 1.75, axis=0)                 # :993 (taken BEFORE the formant warps)
    T_env = env.shape[1]
    F = formant_rows(formants, T_env)                              # :999-1002
    if any(s != 1.0 for s in F_shifts):                            # :1004-1014
        env = warp_by_formants(env, F, F_shifts, sr)
    if formant_shift != 1.0:                                       # :1016-1017
        env = shift_all_formants(env, formant_shift, sr)

    if f0_jitter:                                                  # :1069-1071
        j = jitter_curve(noise["sh"], sr, f0_jitter_speed, f0_jitter_strength)
        f0 *= 1.0 + ((j - 1.0) * vm)

    pulse = dsp.pulse_train(f0, sr, Ra=0.02, Rg=1.7, Rk=0.8)       # :1074
    if add_subharm:                                                # :1076-1097
        f0s = growl_f0(f0, sr, subharm_vibrato_rate, subharm_vibrato_depth, subharm_vibrato_delay) \
            if subharm_vibrato else f0
        pulse += dsp.subharm_layer(f0s, sr, subharm_weight, subharm_semitones, vm)
    if taps is not None:
        taps["f0"] = f0.copy()
        taps["pulse"] = pulse.copy()

    S = dsp.stft(pulse, n_fft, hop, win)                           # :1099
    T = S.shape[1]
    f0_fr = f0[::hop]                                              # :1104-1106
    f0_fr = np.pad(f0_fr, (0, max(0, T - len(f0_fr))), mode="edge")[:T]
    hp = 1.0 / (1.0 + np.exp(-np.clip((tb["freqs"] - f0_fr.reshape(1, -1)) / 5, -60, 60)))   # :1111
    S *= hp                                                        # :1114
    if env.shape[1] > T:                                           # :1115-1119
        env = env[:, :T]
    elif env.shape[1] < T:
        env = np.pad(env, ((0, 0), (0, T - env.shape[1])), mode="edge")
    mag = np.max(np.abs(S) + 1e-8)                                 # :1121
    S = (S / mag) * env                                            # :1128
    S *= tb["boost"]                                               # :1129

    vf = vm[::hop]                                                 # :1132-1136
    vf = np.pad(vf, (0, T - vf.size), mode="edge") if vf.size < T else vf[:T]
    cols = np.nonzero(vf > 0)[0]
    if cols.size:                                                  # :1138-1144
        blk = S[:, cols] * tb["bright_harm"]
        S[:, cols] = dsp.gauss1d(blk, 0.5, axis=0)
    if taps is not None:
        taps["S_harm"] = S.copy()
        taps["mag"] = float(mag)
    harmonic = dsp.istft(S, hop, win, n_out)                       # :1146

    envn = env_noise_src                                           # :1148 match_env_frames
    if envn.shape[1] > T:
        envn = envn[:, :T]
    elif envn.shape[1] < T:
        envn = np.pad(envn, ((0, 0), (0, T - envn.shape[1])), mode="edge")
    envn = envn.astype(np.float32)
    phi = np.asarray(noise["phi"], dtype=np.float32)               # :1151-1152
    assert phi.shape == envn.shape, (phi.shape, envn.shape)
    U = np.cos(phi) + 1j * np.sin(phi)                             # :1153
    S_uv = U * envn                                                # :1156
    S_br = (U * envn) * hp                                         # :1157
    if cols.size:                                                  # :1159-1173
        blk = S_br[:, cols] * tb["bright_breath"]
        S_br[:, cols] = dsp.gauss1d(blk, 0.5, axis=0)
    aper_breath = dsp.istft(S_br, hop, win, n_out)                 # :1175
    aper_uv = dsp.istft(S_uv, hop, win, n_out)                     # :1176

    ms = smooth_mask(vm, sigma=noise_transition_smoothness, ds=4)  # :1179
    aper_bre = aper_breath * ms * breath_strength                  # :1180
    aper_uv = aper_uv * (1.0 - ms) * uv_strength                   # :1181
    if taps is not None:
        taps["mask_smooth"] = ms.copy()

    if volume_jitter:                                              # :1185-1191
        hj = jitter_curve(noise["sr_h"], sr, volume_jitter_speed, volume_jitter_strength_harm)
        bj = jitter_curve(noise["sr_b"], sr, volume_jitter_speed, volume_jitter_strength_breath)
        vjm = dsp.gauss1d(vm, 20)
        harmonic *= 1.0 + (hj - 1.0) * vjm
        aper_bre *= 1.0 + (bj - 1.0) * vjm

    combined = harmonic + aper_uv + aper_bre                       # :1193
    peak = float(np.max(np.abs(combined)) + 1e-12)                 # :1210
    gain = (1.0 / peak) ** float(np.clip(normalize, 0.0, 1.0))     # :1208-1213
    harmonic *= gain
    aper_uv *= gain
    aper_bre *= gain
    return combined * gain, harmonic, aper_uv, aper_bre
