"""ORACLE primitives -- CPU restatement of the reference's DSP building blocks.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package ``goofer_b200``.

Parity status: PINNED BY EXECUTION -- the reference ships no tests or golden vectors
(SURVEY.md section 4); every function here is checked against the unmodified reference imported from
/root/reference by tests/golden/make_golden.py + tests/test_oracle_vs_reference.py (run in the build
container) and against the committed outputs of that run under tests/golden/.

Each function cites the reference lines it restates.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

N_FFT = 1024            # SillySampler.py:14
HOP = N_FFT // 4        # SillySampler.py:15
N_BINS = N_FFT // 2 + 1

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libgoofer_oracle_seq.so")
_lib = None


def build_seq_lib(force: bool = False) -> str:
    """Compile oracle/seq_kernels.c (plain gcc, no fast-math, no FMA contraction)."""
    src = os.path.join(_HERE, "seq_kernels.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


def _seq():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build_seq_lib())
        c = ctypes
        fp, dp, ip = c.POINTER(c.c_float), c.POINTER(c.c_double), c.POINTER(c.c_int32)
        lib.orc_pulse_train.restype = c.c_long
        lib.orc_pulse_train.argtypes = [fp, c.c_long, c.c_double, c.c_double, c.c_double, c.c_double,
                                        fp, ip, ip, c.c_long]
        lib.orc_sub_events.restype = c.c_long
        lib.orc_sub_events.argtypes = [dp, dp, c.c_long, c.c_double, c.c_double, ip, dp, c.c_long]
        lib.orc_onepole_cascade.restype = None
        lib.orc_onepole_cascade.argtypes = [fp, fp, c.c_long, c.c_int, c.c_int]
        lib.orc_overlap_add.restype = None
        lib.orc_overlap_add.argtypes = [fp, fp, c.c_long, c.c_long, c.c_long, fp, c.c_long]
        _lib = lib
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


# ----------------------------------------------------------------------------------------------
# windows / constant tables   (GOOFER.py:12-46, 585-595)
# ----------------------------------------------------------------------------------------------
def sqrt_hann(n_fft: int = N_FFT) -> np.ndarray:
    """GOOFER.py:16 -- symmetric Hann cast to f32 and then square-rooted in f32."""
    return np.hanning(n_fft).astype(np.float32) ** 0.5


def bin_freqs_f32(sr: int, n_fft: int = N_FFT) -> np.ndarray:
    """GOOFER.py:24 -- rfftfreq as an f32 column (used by the high-pass sigmoid)."""
    return np.fft.rfftfreq(n_fft, 1.0 / sr).astype(np.float32)


def boost_curve(n_bins: int = N_BINS) -> np.ndarray:
    """GOOFER.py:33 -- linspace(1, 100) in f32."""
    return np.linspace(1, 100, n_bins, dtype=np.float32)


def brightness_curve(n_bins: int, sr: int, start_hz: float, end_hz: float, gain_db: float) -> np.ndarray:
    """GOOFER.py:585-595 (cast to f32 at :42-43)."""
    f = np.linspace(0, sr / 2, n_bins)
    g = np.ones_like(f)
    a = int(np.searchsorted(f, start_hz))
    b = int(np.searchsorted(f, end_hz))
    top = 10 ** (gain_db / 20)
    g[a:b] = 1 + np.linspace(0, 1, b - a) * (top - 1)
    g[b:] = top
    return g.astype(np.float32)


# ----------------------------------------------------------------------------------------------
# Gaussian FIR with numpy-'reflect' padding   (GOOFER.py:241-261)
# ----------------------------------------------------------------------------------------------
def gauss_taps(sigma: float, truncate: float = 4.0):
    """Kernel of GOOFER.py:247-252; returns (radius, fp64 taps) or (0, None) when it is a no-op."""
    if sigma <= 0.0:
        return 0, None
    radius = int(truncate * sigma + 0.5)
    if radius <= 0:
        return 0, None
    t = np.arange(-radius, radius + 1)
    k = np.exp(-0.5 * (t / sigma) ** 2)
    k /= k.sum()
    return radius, k


def gauss1d(x, sigma: float, axis: int = -1) -> np.ndarray:
    """GOOFER.py:241-261.  Result dtype follows numpy promotion with the fp64 taps
    (f32 -> f64, c64 -> c128), exactly like np.convolve in the reference."""
    a = np.asarray(x)
    if a.size == 0 or a.shape[axis] == 0:
        return a.copy()
    radius, k = gauss_taps(float(sigma))
    if radius == 0:
        return a.copy()
    m = np.moveaxis(a, axis, -1)
    padded = np.pad(m, [(0, 0)] * (m.ndim - 1) + [(radius, radius)], mode="reflect")
    n = m.shape[-1]
    if m.ndim == 1:
        out = np.convolve(padded, k, mode="valid")
    else:
        out = np.zeros(m.shape, dtype=np.result_type(m.dtype, np.float64))
        # symmetric taps: convolution == correlation; accumulate tap by tap over the whole array
        for j in range(2 * radius + 1):
            out += k[j] * padded[..., j:j + n]
    return np.moveaxis(out, -1, axis)


# ----------------------------------------------------------------------------------------------
# linear interpolation with linear extrapolation   (GOOFER.py:173-239)
# ----------------------------------------------------------------------------------------------
def lerp_extrap(x, y, xq) -> np.ndarray:
    """interp1d(kind='linear', fill_value='extrapolate') of GOOFER.py:173-239 evaluated at xq."""
    x = np.asarray(x)
    y = np.asarray(y)
    xq = np.asarray(xq)
    if x.size == 0:
        raise ValueError("x cannot be empty")
    if x.size == 1:
        return np.full_like(xq, y[0], dtype=y.dtype)           # :183-191
    sl = (y[1] - y[0]) / (x[1] - x[0] + 1e-10)                 # :204
    sr_ = (y[-1] - y[-2]) / (x[-1] - x[-2] + 1e-10)            # :205
    out = np.interp(xq, x, y)                                  # :227 (always fp64)
    lo = xq < x[0]
    if np.any(lo):
        out[lo] = y[0] + sl * (xq[lo] - x[0])
    hi = xq > x[-1]
    if np.any(hi):
        out[hi] = y[-1] + sr_ * (xq[hi] - x[-1])
    return out


def stretch_rows(feature: np.ndarray, stretch: float) -> np.ndarray:
    """GOOFER.py:597-616 stretch_feature (1-D or 2-D along the last axis)."""
    if stretch == 1.0:
        return feature.copy()
    n = feature.shape[-1]
    target = int(n * stretch)
    xo = np.linspace(0, 1, n)
    xn = np.linspace(0, 1, target)
    if feature.ndim == 1:
        return lerp_extrap(xo, feature, xn)
    return np.stack([lerp_extrap(xo, row, xn) for row in feature], axis=0)


# ----------------------------------------------------------------------------------------------
# STFT / iSTFT   (GOOFER.py:355-413)
# ----------------------------------------------------------------------------------------------
def stft(x, n_fft: int = N_FFT, hop: int = HOP, window=None) -> np.ndarray:
    """GOOFER.py:355-370: reflect-pad n_fft/2, sqrt-Hann, rfft (f32 in => c64 out on numpy >= 2)."""
    if window is None:
        window = np.hanning(n_fft) ** 0.5
    x = np.asarray(x, dtype=np.float32)
    pad = n_fft // 2
    xp = np.pad(x, pad, mode="reflect") if len(x) >= 2 else np.pad(x, pad, mode="edge")
    if len(xp) < n_fft:
        xp = np.pad(xp, (0, n_fft - len(xp)), mode="edge")
    T = max(1, 1 + (len(xp) - n_fft) // hop)
    idx = np.arange(n_fft)[:, None] + hop * np.arange(T)[None, :]
    frames = xp[idx]
    frames *= window[:, None]
    return np.fft.rfft(frames, axis=0)


def istft(S, hop: int = HOP, window=None, length=None) -> np.ndarray:
    """GOOFER.py:392-413 with the sequential f32 overlap-add of :372-390."""
    n_fft = (S.shape[0] - 1) * 2
    window = sqrt_hann(n_fft) if window is None else np.asarray(window, dtype=np.float32)
    S = np.asarray(S, dtype=np.complex64)
    frames = np.ascontiguousarray(np.fft.irfft(S, axis=0, n=n_fft).astype(np.float32))
    T = frames.shape[1]
    expected = n_fft + hop * (T - 1)
    y = np.empty(expected, dtype=np.float32)
    w = np.ascontiguousarray(window, dtype=np.float32)
    _seq().orc_overlap_add(_p(frames, ctypes.c_float), _p(w, ctypes.c_float), n_fft, T, hop,
                           _p(y, ctypes.c_float), expected)
    pad = n_fft // 2
    y = y[pad:expected - pad]
    if length is not None:
        if y.shape[0] < length:
            y = np.pad(y, (0, length - y.shape[0]), mode="constant")
        else:
            y = y[:length]
    return y


# ----------------------------------------------------------------------------------------------
# mel-knot envelope codec   (GOOFER.py:74-168)
# ----------------------------------------------------------------------------------------------
def mel_knots_hz(sr: int, K: int) -> np.ndarray:
    """GOOFER.py:77-82 (f32 knots on a mel-uniform grid)."""
    mel_max = 2595.0 * np.log10(1.0 + (sr / 2.0) / 700.0)
    mel = np.linspace(2595.0 * np.log10(1.0), mel_max, K, dtype=np.float32)
    return (700.0 * (10 ** (mel / 2595.0) - 1.0)).astype(np.float32)


def knot_lerp_table(freqs_f32: np.ndarray, hz_knots: np.ndarray):
    """GOOFER.py:84-95 as (idx, w0, w1) instead of the dense (N, K) matrix (<= 2 nnz per row)."""
    K = len(hz_knots)
    idx = np.clip(np.searchsorted(hz_knots, freqs_f32, side="right") - 1, 0, K - 2)
    x0 = hz_knots[idx]
    x1 = hz_knots[idx + 1]
    w1 = (freqs_f32 - x0) / np.maximum(x1 - x0, 1e-12)
    w0 = 1.0 - w1
    return idx, w0.astype(np.float32), w1.astype(np.float32)


def decode_knots(pack: dict) -> np.ndarray:
    """GOOFER.py:149-168: env = exp(W @ log-knots) in f32, W materialised like the reference so the
    sgemm (and therefore the rounding) is the same library call."""
    kv = np.asarray(pack["knot_vals_log"]).astype(np.float32)
    hz = np.asarray(pack["hz_knots"]).astype(np.float32)
    n_fft, sr, n_bins = int(pack["n_fft"]), int(pack["sr"]), int(pack["n_bins"])
    freqs = np.fft.rfftfreq(n_fft, 1.0 / sr).astype(np.float32)
    idx, w0, w1 = knot_lerp_table(freqs, hz)
    W = np.zeros((len(freqs), len(hz)), dtype=np.float32)
    r = np.arange(len(freqs))
    W[r, idx] = w0
    W[r, idx + 1] = w1
    env = np.exp(W @ kv).astype(np.float32)
    return env[:n_bins]


def compress_to_knots(env_spec, sr: int, n_fft: int = N_FFT, eps=1e-2, K_start=32, K_step=16, K_max=192,
                      smooth_sigma_bins=0.5) -> dict:
    """GOOFER.py:97-147: smallest K in 32,48,...,192 whose 2-tap log-lerp reconstructs the envelope to
    max relative error < eps on <= 256 probe frames; falls back to K_max."""
    env = np.asarray(env_spec, dtype=np.float32)
    if smooth_sigma_bins > 0:
        env = gauss1d(env, smooth_sigma_bins, axis=0)
    log_env = np.log(np.maximum(env, 1e-8)).astype(np.float32)
    n_bins, T = log_env.shape
    freqs = np.fft.rfftfreq(n_fft, 1.0 / sr).astype(np.float32)
    res = sr / n_fft
    probe = np.linspace(0, T - 1, min(256, T), dtype=int)
    env_probe = env[:, probe]

    def pack_for(K):
        hz = mel_knots_hz(sr, K)
        b = np.clip(np.round(hz / res).astype(int), 0, n_bins - 1)
        return hz, log_env[b, :]

    chosen = None
    for K in range(K_start, K_max + 1, K_step):
        hz, kv = pack_for(K)
        idx, w0, w1 = knot_lerp_table(freqs, hz)
        W = np.zeros((n_bins, K), dtype=np.float32)
        r = np.arange(n_bins)
        W[r, idx] = w0
        W[r, idx + 1] = w1
        rec = W @ kv[:, probe]
        err = np.max(np.abs(np.exp(rec) - env_probe) / (env_probe + 1e-8))
        if err < eps:
            chosen = (hz, kv)
            break
    if chosen is None:
        chosen = pack_for(K_max)
    hz, kv = chosen
    return {"mode": "knots", "knot_vals_log": kv.astype(np.float16), "hz_knots": hz.astype(np.float32),
            "n_bins": int(n_bins), "n_fft": int(n_fft), "sr": int(sr)}


def analyse_envelope(y, sr: int):
    """Envelope half of extract_features (GOOFER.py:942-946, 968): STFT -> |S|+1e-8 -> Gaussian
    sigma=2 along frequency -> knots.  (f0 / formants come from Praat in the reference: out of scope.)"""
    S = stft(y, N_FFT, HOP, sqrt_hann(N_FFT))
    env = gauss1d(np.abs(S) + 1e-8, 2.0, axis=0)
    return env, compress_to_knots(env, sr)


# ----------------------------------------------------------------------------------------------
# sequential kernels (C)   (GOOFER.py:473-554, 672-698; SillySampler.py:95-174)
# ----------------------------------------------------------------------------------------------
def pulse_train(f0_f32: np.ndarray, sr: int, Ra=0.02, Rg=1.7, Rk=0.8, want_onsets=False):
    """GOOFER.py:473-554."""
    f0 = np.ascontiguousarray(f0_f32, dtype=np.float32)
    n = f0.size
    out = np.empty(n, dtype=np.float32)
    cap = n + 8 if want_onsets else 0
    oi = np.zeros(max(cap, 1), dtype=np.int32)
    ot = np.zeros(max(cap, 1), dtype=np.int32)
    cnt = _seq().orc_pulse_train(_p(f0, ctypes.c_float), n, float(sr), Ra, Rg, Rk, _p(out, ctypes.c_float),
                                 _p(oi, ctypes.c_int32) if want_onsets else None,
                                 _p(ot, ctypes.c_int32) if want_onsets else None, cap)
    if want_onsets:
        return out, oi[:cnt].copy(), ot[:cnt].copy()
    return out


def lf_pulse_f32(T, sr: int, Ra=0.02, Rg=1.7, Rk=1.0) -> np.ndarray:
    """GOOFER.py:437-471 (smoothing=False) as add_subharms calls it: T arrives as an np.float64 scalar
    (GOOFER.py:717, sub_f0 is an element of an fp64 array), which under NEP 50 is a *strong* type, so
    the f32 time axis is promoted: comparisons and the divisions by Tp / (Tc - Tp) run in fp64, only
    ``np.pi * t`` is rounded to f32 first.  Written with explicit dtypes so the result does not depend
    on whether the caller passes a Python float or a numpy scalar."""
    T = np.float64(T)
    n = int(round(sr * T))
    if n <= 3:
        n = 3
    t = np.linspace(0, T, n, endpoint=False, dtype=np.float32)
    Tp = np.float64(Ra) * T
    Tc = Tp + np.float64(Rk) * (T - Tp)
    t64 = t.astype(np.float64)
    p = np.zeros(n, dtype=np.float32)
    rise = t64 < Tp
    if np.any(rise):
        a = (np.float32(np.pi) * t[rise]).astype(np.float64)      # f32 product, then promoted
        p[rise] = np.sin(a / (2 * Tp)) ** 2
    fall = (t64 >= Tp) & (t64 < Tc)
    if np.any(fall):
        tau = (t64[fall] - Tp) / (Tc - Tp)
        p[fall] = np.exp(-Rg * tau) * np.cos(np.pi * tau / 2)
    m = np.max(np.abs(p))
    if m > 0:
        p /= m
    return p


def subharm_layer(f0_f64: np.ndarray, sr: int, weight: float, semitones: float, mask) -> np.ndarray:
    """GOOFER.py:700-736 add_subharms for a single semitone offset (SillySampler passes +12)."""
    f0 = np.ascontiguousarray(f0_f64, dtype=np.float64)
    vm = np.ascontiguousarray(mask, dtype=np.float64)
    n = f0.size
    ratio = 2.0 ** (float(semitones) / 12.0)
    ei = np.zeros(n + 1, dtype=np.int32)
    ef = np.zeros(n + 1, dtype=np.float64)
    ne = _seq().orc_sub_events(_p(f0, ctypes.c_double), _p(vm, ctypes.c_double), n, float(sr), ratio,
                               _p(ei, ctypes.c_int32), _p(ef, ctypes.c_double), n + 1)
    sub = np.zeros(n, dtype=np.float64)
    bank = {}
    for i, sf0 in zip(ei[:ne].tolist(), ef[:ne]):
        key = f"{sf0:.2f}_sub{ratio:.3f}"                 # :718 first occurrence fixes the pulse
        pl = bank.get(key)
        if pl is None:
            pl = lf_pulse_f32(1.0 / sf0, sr, Ra=0.02, Rg=1.7, Rk=1).astype(np.float64)
            bank[key] = pl
        e = min(n, i + len(pl))
        sub[i:e] += pl[:e - i]
    sub *= vm
    mx = np.max(np.abs(sub)) if n else 0.0
    if mx > 1e-6:
        sub /= mx
    sub *= weight
    return sub


def dyn_onepole(signal, f0, sr: int, cutoff_factor, order=4, btype="lowpass") -> np.ndarray:
    """SillySampler.py:95-174 dynamic_butter_filter (f0.size == n on every call site)."""
    x = np.asarray(signal, dtype=np.float32)
    n = len(x)
    if n == 0:
        return x
    f0 = np.asarray(f0, dtype=np.float32)
    if f0.size != n:
        pos = np.linspace(0, n - 1, num=f0.size, dtype=np.float64)
        f0 = lerp_extrap(pos, f0.astype(np.float64), np.arange(n, dtype=np.float64)).astype(np.float32)
    if np.any(f0 > 0):
        f0s = np.convolve(np.pad(f0, (2, 2), mode="edge"), np.ones(5, dtype=np.float32) / 5, mode="valid")
    else:
        f0s = f0
    # :128-152 -- products with the (python float / int) factor are formed in f64, stored to f32
    fc = np.where(f0s > 0.0, f0s.astype(np.float64) * float(cutoff_factor), float(cutoff_factor)).astype(np.float32)
    fc = np.maximum(fc, np.float32(60.0 if btype == "lowpass" else 20.0))
    fc = np.minimum(fc.astype(np.float64), 0.45 * sr).astype(np.float32)
    w = (2.0 * np.pi) * fc.astype(np.float64)
    alpha = (w / (w + sr) if btype == "lowpass" else sr / (w + sr)).astype(np.float32)
    y = np.ascontiguousarray(x.copy())
    alpha = np.ascontiguousarray(alpha)
    _seq().orc_onepole_cascade(_p(y, ctypes.c_float), _p(alpha, ctypes.c_float), n, max(1, int(order)),
                               0 if btype == "lowpass" else 1)
    return y


def rms(x) -> float:
    """GOOFER.py:170-171."""
    return float(np.sqrt(np.mean(np.square(x)) + 1e-12))
