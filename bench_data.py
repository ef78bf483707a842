"""Synthetic workloads of BASELINE.json for bench.py -- numpy only, no oracle import.

A voicebank source is generated directly in the reference's cached-feature form (what gf.load_features
returns, GOOFER.py:319-339): fp16 log-envelope knots on the mel grid of GOOFER.py:77-82, an fp16-quantised
voicing mask and four constant formant tracks.  The envelope is the analytic resonance curve of the vowel
(five formant sets) with a slow per-frame modulation so that no two frames are equal.
Every 4th source starts with 120 ms of unvoiced fricative (flat high-passed envelope, mask 0).
"""
from __future__ import annotations

import numpy as np

SR = 44100
N_FFT = 1024
HOP = 256
N_BINS = 513
VOWELS = [
    (700.0, 1200.0, 2600.0, 3500.0), (300.0, 2300.0, 3000.0, 3600.0), (320.0, 800.0, 2300.0, 3300.0),
    (500.0, 1900.0, 2600.0, 3500.0), (500.0, 900.0, 2500.0, 3400.0),
]
_BW = (80.0, 90.0, 120.0, 150.0)
_B64 = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/"
_NOTE_NAMES = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]


def mel_knots_hz(sr: int, K: int) -> np.ndarray:
    """Same grid as GOOFER.py:77-82 (f32)."""
    mel_max = 2595.0 * np.log10(1.0 + (sr / 2.0) / 700.0)
    mel = np.linspace(0.0, mel_max, K, dtype=np.float32)
    return (700.0 * (10 ** (mel / 2595.0) - 1.0)).astype(np.float32)


def make_source(index: int, seconds: float = 1.0, K: int = 192, sr: int = SR) -> dict:
    """Cached features of synthetic source `index` (dict of numpy arrays, see module docstring)."""
    F = VOWELS[index % 5]
    n = int(round(seconds * sr))
    T = 1 + n // HOP
    hz = mel_knots_hz(sr, K).astype(np.float64)
    rng = np.random.Generator(np.random.PCG64(5000 + index))
    res = sum((1.0 + ((hz[:, None] - f) / b) ** 2) ** -0.5 for f, b in zip(F, _BW))         # (K, 1)
    roll = 1.0 / (1.0 + hz[:, None] / 220.0)                                                # harmonic 1/k roll-off
    t = np.arange(T)[None, :] / max(1, T - 1)
    wob = 1.0 + 0.15 * np.sin(2 * np.pi * (1.5 * t + rng.uniform(0, 1, (K, 1)))) * rng.uniform(0.2, 1.0, (K, 1))
    env = 40.0 * res * roll * wob + 1e-4
    mask = np.ones(n, dtype=np.float32)
    if index % 4 == 3:
        m = int(0.120 * sr)
        fr = m // HOP
        fric = 0.02 * (hz[:, None] / (sr / 2.0)) + 1e-4
        env[:, :fr] = fric * (1.0 + 0.1 * rng.standard_normal((K, fr)) ** 2)
        mask[:m] = 0.0
    return {
        "knot_vals_log": np.log(env).astype(np.float16), "hz_knots": hz.astype(np.float32),
        "mask": mask.astype(np.float16).astype(np.float32),
        "formants": {k + 1: np.full(T, F[k], dtype=np.float64) for k in range(4)},
        "sr": sr, "ylen": n, "n_bins": N_BINS, "n_fft": N_FFT,
    }


def midi_to_name(m: int) -> str:
    return f"{_NOTE_NAMES[m % 12]}{m // 12 - 1}"


def cents_to_pitch_string(cents) -> str:
    """Inverse of pitch_string_to_cents (SillySampler.py:56-84) without run-length packing."""
    out = []
    for c in cents:
        v = int(c) & 0xFFF
        out.append(_B64[v >> 6] + _B64[v & 63])
    return "".join(out)


def _vibrato_string(i: int, seconds: float, tempo: float = 120.0) -> str:
    ticks = int(np.ceil(seconds / (60.0 / (tempo * 96.0)))) + 1
    tt = np.arange(ticks) * (60.0 / (tempo * 96.0))
    ph = np.random.Generator(np.random.PCG64(30000 + i)).uniform(0, 2 * np.pi)
    return cents_to_pitch_string(np.round(30.0 * np.sin(2 * np.pi * 5.5 * tt + ph)).astype(int))


def formant_flags(i: int) -> str:
    """c2: g, fa-fd, fw, fst, br, es ~ integer U[-100, 100] (fa-fd in [-9, 9]) from PCG64(10000 + i)."""
    r = np.random.Generator(np.random.PCG64(10000 + i))
    v = r.integers(-100, 101, size=9)
    f = r.integers(-9, 10, size=4)
    return (f"g{v[0]}fa{f[0]}fb{f[1]}fc{f[2]}fd{f[3]}fw{v[1]}fst{v[2]}br{v[3]}es{v[4]}")


def full_flags(i: int) -> str:
    """c3: c2 plus B/U/V, sh/sr/sg/sd/sj/sa/su, vf/vh/vl, st, pd."""
    r = np.random.Generator(np.random.PCG64(40000 + i))
    B, U = r.integers(-100, 101, size=2)
    V = r.integers(0, 101)
    s = r.integers(0, 101, size=7)
    vf = r.integers(-100, 101)
    vh = r.integers(20, 101)
    vl = r.integers(0, 101)
    st, pd = r.integers(-100, 101, size=2)
    return (formant_flags(i) + f"B{B}U{U}V{V}sh{s[0]}sr{s[1]}sg{s[2]}sd{s[3]}sj{s[4]}sa{s[5]}su{s[6]}"
            f"vf{vf}vh{vh}vl{vl}st{st}pd{pd}")


def note_cli(i: int, workload: str, n_sources: int = 64):
    """(source index, the 11 CLI strings after the two wav paths) of note i."""
    pitch = midi_to_name(45 + (i % 37))
    bend = "AA" if i % 2 == 0 else _vibrato_string(i, 1.0)
    if workload == "c1":
        flags = ""
    elif workload in ("c2", "c5"):
        flags = formant_flags(i) if (workload == "c2" or i % 2) else ""
    elif workload == "c3":
        flags = full_flags(i)
    elif workload == "c4":
        # long-note sustain: 4 s sources stretched to 16 s, loop modes L0 / L1 / L2 x R0 / R1
        flags = f"L{i % 3}" + ("R1" if (i // 3) % 2 else "")
        return i % n_sources, [pitch, "100", flags, "30", "16000", "150", "200", "100", "0", "!120", "AA" if i % 2 == 0 else _vibrato_string(i, 16.2)]
    else:
        raise ValueError(workload)
    return i % n_sources, [pitch, "100", flags, "0", "1000", "0", "0", "100", "0", "!120", bend]


SOURCE_SECONDS = {"c4": 4.0}

WORKLOADS = {
    "c1": "SillySampler single-note CLI resample: 1 s synthetic 44.1 kHz vowel, default flags",
    "c2": "batch of 1,024 synthetic 1 s notes, formant flags g/fa-fd/fw/fst/br/es",
    "c3": "full-flag stress batch: B/U/V mix + sh/sr/sg/sd/sj/sa/su + vf vocal fry, fixed-seed noise",
    "c4": "long-note sustain: 4 s source stretched to 16 s with L0/L1/L2 loop modes and R1 reverse",
    "c5": "whole-voicebank sweep: default+formant flags, notes sharded across ranks",
}


def algorithmic_bytes(info: dict, T_in: int, N_in: int) -> int:
    """compulsory HBM bytes of one note with inputs in resample()'s layouts."""
    C = sum(info["need_phi"])
    R = sum(info["need_nrm"])
    return (4 * N_BINS * T_in + 4 * N_in + 16 * T_in + 4 * N_BINS * info["t_out"] * C + 8 * info["n_total"] * R
            + 4 * info["n_total"])
