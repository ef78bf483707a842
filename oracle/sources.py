"""ORACLE: deterministic synthetic voicebank sources.  TEST INFRASTRUCTURE ONLY.

The reference obtains f0 / formant tracks from Praat (GOOFER.py:344, 770: third party, not installed,
out of scope) -- for the synthetic sources they are known analytically.  The spectral envelope goes
through the oracle's restatement of the reference's own analysis (STFT -> |S|+1e-8 -> Gaussian sigma 2
-> mel knots, GOOFER.py:942-946, 968, 97-147) and the .goofy fp16 round trip (GOOFER.py:287-339), so
the features carry the same quantisation a real voicebank would.
"""
from __future__ import annotations

import numpy as np

from . import dsp
from .resampler import Features

SR = 44100
VOWELS = {
    "a": (700.0, 1200.0, 2600.0, 3500.0),
    "i": (300.0, 2300.0, 3000.0, 3600.0),
    "u": (320.0, 800.0, 2300.0, 3300.0),
    "e": (500.0, 1900.0, 2600.0, 3500.0),
    "o": (500.0, 900.0, 2500.0, 3400.0),
}
_BW = (80.0, 90.0, 120.0, 150.0)
VOWEL_ORDER = "aiueo"


def vowel_wave(formant_hz, seconds: float, f_src: float = 220.0, seed: int = 0, sr: int = SR) -> np.ndarray:
    """sum_k g(k f)/k sin(2 pi k f t + theta_k), g = sum of resonance magnitudes, peak 0.5."""
    n = int(round(seconds * sr))
    t = np.arange(n) / sr
    rng = np.random.Generator(np.random.PCG64(seed))
    K = int((sr / 2) // f_src)
    th = rng.uniform(0, 2 * np.pi, K)
    y = np.zeros(n)
    for k in range(1, K + 1):
        f = k * f_src
        g = sum((1 + ((f - F) / B) ** 2) ** -0.5 for F, B in zip(formant_hz, _BW))
        y += (g / k) * np.sin(2 * np.pi * f * t + th[k - 1])
    return 0.5 * y / np.max(np.abs(y))


def make_source(index: int, seconds: float = 1.0, fricative: bool | None = None, sr: int = SR):
    """Source ``index`` of the synthetic voicebank: vowel = index mod 5; every 4th source starts
    with 120 ms of first-differenced white noise (unvoiced).  Returns (wave f64, analytic tracks)."""
    v = VOWEL_ORDER[index % 5]
    F = VOWELS[v]
    if fricative is None:
        fricative = (index % 4 == 3)
    y = vowel_wave(F, seconds, 220.0, seed=index, sr=sr)
    n = len(y)
    mask = np.ones(n)
    if fricative:
        m = int(0.120 * sr)
        rng = np.random.Generator(np.random.PCG64(1000 + index))
        w = rng.normal(0.0, 0.05, m + 1)
        y[:m] = np.diff(w)
        mask[:m] = 0.0
    f0 = 220.0 * mask
    return y, {"f0": np.clip(f0, 1e-5, 2000), "mask": mask, "F": F}


def features_from_wave(y: np.ndarray, tracks: dict, sr: int = SR):
    """Analysis + .goofy-equivalent quantisation.  Returns (Features, knot pack)."""
    env, pack = dsp.analyse_envelope(y, sr)
    T = env.shape[1]
    forms = {i + 1: np.full(T, tracks["F"][i], dtype=np.float64) for i in range(4)}
    mask16 = tracks["mask"].astype(np.float16)
    feat = Features(env=dsp.decode_knots(pack), mask=mask16.astype(np.float32), formants=forms,
                    sr=sr, ylen=len(y))
    return feat, pack


_CACHE = {}


def source_features(index: int, seconds: float = 1.0, sr: int = SR):
    key = (index, seconds, sr)
    if key not in _CACHE:
        y, tr = make_source(index, seconds, sr=sr)
        feat, pack = features_from_wave(y, tr, sr)
        _CACHE[key] = (feat, pack, y, tr)
    return _CACHE[key]
