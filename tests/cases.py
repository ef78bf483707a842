"""Shared note lists for the parity tests (the same CLI strings the oracle was pinned on against the
unmodified reference: oracle/validate_against_reference.py)."""
from __future__ import annotations

import numpy as np

from oracle import resampler, sources
from oracle.validate_against_reference import CASES  # noqa: F401  (name, source idx, seconds, cli)
from goofer_b200 import host

SEED_BASE, SEED_LEGACY = 20000, 777


def source_for(idx: int, seconds: float):
    """(oracle Features, goofer_b200 SourceFeatures) of synthetic source idx."""
    feat, pack, y, tr = sources.source_features(idx, seconds)
    forms = {i + 1: np.full(feat.env.shape[1], tr["F"][i], dtype=np.float64) for i in range(4)}
    return feat, host.SourceFeatures.from_knot_pack(pack, feat.mask, forms, feat.sr, feat.ylen)


def oracle_render(feat, cli, taps=None):
    spec = resampler.NoteSpec.from_cli(*cli)
    return resampler.resample(feat, spec, lambda n, T: resampler.noise_for_note(spec, n, T, SEED_BASE, SEED_LEGACY), taps=taps)


def lsd_db(ref: np.ndarray, got: np.ndarray, floor_db: float = -100.0) -> float:
    """Log-spectral distance on the reference STFT grid (1024 / 256, sqrt-Hann) with a floor at
    -100 dB re the reference peak."""
    from oracle import dsp
    A = np.abs(dsp.stft(ref.astype(np.float32)))
    B = np.abs(dsp.stft(got.astype(np.float32)))
    fl = max(A.max(), 1e-30) * 10 ** (floor_db / 20)
    la = 20 * np.log10(np.maximum(A, fl))
    lb = 20 * np.log10(np.maximum(B, fl))
    return float(np.sqrt(np.mean((la - lb) ** 2)))
