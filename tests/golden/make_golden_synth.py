"""Generates tests/golden/synth_direct.npz by calling the UNMODIFIED reference's gf.synthesize directly, the way
SillyEditor.py:227, 559 and test.py:38 do (build container only; /root/reference must exist).

    python tests/golden/make_golden_synth.py

Three calls on synthetic source 2 (1 s): (a) default keyword arguments with a 5.5 Hz vibrato f0 curve and a knots dict,
(b) formant_shift / F-shifts / f0_jitter / volume_jitter / normalize on the dense envelope, (c) continuous keyword
values that no integer flag reaches (test.py:38: formant_shift, breath_strength, uv_strength; pitch_shift).  Noise is drawn under
oracle/ref_harness.seeded_noise; stored: the inputs that are not reproducible from the seed alone (f0 curves) and the
four returned arrays as float32.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import dsp, ref_harness, sources                                # noqa: E402

SEED_BASE, SEED_LEGACY = 31000, 911
SRC, SECS = 2, 1.0
KW_B = dict(formant_shift=1.1, F1_shift=1.05, F2_shift=0.95, F3_shift=1.02, F4_shift=1.0, f0_jitter=True,
            f0_jitter_strength=0.6, volume_jitter=True, volume_jitter_strength_harm=0.8, volume_jitter_strength_breath=1.6,
            normalize=0.5)
# (c) continuous values no integer flag can express: test.py:38's keyword arguments (formant_shift, breath_strength,
# uv_strength) plus pitch_shift, one F-shift, fractional jitter strengths and a partial normalisation
KW_C = dict(formant_shift=1.0123, F2_shift=0.937, pitch_shift=1.03, normalize=0.6, breath_strength=0.05, uv_strength=0.4,
            f0_jitter=True, f0_jitter_strength=0.731, volume_jitter=True, volume_jitter_strength_harm=0.33,
            volume_jitter_strength_breath=0.9)


def inputs():
    feat, pack, y, tr = sources.source_features(SRC, SECS)
    n = len(feat.mask)
    t = np.arange(n) / feat.sr
    f0_a = (196.0 * 2.0 ** (0.3 * np.sin(2 * np.pi * 5.5 * t) / 12.0) * feat.mask).astype(np.float32)
    f0_b = (np.linspace(150.0, 330.0, n) * feat.mask).astype(np.float32)
    forms = {i + 1: np.full(feat.env.shape[1], tr["F"][i], dtype=np.float64) for i in range(4)}
    return feat, pack, forms, f0_a, f0_b


def main():
    gf, _ = ref_harness.load_reference()
    feat, pack, forms, f0_a, f0_b = inputs()
    n = len(feat.mask)
    knots = {"mode": "knots", "knot_vals_log": pack["knot_vals_log"], "hz_knots": pack["hz_knots"], "n_fft": 1024,
             "sr": feat.sr, "n_bins": 513}
    out = {"f0_a": f0_a, "f0_b": f0_b, "seeds": np.array([SEED_BASE, SEED_LEGACY])}
    with ref_harness.seeded_noise(SEED_BASE, SEED_LEGACY):
        ra = gf.synthesize(knots, f0_a.copy(), feat.mask.copy(), np.empty(n, dtype=np.bool_), feat.sr, formants=forms)
    with ref_harness.seeded_noise(SEED_BASE, SEED_LEGACY):
        rb = gf.synthesize(feat.env.copy(), f0_b.copy(), feat.mask.copy(), np.empty(n, dtype=np.bool_), feat.sr,
                           formants=forms, **KW_B)
    with ref_harness.seeded_noise(SEED_BASE, SEED_LEGACY):
        rc = gf.synthesize(knots, f0_a.copy(), feat.mask.copy(), np.empty(n, dtype=np.bool_), feat.sr, formants=forms, **KW_C)
    for tag, r in (("a", ra), ("b", rb), ("c", rc)):
        for name, arr in zip(("reconstruct", "harmonic", "aper_uv", "aper_bre"), r):
            out[f"{tag}_{name}"] = np.asarray(arr, dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "synth_direct.npz"), **out)
    print("wrote synth_direct.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
