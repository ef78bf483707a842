"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, build container only).

    python tests/golden/make_golden.py

What is stored (the reference ships no tests or golden vectors of its own -- SURVEY.md section 4 -- so these
are outputs of the reference itself, run under oracle/ref_harness.py's deterministic noise recipe):

  render_cases.npz   per CLI case of oracle/validate_against_reference.CASES: the float64 array the reference
                     hands to sf.write, stored as float32 (1 s cases in full, long cases every 7th sample) and
                     the sha256 of the float64 bytes; for the cases in TAP_CASES also the (harmonic, aper_uv,
                     aper_bre) tuple gf.synthesize returned for the main pass.
  stages.npz         gf.stft / gf.istft on a seeded signal, gf.pulse_train_numba on rounding-tie pitches
                     (onset lists; sha256 of the pulse arrays), dynamic_butter_filter on seeded noise,
                     gf.gaussian_filter1d, gf.decode_env_from_knots.

The file also records numpy / numba / scipy versions: the reference's float32 FFT path exists only on
numpy >= 2.
"""
from __future__ import annotations

import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import dsp, ref_harness, sources                                # noqa: E402
from oracle.validate_against_reference import CASES                        # noqa: E402

SEED_BASE, SEED_LEGACY = 20000, 777
TAP_CASES = ("formant",)
LONG_STRIDE = 7
TIE_PITCHES = (110.0, 220.0, 440.0, 880.0, 50.0, 261.6255653005986, 97.99885899543733)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def versions():
    import numba
    import scipy
    return np.array([f"numpy {np.__version__}", f"numba {numba.__version__}", f"scipy {scipy.__version__}"])


def render_cases():
    out = {"versions": versions(), "case_names": np.array([c[0] for c in CASES]), "long_stride": np.array([LONG_STRIDE])}
    with tempfile.TemporaryDirectory() as tmp:
        for name, si, secs, cli in CASES:
            feat, pack, y, tr = sources.source_features(si, secs)
            wav = os.path.join(tmp, f"src{si}_{int(secs * 1000)}.wav")
            goofy = wav[:-4] + "_features.goofy"
            if not os.path.exists(goofy):
                forms = {i + 1: np.full(feat.env.shape[1], tr["F"][i]) for i in range(4)}
                ref_harness.write_goofy(goofy, pack, tr["f0"], tr["mask"], forms, feat.sr, len(y))
            ref, sr, cap = ref_harness.render_note(goofy, [wav, os.path.join(tmp, name + ".wav")] + cli, SEED_BASE, SEED_LEGACY,
                                                   taps=name in TAP_CASES)
            out[f"sha_{name}"] = np.array([sha(ref)])
            out[f"n_{name}"] = np.array([len(ref)])
            out[f"out_{name}"] = (ref if len(ref) <= 50000 else ref[::LONG_STRIDE]).astype(np.float32)
            if name in TAP_CASES:
                rec = cap["synth"][0]
                for key, arr in zip(("harm", "uv", "bre"), rec["out"][1:]):
                    out[f"tap_{key}_{name}"] = np.asarray(arr, dtype=np.float32)
            print(f"{name:16s} n={len(ref)} sha={out[f'sha_{name}'][0][:12]}")
    np.savez_compressed(os.path.join(HERE, "render_cases.npz"), **out)


def stages():
    gf, ss = ref_harness.load_reference()
    out = {"versions": versions()}
    rng = np.random.Generator(np.random.PCG64(1234))
    # --- stft / istft (GOOFER.py:355-413) ---
    x = (0.3 * rng.standard_normal(4096)).astype(np.float32)
    win = gf.get_cached_window(44100, 1024)
    S = gf.stft(x, 1024, 256, win)
    out["stft_x"], out["stft_S"] = x, S.astype(np.complex64)
    S2 = (rng.standard_normal(S.shape) + 1j * rng.standard_normal(S.shape)).astype(np.complex64)
    out["istft_S"] = S2
    out["istft_y"] = gf.istft(S2, 256, win, length=4300).astype(np.float32)
    # --- pulse train on rounding-tie pitches (GOOFER.py:473-554) ---
    out["tie_pitches"] = np.array(TIE_PITCHES)
    for k, hz in enumerate(TIE_PITCHES):
        f0 = np.full(2 * 44100, hz, dtype=np.float32)
        f0[:3000] = 0.0                                     # leading unvoiced stretch: last_valid_f0 = 160 default
        ref = gf.pulse_train_numba(f0, 44100, 0.02, 1.7, 0.8)
        orc, oi, ot = dsp.pulse_train(f0, 44100, want_onsets=True)
        assert np.array_equal(ref, orc), f"oracle pulse train differs from the reference at {hz} Hz"
        out[f"pulse_sha_{k}"] = np.array([sha(ref)])
        out[f"pulse_onsets_{k}"] = oi.astype(np.int32)
        out[f"pulse_T0_{k}"] = ot.astype(np.int32)
        out[f"pulse_head_{k}"] = ref[:8192].astype(np.float32)
    # glide: f0 sweeping through several T0 values (exercises the 5-slot cache and overlapping pulses)
    f0 = np.linspace(90.0, 700.0, 44100).astype(np.float32)
    ref = gf.pulse_train_numba(f0, 44100, 0.02, 1.7, 0.8)
    orc, oi, ot = dsp.pulse_train(f0, 44100, want_onsets=True)
    out["glide_max_diff"] = np.array([float(np.max(np.abs(ref - orc)))])
    out["glide_onsets"] = oi.astype(np.int32)
    out["glide_pulse"] = ref.astype(np.float32)
    # --- dynamic one-pole cascades (SillySampler.py:95-174) ---
    xs = (0.2 * rng.standard_normal(8192)).astype(np.float32)
    f0r = np.concatenate([np.zeros(1000), np.linspace(120, 600, 7192)]).astype(np.float32)
    out["op_x"], out["op_f0"] = xs, f0r
    out["op_lp3"] = ss.dynamic_butter_filter(xs, f0r, 44100, 1.4, order=3, btype="lowpass").astype(np.float32)
    out["op_hp6"] = ss.dynamic_butter_filter(xs, f0r, 44100, 1.0, order=6, btype="highpass").astype(np.float32)
    # --- gaussian_filter1d (GOOFER.py:241-261) ---
    g = rng.standard_normal(777)
    out["gauss_x"] = g
    for sig in (0.5, 1.75, 25.0):
        out[f"gauss_{sig}"] = gf.gaussian_filter1d(g, sig)
    # --- knot decode (GOOFER.py:149-168) ---
    feat, pack, y, tr = sources.source_features(2, 1.0)
    env = gf.decode_env_from_knots(pack)
    out["knots_log"], out["hz_knots"] = pack["knot_vals_log"], pack["hz_knots"]
    out["knots_env_cols"] = np.asarray(env, dtype=np.float32)[:, ::16]
    # --- analysis front-end: the envelope half of extract_features + compress_env_to_knots (GOOFER.py:940-946, 97-147) ---
    for tag, idx in (("an_a", 2), ("an_b", 3), ("an_silence", -1)):
        yv = np.zeros(20000, dtype=np.float32) if idx < 0 else sources.make_source(idx, 1.0)[0].astype(np.float32)
        S0 = gf.stft(yv, 1024, 256, win)
        env_spec = gf.gaussian_filter1d(np.abs(S0) + 1e-8, sigma=2.0, axis=0)
        packr = gf.compress_env_to_knots(env_spec, sr=44100, n_fft=1024, eps=1e-2, K_start=32, K_step=16, K_max=192)
        out[f"{tag}_src"] = np.array([idx])
        out[f"{tag}_K"] = np.array([packr["knot_vals_log"].shape[0]])
        out[f"{tag}_hz"] = packr["hz_knots"]
        out[f"{tag}_knots"] = packr["knot_vals_log"]
        env_o, pack_o = dsp.analyse_envelope(yv, 44100)
        assert pack_o["knot_vals_log"].shape == packr["knot_vals_log"].shape, "oracle chose a different K than the reference"
        assert np.array_equal(pack_o["hz_knots"], packr["hz_knots"]) and np.array_equal(pack_o["knot_vals_log"], packr["knot_vals_log"])
    np.savez_compressed(os.path.join(HERE, "stages.npz"), **out)
    print("stages written")


if __name__ == "__main__":
    assert ref_harness.reference_available(), "needs /root/reference"
    render_cases()
    stages()
    for f in ("render_cases.npz", "stages.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
