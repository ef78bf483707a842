"""CPU: the oracle (numpy + C restatement) against the committed outputs of the UNMODIFIED reference
(tests/golden/*.npz, written by tests/golden/make_golden.py in the build container)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import dsp
from tests import cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold_cases():
    return np.load(os.path.join(GOLD, "render_cases.npz"))


@pytest.fixture(scope="module")
def gold_stages():
    return np.load(os.path.join(GOLD, "stages.npz"))


@pytest.mark.parametrize("case", cases.CASES, ids=[c[0] for c in cases.CASES])
def test_render_case_matches_reference(case, gold_cases):
    name, si, secs, cli = case
    feat, _ = cases.source_for(si, secs)
    got = cases.oracle_render(feat, cli)
    n = int(gold_cases[f"n_{name}"][0])
    assert len(got) == n
    ref32 = gold_cases[f"out_{name}"]
    cmp = got if n <= 50000 else got[::int(gold_cases["long_stride"][0])]
    # the fixture is the reference's float64 output rounded to float32 (|out| <= 2.4 -> 1.4e-7);
    # bit-identical float64 in the build container (sha below), 1e-6 leaves room for another libm / BLAS
    assert np.max(np.abs(cmp - ref32.astype(np.float64))) <= 1e-6
    same = hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == str(gold_cases[f"sha_{name}"][0])
    if not same:
        print(f"note: {name} is within 1e-6 of the reference but not bit-identical on this machine")


def test_stft_istft(gold_stages):
    g = gold_stages
    S = dsp.stft(g["stft_x"])
    assert S.dtype == np.complex64 and S.shape == g["stft_S"].shape
    assert np.max(np.abs(S - g["stft_S"])) <= 1e-6 * np.max(np.abs(g["stft_S"]))
    y = dsp.istft(g["istft_S"], length=4300)
    assert np.max(np.abs(y - g["istft_y"])) <= 1e-6 * np.max(np.abs(g["istft_y"]))
    assert np.all(y[256 * (g["istft_S"].shape[1] - 1):] == 0.0)          # zero tail of GOOFER.py:407-409


def test_pulse_onsets_on_tie_pitches(gold_stages):
    g = gold_stages
    for k, hz in enumerate(g["tie_pitches"]):
        f0 = np.full(2 * 44100, hz, dtype=np.float32)
        f0[:3000] = 0.0
        out, oi, ot = dsp.pulse_train(f0, 44100, want_onsets=True)
        assert np.array_equal(oi, g[f"pulse_onsets_{k}"]), f"onsets moved at {hz} Hz"
        assert np.array_equal(ot, g[f"pulse_T0_{k}"])
        assert np.max(np.abs(out[:8192] - g[f"pulse_head_{k}"])) <= 1e-6
    f0 = np.linspace(90.0, 700.0, 44100).astype(np.float32)
    out, oi, _ = dsp.pulse_train(f0, 44100, want_onsets=True)
    assert np.array_equal(oi, g["glide_onsets"])
    assert np.max(np.abs(out - g["glide_pulse"])) <= 1e-5


def test_onepole_gauss_knots(gold_stages):
    g = gold_stages
    lp = dsp.dyn_onepole(g["op_x"], g["op_f0"], 44100, 1.4, order=3, btype="lowpass")
    hp = dsp.dyn_onepole(g["op_x"], g["op_f0"], 44100, 1.0, order=6, btype="highpass")
    assert np.max(np.abs(lp - g["op_lp3"])) <= 1e-6 and np.max(np.abs(hp - g["op_hp6"])) <= 1e-6
    for sig in (0.5, 1.75, 25.0):
        assert np.max(np.abs(dsp.gauss1d(g["gauss_x"], sig) - g[f"gauss_{sig}"])) <= 1e-12
    env = dsp.decode_knots({"knot_vals_log": g["knots_log"], "hz_knots": g["hz_knots"], "n_fft": 1024, "sr": 44100, "n_bins": 513})
    ref = g["knots_env_cols"]
    assert np.max(np.abs(env[:, ::16] - ref) / ref) <= 1e-5


def test_analysis_front_end(gold_stages):
    """Envelope half of gf.extract_features + gf.compress_env_to_knots (GOOFER.py:940-946, 97-147)."""
    from oracle import sources
    g = gold_stages
    for tag in ("an_a", "an_b", "an_silence"):
        idx = int(g[f"{tag}_src"][0])
        y = np.zeros(20000, np.float32) if idx < 0 else sources.make_source(idx, 1.0)[0].astype(np.float32)
        env, pack = dsp.analyse_envelope(y, 44100)
        assert pack["knot_vals_log"].shape[0] == int(g[f"{tag}_K"][0])
        assert np.array_equal(pack["hz_knots"], g[f"{tag}_hz"])
        d = np.abs(pack["knot_vals_log"].astype(np.float32) - g[f"{tag}_knots"].astype(np.float32))
        assert np.max(d) <= 2e-3 * np.max(np.abs(g[f"{tag}_knots"].astype(np.float32)))     # one f16 ulp
    assert int(g["an_silence_K"][0]) == 32


def _synth_direct_inputs():
    from tests.golden import make_golden_synth as mg
    feat, pack, forms, _, _ = mg.inputs()
    g = np.load(os.path.join(GOLD, "synth_direct.npz"))
    return mg, feat, pack, forms, g


def synth_direct_noise(n, T, base, legacy, f0_jitter, volume_jitter):
    """The buffers gf.synthesize draws under ref_harness.seeded_noise, in its own order (GOOFER.py:666, 653, 1151)."""
    leg = np.random.RandomState(legacy)
    nz = {}
    if f0_jitter:
        nz["sh"] = leg.randn(n)
    if volume_jitter:
        nz["sr_h"] = leg.randn(n)
        nz["sr_b"] = leg.randn(n)
    nz["phi"] = np.random.Generator(np.random.PCG64(base)).uniform(0.0, 2.0 * np.pi, size=(513, T)).astype(np.float32)
    return nz


def test_oracle_direct_synthesize_matches_reference():
    """gf.synthesize called directly (SillyEditor.py:227, test.py:38): oracle.synth against the reference's outputs."""
    from oracle import synth
    mg, feat, pack, forms, g = _synth_direct_inputs()
    n, sr = len(feat.mask), feat.sr
    T = 1 + n // 256
    base, legacy = (int(x) for x in g["seeds"])
    knots = {"knot_vals_log": pack["knot_vals_log"], "hz_knots": pack["hz_knots"], "n_fft": 1024, "sr": sr, "n_bins": 513}
    ra = synth.synthesize(knots, g["f0_a"], feat.mask, n, sr, synth_direct_noise(n, T, base, legacy, False, False), formants=forms)
    kw = dict(mg.KW_B)
    shifts = tuple(kw.pop(f"F{i}_shift") for i in (1, 2, 3, 4))
    rb = synth.synthesize(feat.env, g["f0_b"], feat.mask, n, sr, synth_direct_noise(n, T, base, legacy, True, True),
                          formants=forms, F_shifts=shifts, **kw)
    kc = dict(mg.KW_C)
    shifts_c = (1.0, kc.pop("F2_shift"), 1.0, 1.0)
    f0_c = g["f0_a"] * np.float32(kc.pop("pitch_shift"))                  # GOOFER.py:995 on the to_compute'd f32 array
    rc = synth.synthesize(knots, f0_c, feat.mask, n, sr, synth_direct_noise(n, T, base, legacy, True, True),
                          formants=forms, F_shifts=shifts_c, **kc)
    for tag, r in (("a", ra), ("b", rb), ("c", rc)):
        for name, arr in zip(("reconstruct", "harmonic", "aper_uv", "aper_bre"), r):
            assert np.array_equal(np.asarray(arr, dtype=np.float32), g[f"{tag}_{name}"]), (tag, name)
