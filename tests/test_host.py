"""CPU: host-side parsing mirrors SillySampler.GooferResampler.__init__ (checked against the oracle's own
restatement, which is pinned to the reference), .goofy loading, sharding helpers, bench workloads."""
import os

import numpy as np
import pytest

import bench_data
from goofer_b200 import capi, host, shard
from oracle import resampler


def test_flag_regex_and_case_rules():
    s = "g-20fa10/fb-10B20Mt50xyz5ES40Pd-30fstA7l2"
    assert host.parse_flags(s) == resampler.parse_flag_string(s)
    nt, _ = host.NoteArgs.from_cli(0, ["C4", "100", s]).to_struct(0)
    fl = dict(zip(capi.FLAG_NAMES, nt.flag))
    present = {n for i, n in enumerate(capi.FLAG_NAMES) if (nt.present >> i) & 1}
    assert fl["g"] == -20 and fl["fa"] == 10 and fl["fb"] == -10 and fl["B"] == 20
    assert fl["es"] == 40 and "es" in present          # case-insensitive lookup (SillySampler.py:384)
    assert fl["pd"] == -30 and fl["L"] == 2 and fl["fsta"] == 7
    assert "t" not in present                           # 'Mt' is its own (unknown) flag
    with pytest.raises(TypeError):
        host.NoteArgs.from_cli(0, ["C4", "100", "g_B5"]).to_struct(0)     # "g" without a number: None / 200.0 in the reference


@pytest.mark.parametrize("s", ["AA", "AAAB#3#ACADAFAIALAOAQASATATASAQAOALAIAFADACAB#20#", "AA///+/9/7/5/3#30#/5/9AA#40#", "", "//#5#"])
def test_pitch_string(s):
    assert np.array_equal(host.pitch_string_to_cents(s), resampler.bend_cents(s))


def test_note_names_and_roundtrip():
    for nm in ("C4", "A3", "G#5", "C#-1", "B7"):
        assert host.note_to_midi(nm) == resampler.midi_of(nm)
    for m in range(24, 100):
        assert host.note_to_midi(bench_data.midi_to_name(m)) == m
    cents = np.array([0, 1, -1, 2047, -2048, 30, -30])
    assert np.array_equal(host.pitch_string_to_cents(bench_data.cents_to_pitch_string(cents)), cents.astype(np.float32))
    with pytest.raises(ValueError):
        host.note_to_midi("H2")


def test_load_goofy_reads_the_reference_format(tmp_path):
    # the layout gf.save_features writes (GOOFER.py:287-317): npz, fp16 knots, pickled formant dict
    src = bench_data.make_source(3)
    path = os.path.join(tmp_path, "x_features.goofy")
    with open(path, "wb") as fh:
        np.savez_compressed(fh, mode=np.array(["knots"]), knot_vals_log=src["knot_vals_log"], hz_knots=src["hz_knots"],
                            n_bins=np.array([513]), n_fft=np.array([1024]), f0_interp=np.zeros(8, np.float16),
                            voicing_mask=src["mask"].astype(np.float16), formants=np.array(src["formants"], dtype=object),
                            sr=np.array([44100]), y_len=np.array([src["ylen"]]))
    sf = host.load_goofy(path)
    assert sf.knots_log.dtype == np.float16 and sf.knots_log.shape == src["knot_vals_log"].shape
    assert sf.sr == 44100 and sf.ylen == src["ylen"] and set(sf.formants) == {1, 2, 3, 4}
    assert np.array_equal(sf.mask, src["mask"])


def test_shard_helpers():
    for n, w in ((10, 3), (1024, 8), (5, 8), (0, 2)):
        got = [i for r in range(w) for i in shard.contiguous_range(n, r, w)]
        assert got == list(range(n))
        sizes = [len(shard.contiguous_range(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1
    costs = [4, 1, 1, 1, 1, 4, 2, 2]
    bins = shard.balanced_partition(costs, 2)
    assert sorted(i for b in bins for i in b) == list(range(8))
    assert abs(sum(costs[i] for i in bins[0]) - sum(costs[i] for i in bins[1])) <= 1
    with pytest.raises(ValueError):
        shard.contiguous_range(4, 2, 2)


def test_bench_workloads_plan(lib):
    b = host.Batch()
    for s in range(8):
        f = bench_data.make_source(s)
        b.add_source(host.SourceFeatures.from_knot_pack(f, f["mask"], f["formants"], f["sr"], f["ylen"]))
    for i in range(16):
        src, cli = bench_data.note_cli(i, "c3", n_sources=8)
        b.add_note(host.NoteArgs.from_cli(src, cli))
    ab = b.assemble(host.SeededNoise())
    assert all(inf["n_total"] == 44100 and inf["t_out"] == 173 for inf in ab.infos)
    assert bench_data.algorithmic_bytes({"need_phi": [1, 0, 0, 0], "need_nrm": [0] * 4, "t_out": 173, "n_total": 44100}, 172, 44100) == 1063492


def test_cli_argument_errors_without_gpu(tmp_path):
    from goofer_b200 import cli
    assert cli.main(["a.wav", "b.wav", "C4"]) == 1                       # TypeError path: usage text, exit code 1
    assert cli.main([os.path.join(tmp_path, "nope.wav"), "b.wav", "C4", "100", "", "0", "1000", "0", "0", "100", "0", "!120", "AA"]) == 1


def test_pcm16_restates_libsndfile_clip_path():
    """cli.pcm16_like_soundfile == pcm.c d2s_clip_array with normalisation (what sf.write does for a .wav,
    SillySampler.py:1185, python-soundfile enabling SFC_SET_CLIPPING): scalar restatement, rounding ties, saturation."""
    import math
    from goofer_b200 import cli

    def scalar(x):
        s = float(x) * (8.0 * 0x10000000)
        if s >= 1.0 * 0x7FFFFFFF:
            return 0x7FFF
        if s <= -8.0 * 0x10000000:
            return -0x8000
        r = math.floor(s)                                  # lrint: nearest, ties to even
        d = s - r
        if d > 0.5 or (d == 0.5 and r % 2 != 0):
            r += 1
        return int(r) >> 16

    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-1.2, 1.2, 4000), rng.uniform(-1e-4, 1e-4, 2000),
                        np.array([0.0, -0.0, 1.0, -1.0, 0.999999, -0.999999, 2.0, -2.0, 1.0 / 32768, -1.0 / 32768,
                                  0.5 / 32768, 1.5 / 32768, (65536 * 3 - 0.5) / 2 ** 31, (65536 * 3 - 0.4) / 2 ** 31,
                                  (65536 * 5 + 0.5) / 2 ** 31, -(65536 * 7 + 0.5) / 2 ** 31])])
    got = cli.pcm16_like_soundfile(x)
    assert got.dtype == np.dtype("<i2")
    assert np.array_equal(got.astype(np.int64), np.array([scalar(v) for v in x]))
    x32 = x.astype(np.float32)
    assert np.array_equal(cli.pcm16_like_soundfile(x32), cli.pcm16_like_soundfile(x32.astype(np.float64)))


def test_openutau_manifest_covers_the_flag_surface():
    """goofer_b200.manifest: every expression maps to a flag the C ABI knows; nothing of the 34-flag surface is lost
    besides g / B / P, which the reference's own manifest leaves to OpenUtau's built-in expressions."""
    import yaml
    from goofer_b200 import capi, manifest
    doc = yaml.safe_load(manifest.render())["expressions"]
    assert len(doc) == 31
    assert manifest.flags() <= set(capi.FLAG_NAMES)
    assert set(capi.FLAG_NAMES) - manifest.flags() == {"g", "B", "P"}
    assert doc["vfhz"]["default_value"] == 50 and doc["vfsl"]["default_value"] == 15 and doc["Hvoi"]["default_value"] == 100
    assert doc["sust"]["options"] == ["L0", "L1", "L2"]


@pytest.mark.reference
def test_openutau_manifest_equals_the_reference_manifest():
    import os
    import yaml
    from goofer_b200 import manifest
    ref = "/root/reference/SillySampler.yaml"
    if not os.path.exists(ref):
        pytest.skip("reference tree absent")
    with open(ref) as fh:
        assert yaml.safe_load(manifest.render()) == yaml.safe_load(fh)


def test_batch_plan_and_cost_partition():
    """host.Batch.plan() (planner only) equals the infos assemble() gets; shard.balanced_partition splits by cost."""
    import bench_data
    from goofer_b200 import shard
    b = host.Batch()
    for s in range(4):
        f = bench_data.make_source(s)
        b.add_source(host.SourceFeatures.from_knot_pack(f, f["mask"], f["formants"], f["sr"], f["ylen"]))
    for j in range(24):
        src, cli = bench_data.note_cli(j, "c3" if j % 3 == 0 else "c2", n_sources=4)
        b.add_note(host.NoteArgs.from_cli(src, cli))
    infos = b.plan()
    assert infos == b.assemble(host.DeviceNoise()).infos
    costs = [shard.note_cost(i) for i in infos]
    parts = shard.balanced_partition(costs, 4)
    loads = [sum(costs[i] for i in p) for p in parts]
    assert sorted(i for p in parts for i in p) == list(range(24))
    assert max(loads) - min(loads) <= max(costs)


def test_continuous_overrides_reach_the_plan():
    """NoteArgs.overrides -> GooferNote.override_val -> the planner's scalars (direct gf.synthesize seam)."""
    import ctypes as C
    from tests.plan_struct import GfNotePlan
    f = bench_data.make_source(1)
    b = host.Batch()
    b.add_source(host.SourceFeatures.from_knot_pack(f, f["mask"], f["formants"], f["sr"], f["ylen"]))
    b.add_note(host.NoteArgs(source=0, pitch="C4", flags="g1fb1sh1sr1", overrides={
        "formant_shift": 1.0123, "F2_shift": 0.937, "f0_jitter_strength": 0.731, "volume_jitter_strength_harm": 0.33,
        "volume_jitter_strength_breath": 0.9, "normalize": 0.6, "breath_strength": 0.05, "uv_strength": 0.4}))
    ab = b.assemble(host.SeededNoise())
    p = GfNotePlan()
    assert capi.load().goofer_debug_plan(C.byref(ab.desc), 0, C.byref(p), C.sizeof(p)) == C.sizeof(p)
    assert p.formant_shift == 1.0123 and p.F_shift[1] == 0.937 and p.F_shift[0] == 1.0 and p.any_F_shift == 1
    assert p.f0_jitter == 1 and p.f0_jitter_strength == 0.731 and p.vol_jitter_strength == 0.33 and p.vol_jitter_strength_breath == 0.9
    assert p.normalize == 0.6 and abs(p.breath_strength - 0.05) < 1e-8 and abs(p.uv_strength - 0.4) < 1e-7
