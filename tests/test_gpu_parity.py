"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle on identical inputs and identical
host-supplied noise buffers.  Tolerances are BASELINE.json's: max-abs sample error <= 1e-4 of full scale
(full scale = 1.0) and log-spectral distance <= 0.05 dB (floor -100 dB re the reference peak)."""
import os

import numpy as np
import pytest

import bench_data
from goofer_b200 import capi, host, ops
from oracle import dsp, resampler
from tests import cases

pytestmark = pytest.mark.gpu
MAX_ABS, MAX_LSD = 1e-4, 0.05
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def torch_cuda(lib):
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device; goofer_b200 has no CPU fallback"
    return torch


def render_one(sf, cli, taps=False, device=True):
    b = host.Batch()
    b.add_source(sf)
    b.add_note(host.NoteArgs.from_cli(0, cli))
    ab = b.assemble(host.SeededNoise(cases.SEED_BASE, cases.SEED_LEGACY), taps=taps)
    if not device:
        return ab.render_host()
    db = ab.to_device("cuda:0")
    db.render()
    outs = db.outputs()
    if taps:
        return outs, [ab.split(t[:ab.out_total].cpu().numpy()) for t in db.tap]
    return outs


# ---- stage level ------------------------------------------------------------------------------------------
def test_stft_istft_stage(torch_cuda):
    torch = torch_cuda
    g = np.load(os.path.join(GOLD, "stages.npz"))
    rng = np.random.default_rng(3)
    for n in (4096, 44100, 300, 257):
        x = (0.3 * rng.standard_normal((3, n))).astype(np.float32)
        if n == 4096:
            x[0] = g["stft_x"]
        S = ops.stft(torch.from_numpy(x).cuda()).cpu().numpy()
        for k in range(3):
            ref = dsp.stft(x[k])
            assert S[k].shape == ref.shape
            assert np.max(np.abs(S[k] - ref)) <= 2e-6 * np.max(np.abs(ref))
        if n == 4096:
            assert np.max(np.abs(S[0] - g["stft_S"])) <= 2e-6 * np.max(np.abs(g["stft_S"]))     # reference's own output
        T = S.shape[2]
        Sr = (rng.standard_normal((2, 513, T)) + 1j * rng.standard_normal((2, 513, T))).astype(np.complex64)
        for length in (256 * (T - 1), 256 * (T - 1) + 44, n):
            if length < 256 * (T - 1):
                continue
            y = ops.istft(torch.from_numpy(Sr).cuda(), length).cpu().numpy()
            for k in range(2):
                ref = dsp.istft(Sr[k], length=length)
                assert np.max(np.abs(y[k] - ref)) <= 2e-6 * max(1.0, np.max(np.abs(ref)))
    y = ops.istft(torch.from_numpy(g["istft_S"]).cuda(), 4300).cpu().numpy()
    assert np.max(np.abs(y - g["istft_y"])) <= 2e-6 * np.max(np.abs(g["istft_y"]))
    # round trip: istft(stft(x)) == x away from the edges
    x = (0.3 * rng.standard_normal(20000)).astype(np.float32)
    xr = ops.istft(ops.stft(torch.from_numpy(x).cuda()), 20000).cpu().numpy()
    assert np.max(np.abs(xr[:19900] - x[:19900])) <= 5e-6


def test_pulse_train_onsets_bit_exact(torch_cuda):
    """Rounding-tie pitches: a single moved onset shows up as an O(1) error."""
    torch = torch_cuda
    g = np.load(os.path.join(GOLD, "stages.npz"))
    rows = []
    for hz in g["tie_pitches"]:
        f0 = np.full(2 * 44100, hz, dtype=np.float32)
        f0[:3000] = 0.0
        rows.append(f0)
    rows.append(np.concatenate([np.linspace(90.0, 700.0, 44100), np.linspace(700.0, 60.0, 44100)]).astype(np.float32))
    # the phase walk is a scan of integer parity maps (tests/test_walk_scan_cpu.py): gaps, tiny and negative
    # increments, a negative running total, increments larger than the total
    rng = np.random.default_rng(9)
    rows.append((330 * (rng.random(88200) > 0.3)).astype(np.float32))
    rows.append(np.concatenate([1e-5 * rng.random(30000), np.full(58200, 523.25)]).astype(np.float32))
    rows.append((220 * (1 + 1.5 * np.sin(np.arange(88200) / 50.0))).astype(np.float32))
    rows.append(np.concatenate([np.full(500, -80.0), np.full(87700, 300.0)]).astype(np.float32))
    rows.append(np.concatenate([np.full(40000, 0.001), np.full(48200, 1200.0)]).astype(np.float32))
    f0 = np.stack(rows)
    got = ops.pulse_train(torch.from_numpy(f0).cuda(), 44100).cpu().numpy()
    for k in range(f0.shape[0]):
        ref = dsp.pulse_train(f0[k], 44100)
        assert np.max(np.abs(got[k] - ref)) <= 2e-6, f"row {k}"
    for k in range(len(g["tie_pitches"])):
        assert np.max(np.abs(got[k][:8192] - g[f"pulse_head_{k}"])) <= 2e-6


def test_onepole_stage(torch_cuda):
    torch = torch_cuda
    g = np.load(os.path.join(GOLD, "stages.npz"))
    x, f0 = torch.from_numpy(g["op_x"]).cuda(), torch.from_numpy(g["op_f0"]).cuda()
    lp = ops.onepole(x, f0, 44100, 1.4, order=3, btype="lowpass").cpu().numpy()
    hp = ops.onepole(x, f0, 44100, 1.0, order=6, btype="highpass").cpu().numpy()
    assert np.max(np.abs(lp - g["op_lp3"])) <= 2e-6 and np.max(np.abs(hp - g["op_hp6"])) <= 2e-6
    rng = np.random.default_rng(5)
    xs = (0.2 * rng.standard_normal((4, 30000))).astype(np.float32)
    fs = np.abs(rng.standard_normal((4, 30000)) * 200 + 300).astype(np.float32)
    fs[1, :5000] = 0
    for order, bt, cf in ((1, "lowpass", 2.0), (5, "lowpass", 1.25), (4, "highpass", 0.6), (6, "highpass", 4.0)):
        got = ops.onepole(torch.from_numpy(xs).cuda(), torch.from_numpy(fs).cuda(), 44100, cf, order=order, btype=bt).cpu().numpy()
        for k in range(4):
            ref = dsp.dyn_onepole(xs[k], fs[k], 44100, cf, order=order, btype=bt)
            assert np.max(np.abs(got[k] - ref)) <= 5e-6


# ---- whole notes ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", cases.CASES, ids=[c[0] for c in cases.CASES])
def test_render_case(case, torch_cuda):
    name, si, secs, cli = case
    feat, sf = cases.source_for(si, secs)
    ref = cases.oracle_render(feat, cli)
    got = render_one(sf, cli)[0].astype(np.float64)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= MAX_ABS
    assert cases.lsd_db(ref, got) <= MAX_LSD
    gold = np.load(os.path.join(GOLD, "render_cases.npz"))
    g32 = gold[f"out_{name}"].astype(np.float64)
    cmp = got if len(got) <= 50000 else got[::int(gold["long_stride"][0])]
    assert np.max(np.abs(cmp - g32)) <= MAX_ABS          # the reference's own output, committed


def test_host_entry_point_and_taps(torch_cuda):
    name, si, secs, cli = cases.CASES[2]
    feat, sf = cases.source_for(si, secs)
    taps = {}
    ref = cases.oracle_render(feat, cli, taps=taps)
    outs, tp = render_one(sf, cli, taps=True, device=False)
    assert np.max(np.abs(outs[0].astype(np.float64) - ref)) <= MAX_ABS
    gold = np.load(os.path.join(GOLD, "render_cases.npz"))
    for key, arr in zip(("harm", "uv", "bre"), tp):
        assert np.max(np.abs(arr[0] - gold[f"tap_{key}_{name}"])) <= MAX_ABS, key
    st = capi.last_stats()
    assert st["kernel_launches"] >= 10 and st["h2d_bytes"] > 0 and st["d2h_bytes"] == 4 * 4 * len(ref)


def _workload_batch(workload, idx, n_sources=8):
    b = host.Batch()
    feats = [bench_data.make_source(s) for s in range(n_sources)]
    for f in feats:
        b.add_source(host.SourceFeatures.from_knot_pack(f, f["mask"], f["formants"], f["sr"], f["ylen"]))
    for i in idx:
        src, cli = bench_data.note_cli(i, workload, n_sources=n_sources)
        b.add_note(host.NoteArgs.from_cli(src, cli))
    noise = host.SeededNoise(base_seed=lambda j: 20000 + 16 * idx[j], legacy_seed=lambda j: 777 + idx[j])
    return b, feats, noise


def _oracle_workload_note(workload, i, feats, n_sources=8):
    src, cli = bench_data.note_cli(i, workload, n_sources=n_sources)
    f = feats[src]
    env = dsp.decode_knots({"knot_vals_log": f["knot_vals_log"], "hz_knots": f["hz_knots"], "n_fft": 1024, "sr": 44100, "n_bins": 513})
    feat = resampler.Features(env=env, mask=f["mask"], formants=f["formants"], sr=f["sr"], ylen=f["ylen"])
    spec = resampler.NoteSpec.from_cli(*cli)
    return resampler.resample(feat, spec, lambda n, T: resampler.noise_for_note(spec, n, T, 20000 + 16 * i, 777 + i))


@pytest.mark.parametrize("workload,count,check", [("c2", 96, 12), ("c3", 24, 8)])
def test_batch_against_oracle(workload, count, check, torch_cuda):
    """BASELINE.json configs[1] / [2] parameterisation: a batch rendered in one call; a spread of its notes
    is compared with the oracle, every note must be finite."""
    idx = list(range(count))
    b, feats, noise = _workload_batch(workload, idx)
    db = b.assemble(noise).to_device("cuda:0")
    db.render()
    outs = db.outputs()
    assert all(np.all(np.isfinite(o)) for o in outs)
    worst = 0.0
    for i in idx[:: max(1, count // check)]:
        ref = _oracle_workload_note(workload, i, feats)
        err = float(np.max(np.abs(outs[i].astype(np.float64) - ref)))
        worst = max(worst, err)
        assert err <= MAX_ABS, f"{workload} note {i}: {err}"
        assert cases.lsd_db(ref, outs[i].astype(np.float64)) <= MAX_LSD
    print(f"{workload}: worst max-abs {worst:.2e}")


def test_batch_invariance_determinism_and_waves(torch_cuda):
    """A note renders to the same bits alone, inside a batch, on a second run, and when a small workspace
    forces the batch through several waves."""
    import ctypes as C
    torch = torch_cuda
    idx = list(range(40))
    b, feats, noise = _workload_batch("c3", idx)
    ab = b.assemble(noise)
    db = ab.to_device("cuda:0")
    a = db.render().clone()
    b2 = db.render().clone()
    assert torch.equal(a, b2)
    full = db.outputs()
    # the side-stream fork of the excitation chain (default) against everything on the caller's stream (goofer_profile(2))
    capi.profile(True, serial=True)
    c = db.render().clone()
    torch.cuda.synchronize()
    prof = capi.profile_summary()
    capi.profile(False)
    assert torch.equal(a, c)
    assert {"env", "frame", "walk", "mix"} <= set(prof) and all(ms > 0 for _, ms in prof.values())
    # several waves: a workspace that only fits a few notes at a time
    lib = capi.load()
    small = int(lib.goofer_workspace_bytes(C.byref(db.desc), 4))
    assert small < db.workspace.numel()
    db.workspace = torch.empty(small, dtype=torch.uint8, device="cuda:0")
    db.render()
    torch.cuda.synchronize()
    assert capi.last_stats()["waves"] > 1
    waved = db.outputs()
    for x, y in zip(full, waved):
        assert np.array_equal(x, y)
    # alone
    for i in (0, 7, 23):
        bb, _, nz = _workload_batch("c3", [i])
        d1 = bb.assemble(nz).to_device("cuda:0")
        d1.render()
        assert np.array_equal(d1.outputs()[0], full[i])


def test_properties_at_full_size(torch_cuda):
    """BASELINE.json configs[1] at its full size (1,024 notes): size-independent properties."""
    idx = list(range(1024))
    b, feats, noise = _workload_batch("c2", idx, n_sources=64)
    db = b.assemble(noise).to_device("cuda:0")
    db.render()
    outs = db.outputs()
    peaks = np.array([np.max(np.abs(o)) for o in outs])
    assert np.all(np.isfinite(peaks))
    # P absent => every synth pass is peak-normalised to 1 (GOOFER.py:1208-1218); V = B = U = volume = 1
    assert np.max(np.abs(peaks - 1.0)) <= 1e-5
    # istft leaves the last n - 256 * (T - 1) samples exactly zero (GOOFER.py:407-409)
    assert all(np.all(o[256 * 172:] == 0.0) for o in outs)
    # volume is linear: the same note at volume 50 is half the note at volume 100
    src, cli = bench_data.note_cli(5, "c2")
    f = feats[src]
    sf = host.SourceFeatures.from_knot_pack(f, f["mask"], f["formants"], f["sr"], f["ylen"])
    halves = []
    for vol in ("100", "50"):
        c = list(cli)
        c[7] = vol
        bb = host.Batch()
        bb.add_source(sf)
        bb.add_note(host.NoteArgs.from_cli(0, c))
        d = bb.assemble(host.SeededNoise(20000 + 16 * 5, 777 + 5)).to_device("cuda:0")
        d.render()
        halves.append(d.outputs()[0].astype(np.float64))
    assert np.max(np.abs(halves[0] * 0.5 - halves[1])) <= 1e-7
    assert np.array_equal(halves[0].astype(np.float32), outs[5])


def test_error_paths(torch_cuda):
    import ctypes as C
    feat, sf = cases.source_for(0, 1.0)
    b = host.Batch()
    b.add_source(sf)
    b.add_note(host.NoteArgs.from_cli(0, ["C4"]))
    db = b.assemble(host.SeededNoise()).to_device("cuda:0")
    lib = capi.load()
    rc = lib.goofer_render_batch(C.byref(db.desc), db.workspace.data_ptr(), 1024, None)
    assert rc == capi.ERR_WORKSPACE and b"workspace" in lib.goofer_last_error()
    db.desc.phi_total = 10
    rc = lib.goofer_render_batch(C.byref(db.desc), db.workspace.data_ptr(), db.workspace.numel(), None)
    assert rc == capi.ERR_INVALID


# ---- BASELINE.json configs[3]: long-note sustain, 4 s source stretched to 16 s -----------------------------------
@pytest.mark.parametrize("flags", ["L0", "L1", "L2R1", "L0R1"])
def test_long_note_sustain(flags, torch_cuda):
    """4 s source, length 16000 ms, consonant 150 ms: n_total = 705,600 + prefix, T_out ~ 2,757 frames."""
    feat, sf = cases.source_for(3, 4.0)            # fricative-initial source
    cli = ["G3", "100", flags, "30", "16000", "150", "200", "100", "0", "!120", "AA#200#AIAQAY#300#AQAIAA"]
    ref = cases.oracle_render(feat, cli)
    assert len(ref) > 700000
    got = render_one(sf, cli)[0].astype(np.float64)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= MAX_ABS
    assert cases.lsd_db(ref, got) <= MAX_LSD


def test_cli_drop_in(tmp_path, torch_cuda):
    """python -m goofer_b200.cli with the reference's 13 arguments: reads <stem>_features.goofy, writes PCM16."""
    import wave
    from goofer_b200 import cli
    src = bench_data.make_source(1)
    stem = os.path.join(tmp_path, "voice_a")
    with open(stem + "_features.goofy", "wb") as fh:
        np.savez_compressed(fh, mode=np.array(["knots"]), knot_vals_log=src["knot_vals_log"], hz_knots=src["hz_knots"],
                            n_bins=np.array([513]), n_fft=np.array([1024]), f0_interp=np.zeros(8, np.float16),
                            voicing_mask=src["mask"].astype(np.float16), formants=np.array(src["formants"], dtype=object),
                            sr=np.array([44100]), y_len=np.array([src["ylen"]]))
    out = os.path.join(tmp_path, "out.wav")
    argv = [stem + ".wav", out, "C4", "100", "g-10br20", "0", "1000", "0", "0", "100", "0", "!120", "AA"]
    assert cli.main(argv) == 0
    with wave.open(out, "rb") as w:
        assert w.getframerate() == 44100 and w.getsampwidth() == 2 and w.getnframes() == 44100
        pcm = np.frombuffer(w.readframes(44100), dtype="<i2")
    assert np.max(np.abs(pcm)) >= 32000               # P absent: peak-normalised
    # the wav holds the device-encoded PCM: identical to soundfile's conversion of the f32 render of the same note
    # (cli.render_notes draws fresh noise, so compare through a seeded batch instead)
    assert cli.main(argv[:5]) == 1                      # fewer than 13 arguments: usage + exit code 1
    assert cli.main([os.path.join(tmp_path, "missing.wav")] + argv[1:]) == 1


def test_analysis_front_end_gpu(torch_cuda):
    """goofer_analyse_batch against the reference's own stft + gaussian_filter1d + compress_env_to_knots output
    (tests/golden/stages.npz) and against the oracle on a batch."""
    from oracle import sources
    torch = torch_cuda
    g = np.load(os.path.join(GOLD, "stages.npz"))
    ys = [sources.make_source(int(g[f"{t}_src"][0]), 1.0)[0].astype(np.float32) for t in ("an_a", "an_b")]
    packs = ops.analyse_envelope(torch.from_numpy(np.stack(ys)).cuda(), 44100)
    for tag, y, pack in zip(("an_a", "an_b"), ys, packs):
        assert pack["knot_vals_log"].shape == g[f"{tag}_knots"].shape and pack["knot_vals_log"].dtype == np.float16
        # numpy's float32 power (SIMD) is not correctly rounded: 10 ** x differs from powf by one ulp in ~20 % of the
        # knots, which 700 * (10 ** x - 1) turns into <= 0.002 Hz; the knot bins (round(hz / 43.07)) are identical
        assert np.max(np.abs(pack["hz_knots"] - g[f"{tag}_hz"])) <= 4e-3
        a, r = pack["knot_vals_log"].astype(np.float32), g[f"{tag}_knots"].astype(np.float32)
        assert np.max(np.abs(a - r)) <= 2e-3 * np.max(np.abs(r))                 # at most one f16 ulp apart
        assert np.mean(pack["knot_vals_log"].view(np.uint16) == g[f"{tag}_knots"].view(np.uint16)) > 0.99
        # decoded envelopes agree (the render path consumes exactly this)
        env_a = dsp.decode_knots(pack)
        env_r = dsp.decode_knots({"knot_vals_log": g[f"{tag}_knots"], "hz_knots": g[f"{tag}_hz"], "n_fft": 1024, "sr": 44100, "n_bins": 513})
        assert np.max(np.abs(env_a - env_r) / env_r) <= 5e-3          # one f16 ulp of a log value near 5 is 0.4 %
    sil = ops.analyse_envelope(torch.zeros(20000, device="cuda"), 44100)[0]
    assert sil["knot_vals_log"].shape[0] == 32 == int(g["an_silence_K"][0])
    assert np.array_equal(sil["knot_vals_log"].view(np.uint16), g["an_silence_knots"].view(np.uint16))
    # ragged / short inputs
    rng = np.random.default_rng(11)
    for n in (300, 5000):
        y = (0.2 * rng.standard_normal(n)).astype(np.float32)
        p = ops.analyse_envelope(torch.from_numpy(y).cuda(), 44100)[0]
        _, po = dsp.analyse_envelope(y, 44100)
        assert p["knot_vals_log"].shape == po["knot_vals_log"].shape
        assert np.max(np.abs(p["knot_vals_log"].astype(np.float32) - po["knot_vals_log"].astype(np.float32))) <= 2e-2


# ---- randomised argument fuzz: ragged lengths, extreme flags, both feature modes --------------------------------
def _fuzz_cli(rng):
    pitch = bench_data.midi_to_name(int(rng.integers(36, 84)))
    vel = str(int(rng.choice([0, 50, 100, 100, 100, 150, 200])))
    offset = int(rng.integers(0, 300))
    length = int(rng.choice([60, 150, 400, 900, 1500, 3000]))
    consonant = int(rng.choice([0, 0, 40, 120, 250]))
    cutoff = int(rng.choice([0, 50, 200, -300, -600]))
    volume = str(int(rng.choice([100, 100, 60, 130])))
    tempo = "!" + str(int(rng.choice([60, 120, 180, 240])))
    n_ticks = int(rng.integers(1, 400))
    bend = "AA" if rng.random() < 0.4 else bench_data.cents_to_pitch_string(np.round(80 * np.sin(np.arange(n_ticks) / 9.0 + rng.random())).astype(int))
    fl = []
    pool = {"g": (-100, 100), "fa": (-9, 9), "fb": (-9, 9), "fc": (-9, 9), "fd": (-9, 9), "fw": (-100, 100), "fst": (-100, 100),
            "fsta": (-50, 50), "fstb": (-100, 100), "fstc": (-100, 100), "fstd": (-50, 50), "br": (-100, 100), "es": (-100, 100),
            "V": (0, 100), "B": (-100, 100),
            "U": (-100, 100), "P": (0, 100), "t": (-50, 50), "L": (0, 2), "R": (0, 1), "FV": (0, 1), "sd": (0, 100), "st": (-100, 100),
            "su": (0, 100), "sa": (0, 100), "vf": (-100, 100), "vh": (20, 100), "vl": (0, 100), "pd": (-100, 100), "sr": (0, 100),
            "sh": (0, 100), "sg": (0, 100), "sj": (0, 100)}
    for k in rng.choice(list(pool), size=int(rng.integers(0, 9)), replace=False):
        lo, hi = pool[k]
        fl.append(f"{k}{int(rng.integers(lo, hi + 1))}")
    return [pitch, vel, "".join(fl), str(offset), str(length), str(consonant), str(cutoff), volume, "0", tempo, bend]


def test_fuzzed_arguments(torch_cuda):
    """48 seeded random argument lists (ragged lengths 60 ms .. 3 s, velocity 0 .. 200, negative cutoffs, every
    loop / reverse mode, extreme flag values) on knot-coded AND dense-envelope sources, rendered as ONE batch.
    Notes the reference cannot render (empty tail: ZeroDivisionError) must be rejected by the planner too."""
    rng = np.random.default_rng(2024)
    feats, sfs = [], []
    for si in range(4):
        feat, sf = cases.source_for(si, 1.0)
        if si % 2:                                     # 'full' mode: dense (513, T) envelope instead of knots
            sf = host.SourceFeatures.from_dense(feat.env, feat.mask, feat.formants, feat.sr, feat.ylen)
        feats.append(feat)
        sfs.append(sf)
    notes, refs = [], []
    lib = capi.load()
    tried = rejected = 0
    while len(notes) < 48:
        si = int(rng.integers(0, 4))
        cli = _fuzz_cli(rng)
        tried += 1
        spec = resampler.NoteSpec.from_cli(*cli)
        seed = 5000 + tried
        try:
            ref = resampler.resample(feats[si], spec, lambda n, T: resampler.noise_for_note(spec, n, T, seed, seed + 1))
        except ZeroDivisionError:
            b = host.Batch()
            b.add_source(sfs[si])
            b.add_note(host.NoteArgs.from_cli(0, cli))
            with pytest.raises(capi.GooferError):
                b.assemble(host.SeededNoise())
            rejected += 1
            continue
        if len(ref) < 2:
            continue
        notes.append((si, cli, seed))
        refs.append(ref)
    b = host.Batch()
    for sf in sfs:
        b.add_source(sf)
    for si, cli, seed in notes:
        b.add_note(host.NoteArgs.from_cli(si, cli))
    seeds = [s for _, _, s in notes]
    db = b.assemble(host.SeededNoise(base_seed=lambda j: seeds[j], legacy_seed=lambda j: seeds[j] + 1)).to_device("cuda:0")
    db.render()
    outs = db.outputs()
    worst = 0.0
    for k, (ref, got) in enumerate(zip(refs, outs)):
        assert got.shape == ref.shape, notes[k]
        err = float(np.max(np.abs(got.astype(np.float64) - ref))) if len(ref) else 0.0
        worst = max(worst, err)
        assert err <= MAX_ABS, (notes[k], err)
    print(f"fuzz: {len(notes)} notes, {rejected} rejected by both, worst max-abs {worst:.2e}")


def test_pcm16_encoded_on_device(torch_cuda):
    """GooferBatch.out_pcm16 (the step after the path, SillySampler.py:1185): the device encoder equals
    libsndfile's clip-path conversion of the f32 output bit for bit, through both entry points; `out` may be NULL."""
    from goofer_b200 import cli
    notes = [["A3", "100", "g-20fa5br20", "0", "1000", "0", "0", "100", "0", "!120", "AA"],
             ["C4", "100", "P0V100", "0", "700", "0", "0", "200", "0", "!120", "AA"],       # unnormalised, volume 2: saturates
             ["E4", "90", "B50sh20sr30", "0", "900", "50", "0", "3", "0", "!120", "AA"]]      # quiet: rounding near zero
    b = host.Batch()
    b.add_source(cases.source_for(1, 1.0)[1])
    for c in notes:
        b.add_note(host.NoteArgs.from_cli(0, c))
    ab = b.assemble(host.SeededNoise(cases.SEED_BASE, cases.SEED_LEGACY))
    db = ab.to_device("cuda:0").enable_pcm16()
    db.render()
    torch_cuda.cuda.synchronize()
    f32, pcm = db.outputs(), db.outputs_pcm16()
    for x, q in zip(f32, pcm):
        assert q.dtype == np.int16 and np.array_equal(q, cli.pcm16_like_soundfile(x))
    assert any(np.max(q) == 32767 for q in pcm)                       # the saturating note did saturate
    ab.pin()
    host_pcm = ab.render_host(pcm16=True)
    st = capi.last_stats()
    assert st["d2h_bytes"] == 2 * sum(len(q) for q in pcm)            # int16 only: half of the f32 download
    for q, h in zip(pcm, host_pcm):
        assert np.array_equal(q, h)


def test_direct_synthesize_seam(torch_cuda):
    """ops.synthesize = gf.synthesize called directly (SillyEditor.py:227, 559; test.py:38; SURVEY 8b "Python DSP API"):
    explicit per-sample f0 curve, whole source, keyword arguments carried by the flag columns -- against the committed
    outputs of the reference itself (tests/golden/synth_direct.npz, made by tests/golden/make_golden_synth.py)."""
    from tests.golden import make_golden_synth as mg
    from tests.test_oracle_golden import synth_direct_noise
    feat, pack, forms, _, _ = mg.inputs()
    g = np.load(os.path.join(GOLD, "synth_direct.npz"))
    n, sr = len(feat.mask), feat.sr
    base, legacy = (int(x) for x in g["seeds"])

    def provider(f0_jitter, volume_jitter):
        return lambda i, info: synth_direct_noise(info["n_total"], info["t_out"], base, legacy, f0_jitter, volume_jitter)

    knots = {"mode": "knots", "knot_vals_log": pack["knot_vals_log"], "hz_knots": pack["hz_knots"], "n_fft": 1024, "sr": sr,
             "n_bins": 513}
    ra = ops.synthesize(knots, g["f0_a"], feat.mask, np.empty(n, dtype=np.bool_), sr, formants=forms, noise=provider(False, False))
    rb = ops.synthesize(feat.env, g["f0_b"], feat.mask, None, sr, formants=forms, noise=provider(True, True), **mg.KW_B)
    # continuous keyword values (GooferNote.override_val): test.py:38's formant_shift / breath_strength / uv_strength ...
    rc = ops.synthesize(knots, g["f0_a"], feat.mask, None, sr, formants=forms, noise=provider(True, True), **mg.KW_C)
    for tag, r in (("a", ra), ("b", rb), ("c", rc)):
        for name, arr in zip(("reconstruct", "harmonic", "aper_uv", "aper_bre"), r):
            ref = g[f"{tag}_{name}"]
            assert arr.shape == ref.shape and arr.dtype == np.float32
            assert np.max(np.abs(arr.astype(np.float64) - ref)) <= MAX_ABS, (tag, name)
        assert cases.lsd_db(g[f"{tag}_reconstruct"].astype(np.float64), r[0].astype(np.float64)) <= MAX_LSD
    with pytest.raises(NotImplementedError):
        ops.synthesize(feat.env, g["f0_b"], feat.mask, None, sr, roughness_on=True)
    with pytest.raises(NotImplementedError):
        ops.synthesize(feat.env, g["f0_b"], feat.mask, None, sr, stretch_factor=1.5)


def test_http_server_batches_on_the_gpu(tmp_path, torch_cuda):
    """The reference's server mode (SillySampler.py:1187-1224) with the real renderer: concurrent POSTs come back
    200, are rendered as ONE GPU batch, and every out.wav holds a peak-normalised 16-bit note."""
    import threading
    import urllib.request
    import wave
    from goofer_b200 import server
    src = bench_data.make_source(2)
    stem = os.path.join(tmp_path, "voice_b")
    with open(stem + "_features.goofy", "wb") as fh:
        np.savez_compressed(fh, mode=np.array(["knots"]), knot_vals_log=src["knot_vals_log"], hz_knots=src["hz_knots"],
                            n_bins=np.array([513]), n_fft=np.array([1024]), f0_interp=np.zeros(8, np.float16),
                            voicing_mask=src["mask"].astype(np.float16), formants=np.array(src["formants"], dtype=object),
                            sr=np.array([44100]), y_len=np.array([src["ylen"]]))
    batcher = server.Batcher(window_ms=300.0)
    batcher.start()
    httpd = server.ThreadedHTTPServer(("127.0.0.1", 0), server.make_handler(batcher))
    port = httpd.server_address[1]
    th = threading.Thread(target=httpd.serve_forever, daemon=True)
    th.start()
    try:
        pitches = ["C4", "E4", "G4", "A3", "D4", "F4"]
        outs = [os.path.join(tmp_path, f"out_{i}.wav") for i in range(len(pitches))]
        codes = [None] * len(pitches)

        def post(i):
            body = " ".join([stem + ".wav", outs[i], pitches[i], "100", "g-5br10", "0", "600", "0", "0", "100", "0", "!120", "AA"])
            req = urllib.request.Request(f"http://127.0.0.1:{port}/", data=body.encode("utf-8"), method="POST")
            with urllib.request.urlopen(req, timeout=120) as r:
                codes[i] = r.status

        threads = [threading.Thread(target=post, args=(i,)) for i in range(len(pitches))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert codes == [200] * len(pitches)
        assert sum(batcher.batches) == len(pitches) and len(batcher.batches) <= 2       # coalesced, not one render per request
        for o in outs:
            with wave.open(o, "rb") as w:
                assert w.getframerate() == 44100 and w.getsampwidth() == 2 and w.getnframes() == 26460
                pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
            assert np.max(np.abs(pcm.astype(np.int32))) >= 32000
        bad = urllib.request.Request(f"http://127.0.0.1:{port}/", data=b"no paths here C4 100 g0 0 600 0 0 100 0 !120 AA", method="POST")
        try:
            urllib.request.urlopen(bad, timeout=60)
            assert False, "expected HTTP 500"
        except urllib.error.HTTPError as e:
            assert e.code == 500
    finally:
        httpd.shutdown()
        batcher.stop()


def test_device_source_cache(tmp_path, torch_cuda):
    """a feature file is uploaded once per process; later batches reuse the resident tensors and
    render the same samples."""
    from goofer_b200 import cli
    src = bench_data.make_source(4)
    stem = os.path.join(tmp_path, "voice_c")
    with open(stem + "_features.goofy", "wb") as fh:
        np.savez_compressed(fh, mode=np.array(["knots"]), knot_vals_log=src["knot_vals_log"], hz_knots=src["hz_knots"],
                            n_bins=np.array([513]), n_fft=np.array([1024]), f0_interp=np.zeros(8, np.float16),
                            voicing_mask=src["mask"].astype(np.float16), formants=np.array(src["formants"], dtype=object),
                            sr=np.array([44100]), y_len=np.array([src["ylen"]]))
    args = [[stem + ".wav", "x.wav", "D4", "100", "g10fst20", "0", "800", "0", "0", "100", "0", "!120", "AA"],
            [stem + ".wav", "y.wav", "F4", "100", "", "0", "500", "0", "0", "100", "0", "!120", "AA"]]
    h0, m0 = cli.SOURCE_CACHE.hits, cli.SOURCE_CACHE.misses
    a = cli.render_notes(args, noise=host.SeededNoise(5, 6))
    b = cli.render_notes(args, noise=host.SeededNoise(5, 6))
    assert cli.SOURCE_CACHE.misses == m0 + 1 and cli.SOURCE_CACHE.hits == h0 + 1     # one file, two batches
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_edge_lengths_and_empty_batch(torch_cuda):
    """One batch with a 10 ms note (441 samples: a single STFT frame pair), a 20 s looped note, odd lengths (ragged,
    unaligned output offsets: the 4-byte store path of the mix kernel) and an unnormalised quiet note, against the
    oracle, with the device PCM; the host entry point returns the same samples; an empty batch is a no-op."""
    from goofer_b200 import cli
    feat, sf = cases.source_for(1, 1.0)
    clis = [["C4", "100", "", "0", "333", "0", "0", "100", "0", "!120", "AA"], ["D4", "100", "g5", "0", "10", "0", "0", "100", "0", "!120", "AA"],
            ["E4", "100", "L1", "0", "20000", "50", "0", "100", "0", "!120", "AA"], ["F4", "100", "P0", "0", "777", "0", "0", "37", "0", "!120", "AA"]]
    b = host.Batch()
    b.add_source(sf)
    for c in clis:
        b.add_note(host.NoteArgs.from_cli(0, c))
    ab = b.assemble(host.SeededNoise(cases.SEED_BASE, cases.SEED_LEGACY))
    db = ab.to_device("cuda:0").enable_pcm16()
    db.render()
    torch_cuda.cuda.synchronize()
    outs, pcm = db.outputs(), db.outputs_pcm16()
    assert [len(o) for o in outs] == [14685, 441, 884205, 34265]
    for c, o, q in zip(clis, outs, pcm):
        ref = cases.oracle_render(feat, c)
        assert np.max(np.abs(o.astype(np.float64) - ref)) <= MAX_ABS, c
        assert np.array_equal(q, cli.pcm16_like_soundfile(o))
    for h, o in zip(ab.render_host(), outs):
        assert np.array_equal(h, o)
    empty = host.Batch()
    empty.add_source(sf)
    ab0 = empty.assemble(host.SeededNoise(1, 2))
    assert ab0.render_host() == []


def test_long_kernels_on_short_notes_and_long_passes(torch_cuda):
    """The overlap-save smoothing (sigma 441 / 73.5 / 49: halos of 1,764 / 294 / 196 samples) on notes SHORTER than the halo
    (numpy's 'reflect' padding folds several times there), on notes between one and two halos, and a 5 s note (its pulse
    onsets skip the fixed-point scan: passes over 4 s go straight to the bit-exact walk), all flags that use long kernels on."""
    import bench_data
    feat, sf = cases.source_for(2, 1.0)
    flags = "pd80sh70sr60sg40"
    clis = [["C4", "100", flags, "0", "10", "0", "0", "100", "0", "!120", bench_data._vibrato_string(3, 0.1)],
            ["A3", "100", flags, "0", "60", "0", "0", "100", "0", "!120", bench_data._vibrato_string(4, 0.2)],
            ["E4", "100", "pd-60sh30", "0", "5000", "40", "0", "100", "0", "!120", bench_data._vibrato_string(5, 5.2)],
            ["A4", "100", "pd50sr90", "0", "700", "0", "0", "100", "0", "!120", "AA"]]
    b = host.Batch()
    b.add_source(sf)
    for c in clis:
        b.add_note(host.NoteArgs.from_cli(0, c))
    ab = b.assemble(host.SeededNoise(cases.SEED_BASE, cases.SEED_LEGACY))
    db = ab.to_device("cuda:0")
    db.render()
    torch_cuda.cuda.synchronize()
    outs = db.outputs()
    assert len(outs[0]) < 1764 and 1764 < len(outs[1]) < 2 * 1764 and len(outs[2]) > 4 * 44100
    for c, o in zip(clis, outs):
        ref = cases.oracle_render(feat, c)
        assert len(ref) == len(o)
        assert np.max(np.abs(o.astype(np.float64) - ref)) <= MAX_ABS, c


def test_noise_phases_drawn_on_the_device(torch_cuda):
    """GooferNote.phi_rng: the device draws numpy's Generator(PCG64).uniform(0, 2 pi, (513, T)).astype(float32) stream bit
    for bit (128-bit LCG jump-ahead per thread) -- the renders with host-supplied buffers (SeededNoise) and with
    device-drawn phases (DeviceNoise, same seeds) are identical sample for sample, through both entry points, for the
    main, su, sj and sa passes; nothing but the seeds crosses PCIe for the phases."""
    feat, sf = cases.source_for(2, 1.0)
    clis = [["A3", "100", "g-20fa5br20", "0", "1000", "0", "0", "100", "0", "!120", "AA"],
            ["C4", "100", "su40sj30sa50sh20sr10", "0", "800", "50", "0", "100", "0", "!120", "AA"],
            ["E4", "100", "L1", "0", "2500", "50", "0", "100", "0", "!120", "AA"]]

    def render(noise):
        b = host.Batch()
        b.add_source(sf)
        for c in clis:
            b.add_note(host.NoteArgs.from_cli(0, c))
        ab = b.assemble(noise)
        db = ab.to_device("cuda:0")
        db.render()
        torch_cuda.cuda.synchronize()
        return ab, db.outputs()

    ab_h, ref = render(host.SeededNoise(cases.SEED_BASE, cases.SEED_LEGACY))
    ab_d, got = render(host.DeviceNoise(cases.SEED_BASE, cases.SEED_LEGACY))
    assert ab_d.phi.size == 1 and ab_h.phi.size > 3 * 513 * 100                 # no phase buffer on the host side
    for r, g in zip(ref, got):
        assert np.array_equal(r, g)
    host_out = ab_d.render_host()
    h2d_device_noise = capi.last_stats()["h2d_bytes"]
    for r, g in zip(ref, host_out):
        assert np.array_equal(r, g)
    ab_h.render_host()
    assert capi.last_stats()["h2d_bytes"] - h2d_device_noise >= ab_h.phi.nbytes - 64     # only the phases differ
