"""CPU: the host planner (plan.cpp) and the __host__ __device__ index maps / FFT passes of goofer_b200/csrc
(driven serially by tests/cpu_emul) against the oracle's restatement of SillySampler's slicing, looping,
velocity-stretch and formant-track logic."""
import ctypes as C

import numpy as np
import pytest

from goofer_b200 import capi, host
from tests import cases, cpu_emul
from tests.plan_struct import GfNotePlan


def one_note_batch(sf, cli):
    b = host.Batch()
    b.add_source(sf)
    b.add_note(host.NoteArgs.from_cli(0, cli))
    return b


def plan_of(lib, ab) -> GfNotePlan:
    p = GfNotePlan()
    rc = lib.goofer_debug_plan(C.byref(ab.desc), 0, C.byref(p), C.sizeof(p))
    assert rc == C.sizeof(p), "tests/plan_struct.py is out of sync with csrc/gf_plan.h"
    return p


@pytest.fixture(scope="module")
def emul():
    L = cpu_emul.load()
    assert L.emul_plan_size() == C.sizeof(GfNotePlan)
    return L


@pytest.mark.parametrize("case", cases.CASES, ids=[c[0] for c in cases.CASES])
def test_plan_lengths_and_maps(case, lib, emul):
    name, si, secs, cli = case
    feat, sf = cases.source_for(si, secs)
    taps = {}
    cases.oracle_render(feat, cli, taps=taps)
    ab = one_note_batch(sf, cli).assemble(host.SeededNoise(cases.SEED_BASE, cases.SEED_LEGACY))
    inf = ab.infos[0]
    assert inf["n_total"] == taps["n_total"]
    assert inf["t_env"] == taps["env_new"].shape[1]
    assert inf["t_out"] == 1 + taps["n_total"] // 256
    p = plan_of(lib, ab)
    assert p.status == 0 and p.n_total == inf["n_total"]

    # mask_new (tile + velocity stretch), SillySampler.py:699-712, 788
    mask_src = np.ascontiguousarray(sf.mask, dtype=np.float32)
    got = np.zeros(p.n_total, dtype=np.float64)
    emul.emul_mask_new(C.byref(p), mask_src.ctypes.data_as(C.POINTER(C.c_float)), got.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.array_equal(got.astype(np.float32), np.asarray(taps["mask_new"], dtype=np.float32))

    # canonical formant tracks, SillySampler.py:714-763, 776-792
    for k in range(4):
        trk = np.ascontiguousarray(sf.formants[k + 1], dtype=np.float64)
        out = np.zeros(p.T_env, dtype=np.float32)
        emul.emul_track_canon(C.byref(p), trk.ctypes.data_as(C.POINTER(C.c_double)), k, out.ctypes.data_as(C.POINTER(C.c_float)))
        ref = taps["formants"][f"F{k + 1}"]
        m = min(len(ref), p.T_env)
        assert np.max(np.abs(out[:m] - ref[:m])) <= 1e-3 * max(1.0, np.max(np.abs(ref)))

    # envelope frame map: only when no flag reshapes the envelope (br / es / fw / fst / vf change values)
    fl = host.parse_flags(cli[2])
    if not any(k in fl for k in ("br", "es", "fw", "fst", "fsta", "fstb", "fstc", "fstd", "vf")):
        env_src = np.asarray(feat.env, dtype=np.float64)
        T = p.T_env
        f = (C.c_int * 4)()
        w = (C.c_double * 4)()
        rec = np.zeros((513, T))
        for t in range(T):
            n = emul.emul_env_mix(C.byref(p), t, f, w)
            assert 1 <= n <= 4
            for j in range(n):
                rec[:, t] += w[j] * env_src[:, f[j]]
        ref = np.asarray(taps["env_new"], dtype=np.float64)
        assert rec.shape == ref.shape
        assert np.max(np.abs(rec - ref) / (np.abs(ref) + 1e-12)) <= 1e-6


def test_fft_passes_match_numpy(emul):
    rng = np.random.default_rng(7)
    x = rng.standard_normal(1024).astype(np.float32)
    X = np.zeros(513 * 2, dtype=np.float32)
    fp = C.POINTER(C.c_float)
    emul.emul_rfft1024(x.ctypes.data_as(fp), X.ctypes.data_as(fp))
    got = X[0::2] + 1j * X[1::2]
    ref = np.fft.rfft(x.astype(np.float64))
    assert np.max(np.abs(got - ref)) <= 2e-6 * np.max(np.abs(ref))
    y = np.zeros(1024, dtype=np.float32)
    Xin = np.ascontiguousarray(np.stack([ref.real, ref.imag], axis=1).astype(np.float32).reshape(-1))
    emul.emul_irfft1024(Xin.ctypes.data_as(fp), y.ctypes.data_as(fp))
    assert np.max(np.abs(y - x)) <= 2e-6 * np.max(np.abs(x))


def _gauss_direct(x, sigma):
    """gaussian_filter1d as the reference defines it (GOOFER.py:241-261), fp64"""
    radius = int(4.0 * sigma + 0.5)
    t = np.arange(-radius, radius + 1, dtype=np.float64)
    k = np.exp(-0.5 * (t / sigma) ** 2)
    k /= k.sum()
    return np.convolve(np.pad(x.astype(np.float64), radius, mode="reflect"), k, mode="valid")


@pytest.mark.parametrize("n,sigma", [(44100, 441.0), (5000, 441.0), (9329, 441.0), (44100, 73.5), (70001, 49.0), (3000, 73.5)])
def test_overlap_save_passes_match_direct_convolution(emul, n, sigma):
    """k_conv.cu's transform passes, block layout and spectrum product (run serially) against np.convolve"""
    rng = np.random.default_rng(n)
    fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
    x64 = rng.standard_normal(n)
    ref = _gauss_direct(x64, sigma)
    y64 = np.zeros(n)
    assert emul.emul_fftconv_f64(x64.ctypes.data_as(dp), n, sigma, y64.ctypes.data_as(dp)) == (1 if sigma < 256 else 0)
    if sigma < 256:
        assert np.max(np.abs(y64 - ref)) <= 1e-14
    # f32: a mask-like 0/1 signal and a smooth curve plus noise, like the pitch-deviation input
    for x in ((rng.random(n) < 0.5).astype(np.float32), (np.sin(np.arange(n) / 3000.0) * 2 + 0.01 * rng.standard_normal(n)).astype(np.float32)):
        y = np.zeros(n, dtype=np.float32)
        assert emul.emul_fftconv_f32(x.ctypes.data_as(fp), n, sigma, y.ctypes.data_as(fp)) == 1
        assert np.max(np.abs(y - _gauss_direct(x, sigma))) <= 2e-6


def test_planner_statuses(lib):
    feat, sf = cases.source_for(0, 1.0)
    # empty tail: offset past the end -> the reference dies with ZeroDivisionError (SillySampler.py:634)
    for cli, status in ((["C4", "100", "", "1000", "1000", "0", "0", "100", "0", "!120", "AA"], capi.NOTE_EMPTY_TAIL),
                        (["C4", "100", "SE1", "0", "1000", "0", "0", "100", "0", "!120", "AA"], capi.NOTE_EDITOR)):
        b = one_note_batch(sf, cli)
        with pytest.raises(capi.GooferError) as ei:
            b.assemble(host.SeededNoise())
        assert ei.value.code == capi.ERR_NOTE
    with pytest.raises(ZeroDivisionError):                 # what the reference does with the first of them
        cases.oracle_render(feat, ["C4", "100", "", "1000", "1000", "0", "0", "100", "0", "!120", "AA"])
    b = host.Batch()
    b.add_source(sf)
    b.add_note(host.NoteArgs.from_cli(3, ["C4"]))          # source index out of range
    with pytest.raises(capi.GooferError):
        b.assemble(host.SeededNoise())
