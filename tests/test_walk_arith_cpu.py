"""CPU: the integer arithmetic behind gf_walk_kernel (goofer_b200/csrc/k_excite.cu), restated with Python integers
and exact rationals and checked against the scalar fp64 loop of pulse_train_numba (GOOFER.py:479-493).

1. While the exponent e of the running total is fixed, one round-to-nearest-even addition of an increment is the
   integer map M -> M + delta[M & 1] (positive AND negative increments); the maps compose associatively, so the
   running mantissas come out of a scan.  Steps that leave the binade ("events", ~16 per note) are real additions.
2. Division by the sample rate with the correctly rounded reciprocal (Markstein): q = RN(a y), r = a - b q,
   RN(q + r y) equals RN(a / b)."""
import fractions
import struct

import numpy as np
import pytest

F = fractions.Fraction
ONE52, ONE53 = 1 << 52, 1 << 53


def _split(x):
    b = struct.unpack("<Q", struct.pack("<d", abs(float(x))))[0]
    e, m = (b >> 52) & 0x7FF, b & (ONE52 - 1)
    if e == 0:
        return (0, None) if m == 0 else (None, None)
    return m | ONE52, e - 1023


def delta(x, e):
    """(d0, d1, guard, is_event) of gf_walk_delta for a positive normal running total with exponent e."""
    if x == 0:
        return 0, 0, 0, False
    mk, ek = _split(x)
    if mk is None or e - ek < 0:
        return 0, 0, 0, True
    s = e - ek
    if s >= 64:
        return 0, 0, 0, False
    q, r = mk >> s, mk & ((1 << s) - 1)
    half = (1 << (s - 1)) if s > 0 else 0
    tie = 1 if (s > 0 and r == half) else 0
    if x > 0:
        c = 1 if (s > 0 and r > half) else 0
        return q + c + tie * (q & 1), q + c + tie * ((q + 1) & 1), 0, False
    guard = q + (1 if r > 0 else 0)
    if r == 0:
        return -q, -q, guard, False
    c = 1 if r < half else 0
    return -q - 1 + c + tie * ((q + 1) & 1), -q - 1 + c + tie * (q & 1), guard, False


def delta_fp64(x, e):
    """gf_walk_delta as the kernel computes it: the delta pair read off TWO REAL fp64 additions on representative
    totals of the binade -- 1.5 * 2^e (even mantissa) and its successor (odd mantissa).  The delta only depends on the
    parity of M while M + delta stays in the binade, which holds for the representatives when |x| < 2^(e-1); larger
    increments (the first samples of a note) are events.  Python floats are IEEE doubles with round-to-nearest-even."""
    if x == 0:
        return 0, 0, 0, False
    mk, ek = _split(x)
    if mk is None or ek >= e - 1:
        return 0, 0, 0, True
    base0 = 1.5 * 2.0 ** e
    base1 = struct.unpack("<d", struct.pack("<Q", struct.unpack("<Q", struct.pack("<d", base0))[0] + 1))[0]
    scale = 2.0 ** (52 - e)
    d0 = int(((base0 + x) - base0) * scale)
    d1 = int(((base1 + x) - base1) * scale)
    guard = 0
    if x < 0:
        g = F(-x) * F(2) ** (52 - e)                           # ceil(-x * 2^(52 - e)); the product is exact in fp64
        guard = -((-g.numerator) // g.denominator)
    return d0, d1, guard, False


def then(a, b):
    return a[0] + b[(0 + a[0]) & 1], a[1] + b[(1 + a[1]) & 1]


def walk(inc, block=256, delta=None):
    """attempt / commit loop of the kernel (scan written as a Hillis-Steele pass over the block)."""
    delta = delta or globals()["delta"]
    n = len(inc)
    out = np.zeros(n)
    started, raw_mode, raw, M, e = False, False, 0.0, 0, 0
    attempts = events = 0
    for blk in range(0, n, block):
        blk_end, lo = min(n, blk + block), blk
        while lo < blk_end:
            attempts += 1
            cnt = blk_end - lo
            maps, guards, evs = [], [], []
            for k in range(lo, blk_end):
                x = inc[k]
                if not started or raw_mode:
                    maps.append((0, 0)); guards.append(0); evs.append(x != 0)
                else:
                    d0, d1, g, ev = delta(x, e)
                    maps.append((d0, d1)); guards.append(g); evs.append(ev)
            Fs = list(maps)
            o = 1
            while o < cnt:
                Fs = [Fs[k] if k < o else then(Fs[k - o], Fs[k]) for k in range(cnt)]
                o <<= 1
            kstar, Mv = blk_end, []
            for k in range(cnt):
                Mprev = M if k == 0 else M + Fs[k - 1][M & 1]
                Mk = M + Fs[k][M & 1]
                bad = evs[k]
                if started and not raw_mode:
                    bad = bad or (Mprev - guards[k] < ONE52) or Mk >= ONE53 or Mk < ONE52
                if bad:
                    kstar = lo + k
                    break
                Mv.append(Mk)
            for j in range(kstar - lo):
                out[lo + j] = raw if raw_mode else (Mv[j] * 2.0 ** (e - 52) if started else 0.0)
            if kstar > lo and started and not raw_mode:
                M = Mv[-1]
            if kstar < blk_end:
                prev = raw if raw_mode else (M * 2.0 ** (e - 52) if started else 0.0)
                tot = prev + inc[kstar]                               # the one real fp64 addition
                out[kstar] = tot
                events += 1
                mk, ek = _split(tot) if tot > 0 else (None, None)
                if tot > 0 and mk:
                    started, raw_mode, M, e = True, False, mk, ek
                elif tot == 0:
                    started, raw_mode, M = False, False, 0
                else:
                    started, raw_mode, raw, M = False, True, tot, 0
                lo = kstar + 1
            else:
                lo = blk_end
    return out, attempts, events


def scalar(inc):
    out = np.zeros(len(inc))
    t = 0.0
    for i, x in enumerate(inc):
        t = t + x
        out[i] = t
    return out


def _cases():
    rng = np.random.default_rng(1)
    c = {}
    for hz in (110.0, 220.0, 440.0, 50.0, 261.6255653005986):
        f = np.full(12000, hz, dtype=np.float32)
        f[:1500] = 0
        c[f"flat {hz:g}"] = f.astype(np.float64) / 44100.0
    c["vibrato"] = (220 * 2 ** (0.3 * np.sin(np.arange(16000) / 800.0) / 12)).astype(np.float32).astype(np.float64) / 44100.0
    c["glide"] = np.linspace(60, 900, 12000).astype(np.float32).astype(np.float64) / 44100.0
    c["gappy"] = (330 * (rng.random(12000) > 0.3)).astype(np.float32).astype(np.float64) / 44100.0
    c["tiny"] = (1e-5 * rng.random(8000)).astype(np.float32).astype(np.float64) / 44100.0
    c["negative jitter"] = (220 * (1 + 2.0 * np.clip(rng.standard_normal(12000), -1, 1))).astype(np.float32).astype(np.float64) / 44100.0
    c["negative start"] = np.concatenate([np.full(300, -80.0), np.full(4000, 300.0)]).astype(np.float32).astype(np.float64) / 44100.0
    c["saw"] = np.tile(np.concatenate([np.full(300, 500.0), np.full(290, -500.0)]), 12).astype(np.float32).astype(np.float64) / 44100.0
    c["binary fractions"] = np.tile(np.array([0.5, 0.25, 0.125, -0.25]), 1500)
    # subtracting between a quarter and half an ulp from an exact power of two lands in the finer binade below
    c["power of two"] = np.array([1.0, -1.5 * 2.0 ** -54, 0.0, 2.0 ** -60, 1.0, -2.0 ** -54, -2.0 ** -55, -1.2 * 2.0 ** -53, 3.0, -0.75 * 2.0 ** -52] * 40)
    return c


@pytest.mark.parametrize("name", list(_cases()))
def test_parity_map_scan_equals_the_scalar_fp64_chain(name):
    inc = _cases()[name]
    got, attempts, events = walk(inc)
    assert np.array_equal(got, scalar(inc)), name
    if name.startswith("flat") or name in ("vibrato", "glide", "gappy"):
        assert events <= 24                       # one real addition per power of two the total crosses


def test_markstein_division_by_the_sample_rate_is_exact():
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.uniform(20, 2000, 20000).astype(np.float32), (rng.uniform(0, 1, 5000) ** 4 * 1e-3).astype(np.float32),
                           np.arange(1, 3001, dtype=np.float32), -rng.uniform(1, 900, 3000).astype(np.float32),
                           np.array([1e-30, 3e-38, 1e30, 65504.0, 0.5, 1.0, 44100.0, 22050.0, 1e-5], dtype=np.float32)])
    for b in (44100.0, 48000.0, 100.0, 12.0):
        y = float(F(1) / F(b))
        for f in vals[:: (1 if b == 44100.0 else 7)].tolist():
            a = F(f)
            q0 = float(a * F(y))
            r = float(a - F(b) * F(q0))
            assert float(F(q0) + F(r) * F(y)) == float(a / F(b)), (f, b)


def test_fp64_delta_pairs_equal_the_integer_rule():
    """Wherever the kernel's fp64 formulation does not call a step an event, its delta pair and guard equal the
    integer rounding rule; and it calls a step an event only for increments of at least a quarter of the binade."""
    rng = np.random.default_rng(7)
    n_checked = 0
    for e in (-20, -8, -3, 0, 1, 5, 11, 30):
        mags = 2.0 ** rng.uniform(e - 70, e + 2, 6000)
        xs = np.concatenate([mags * rng.choice([-1.0, 1.0], mags.size),
                             # exact ties and near-ties at every shift: (k + 1/2) ulp, +- one unit in the last place of x
                             [(k + 0.5) * 2.0 ** (e - 52) * sg for k in (0, 1, 2, 3, 1000, 2 ** 30 + 1) for sg in (1, -1)],
                             [np.nextafter((k + 0.5) * 2.0 ** (e - 52), dirn) * sg for k in (1, 2, 7) for dirn in (0, np.inf) for sg in (1, -1)],
                             [2.0 ** (e - 52) * sg * k for k in (1, 2, 3) for sg in (1, -1)], [0.0]])
        for x in xs:
            x = float(x)
            a, b = delta(x, e), delta_fp64(x, e)
            if b[3]:
                assert x != 0 and (not np.isfinite(x) or abs(x) >= 2.0 ** (e - 1) or abs(x) < 2.0 ** -1022)
                continue
            assert not a[3]
            assert a[:2] == b[:2], (x, e, a, b)
            # the integer rule drops the guard of a negative increment 2^64 times below the total (the step cannot
            # change it); the fp64 form keeps ceil(|x| / ulp) = 1 there: an extra event exactly on a power of two
            mk, ek = _split(x) if x else (0, 0)
            assert a[2] == b[2] or (x < 0 and e - ek >= 64 and (a[2], b[2]) == (0, 1)), (x, e, a, b)
            n_checked += 1
    assert n_checked > 40000


@pytest.mark.parametrize("name", ["flat 220", "vibrato", "negative jitter", "saw", "power of two", "tiny"])
def test_walk_with_fp64_deltas_equals_the_scalar_chain(name):
    inc = _cases()[name]
    got, attempts, events = walk(inc, delta=delta_fp64)
    assert np.array_equal(got, scalar(inc))


# ---- gf_sg_scan_kernel (k_fx.cu): growl events as integer crossings of the EXACT running sum ----------------------
def _sg_scan(f0, mask, sr):
    """the kernel's arithmetic in Python integers: (events, flagged)"""
    n = len(f0)
    U = 88
    margin = (n + 2) << 12                       # units of 2^-64
    S, Fp, flagged, ev = 0, 0, False, []
    for i in range(n):
        f = np.float32(f0[i])
        sub = float(f) * 2.0
        if not (mask[i] > 0 and f > 0 and not sub < 1e-2):
            continue
        bits = struct.unpack("<Q", struct.pack("<d", sub / sr))[0]
        ex = (bits >> 52) & 0x7FF
        if ex < 1023 - 24 or ex > 1022:
            return [], True
        S += ((bits & (ONE52 - 1)) | ONE52) << (ex - 1075 + U)
        Fk = S >> U
        frac = (S & ((1 << U) - 1)) >> (U - 64)
        if frac <= margin or frac >= (1 << 64) - 1 - margin:
            flagged = True
        if Fk != Fp:
            if Fk != Fp + 1:
                flagged = True
            ev.append((i, sub))
        Fp = Fk
    return ev, flagged


def _sg_sequential(f0, mask, sr):
    """_detect_pulse_events with one ratio of 2.0 (GOOFER.py:672-698)"""
    phase, ev = 0.0, []
    for i in range(len(f0)):
        f = np.float32(f0[i])
        if mask[i] <= 0 or f <= 0:
            continue
        sub = float(f) * 2.0
        if sub < 1e-2:
            continue
        phase += sub / sr
        if phase >= 1.0:
            ev.append((i, sub))
            phase -= 1.0
    return ev


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_growl_events_are_crossings_of_the_exact_sum(seed):
    rng = np.random.default_rng(seed)
    n, sr = 30000, 44100
    t = np.arange(n) / sr
    base = 110.0 * 2 ** rng.uniform(0, 2) * (1 + 0.02 * np.sin(2 * np.pi * 5 * t))
    f0 = (base * (1 + 3.0 * np.sin(2 * np.pi * 75 * t))).astype(np.float32)       # apply_subharm_vibrato, depth 3
    mask = (rng.random(n) < 0.9).astype(np.float32)
    ev, flagged = _sg_scan(f0, mask, sr)
    assert not flagged
    assert ev == _sg_sequential(f0, mask, sr)


def test_growl_scan_flags_borderline_sums_instead_of_guessing():
    # a flat 441 Hz sub-harmonic at 44.1 kHz: the sum sits on an integer every 100 samples (within rounding)
    n, sr = 4000, 44100
    f0 = np.full(n, 220.5, dtype=np.float32)
    ev, flagged = _sg_scan(f0, np.ones(n, dtype=np.float32), sr)
    assert flagged
    # increments of 1 or more cannot be placed by the scan either
    assert _sg_scan(np.full(8, 30000.0, dtype=np.float32), np.ones(8, dtype=np.float32), sr)[1]


# ---- gf_walk_scan_kernel (k_excite.cu): pulse onsets from a truncated fixed-point sum, flagged when rounding could matter ----
def _walk_scan(f0, sr):
    """the kernel's arithmetic in Python integers: (onsets, flagged)"""
    U = 44
    A, R, flagged, ons = 0, 0, False, []
    for i, f in enumerate(np.asarray(f0, dtype=np.float32)):
        if f == 0:
            continue
        if not abs(f) < np.float32(0.25) * np.float32(sr):
            return [], True
        A += int(np.floor((float(f) / sr) * 2.0 ** U))
        if A < 0:
            return [], True
        Fk = A >> U
        Rn = max(R, Fk)
        k2 = i + 2
        lz = 64 - ((Rn + 1) << U).bit_length()
        m = k2 << max(0, 10 - lz)
        lo, hi = max(A - m, 0), A + m + k2
        if (lo >> U) != (hi >> U) and (hi >> U) > R:
            flagged = True
        if Rn > R:
            if not f > np.float32(1e-6):
                flagged = True
            ons += [(i, float(f))] * (Rn - R)
            R = Rn
    return ons, flagged


def _walk_sequential(f0, sr):
    """the onset part of pulse_train_numba (GOOFER.py:479-493)"""
    total, next_k, lv, ons = 0.0, 1.0, 160.0, []
    for i, f in enumerate(np.asarray(f0, dtype=np.float32)):
        if f > 1e-6:
            lv = float(f)
        total += float(f) / sr
        while total >= next_k:
            ons.append((i, lv))
            next_k += 1.0
    return ons


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_pulse_onsets_from_the_fixed_point_scan(seed):
    rng = np.random.default_rng(seed)
    n, sr = 44100, 44100
    t = np.arange(n) / sr
    f0 = (220.0 * 2 ** rng.uniform(-1, 2) * 2 ** (rng.uniform(0, 1) * np.sin(2 * np.pi * rng.uniform(3, 7) * t) / 12)).astype(np.float32)
    f0[rng.integers(0, n, 50)] = 0.0
    f0[10000:12000] = 0.0                                    # an unvoiced stretch
    if seed >= 2:                                            # f0 jitter beyond 100 %: the total runs backwards for a while
        f0[20000:20400] *= np.float32(-1.5)
        f0[30000:30050] *= np.float32(-0.3)
    ons, flagged = _walk_scan(f0, sr)
    assert not flagged
    assert ons == _walk_sequential(f0, sr)


def test_pulse_scan_flags_what_it_cannot_decide():
    sr = 44100
    # flat A4: 440 / 44100 * 2205 = 22 up to rounding -- the scan must hand the pass to the bit-exact walk
    assert _walk_scan(np.full(5000, 440.0, dtype=np.float32), sr)[1]
    # a negative total, NaN
    assert _walk_scan(np.array([100.0, -500.0, 100.0], dtype=np.float32), sr)[1]
    assert _walk_scan(np.array([100.0, np.nan], dtype=np.float32), sr)[1]
    # a flat note off the grid is decided by the scan
    f0 = np.full(20000, 261.6256, dtype=np.float32)
    ons, flagged = _walk_scan(f0, sr)
    assert not flagged and ons == _walk_sequential(f0, sr)
