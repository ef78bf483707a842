"""EXPERIMENT (not used by the product): a bit-exact PARALLEL formulation of the fp64 phase walk.

Tried as a replacement for gf_walk_kernel's serial DADD chain in round 1: bit-identical on every test, but with one
warp per note (about two warps per SM sub-partition) the 64-bit pair scan costs more issue slots than the chain's
latency (1.74 ms vs 0.82 ms per 1,024 notes).  Kept because it would win with several samples per lane.

Original header: the arithmetic behind gf_walk_kernel (goofer_b200/csrc/k_excite.cu).

pulse_train_numba accumulates `total_phase += f0[i] / sr` in fp64, sample by sample (GOOFER.py:479-493), and the
rounding of that chain decides where pulses start.  The CUDA kernel does not run the chain serially: while the
exponent of the running total is fixed, one round-to-nearest-even addition is the integer map
M -> M + delta[M & 1]; those maps compose associatively, so the running mantissas come out of a scan.  This test
restates that scheme with Python integers and checks it bit for bit against the scalar fp64 loop on tie pitches,
glides, gaps, tiny and negative increments."""
import struct

import numpy as np
import pytest


def _bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def _split(x):
    b = _bits(x)
    e, m = (b >> 52) & 0x7FF, b & ((1 << 52) - 1)
    if e == 0:
        return (0, None) if m == 0 else (None, None)          # zero / subnormal
    return m | (1 << 52), e - 1023


def _then(a, b):
    """composition of two parity maps: first a, then b."""
    return (a[0] + b[(0 + a[0]) & 1], a[1] + b[(1 + a[1]) & 1])


def walk_scan(inc, group=32):
    n = len(inc)
    out = np.zeros(n)
    started, raw_mode, raw, M, e = False, False, 0.0, 0, 0
    pos = events = 0
    while pos < n:
        cnt = min(group, n - pos)
        deltas, ev = [], []
        for k in range(cnt):
            x = inc[pos + k]
            mk, ek = _split(x) if x >= 0 else (None, None)
            d, is_ev = (0, 0), False
            if mk == 0:
                pass                                            # zero increment: identity
            elif mk is None or not started or raw_mode or e - ek < 0:
                is_ev = True
            elif e - ek < 64:
                s = e - ek
                q, r = mk >> s, mk & ((1 << s) - 1)
                c = t = 0
                if s > 0:
                    half = 1 << (s - 1)
                    c, t = int(r > half), int(r == half)
                d = (q + c + t * (q & 1), q + c + t * ((q + 1) & 1))
            deltas.append(d)
            ev.append(is_ev)
        # Hillis-Steele inclusive scan, exactly as the warp does it
        F = list(deltas)
        o = 1
        while o < cnt:
            F = [F[k] if k < o else _then(F[k - o], F[k]) for k in range(cnt)]
            o <<= 1
        Mk = [M + F[k][M & 1] for k in range(cnt)]
        first = next((k for k in range(cnt) if ev[k] or (started and Mk[k] >= (1 << 53))), None)
        lim = cnt if first is None else first
        for k in range(lim):
            out[pos + k] = Mk[k] * 2.0 ** (e - 52) if started else (raw if raw_mode else 0.0)
        if lim > 0 and started:
            M = Mk[lim - 1]
        if first is None:
            pos += cnt
            continue
        prev = raw if raw_mode else (M * 2.0 ** (e - 52) if started else 0.0)
        tot = prev + inc[pos + first]                           # the one real fp64 addition
        out[pos + first] = tot
        events += 1
        mk, ek = _split(tot) if tot > 0 else (None, None)
        if tot > 0 and mk:
            started, raw_mode, M, e = True, False, mk, ek
        elif tot == 0:
            started, raw_mode, M = False, False, 0
        else:
            started, raw_mode, raw, M = False, True, tot, 0
        pos += first + 1
    return out, events


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
        f = np.full(30000, hz, dtype=np.float32)
        f[:3000] = 0
        c[f"flat {hz:g}"] = f
    c["vibrato"] = (220 * 2 ** (0.3 * np.sin(np.arange(40000) / 800.0) / 12)).astype(np.float32)
    c["glide"] = np.linspace(60, 900, 30000).astype(np.float32)
    c["gappy"] = (330 * (rng.random(30000) > 0.3)).astype(np.float32)
    c["tiny"] = (1e-5 * rng.random(20000)).astype(np.float32)
    c["negative jitter"] = (220 * (1 + 1.5 * np.sin(np.arange(30000) / 50.0))).astype(np.float32)
    c["negative start"] = np.concatenate([np.full(500, -80.0), np.full(8000, 300.0)]).astype(np.float32)
    return c


@pytest.mark.parametrize("name", list(_cases()))
def test_integer_scan_equals_the_scalar_fp64_chain(name):
    f = _cases()[name]
    inc = f.astype(np.float64) / 44100.0
    got, events = walk_scan(inc)
    assert np.array_equal(got, scalar(inc)), name
    if "negative" not in name:
        assert events <= 24                                    # one real addition per power of two the total crosses
