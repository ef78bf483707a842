"""ORACLE: CPU restatement of ``SillySampler.GooferResampler`` (SillySampler.py:285-1185).

TEST INFRASTRUCTURE ONLY (see oracle/dsp.py header).  Parity status: pinned by execution against
the unmodified reference (tests/test_oracle_vs_reference.py) and by tests/golden/.

Form differs from the reference on purpose (this is a restatement, not a copy): the CLI strings are
parsed into a ``NoteSpec`` once, features arrive as a ``Features`` record instead of being read from
disk, the noise the reference would draw is passed in (see ``noise_for_note``), and the result is
returned instead of written.  Draw order of the reference for a full-flag note:
legacy randn(N) x3 (sh, sr harm, sr breath); default_rng() x5 (main phi, su phi, sj normal, sj phi, sa phi).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np

from . import dsp
from .synth import synthesize

HOP = dsp.HOP

_SEMITONE = {"C": 0, "C#": 1, "D": 2, "D#": 3, "E": 4, "F": 5, "F#": 6, "G": 7, "G#": 8, "A": 9, "A#": 10, "B": 11}
_NOTE = re.compile(r"([A-G]#?)(-?\d+)")
_FLAG = re.compile(r"([A-Za-z]{1,4})([+-]?\d+)?")


def parse_flag_string(s: str) -> dict:
    """SillySampler.py:48-54."""
    return {k: (int(v) if v else None) for k, v in _FLAG.findall(s.replace("/", ""))}


def midi_of(name: str) -> int:
    """SillySampler.py:86-90."""
    m = _NOTE.match(name)
    if not m:
        raise ValueError(f"Bad note '{name}'")
    return (int(m.group(2)) + 1) * 12 + _SEMITONE[m.group(1)]


def _b64(c: str) -> int:
    o = ord(c)
    if o >= 97:
        return o - 71
    if o >= 65:
        return o - 65
    if o >= 48:
        return o + 4
    if o == 43:
        return 62
    if o == 47:
        return 63
    raise ValueError(f"Bad b64 '{c}'")


def bend_cents(s: str) -> np.ndarray:
    """SillySampler.py:56-84: base64 12-bit signed pairs with '#n#' run-length."""
    def pairs(ps):
        out = []
        for i in range(0, len(ps), 2):
            p = ps[i:i + 2]
            v = (_b64(p[0]) << 6) | _b64(p[1])
            out.append(v - 4096 if (v & 0x800) else v)
        return out
    parts = s.split("#")
    vals = []
    for i in range(0, len(parts), 2):
        chunk = parts[i:i + 2]
        vals += pairs(chunk[0])
        if len(chunk) == 2:
            vals += [vals[-1]] * int(chunk[1])
    a = np.array(vals, dtype=np.float32)
    return a if a.size else np.array([0.0], dtype=np.float32)


def _ci(flags: dict, name: str, default=0):
    """Case-insensitive lookup used for SE, L, es, pd, fst* (SillySampler.py:309,346,384,391,399-405)."""
    return next((v for k, v in flags.items() if k.lower() == name), default)


@dataclass
class NoteSpec:
    """Everything GooferResampler.__init__ derives from the 13 CLI strings (SillySampler.py:286-410)."""
    pitch_m: int
    velocity: float
    flags: dict
    offset: float
    length: float
    consonant: float
    cutoff: float
    volume: float
    tempo: float
    bend: np.ndarray
    d: dict = field(default_factory=dict)

    @classmethod
    def from_cli(cls, pitch, velocity, flags="", offset=0, length=1000, consonant=0, cutoff=0,
                 volume=100, modulation=0, tempo="!120", pitch_string="AA"):
        fl = parse_flag_string(flags)
        self = cls(midi_of(pitch), float(velocity), fl, float(offset) / 1000.0, float(length) / 1000.0,
                   float(consonant) / 1000.0, float(cutoff) / 1000.0, float(volume) / 100.0,
                   float(tempo.lstrip("!")), bend_cents(pitch_string))
        float(modulation)                                          # parsed, unused (:304)
        g = fl.get
        d = self.d
        d["formant_shift"] = 1.0 + (g("g", 0) / 200.0)             # :313
        d["brightness_env"] = (g("br", 0) + 100) / 100.0           # :316
        d["F_shifts"] = tuple(1.0 + (g(k, 0) / 100.0) for k in ("fa", "fb", "fc", "fd"))   # :319-322
        sh = g("sh", None)
        d["f0_jitter"] = sh is not None and sh > 0                 # :325-327
        d["f0_jitter_strength"] = (sh or 0) / 50.0
        srv = g("sr", None)
        d["volume_jitter"] = srv is not None and srv > 0           # :328-330
        d["volume_jitter_strength"] = (srv or 0) / 50.0
        d["sd"] = float(g("sd", None) or 0)                        # :333-334
        d["B"] = (g("B", 0) + 100) / 100.0                         # :337
        d["U"] = (g("U", 0) + 100) / 100.0                         # :340
        d["V"] = np.clip(g("V", 100), 0, 100) / 100.0              # :343
        lk = next((k for k in fl if k.lower() == "l"), None)       # :346-358
        d["loop"] = {1: "avg", 2: "stretch"}.get(fl[lk], "concat") if lk else "concat"
        d["tension"] = g("st", 0) / 100.0                          # :361
        sg = g("sg", 0)
        d["subharm_weight"] = (sg / 100.0) * 1.5                   # :364-366
        d["add_subharm"] = sg > 0
        d["reverse"] = g("R", 0) == 1                              # :369
        d["sj"] = np.clip(g("sj", 0) or 0, 0, 100) / 100.0         # :372
        d["sa"] = np.clip(g("sa", 0) or 0, 0, 100) / 100.0         # :375
        d["su"] = np.clip(g("su", 0) or 0, 0, 100) / 100.0         # :378
        d["normalize"] = (np.clip(fl["P"], 0, 100) / 100.0) if "P" in fl else 1.0   # :381
        d["es"] = float(np.clip(_ci(fl, "es") or 0, -100, 100)) / 100.0              # :384-385
        d["FV"] = g("FV", 0) == 1                                  # :388
        d["pd"] = float(int(np.clip(_ci(fl, "pd") or 0, -100, 100))) / 100.0         # :391-393
        d["fw"] = ((g("fw", 0) or 0) / 100.0) * 0.1                # :396
        fst = float(np.clip(_ci(fl, "fst") or 0, -100, 100)) / 100.0                 # :399-400
        d["fst"] = tuple(float(np.clip(fst + ((_ci(fl, "fst" + c) or 0) / 100.0), -1.0, 1.0)) for c in "abcd")
        d["SE"] = _ci(fl, "se") == 1                               # :309-310 (GUI: out of scope)
        return self


@dataclass
class Features:
    """What gf.load_features returns (GOOFER.py:319-339), envelope already decoded to (513, T) f32."""
    env: np.ndarray
    mask: np.ndarray          # (N,) f32 (f16-quantised values)
    formants: dict            # {1..4: (T_f,) float}
    sr: int
    ylen: int


def noise_for_note(spec: NoteSpec, n_total: int, T_out: int, base_seed: int, legacy_seed: int) -> dict:
    """The buffers the reference draws under oracle/ref_harness.seeded_noise(base_seed, legacy_seed),
    in the reference's own draw order, as standalone arrays (what the host hands to the CUDA path)."""
    d = spec.d
    leg = np.random.RandomState(legacy_seed)
    out = {}
    if d["f0_jitter"]:
        out["sh"] = leg.randn(n_total)
    if d["volume_jitter"]:
        out["sr_h"] = leg.randn(n_total)
        out["sr_b"] = leg.randn(n_total)
    k = 0

    def gen():
        nonlocal k
        g = np.random.Generator(np.random.PCG64(base_seed + k))
        k += 1
        return g

    def phases():
        return gen().uniform(0.0, 2.0 * np.pi, size=(dsp.N_BINS, T_out)).astype(np.float32)

    out["phi"] = phases()
    if d["su"] > 0.0:
        out["phi_su"] = phases()
    if d["sj"] > 0.0:
        out["sj_z"] = gen().standard_normal(n_total)    # rng.normal(0, s) == 0 + s * standard_normal
        out["phi_sj"] = phases()
    if d["sa"] > 0.0:
        out["phi_sa"] = phases()
    return out


# ----------------------------------------------------------------------------------------------
def _tilt(sr: int, n_bins: int, brightness_env: float) -> np.ndarray:
    """SillySampler.py:506-510 (br)."""
    f = np.linspace(1e-6, sr * 0.5, n_bins, dtype=np.float32)
    nf = np.clip(f / (sr * 0.5), 0.02, 1.0)
    a = np.clip(brightness_env - 1.0, -0.9, 1.0)
    t = nf ** a
    t /= (t.mean() + 1e-12)
    return t


def _env_shape(block: np.ndarray, es: float) -> np.ndarray:
    """SillySampler.py:518-551 (es): blur + mean-match (es<0) or unsharp mask + mean-match (es>0)."""
    if not block.size:
        return block
    s = abs(es)

    def mean_match(orig, mod):
        m0 = np.mean(orig, axis=0, keepdims=True)
        m1 = np.mean(mod, axis=0, keepdims=True)
        return (mod * (m0 / (m1 + 1e-12))).astype(orig.dtype)

    if es < 0.0:
        blur = dsp.gauss1d(block, 1.0 + 6.0 * s, axis=0)
        return np.maximum(0.0, mean_match(block, blur))
    blur = dsp.gauss1d(block, 0.8 + 4.0 * s, axis=0)
    out = np.maximum(0.0, block + (5 * s) * (block - blur))
    return mean_match(block, out)


def _bandwidth_warp(env: np.ndarray, amount: float) -> np.ndarray:
    """SillySampler.py:555-569 (fw)."""
    nb = env.shape[0]
    pos = np.clip((np.arange(nb, dtype=np.float64) - nb / 2.0) * (1.0 + amount) + nb / 2.0, 0, nb - 1)
    lo = np.floor(pos).astype(int)
    hi = np.minimum(lo + 1, nb - 1)
    fr = (pos - lo)[:, None]
    out = np.empty_like(env)
    out[...] = (1 - fr) * env[lo, :] + fr * env[hi, :]
    return out


def _loop_env(tail: np.ndarray, want: int, mode: str, n_bins: int) -> np.ndarray:
    """SillySampler.py:628-696."""
    have = tail.shape[1]
    if have >= want:
        return tail[:, :want]
    reps, rem = want // have, want % have            # ZeroDivisionError when have == 0, like :634
    if mode == "stretch":
        return dsp.stretch_rows(tail, want / have)
    if mode == "avg":
        tile = (tail + tail[:, ::-1]) / 2.0
        return np.concatenate([tile] * reps + ([tile[:, :rem]] if rem else []), axis=1)
    # concat: 8-frame linear cross-fade between successive copies (Appendix B of SURVEY.md)
    chain = [tail.copy()]
    for _ in range(reps - 1):
        fade = min(8, have // 2)
        if fade == 0:                      # have == 1: prev[:, :-0] is empty in the reference (:668)
            chain[-1] = tail.copy()
            chain.append(tail.copy())
            continue
        up = np.linspace(0, 1, fade)[None, :]
        dn = np.linspace(1, 0, fade)[None, :]
        prev = chain[-1]
        xf = prev[:, prev.shape[1] - fade:] * dn + tail[:, :fade] * up
        chain[-1] = np.concatenate([prev[:, :prev.shape[1] - fade], xf, tail[:, fade:]], axis=1)
        chain.append(tail.copy())
    if rem:
        last = tail[:, :rem]
        prev = chain[-1]
        fade = min(8, rem // 2)
        if fade > 0:
            up = np.linspace(0, 1, fade)[None, :]
            dn = np.linspace(1, 0, fade)[None, :]
            xf = prev[:, prev.shape[1] - fade:] * dn + last[:, :fade] * up
            chain[-1] = np.concatenate([prev[:, :prev.shape[1] - fade], xf, last[:, fade:]], axis=1)
        else:
            chain[-1] = np.concatenate([prev, last], axis=1)
    return np.concatenate(chain, axis=1)


def _loop_samples(x: np.ndarray, want: int) -> np.ndarray:
    """SillySampler.py:699-712: plain tiling for every loop mode."""
    n = len(x)
    if n >= want:
        return x[:want]
    reps, rem = want // n, want % n
    return np.concatenate([x] * reps + ([x[:rem]] if rem else []))


def _loop_track(track, want: int, mode: str) -> np.ndarray:
    """SillySampler.py:717-744."""
    tr = np.asarray(track, dtype=np.float32)
    if tr.size == 0:
        return np.zeros(want, dtype=np.float32)
    if mode == "stretch":
        return dsp.stretch_rows(tr, want / float(tr.size)).astype(np.float32)
    reps, rem = want // tr.size, want % tr.size
    tile = (tr + tr[::-1]) * 0.5 if mode == "avg" else tr
    base = np.tile(tile, reps)
    if rem > 0:
        base = np.concatenate([base, tile[:rem]])
    return base.astype(np.float32)


def _prefix_positions(n: int, pre_len: int, factor: float):
    """SillySampler.py:176-204 shared index arithmetic; None when the stretch is a no-op."""
    if pre_len <= 1 or n <= 1 or abs(factor - 1.0) < 1e-6:
        return None
    pre_new = max(1, int(round(pre_len * factor)))
    idx = np.arange(pre_new + (n - pre_len), dtype=np.float64)
    return np.where(idx < pre_new, idx / factor, (idx - pre_new) + pre_len)


def _stretch_prefix(x: np.ndarray, pre_len: int, factor: float) -> np.ndarray:
    pos = _prefix_positions(x.shape[-1], pre_len, factor)
    if pos is None:
        return x
    grid = np.arange(x.shape[-1], dtype=np.float64)
    if x.ndim == 1:
        return dsp.lerp_extrap(grid, x, pos)
    return np.stack([dsp.lerp_extrap(grid, row, pos) for row in x], axis=0)


def _clean_track(track, T: int, sr: int, min_hz: float, sigma_frames: float) -> np.ndarray:
    """SillySampler.py:264-283 sanitize_smooth_formant."""
    x = np.array(track, dtype=np.float32)
    if len(x) < T:
        x = np.pad(x, (0, T - len(x)), mode="edge")
    elif len(x) > T:
        x = x[:T]
    bad = (~np.isfinite(x)) | (x < min_hz) | (x > sr * 0.48)
    if np.any(bad):
        good = np.where(~bad)[0]
        if good.size:
            x[bad] = dsp.lerp_extrap(good.astype(np.float32), x[~bad], np.where(bad)[0].astype(np.float32))
        else:
            x = np.full_like(x, 300.0)
    if sigma_frames > 0:
        x = dsp.gauss1d(x, sigma_frames)
    return x.astype(np.float32)


def _strength_gain(tracks, strengths, n_bins: int, T: int, sr: int) -> np.ndarray:
    """SillySampler.py:808-830 (fst): product of per-formant Gaussian bells, all f32."""
    f = np.linspace(0.0, sr / 2.0, n_bins, dtype=np.float32)[:, None]
    gain = np.ones((n_bins, T), dtype=np.float32)
    for Ft, s, sig in zip(tracks, strengths, (100.0, 200.0, 350.0, 500.0)):
        if abs(s) < 1e-6:
            continue
        ok = np.isfinite(Ft) & (Ft > 50.0) & (Ft < sr * 0.5)
        if not np.any(ok):
            continue
        c = Ft[None, ok]
        w = np.exp(-0.5 * ((f - c) / np.float32(sig)) ** 2).astype(np.float32)
        gain[:, ok] *= np.float32(1.0) + np.float32((1.0 + s) - 1.0) * w
    return gain


def _fry_f0(f0: np.ndarray, mask: np.ndarray, vf: float, vh: float, vl: float) -> None:
    """SillySampler.py:890-934 (in place on the fp64 f0)."""
    n = len(f0)
    L = int(round(n * (abs(vf) / 100.0)))
    if L <= 0:
        return
    glide = int(np.clip(int(round(L * (vl / 100.0))), 0, L))
    const = L - glide
    base = vh * (mask > 0)
    if vf > 0:
        if const > 0:
            f0[:const] = base[:const]
        if glide > 0:
            w = np.linspace(0.0, 1.0, glide, endpoint=True)
            f0[const:L] = (1.0 - w) * base[const:L] + w * f0[const:L]
    else:
        s = n - L
        if glide > 0:
            w = np.linspace(1.0, 0.0, glide, endpoint=True)
            f0[s:s + glide] = (1.0 - w) * base[s:s + glide] + w * f0[s:s + glide]
        if const > 0:
            f0[s + glide:n] = base[s + glide:n]


def _fry_mask(n: int, vf: float, sr: int):
    """SillySampler.py:937-965."""
    mid = n // 2
    if vf > 0:
        L = int(round(mid * (vf / 100.0)))
        a, b = 0, max(0, min(n, L))
    else:
        L = int(round((n - mid) * (abs(vf) / 100.0)))
        a, b = max(0, n - L), n
    if b <= a:
        return None
    m = np.zeros(n, dtype=np.float32)
    m[a:b] = 1.0
    fade = int(0.01 * sr)
    if fade > 0:
        a1 = min(b, a + fade)
        if a1 > a:
            m[a:a1] *= np.linspace(0.0, 1.0, a1 - a, endpoint=True)
        b0 = max(a, b - fade)
        if b > b0:
            m[b0:b] *= np.linspace(1.0, 0.0, b - b0, endpoint=True)
    return m


def _fry_env(env: np.ndarray, fry_mask: np.ndarray) -> None:
    """SillySampler.py:970-995: frames under the fry mask are compressed towards DC by up to 8 %."""
    nb, T = env.shape
    centers = np.minimum(len(fry_mask) - 1, np.arange(T) * HOP + HOP // 2).astype(int)
    wfr = fry_mask[centers]
    bins = np.arange(nb, dtype=np.float64)
    for j in np.nonzero(wfr > 1e-6)[0]:
        s = 1.0 - float(wfr[j]) * (1.0 - 0.92)
        if abs(s - 1.0) < 1e-6:
            continue
        src = np.clip(bins / s, 0.0, nb - 1.0)
        lo = np.floor(src).astype(np.int32)
        hi = np.minimum(lo + 1, nb - 1)
        fr = src - lo
        col = env[:, j]
        env[:, j] = (1.0 - fr) * col[lo] + fr * col[hi]


def resample(feat: Features, spec: NoteSpec, noise, taps=None) -> np.ndarray:
    """SillySampler.py:415-447 (reverse) + :449-1185.  ``noise`` is a dict (see noise_for_note) or a
    callable ``noise(n_total, T_out) -> dict`` (lengths are only known after the slicing logic).
    Returns the fp64 output array the reference hands to sf.write."""
    d = spec.d
    if d["SE"]:
        raise NotImplementedError("SE (Tk voicing editor) is out of scope")
    sr, ylen = feat.sr, feat.ylen
    env_src = np.array(feat.env, dtype=np.float32)
    mask_src = np.array(feat.mask, dtype=np.float32)
    forms = {k: np.asarray(v) for k, v in feat.formants.items()}
    if d["reverse"]:                                               # :438-444
        env_src = env_src[:, ::-1]
        mask_src = mask_src[::-1]
        forms = {k: v[::-1] for k, v in forms.items()}

    dur = ylen / sr                                                # :453-476
    end_base = (spec.offset - spec.cutoff) if spec.cutoff < 0 else (dur - spec.cutoff)
    if d["reverse"]:
        Lsec = end_base - spec.offset
        off_u = dur - end_base
        cut_u = dur - (off_u + Lsec)
    else:
        off_u, cut_u = spec.offset, spec.cutoff
    s0 = int(off_u * sr)
    s1 = s0 + int(spec.consonant * sr)
    s2 = int(((off_u - cut_u) if cut_u < 0 else (dur - cut_u)) * sr)
    f0_, f1_, f2_ = s0 // HOP, s1 // HOP, s2 // HOP                # :485-487

    env_pre, env_tail = env_src[:, f0_:f1_], env_src[:, f1_:f2_]   # :494-500
    mask_pre, mask_tail = mask_src[s0:s1], mask_src[s1:s2]
    n_bins = env_src.shape[0]

    if d["brightness_env"] != 1.0 and (env_pre.size or env_tail.size):          # :503-515
        tilt = _tilt(sr, n_bins, d["brightness_env"])[:, None].astype(np.float32)
        env_pre = env_pre * tilt
        env_tail = env_tail * tilt
    if d["es"] != 0.0 and (env_pre.size or env_tail.size):                      # :518-551
        env_pre, env_tail = _env_shape(env_pre, d["es"]), _env_shape(env_tail, d["es"])
    if d["fw"] != 0.0 and env_src.size:                                         # :554-574
        if env_pre.size:
            env_pre = _bandwidth_warp(env_pre, d["fw"])
        if env_tail.size:
            env_tail = _bandwidth_warp(env_tail, d["fw"])
    if d["FV"]:                                                                 # :619-623
        mask_pre, mask_tail = np.ones_like(mask_pre), np.ones_like(mask_tail)

    want_samples = int(spec.length * sr)                                        # :625-629
    want_frames = int(np.ceil(spec.length * sr / HOP))
    mode = d["loop"]
    env_loop = _loop_env(env_tail, want_frames, mode, n_bins)
    mask_loop = _loop_samples(mask_tail, want_samples)
    tracks = {}
    for k, v in forms.items():                                                  # :714-749
        tracks[k] = np.concatenate([v[f0_:f1_], _loop_track(v[f1_:f2_], want_frames, mode)])

    env_new = np.concatenate([env_pre, env_loop], axis=1)                       # :752-754
    mask_new = np.concatenate([mask_pre, mask_loop])
    T0_frames = env_new.shape[1]
    for k, f in tracks.items():                                                 # :756-763
        tracks[k] = np.pad(f, (0, T0_frames - len(f)), mode="edge") if len(f) < T0_frames else f[:T0_frames]

    vel = float(2.0 ** (1.0 - (spec.velocity / 100.0)))                         # :766-788
    pre_frames, pre_samples = env_pre.shape[1], len(mask_pre)
    if abs(vel - 1.0) > 1e-6 and pre_frames > 1 and pre_samples > 1:
        env_new = _stretch_prefix(env_new, pre_frames, vel)
        Tn = env_new.shape[1]
        for k, tr in tracks.items():
            f = _stretch_prefix(np.asarray(tr, dtype=np.float64), pre_frames, vel)
            tracks[k] = np.pad(f, (0, Tn - len(f)), mode="edge") if len(f) < Tn else f[:Tn]
        mask_new = _stretch_prefix(mask_new, pre_samples, vel)

    canon = {}                                                                  # :792, 242-262
    for k, v in tracks.items():
        a = np.asarray(v, dtype=np.float32)
        a = np.pad(a, (0, T0_frames - len(a)), mode="edge") if len(a) < T0_frames else a[:T0_frames]
        canon[f"F{k}"] = a
    T = env_new.shape[1]
    if any(abs(s) >= 1e-6 for s in d["fst"]):                                   # :793-832
        clean = [_clean_track(canon.get(f"F{i + 1}", np.zeros(T)), T, sr, mn, 4)
                 for i, mn in enumerate((120.0, 300.0, 1500.0, 2000.0))]
        env_new = env_new * _strength_gain(clean, d["fst"], n_bins, T, sr)

    n_total = len(mask_new)                                                     # :836-855
    t_s = np.arange(n_total) / sr
    semis = spec.bend.astype(np.float64) / 100.0 + spec.pitch_m
    t_cents = spec.flags.get("t", 0)
    if t_cents:
        semis = semis + (t_cents / 100.0)
    ticks = np.arange(len(semis)) * (60.0 / (spec.tempo * 96.0))
    midi = dsp.lerp_extrap(ticks, semis, np.clip(t_s, ticks[0], ticks[-1]))
    f0_new = mask_new * (440.0 * 2 ** ((midi - 69) / 12))

    dyn = None
    if d["pd"] != 0.0:                                                          # :858-881
        base = spec.pitch_m + ((spec.flags.get("t", 0) or 0) / 100.0)
        dev = dsp.gauss1d((midi - base).astype(np.float32), max(1, int(0.010 * sr)))
        ref = float(np.percentile(np.abs(dev), 95)) + 1e-8
        v = np.clip(dev / ref, -1.0, 1.0)
        db = (12.0 * abs(d["pd"])) * (v if d["pd"] > 0 else -v)
        dyn = np.clip(np.power(10.0, db / 20.0).astype(np.float32), 1e-3, 1e3)
        dyn = 1.0 + (dyn - 1.0) * dsp.gauss1d(mask_new.astype(np.float32), int(0.01 * sr))

    vf = float(spec.flags.get("vf", 0))                                         # :884-965
    vh = max(1.0, float(spec.flags.get("vh", 50)))
    vl = float(np.clip(float(spec.flags.get("vl", 15)), 0.0, 100.0))
    fry = None
    if vf != 0:
        vf = float(np.clip(vf, -100.0, 100.0))
        _fry_f0(f0_new, mask_new, vf, vh, vl)
        fry = _fry_mask(len(f0_new), vf, sr)
    if fry is not None and env_new.size:                                        # :967-995
        env_new = np.array(env_new)
        _fry_env(env_new, fry)

    T_out = 1 + n_total // HOP
    if callable(noise):
        noise = noise(n_total, T_out)
    common = dict(formant_shift=d["formant_shift"], formants=canon, F_shifts=d["F_shifts"],
                  normalize=d["normalize"])
    if taps is not None:
        taps.update(env_new=np.array(env_new), f0_new=f0_new.copy(), mask_new=np.array(mask_new),
                    formants={k: v.copy() for k, v in canon.items()}, n_total=n_total)
        taps["main"] = {}
    _, harm, uv, bre = synthesize(                                              # :1006-1035
        env_new, f0_new, mask_new, n_total, sr, noise,
        f0_jitter=d["f0_jitter"], f0_jitter_strength=d["f0_jitter_strength"],
        volume_jitter=d["volume_jitter"], volume_jitter_strength_harm=d["volume_jitter_strength"],
        volume_jitter_strength_breath=d["volume_jitter_strength"] * 2,
        add_subharm=d["add_subharm"], subharm_weight=d["subharm_weight"], subharm_semitones=12,
        subharm_vibrato=True, subharm_vibrato_rate=75, subharm_vibrato_depth=3, subharm_vibrato_delay=0.01,
        taps=None if taps is None else taps["main"], **common)

    def hp12(x, f0c):                                                           # :1052-1058, 1078-1080
        x = dsp.dyn_onepole(x, f0c, sr, 1.0, order=6, btype="highpass")
        return dsp.dyn_onepole(x, f0c, sr, 1.0, order=6, btype="highpass")

    if d["su"] > 0.0:                                                           # :1038-1059
        _, h2, _, _ = synthesize(env_new, f0_new * 0.5, mask_new, n_total, sr, {"phi": noise["phi_su"]}, **common)
        harm += hp12(h2, np.maximum(f0_new, 120.0)) * d["su"]
    if d["sj"] > 0.0:                                                           # :1062-1081
        z = 0.0 + (d["sj"] ** 2) * np.asarray(noise["sj_z"], dtype=np.float64)
        _, h3, _, _ = synthesize(env_new, f0_new * (0.5 * (2.0 ** z)), mask_new, n_total, sr,
                                 {"phi": noise["phi_sj"]}, **common)
        harm = (1.0 - d["sj"]) * harm + d["sj"] * hp12(h3, np.maximum(f0_new, 120.0))
    if fry is not None:                                                         # :1084-1098
        one = np.ones_like(f0_new)
        hh = dsp.dyn_onepole(harm, one, sr, 200, order=6, btype="highpass")
        bh = dsp.dyn_onepole(bre, one, sr, 200, order=6, btype="highpass")
        harm = harm * (1.0 - fry) + hh * fry
        bre = bre * (1.0 - fry) + bh * fry
    if d["sd"] > 0:                                                             # :1102-1112
        env_j = _tremolo(len(bre), sr, d["sd"])
        bre *= 1.0 + (env_j - 1.0) * dsp.gauss1d(mask_new.astype(float), 20)   # in place: stays f32
        bre *= 1.0 + (d["sd"] / 100.0) * 10
    ten = d["tension"]
    if ten != 0:                                                                # :1115-1140
        r0 = dsp.rms(harm + bre)
        a = abs(ten)
        if ten < 0:
            order = int(np.clip(int(np.round(1 + (a * 4))), 1, 6))
            harm = dsp.dyn_onepole(harm, f0_new, sr, 2.0 - a * 0.75, order=order, btype="lowpass")
            bre = dsp.dyn_onepole(bre, f0_new, sr, a, order=4, btype="highpass")
        else:
            hi = dsp.dyn_onepole(harm, f0_new, sr, a * 4, order=4, btype="highpass")
            harm = harm + hi * (1.0 + a * 20.0)
            bre = dsp.dyn_onepole(bre, f0_new, sr, (2.0 - a) / 0.5, order=6, btype="lowpass")
            bre = bre * (1.0 - a)
        r1 = dsp.rms(harm + bre)
        if r1 > 0:
            harm = harm * (r0 / r1)
            bre = bre * (r0 / r1)

    out = ((harm * d["V"] + bre * d["B"]) + uv * d["U"]) * spec.volume          # :1143-1151
    if d["sa"] > 0.0:                                                           # :1153-1172
        _, _, uv_u, bre_u = synthesize(env_new, f0_new, np.ones_like(mask_new), n_total, sr,
                                       {"phi": noise["phi_sa"]}, uv_strength=1.0, breath_strength=1.0,
                                       noise_transition_smoothness=1, **common)
        out = out * (1.0 - d["sa"]) + ((uv_u + bre_u) * spec.volume) * d["sa"]
    if dyn is not None:                                                         # :1174-1182
        out = out * dyn
    return np.asarray(out, dtype=np.float64)


def _tremolo(n: int, sr: int, sd: float) -> np.ndarray:
    from .synth import tremolo_curve
    return tremolo_curve(n, sr, 150.0, sd / 200.0)
