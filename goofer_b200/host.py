"""Host side of the B200 render path: CLI-string parsing, feature loading, batch assembly.

Mirrors the *host* half of SillySampler.GooferResampler (/root/reference/SillySampler.py):
  parse_flags            :48-54     flag regex, last duplicate wins
  note_to_midi           :86-90
  pitch_string_to_cents  :56-84     base64 12-bit pairs with '#n#' run-length
  NoteArgs               :286-306   the 13 positional resampler arguments
  load_goofy             GOOFER.py:319-339  (.goofy = npz of fp16 knots / f0 / mask + pickled formants)
Everything numeric (slicing lengths excepted, they live in the C++ planner) happens on the GPU behind
goofer_render_batch; this module only packs arrays and descriptors.
"""
from __future__ import annotations

import os

import ctypes as C
import re
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import capi

N_BINS = 513
HOP = 256

_FLAG_RE = re.compile(r"([A-Za-z]{1,4})([+-]?\d+)?")
_NOTE_RE = re.compile(r"([A-G]#?)(-?\d+)")
_SEMI = {"C": 0, "C#": 1, "D": 2, "D#": 3, "E": 4, "F": 5, "F#": 6, "G": 7, "G#": 8, "A": 9, "A#": 10, "B": 11}
# flags whose value may be missing (the reference evaluates `x or 0` / `is not None` for them)
_NONE_OK = {"sh", "sr", "sd", "sj", "sa", "su", "es", "pd", "fst", "fsta", "fstb", "fstc", "fstd", "fw",
            "L", "SE", "FV", "R", "t"}


def parse_flags(flag_string: str) -> Dict[str, Optional[int]]:
    out: Dict[str, Optional[int]] = {}
    for name, num in _FLAG_RE.findall(flag_string.replace("/", "")):
        out[name] = int(num) if num else None
    return out


def note_to_midi(name: str) -> int:
    m = _NOTE_RE.match(name)
    if m is None:
        raise ValueError(f"Bad note '{name}'")
    return (int(m.group(2)) + 1) * 12 + _SEMI[m.group(1)]


def _sextet(ch: str) -> int:
    o = ord(ch)
    if o >= ord("a"):
        return o - ord("a") + 26
    if o >= ord("A"):
        return o - ord("A")
    if o >= ord("0"):
        return o - ord("0") + 52
    if ch == "+":
        return 62
    if ch == "/":
        return 63
    raise ValueError(f"Bad pitch-bend character '{ch}'")


def pitch_string_to_cents(s: str) -> np.ndarray:
    fields = s.split("#")
    vals: List[int] = []
    for i in range(0, len(fields), 2):
        body = fields[i]
        for j in range(0, len(body), 2):
            v = (_sextet(body[j]) << 6) | _sextet(body[j + 1])
            vals.append(v - 4096 if v & 0x800 else v)
        if i + 1 < len(fields):
            vals.extend([vals[-1]] * int(fields[i + 1]))
    if not vals:
        vals = [0]
    return np.asarray(vals, dtype=np.float32)


@dataclass
class SourceFeatures:
    """Cached features of one voicebank sample (what gf.load_features returns)."""
    mask: np.ndarray                          # (N,) float32
    formants: Dict[int, np.ndarray]           # {1..4: (T,) float64}
    sr: int
    ylen: int
    knots_log: Optional[np.ndarray] = None    # (K, T) float16
    hz_knots: Optional[np.ndarray] = None     # (K,) float32
    env_dense: Optional[np.ndarray] = None    # (513, T) float32
    cache_key: Optional[tuple] = None         # identity of the feature file (path, mtime, size): DeviceSourceCache

    @property
    def n_frames(self) -> int:
        return int(self.knots_log.shape[1] if self.knots_log is not None else self.env_dense.shape[1])

    @classmethod
    def from_knot_pack(cls, pack: dict, mask, formants, sr: int, ylen: int) -> "SourceFeatures":
        return cls(mask=np.ascontiguousarray(mask, dtype=np.float32), formants=_int_key_formants(formants), sr=int(sr),
                   ylen=int(ylen), knots_log=np.ascontiguousarray(pack["knot_vals_log"], dtype=np.float16),
                   hz_knots=np.ascontiguousarray(pack["hz_knots"], dtype=np.float32))

    @classmethod
    def from_dense(cls, env, mask, formants, sr: int, ylen: int) -> "SourceFeatures":
        return cls(mask=np.ascontiguousarray(mask, dtype=np.float32), formants=_int_key_formants(formants), sr=int(sr),
                   ylen=int(ylen), env_dense=np.ascontiguousarray(env, dtype=np.float32))


def _int_key_formants(formants) -> Dict[int, np.ndarray]:
    out: Dict[int, np.ndarray] = {}
    for k, v in (formants or {}).items():
        if isinstance(k, str) and k.upper().startswith("F"):
            try:
                k = int(k[1:])
            except ValueError:
                continue
        try:
            k = int(k)
        except (TypeError, ValueError):
            continue
        if 1 <= k <= 4:
            out[k] = np.ascontiguousarray(v, dtype=np.float64)
    return out


def load_goofy(path) -> SourceFeatures:
    """Read a `<stem>_features.goofy` written by the reference (npz; formants are a pickled dict)."""
    feat = _load_goofy(path)
    try:
        st = os.stat(path)
        feat.cache_key = (os.path.abspath(str(path)), st.st_mtime_ns, st.st_size)
    except OSError:
        pass
    return feat


class DeviceSourceCache:
    """Voicebank features kept resident in HBM across render calls (the on-disk format either side of
    the path): the fp16 knots, mel-knot frequencies, voicing mask and formant tracks of a `.goofy` file are uploaded
    once per (path, mtime, size) and device, and later batches point their GooferSource records at the same tensors.
    A server renders thousands of notes from a few hundred sources; the reference re-reads and re-decodes the file for
    every note (GOOFER.py:319-339)."""

    def __init__(self, max_sources: int = 4096):
        self.max_sources, self.entries, self.hits, self.misses = int(max_sources), {}, 0, 0

    def get(self, key, device, make):
        k = (key, str(device))
        e = self.entries.get(k)
        if e is None:
            self.misses += 1
            if len(self.entries) >= self.max_sources:
                self.entries.pop(next(iter(self.entries)))          # oldest first
            e = self.entries[k] = make()
        else:
            self.hits += 1
        return e


def _load_goofy(path) -> SourceFeatures:
    with np.load(path, allow_pickle=True) as z:
        mode = str(z["mode"][0])
        mask = z["voicing_mask"].astype(np.float32)
        formants = z["formants"].item()
        sr = int(z["sr"][0])
        ylen = int(z["y_len"][0])
        if mode == "knots":
            if int(z["n_bins"][0]) != N_BINS:
                raise ValueError("only n_fft = 1024 feature files are supported (SillySampler.py:14)")
            return SourceFeatures.from_knot_pack({"knot_vals_log": z["knot_vals_log"], "hz_knots": z["hz_knots"]},
                                                 mask, formants, sr, ylen)
        return SourceFeatures.from_dense(z["env_spec"].astype(np.float32), mask, formants, sr, ylen)


@dataclass
class NoteArgs:
    """The resampler arguments after the two wav paths (SillySampler.py:286-293 defaults)."""
    source: int
    pitch: str
    velocity: float = 100.0
    flags: str = ""
    offset: float = 0.0
    length: float = 1000.0
    consonant: float = 0.0
    cutoff: float = 0.0
    volume: float = 100.0
    modulation: float = 0.0
    tempo: str = "!120"
    pitch_string: str = "AA"
    # direct gf.synthesize call (GooferNote.f0_off): per-sample f0 curve; the source is then used whole and the
    # slicing / pitch arguments above are ignored (include/goofer_b200.h)
    f0_curve: Optional[np.ndarray] = None
    # continuous gf.synthesize keyword arguments that replace the flag-derived scalars (capi.OVERRIDES -> GooferNote.override_val)
    overrides: Optional[Dict[str, float]] = None

    @classmethod
    def from_cli(cls, source: int, args: Sequence[str]) -> "NoteArgs":
        a = list(args)
        if len(a) > 11:
            raise TypeError(f"Expected at most 11 arguments after the wav paths but got {len(a)}")
        names = ["pitch", "velocity", "flags", "offset", "length", "consonant", "cutoff", "volume", "modulation",
                 "tempo", "pitch_string"]
        return cls(source=source, **{n: v for n, v in zip(names, a)})

    def to_struct(self, bend_off: int) -> "tuple[capi.GooferNote, np.ndarray]":
        fl = parse_flags(str(self.flags))
        nt = capi.GooferNote()
        nt.source = int(self.source)
        nt.pitch_midi = note_to_midi(str(self.pitch))
        nt.velocity = float(self.velocity)
        nt.offset_s = float(self.offset) / 1000.0
        nt.length_s = float(self.length) / 1000.0
        nt.consonant_s = float(self.consonant) / 1000.0
        nt.cutoff_s = float(self.cutoff) / 1000.0
        nt.volume = float(self.volume) / 100.0
        float(self.modulation)                              # parsed, unused (SillySampler.py:304)
        nt.tempo = float(str(self.tempo).lstrip("!"))
        present = 0
        resolved: Dict[str, Optional[int]] = {}
        for key, val in fl.items():
            if key in capi.FLAG_SLOT and key.lower() not in capi.CASE_INSENSITIVE:
                resolved[key] = val
        for lower, canon in capi.CASE_INSENSITIVE.items():
            for key, val in fl.items():                     # first key in flag-string order wins
                if key.lower() == lower:
                    resolved[canon] = val
                    break
        for key, val in resolved.items():
            if val is None:
                if key not in _NONE_OK:
                    raise TypeError(f"flag '{key}' needs a number")
                continue                                    # `x or 0` / `is not None` semantics == absent
            slot = capi.FLAG_SLOT[key]
            nt.flag[slot] = int(val)
            present |= 1 << slot
        nt.present = present
        bend = pitch_string_to_cents(str(self.pitch_string))
        nt.bend_off = int(bend_off)
        nt.bend_len = int(bend.size)
        for k in range(4):
            nt.phi_off[k] = -1
            nt.nrm_off[k] = -1
        nt.f0_off = -1
        for name, val in (self.overrides or {}).items():
            k = capi.OVERRIDE_SLOT[name]
            nt.override_val[k] = float(val)
            nt.override_mask |= 1 << k
        return nt, bend


# ----------------------------------------------------------------------------------------------------
# noise providers: the reference draws unseeded noise (GOOFER.py:1151, :653, :666; SillySampler.py:1063);
# the host supplies the very same kind of buffers, seeded or not.
# ----------------------------------------------------------------------------------------------------
class SeededNoise:
    """Draws the buffers in the reference's own order from per-note seeds:
    legacy MT19937 randn x3 (sh, sr harm, sr breath), then k-th default_rng() -> PCG64(base + k) for
    main phases, su phases, sj normal, sj phases, sa phases."""

    def __init__(self, base_seed: Callable[[int], int] | int = 20000, legacy_seed: Callable[[int], int] | int = 777):
        self._base = base_seed if callable(base_seed) else (lambda i, b=base_seed: b)
        self._legacy = legacy_seed if callable(legacy_seed) else (lambda i, s=legacy_seed: s)

    def __call__(self, index: int, info: dict) -> dict:
        n, T = info["n_total"], info["t_out"]
        out = {}
        leg = np.random.RandomState(self._legacy(index))
        if info["need_nrm"][0]:
            out["sh"] = leg.randn(n)
        if info["need_nrm"][1]:
            out["sr_h"] = leg.randn(n)
            out["sr_b"] = leg.randn(n)
        k = [0]
        base = self._base(index)

        def gen():
            g = np.random.Generator(np.random.PCG64(base + k[0]))
            k[0] += 1
            return g

        def phases():
            return gen().uniform(0.0, 2.0 * np.pi, size=(N_BINS, T)).astype(np.float32)

        out["phi"] = phases()
        if info["need_phi"][1]:
            out["phi_su"] = phases()
        if info["need_phi"][2]:
            out["sj_z"] = gen().standard_normal(n)
            out["phi_sj"] = phases()
        if info["need_phi"][3]:
            out["phi_sa"] = phases()
        return out


class DeviceNoise(SeededNoise):
    """Same draws as SeededNoise, but the noise phases are NOT materialised on the host: for every phi slot the
    provider hands over the 128-bit state and increment of np.random.PCG64(base + k) and the library draws
    rng.uniform(0, 2 pi, (513, T)).astype(float32) on the device, bit for bit (GooferNote.phi_rng, gf_phi_kernel).
    The normals (sh / sr / sj: legacy MT19937 randn and Generator.standard_normal) stay host arrays."""

    def __call__(self, index: int, info: dict) -> dict:
        n = info["n_total"]
        out = {}
        leg = np.random.RandomState(self._legacy(index))
        if info["need_nrm"][0]:
            out["sh"] = leg.randn(n)
        if info["need_nrm"][1]:
            out["sr_h"] = leg.randn(n)
            out["sr_b"] = leg.randn(n)
        base = self._base(index)
        k = 0

        def rng_state():
            nonlocal k
            st = np.random.PCG64(base + k).state["state"]
            k += 1
            return int(st["state"]), int(st["inc"])

        out["phi_rng"] = rng_state()
        if info["need_phi"][1]:
            out["phi_su_rng"] = rng_state()
        if info["need_phi"][2]:
            out["sj_z"] = np.random.Generator(np.random.PCG64(base + k)).standard_normal(n)
            k += 1
            out["phi_sj_rng"] = rng_state()
        if info["need_phi"][3]:
            out["phi_sa_rng"] = rng_state()
        return out


class FreshNoise(SeededNoise):
    """Unseeded noise, like the reference CLI."""

    def __init__(self):
        ss = np.random.SeedSequence()
        words = ss.generate_state(2)
        super().__init__(base_seed=lambda i, w=int(words[0]): w + 16 * i, legacy_seed=lambda i, w=int(words[1]): (w + i) % (2 ** 32))


class FreshDeviceNoise(DeviceNoise):
    """Unseeded noise like the reference CLI, with the phases drawn on the device (what cli / server use: drawing
    88,749 uniforms per note and pass with numpy costs about a millisecond of host time per note)."""

    def __init__(self):
        ss = np.random.SeedSequence()
        words = ss.generate_state(2)
        super().__init__(base_seed=lambda i, w=int(words[0]): w + 16 * i, legacy_seed=lambda i, w=int(words[1]): (w + i) % (2 ** 32))


_PHI_KEYS = ("phi", "phi_su", "phi_sj", "phi_sa")
_NRM_KEYS = ("sh", "sr_h", "sr_b", "sj_z")


@dataclass
class Batch:
    """A render batch: sources + notes -> the GooferBatch descriptor of include/goofer_b200.h."""
    sources: List[SourceFeatures] = field(default_factory=list)
    notes: List[NoteArgs] = field(default_factory=list)

    def add_source(self, feat: SourceFeatures) -> int:
        self.sources.append(feat)
        return len(self.sources) - 1

    def add_note(self, note: NoteArgs) -> int:
        self.notes.append(note)
        return len(self.notes) - 1

    def plan(self) -> List[dict]:
        """goofer_plan_batch only (host arithmetic, no noise, no device): the lengths / pass counts of every note, e.g. to
        balance a batch over several GPUs before any shard is assembled."""
        lib = capi.load()
        n_notes = len(self.notes)
        src_arr = (capi.GooferSource * max(1, len(self.sources)))()
        for i, s in enumerate(self.sources):                     # the planner reads sizes only, never the arrays
            g = src_arr[i]
            g.T, g.N, g.sr, g.ylen = s.n_frames, int(s.mask.size), int(s.sr), int(s.ylen)
            g.K = int(s.knots_log.shape[0]) if s.knots_log is not None else 0
            for k in range(4):
                tr = s.formants.get(k + 1)
                if tr is not None and tr.size:
                    g.formants[k] = tr.ctypes.data
                    g.formant_len[k] = int(tr.size)
        note_arr = (capi.GooferNote * max(1, n_notes))()
        off = 0
        for i, nt in enumerate(self.notes):
            st, bend = nt.to_struct(off)
            if nt.f0_curve is not None:
                st.f0_off = 0
            note_arr[i] = st
            off += bend.size
        b = capi.GooferBatch()
        b.n_sources, b.sources, b.n_notes, b.notes = len(self.sources), src_arr, n_notes, note_arr
        b.bend_total = off
        info_arr = (capi.GooferNotePlanInfo * max(1, n_notes))()
        capi.check(lib.goofer_plan_batch(C.byref(b), info_arr))
        return [{"n_total": int(x.n_total), "t_out": int(x.t_out), "t_env": int(x.t_env), "n_passes": int(x.n_passes),
                 "need_phi": list(x.need_phi), "need_nrm": list(x.need_nrm)} for x in info_arr[:n_notes]]

    # ---- assembly ---------------------------------------------------------------------------------
    def assemble(self, noise: Callable[[int, dict], dict], taps: bool = False) -> "AssembledBatch":
        lib = capi.load()
        n_src, n_notes = len(self.sources), len(self.notes)
        src_arr = (capi.GooferSource * max(1, n_src))()
        keep = []
        for i, s in enumerate(self.sources):
            g = src_arr[i]
            if s.knots_log is not None:
                g.knots_log_f16 = s.knots_log.ctypes.data
                g.hz_knots = s.hz_knots.ctypes.data
                g.K = int(s.knots_log.shape[0])
            if s.env_dense is not None:
                g.env_dense = s.env_dense.ctypes.data
            g.T = s.n_frames
            g.mask = s.mask.ctypes.data
            g.N = int(s.mask.size)
            for k in range(4):
                tr = s.formants.get(k + 1)
                if tr is not None and tr.size:
                    g.formants[k] = tr.ctypes.data
                    g.formant_len[k] = int(tr.size)
            g.sr = int(s.sr)
            g.ylen = int(s.ylen)
        note_arr = (capi.GooferNote * max(1, n_notes))()
        bends = []
        off = 0
        for i, nt in enumerate(self.notes):
            st, bend = nt.to_struct(off)
            note_arr[i] = st
            bends.append(bend)
            off += bend.size
        bend_all = np.concatenate(bends) if bends else np.zeros(1, np.float32)
        f0_parts, f0_off = [], 0
        for i, nt in enumerate(self.notes):
            if nt.f0_curve is not None:
                c = np.ascontiguousarray(nt.f0_curve, dtype=np.float32).reshape(-1)
                if c.size != self.sources[nt.source].mask.size:
                    raise ValueError(f"note {i}: the f0 curve has {c.size} samples, the source's voicing mask "
                                     f"{self.sources[nt.source].mask.size} (gf.synthesize: len(f0_interp) == len(voicing_mask))")
                note_arr[i].f0_off = f0_off
                f0_parts.append(c)
                f0_off += c.size
        f0_all = np.concatenate(f0_parts) if f0_parts else None
        b = capi.GooferBatch()
        b.n_sources, b.sources = n_src, src_arr
        b.n_notes, b.notes = n_notes, note_arr
        b.bend_cents, b.bend_total = bend_all.ctypes.data, int(bend_all.size)
        if f0_all is not None:
            b.f0_curves, b.f0_total = f0_all.ctypes.data, int(f0_all.size)
        info_arr = (capi.GooferNotePlanInfo * max(1, n_notes))()
        capi.check(lib.goofer_plan_batch(C.byref(b), info_arr))
        infos = [{"n_total": int(x.n_total), "t_out": int(x.t_out), "t_env": int(x.t_env), "n_passes": int(x.n_passes),
                  "need_phi": list(x.need_phi), "need_nrm": list(x.need_nrm)} for x in info_arr[:n_notes]]
        phi_parts, nrm_parts = [], []
        phi_off = nrm_off = out_off = 0
        for i, inf in enumerate(infos):
            nz = noise(i, inf)
            st = note_arr[i]
            for k, key in enumerate(_PHI_KEYS):
                if inf["need_phi"][k] and key + "_rng" in nz:
                    state, inc = nz[key + "_rng"]               # drawn on the device from this PCG64 stream
                    m64 = (1 << 64) - 1
                    st.phi_rng[k][0], st.phi_rng[k][1] = (state >> 64) & m64, state & m64
                    st.phi_rng[k][2], st.phi_rng[k][3] = (inc >> 64) & m64, inc & m64
                    st.phi_rng_mask |= 1 << k
                elif inf["need_phi"][k]:
                    a = np.ascontiguousarray(nz[key], dtype=np.float32)
                    if a.shape != (N_BINS, inf["t_out"]):
                        raise ValueError(f"note {i}: noise['{key}'] must have shape ({N_BINS}, {inf['t_out']})")
                    st.phi_off[k] = phi_off
                    phi_parts.append(a.reshape(-1))
                    phi_off += a.size
            for k, key in enumerate(_NRM_KEYS):
                if inf["need_nrm"][k]:
                    a = np.ascontiguousarray(nz[key], dtype=np.float64)
                    if a.shape != (inf["n_total"],):
                        raise ValueError(f"note {i}: noise['{key}'] must have shape ({inf['n_total']},)")
                    st.nrm_off[k] = nrm_off
                    nrm_parts.append(a)
                    nrm_off += a.size
            st.out_off = out_off
            out_off += inf["n_total"]
        phi_all = np.concatenate(phi_parts) if phi_parts else np.zeros(1, np.float32)
        nrm_all = np.concatenate(nrm_parts) if nrm_parts else None
        ab = AssembledBatch(self, b, src_arr, note_arr, infos, bend_all, phi_all, nrm_all, out_off, taps)
        ab.f0_curves = f0_all
        return ab


class AssembledBatch:
    """Host arrays + descriptor, ready for goofer_render_batch_host or for upload through torch."""

    def __init__(self, batch, desc, src_arr, note_arr, infos, bend, phi, normals, out_total, taps):
        self.batch, self.desc, self.src_arr, self.note_arr, self.infos = batch, desc, src_arr, note_arr, infos
        self.bend, self.phi, self.normals, self.out_total, self.taps = bend, phi, normals, int(out_total), taps
        d = self.desc
        d.phi, d.phi_total = phi.ctypes.data, int(phi.size)
        if normals is not None:
            d.normals, d.nrm_total = normals.ctypes.data, int(normals.size)
        else:
            d.normals, d.nrm_total = None, 0
        d.out_total = self.out_total

    def pin(self) -> "AssembledBatch":
        """Move every host array the C library reads or writes into page-locked memory (torch owns it),
        so that goofer_render_batch_host's copies run at PCIe speed and asynchronously."""
        import torch
        keep = self._pinned = []

        def pinned(a: np.ndarray) -> np.ndarray:
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            keep.append(t)
            return t.numpy()

        d = self.desc
        self.phi = pinned(self.phi)
        d.phi = self.phi.ctypes.data
        self.bend = pinned(self.bend)
        d.bend_cents = self.bend.ctypes.data
        if self.normals is not None:
            self.normals = pinned(self.normals)
            d.normals = self.normals.ctypes.data
        if getattr(self, "f0_curves", None) is not None:
            self.f0_curves = pinned(self.f0_curves)
            d.f0_curves = self.f0_curves.ctypes.data
        for i, s in enumerate(self.batch.sources):
            g = self.src_arr[i]
            if s.knots_log is not None:
                g.knots_log_f16 = pinned(s.knots_log.view(np.int16)).ctypes.data
                g.hz_knots = pinned(s.hz_knots).ctypes.data
            if s.env_dense is not None:
                g.env_dense = pinned(s.env_dense).ctypes.data
            g.mask = pinned(s.mask).ctypes.data
            for k in range(4):
                tr = s.formants.get(k + 1)
                if tr is not None and tr.size:
                    g.formants[k] = pinned(tr).ctypes.data
        # the output buffers are page-locked on first use (render_host): a float render never pays for the PCM buffer and
        # vice versa (65,536 notes: 11.6 GB + 5.8 GB)
        self._out_bufs = None
        self._pcm_buf = None
        return self

    def _pinned_out(self, pcm16: bool):
        """Persistent output buffers of render_host: page-locked when pin() was called, plain numpy otherwise."""
        def make(dtype):
            a = np.empty(max(1, self.out_total), dtype=dtype)
            if getattr(self, "_pinned", None) is None:
                return a
            import torch
            t = torch.from_numpy(a).pin_memory()
            self._pinned.append(t)
            return t.numpy()

        if pcm16:
            if getattr(self, "_pcm_buf", None) is None:
                self._pcm_buf = make(np.int16)
            return self._pcm_buf
        if getattr(self, "_out_bufs", None) is None:
            self._out_bufs = [make(np.float32) for _ in range(4 if self.taps else 1)]
        return self._out_bufs

    def split(self, flat: np.ndarray) -> List[np.ndarray]:
        """Per-note views of a concatenated array.  The views of a buffer that outlives the call (the pinned output
        buffers) are made once: slicing 1,024 notes costs 0.6 ms of Python, a tenth of a whole render call."""
        pinned = getattr(self, "_pinned", None)               # the buffers pin() made live as long as this object
        key = (flat.__array_interface__["data"][0], flat.dtype.str, flat.size)
        cache = self.__dict__.setdefault("_split_cache", {})
        if pinned and key in cache:
            return list(cache[key])
        outs, off = [], 0
        for inf in self.infos:
            outs.append(flat[off:off + inf["n_total"]])
            off += inf["n_total"]
        mine = (getattr(self, "_out_bufs", None) or []) + ([self._pcm_buf] if getattr(self, "_pcm_buf", None) is not None else [])
        if pinned and any(key[0] == b.__array_interface__["data"][0] for b in mine):
            cache[key] = outs
            return list(outs)
        return outs

    # ---- host-buffer entry point (numpy in, numpy out; copies inside the C library) ---------------
    def render_host(self, pcm16: bool = False):
        """numpy in, numpy out through goofer_render_batch_host.  pcm16=True returns the notes as int16 PCM encoded
        on the device (what SillySampler.py:1185 writes to the .wav) and downloads half the bytes."""
        lib = capi.load()
        if pcm16:
            if self.taps:
                raise ValueError("stage taps are f32: render with pcm16=False")
            pcm = self._pinned_out(True)
            d = self.desc
            d.out, d.out_pcm16 = None, pcm.ctypes.data
            d.tap_harm = d.tap_uv = d.tap_bre = None
            try:
                capi.check(lib.goofer_render_batch_host(C.byref(d)))
            finally:
                d.out_pcm16 = None
            return self.split(pcm[:self.out_total])
        bufs = self._pinned_out(False)
        out = bufs[0]
        d = self.desc
        d.out = out.ctypes.data
        tap_arrays = None
        if self.taps:
            tap_arrays = bufs[1:4]
            d.tap_harm, d.tap_uv, d.tap_bre = (a.ctypes.data for a in tap_arrays)
        else:
            d.tap_harm = d.tap_uv = d.tap_bre = None
        capi.check(lib.goofer_render_batch_host(C.byref(d)))
        res = self.split(out[:self.out_total])
        if self.taps:
            return res, [self.split(a[:self.out_total]) for a in tap_arrays]
        return res

    # ---- device-resident entry point (torch owns the memory and the stream) ------------------------
    def to_device(self, device="cuda:0", source_cache: "Optional[DeviceSourceCache]" = None):
        return DeviceBatch(self, device, source_cache)


class DeviceBatch:
    """All arrays of an AssembledBatch resident in HBM as torch tensors; render() launches on torch's
    current stream and returns the flat f32 output tensor."""

    def __init__(self, ab: AssembledBatch, device, source_cache: "Optional[DeviceSourceCache]" = None):
        import torch
        self.torch = torch
        self.ab = ab
        self.device = torch.device(device)
        dev = self.device
        self._keep = []

        def up(a: np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=False)
            self._keep.append(t)
            return t

        src = ab.batch.sources
        self.src_arr = (capi.GooferSource * max(1, len(src)))()
        def upload_source(s):
            t = {}
            if s.knots_log is not None:
                t["knots"], t["hz"] = up(s.knots_log.view(np.int16)), up(s.hz_knots)
            if s.env_dense is not None:
                t["dense"] = up(s.env_dense)
            t["mask"] = up(s.mask)
            for k in range(4):
                tr = s.formants.get(k + 1)
                if tr is not None and tr.size:
                    t[f"F{k}"] = up(tr)
            return t

        for i, s in enumerate(src):
            g = self.src_arr[i]
            h = ab.src_arr[i]
            g.K, g.T, g.N, g.sr, g.ylen = h.K, h.T, h.N, h.sr, h.ylen
            if source_cache is not None and s.cache_key is not None:
                t = source_cache.get(s.cache_key, dev, lambda s=s: upload_source(s))
                self._keep.append(t)
            else:
                t = upload_source(s)
            if "knots" in t:
                g.knots_log_f16, g.hz_knots = t["knots"].data_ptr(), t["hz"].data_ptr()
            if "dense" in t:
                g.env_dense = t["dense"].data_ptr()
            g.mask = t["mask"].data_ptr()
            for k in range(4):
                if f"F{k}" in t:
                    g.formants[k] = t[f"F{k}"].data_ptr()
                    g.formant_len[k] = int(t[f"F{k}"].numel())
        self.bend = up(ab.bend)
        self.phi = up(ab.phi)
        self.normals = up(ab.normals) if ab.normals is not None else None
        self.f0_curves = up(ab.f0_curves) if getattr(ab, "f0_curves", None) is not None else None
        self.out = torch.empty(max(1, ab.out_total), dtype=torch.float32, device=dev)
        self.pcm = None                                    # int16 PCM copy of out, allocated by enable_pcm16()
        self.tap = [torch.empty_like(self.out) for _ in range(3)] if ab.taps else None
        d = capi.GooferBatch()
        h = ab.desc
        d.n_sources, d.sources = h.n_sources, self.src_arr
        d.n_notes, d.notes = h.n_notes, ab.note_arr
        d.bend_cents, d.bend_total = self.bend.data_ptr(), h.bend_total
        d.phi, d.phi_total = self.phi.data_ptr(), h.phi_total
        d.normals, d.nrm_total = (self.normals.data_ptr() if self.normals is not None else None), h.nrm_total
        d.out, d.out_total = self.out.data_ptr(), h.out_total
        if self.f0_curves is not None:
            d.f0_curves, d.f0_total = self.f0_curves.data_ptr(), h.f0_total
        if self.tap:
            d.tap_harm, d.tap_uv, d.tap_bre = (t.data_ptr() for t in self.tap)
        self.desc = d
        lib = capi.load()
        with torch.cuda.device(dev):
            ws = int(lib.goofer_workspace_bytes(C.byref(d), 0))
        if ws == 0:
            capi.check(capi.ERR_NOTE)
        self.workspace = torch.empty(ws, dtype=torch.uint8, device=dev)

    def render(self):
        torch = self.torch
        lib = capi.load()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            capi.check(lib.goofer_render_batch(C.byref(self.desc), self.workspace.data_ptr(), self.workspace.numel(),
                                               C.c_void_p(stream)))
        return self.out

    def status(self) -> int:
        """Waits for the stream and returns goofer_render_status: 0, or GOOFER_ERR_NOTE when a note's pulse list
        overflowed (capi.check() turns it into an exception naming the note)."""
        lib = capi.load()
        with self.torch.cuda.device(self.device):
            stream = self.torch.cuda.current_stream(self.device).cuda_stream
            return int(lib.goofer_render_status(self.workspace.data_ptr(), C.c_void_p(stream), None))

    def enable_pcm16(self) -> "DeviceBatch":
        """Also encode the output as 16-bit PCM on the device (GooferBatch.out_pcm16); read it with outputs_pcm16()."""
        if self.pcm is None:
            self.pcm = self.torch.empty(max(1, self.ab.out_total), dtype=self.torch.int16, device=self.device)
            self.desc.out_pcm16 = self.pcm.data_ptr()
        return self

    def outputs(self) -> List[np.ndarray]:
        flat = self.out[:self.ab.out_total].cpu().numpy()
        return self.ab.split(flat)

    def outputs_pcm16(self) -> List[np.ndarray]:
        if self.pcm is None:
            raise RuntimeError("call enable_pcm16() before render()")
        return self.ab.split(self.pcm[:self.ab.out_total].cpu().numpy())
