"""Deterministic harness around the UNMODIFIED reference (test infrastructure only).

This file is part of the ORACLE: it is imported only by tests/, tests/golden/make_golden.py
and oracle validation scripts.  It never runs on the GPU box (``/root/reference`` does not
exist there) -- it exists to (a) pin ``oracle/goofer_oracle.py`` against the real reference and
(b) generate the golden fixtures under ``tests/golden/``.

What it does:
  * inserts stub modules for ``soundfile``, ``parselmouth``, ``sounddevice``, ``tkinter`` so that
    ``GOOFER.py:3,5`` / ``SillySampler.py:12`` / ``SillyEditor.py:6-8`` import in this container;
  * gives the ``soundfile`` stub an in-memory ``read``/``write`` so that
    ``GooferResampler.render`` (``SillySampler.py:415-447``) and ``sf.write`` (``:1185``) work;
  * replaces ``np.random.default_rng`` (``GOOFER.py:1151``, ``SillySampler.py:1063``) by a seeded
    factory and seeds the legacy global RNG (``GOOFER.py:653,666``) per note, recording every drawn
    buffer so the very same buffers can be handed to the CUDA path.
"""
from __future__ import annotations

import os
import sys
import types
import tempfile
import contextlib

import numpy as np

REFERENCE_DIR = os.environ.get("GOOFER_REFERENCE_DIR", "/root/reference")
# numba reads its configuration when it is first imported: cache=True at GOOFER.py:473 would otherwise
# write __pycache__/*.nbi into the (read-only) reference tree
os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "goofer_numba_cache"))
sys.dont_write_bytecode = True

_state = {"loaded": False, "gf": None, "ss": None}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "GOOFER.py"))


class _MemSoundFile(types.ModuleType):
    """In-memory stand-in for the two soundfile calls the render path makes."""

    def __init__(self):
        super().__init__("soundfile")
        self.files = {}      # path -> (float64 array, sr)
        self.written = {}    # path -> (array, sr)

    def read(self, path, *a, **k):
        y, sr = self.files[str(path)]
        return np.array(y, dtype=np.float64), int(sr)

    def write(self, path, data, sr, *a, **k):
        self.written[str(path)] = (np.array(data), int(sr))


def load_reference():
    """Import GOOFER and SillySampler from the reference tree (stubs installed first)."""
    if _state["loaded"]:
        return _state["gf"], _state["ss"]
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_DIR}")
    # numba cache=True at GOOFER.py:473 would otherwise write into the read-only reference tree
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "goofer_numba_cache"))
    sys.dont_write_bytecode = True
    sfmod = _MemSoundFile()
    sys.modules["soundfile"] = sfmod
    for name in ("parselmouth", "sounddevice", "tkinter", "tkinter.ttk"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["tkinter"].ttk = sys.modules["tkinter.ttk"]
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import logging
    import GOOFER as gf            # noqa: E402
    import SillySampler as ss      # noqa: E402
    logging.getLogger().setLevel(logging.WARNING)
    _state.update(loaded=True, gf=gf, ss=ss, sf=sfmod)
    return gf, ss


def soundfile_stub() -> _MemSoundFile:
    load_reference()
    return _state["sf"]


class NoiseRecorder:
    """Seeded replacement for the reference's unseeded RNG draws.

    k-th ``np.random.default_rng()`` call -> ``Generator(PCG64(base_seed + k))``.
    The legacy global RNG is seeded with ``legacy_seed`` (``np.random.seed``).
    """

    def __init__(self, base_seed: int, legacy_seed: int):
        self.base_seed = int(base_seed)
        self.legacy_seed = int(legacy_seed)
        self.n_rng_calls = 0

    def factory(self, *a, **k):
        g = np.random.Generator(np.random.PCG64(self.base_seed + self.n_rng_calls))
        self.n_rng_calls += 1
        return g


@contextlib.contextmanager
def seeded_noise(base_seed: int, legacy_seed: int):
    rec = NoiseRecorder(base_seed, legacy_seed)
    orig = np.random.default_rng
    np.random.default_rng = rec.factory
    np.random.seed(rec.legacy_seed)
    try:
        yield rec
    finally:
        np.random.default_rng = orig


def write_goofy(path, env_knots, f0_interp, voicing_mask, formants, sr, y_len):
    """Store features with the reference's own writer (``GOOFER.py:287-317``)."""
    gf, _ = load_reference()
    gf.save_features(path, env_knots, f0_interp, voicing_mask, formants, sr, y_len)


def render_note(goofy_path, args, base_seed, legacy_seed, taps=False):
    """Run ``GooferResampler(*args)`` (``SillySampler.py:285-1185``) on cached features.

    ``args`` = the 13 CLI strings; args[0] must be ``<stem>.wav`` with ``<stem>_features.goofy``
    next to it (= ``goofy_path``).  Returns (out float64 array, sr, dict of taps).
    """
    gf, ss = load_reference()
    sfm = soundfile_stub()
    in_wav = str(args[0])
    assert os.path.abspath(goofy_path) == os.path.abspath(in_wav[:-4] + "_features.goofy")
    # the render path reads the wav only for SE (SillySampler.py:420,434,583): any array will do
    sfm.files[in_wav] = (np.zeros(8, dtype=np.float64), 44100)
    captured = {"synth": []}
    orig_syn = gf.synthesize
    if taps:
        def spy(*a, **k):
            r = orig_syn(*a, **k)
            captured["synth"].append({
                "env": np.array(a[0], dtype=np.float32, copy=True),
                "f0": np.array(a[1], dtype=np.float64, copy=True),
                "mask": np.array(a[2], dtype=np.float32, copy=True),
                "kwargs": {kk: vv for kk, vv in k.items() if kk != "formants"},
                "formants": {kk: np.array(vv, copy=True) for kk, vv in (k.get("formants") or {}).items()},
                "out": [np.array(x, copy=True) for x in r],
            })
            return r
        gf.synthesize = spy
    try:
        with seeded_noise(base_seed, legacy_seed):
            ss.GooferResampler(*[str(a) for a in args])
    finally:
        gf.synthesize = orig_syn
    out, sr = sfm.written[str(args[1])]
    return np.asarray(out, dtype=np.float64), sr, captured
