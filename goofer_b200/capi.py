"""ctypes binding of include/goofer_b200.h -- the thin C-ABI the host code calls.

There is no CPU fallback: if libgoofer_b200.so is missing this module raises at load time, and every
compute entry point returns GOOFER_ERR_CUDA without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

GOOFER_OK = 0
ERR_INVALID, ERR_WORKSPACE, ERR_CUDA, ERR_NOTE = -1, -2, -3, -4
NOTE_OK, NOTE_EMPTY_TAIL, NOTE_TOO_SHORT, NOTE_BAD_SOURCE, NOTE_EDITOR = 0, 1, 2, 3, 4

# flag slots: order of the enum in include/goofer_b200.h
FLAG_NAMES = ["t", "g", "fa", "fb", "fc", "fd", "fw", "fst", "fsta", "fstb", "fstc", "fstd",
              "V", "B", "U", "sh", "sr", "st", "sg", "sd", "sj", "sa", "su", "br", "es", "pd",
              "FV", "L", "R", "P", "vf", "vh", "vl", "SE"]
FLAG_SLOT = {n: i for i, n in enumerate(FLAG_NAMES)}
GF_NFLAGS = len(FLAG_NAMES)
# GF_OVR_* of include/goofer_b200.h: continuous gf.synthesize keyword arguments (GooferNote.override_val)
OVERRIDES = ["formant_shift", "F1_shift", "F2_shift", "F3_shift", "F4_shift", "f0_jitter_strength",
             "volume_jitter_strength_harm", "volume_jitter_strength_breath", "normalize", "breath_strength", "uv_strength"]
OVERRIDE_SLOT = {n: i for i, n in enumerate(OVERRIDES)}
# SillySampler looks these up case-insensitively (SillySampler.py:309,346,384,391,399-405)
CASE_INSENSITIVE = {"se": "SE", "l": "L", "es": "es", "pd": "pd", "fst": "fst",
                    "fsta": "fsta", "fstb": "fstb", "fstc": "fstc", "fstd": "fstd"}


class GooferSource(C.Structure):
    _fields_ = [
        ("knots_log_f16", C.c_void_p), ("hz_knots", C.c_void_p), ("K", C.c_int32),
        ("env_dense", C.c_void_p), ("T", C.c_int32),
        ("mask", C.c_void_p), ("N", C.c_int32),
        ("formants", C.c_void_p * 4), ("formant_len", C.c_int32 * 4),
        ("sr", C.c_int32), ("ylen", C.c_int64),
    ]


class GooferNote(C.Structure):
    _fields_ = [
        ("source", C.c_int32), ("pitch_midi", C.c_int32), ("velocity", C.c_double),
        ("offset_s", C.c_double), ("length_s", C.c_double), ("consonant_s", C.c_double), ("cutoff_s", C.c_double),
        ("volume", C.c_double), ("tempo", C.c_double),
        ("bend_off", C.c_int64), ("bend_len", C.c_int32),
        ("flag", C.c_int32 * GF_NFLAGS), ("present", C.c_uint64),
        ("phi_off", C.c_int64 * 4), ("nrm_off", C.c_int64 * 4), ("out_off", C.c_int64),
        ("f0_off", C.c_int64),
        ("phi_rng", (C.c_uint64 * 4) * 4), ("phi_rng_mask", C.c_uint32), ("override_mask", C.c_uint32),
        ("override_val", C.c_double * 12),
    ]


class GooferNotePlanInfo(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("n_total", C.c_int32), ("t_out", C.c_int32), ("t_env", C.c_int32),
        ("n_passes", C.c_int32), ("need_phi", C.c_int32 * 4), ("need_nrm", C.c_int32 * 4),
    ]


class GooferBatch(C.Structure):
    _fields_ = [
        ("n_sources", C.c_int32), ("sources", C.POINTER(GooferSource)),
        ("n_notes", C.c_int32), ("notes", C.POINTER(GooferNote)),
        ("bend_cents", C.c_void_p), ("bend_total", C.c_int64),
        ("phi", C.c_void_p), ("phi_total", C.c_int64),
        ("normals", C.c_void_p), ("nrm_total", C.c_int64),
        ("out", C.c_void_p), ("out_total", C.c_int64),
        ("tap_harm", C.c_void_p), ("tap_uv", C.c_void_p), ("tap_bre", C.c_void_p),
        ("out_pcm16", C.c_void_p),
        ("f0_curves", C.c_void_p), ("f0_total", C.c_int64),
    ]


class GooferStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("waves", C.c_int32)]


class GooferError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libgoofer_b200 error {code}: {msg}")
        self.code = code


_lib = None

EXPORTS = [
    "goofer_version", "goofer_last_error", "goofer_plan_batch", "goofer_workspace_bytes", "goofer_render_batch",
    "goofer_render_batch_host", "goofer_host_release", "goofer_last_stats", "goofer_stft_batch", "goofer_istft_batch",
    "goofer_pulse_work_bytes", "goofer_pulse_train_batch", "goofer_onepole_batch", "goofer_debug_plan",
    "goofer_profile", "goofer_profile_summary", "goofer_struct_size",
    "goofer_analyse_work_bytes", "goofer_analyse_batch", "goofer_render_status",
]


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """dlopen the in-tree library (no build here: __graft_entry__.build() / goofer_b200._build.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.LIB_PATH):
        raise ImportError(f"{_build.LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  goofer_b200 has no CPU fallback.")
    L = C.CDLL(_build.LIB_PATH)
    vp, i32, i64, dbl, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_size_t
    L.goofer_version.restype = C.c_int
    L.goofer_last_error.restype = C.c_char_p
    L.goofer_plan_batch.restype = C.c_int
    L.goofer_plan_batch.argtypes = [C.POINTER(GooferBatch), C.POINTER(GooferNotePlanInfo)]
    L.goofer_workspace_bytes.restype = sz
    L.goofer_workspace_bytes.argtypes = [C.POINTER(GooferBatch), i32]
    L.goofer_render_batch.restype = C.c_int
    L.goofer_render_batch.argtypes = [C.POINTER(GooferBatch), vp, sz, vp]
    L.goofer_render_batch_host.restype = C.c_int
    L.goofer_render_batch_host.argtypes = [C.POINTER(GooferBatch)]
    L.goofer_render_status.restype = C.c_int
    L.goofer_render_status.argtypes = [vp, vp, C.POINTER(C.c_int32)]
    L.goofer_host_release.restype = None
    L.goofer_last_stats.restype = None
    L.goofer_last_stats.argtypes = [C.POINTER(GooferStats)]
    L.goofer_stft_batch.restype = C.c_int
    L.goofer_stft_batch.argtypes = [vp, i32, i32, vp, vp]
    L.goofer_istft_batch.restype = C.c_int
    L.goofer_istft_batch.argtypes = [vp, i32, i32, i32, vp, vp]
    L.goofer_pulse_work_bytes.restype = sz
    L.goofer_pulse_work_bytes.argtypes = [i32, i32]
    L.goofer_pulse_train_batch.restype = C.c_int
    L.goofer_pulse_train_batch.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    L.goofer_onepole_batch.restype = C.c_int
    L.goofer_onepole_batch.argtypes = [vp, vp, i32, i32, i32, dbl, i32, i32, vp, vp]
    L.goofer_debug_plan.restype = C.c_int
    L.goofer_debug_plan.argtypes = [C.POINTER(GooferBatch), i32, vp, sz]
    L.goofer_analyse_work_bytes.restype = sz
    L.goofer_analyse_work_bytes.argtypes = [i32, i32]
    L.goofer_analyse_batch.restype = C.c_int
    L.goofer_analyse_batch.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp]
    L.goofer_struct_size.restype = sz
    L.goofer_struct_size.argtypes = [C.c_int]
    for which, rec in enumerate((GooferSource, GooferNote, GooferNotePlanInfo, GooferBatch, GooferStats)):
        if int(L.goofer_struct_size(which)) != C.sizeof(rec):
            raise ImportError(f"ABI mismatch: sizeof({rec.__name__}) is {C.sizeof(rec)} in capi.py but "
                              f"{int(L.goofer_struct_size(which))} in libgoofer_b200.so")
    L.goofer_profile.restype = None
    L.goofer_profile.argtypes = [C.c_int]
    L.goofer_profile_summary.restype = C.c_char_p
    _lib = L
    return L


def check(rc: int):
    if rc != GOOFER_OK:
        raise GooferError(rc, load().goofer_last_error().decode("utf-8", "replace"))


def last_stats() -> dict:
    s = GooferStats()
    load().goofer_last_stats(C.byref(s))
    return {"kernel_launches": int(s.kernel_launches), "h2d_bytes": int(s.h2d_bytes), "d2h_bytes": int(s.d2h_bytes),
            "waves": int(s.waves)}


def profile(enable: bool, serial: bool = False) -> None:
    """Per-kernel CUDA-event timing; serial=True also keeps both preparation chains on one stream (spans add up)."""
    load().goofer_profile((2 if serial else 1) if enable else 0)


def profile_summary() -> dict:
    """{kernel: (launches, total_ms)} since profile(True)."""
    txt = load().goofer_profile_summary().decode()
    out = {}
    for part in txt.split(";"):
        if part:
            name, cnt, ms = part.split(":")
            out[name] = (int(cnt), float(ms))
    return out
