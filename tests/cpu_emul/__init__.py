"""TEST-ONLY: builds tests/cpu_emul/emul.cpp (g++), which drives the __host__ __device__ index arithmetic of
goofer_b200/csrc (FFT passes, loop / stretch / velocity index maps) with serial loops so that it can be
compared with the oracle on a machine without a GPU.  Never part of the product."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_emul.so")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(HERE, "emul.cpp")
    csrc = os.path.join(HERE, "..", "..", "goofer_b200", "csrc")
    newest = max([os.path.getmtime(src)] + [os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc)])
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", SO, src])
    L = C.CDLL(SO)
    fp, dp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    L.emul_rfft1024.argtypes = [fp, fp]
    L.emul_irfft1024.argtypes = [fp, fp]
    L.emul_plan_size.restype = C.c_int
    L.emul_env_mix.restype = C.c_int
    L.emul_env_mix.argtypes = [vp, C.c_int, ip, dp]
    L.emul_mask_new.argtypes = [vp, fp, dp]
    L.emul_track_canon.argtypes = [vp, dp, C.c_int, fp]
    L.emul_fftconv_f32.argtypes = [fp, C.c_int, C.c_double, fp]
    L.emul_fftconv_f32.restype = C.c_int
    L.emul_fftconv_f64.argtypes = [dp, C.c_int, C.c_double, dp]
    L.emul_fftconv_f64.restype = C.c_int
    _lib = L
    return L
