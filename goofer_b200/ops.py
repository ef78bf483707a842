"""Stage-level mirrors of the reference's DSP primitives, running as CUDA kernels on torch tensors.

    stft(x)             gf.stft               GOOFER.py:355-370   (n_fft 1024, hop 256, sqrt-Hann)
    istft(S, length)    gf.istft              GOOFER.py:392-413
    pulse_train(f0, sr) gf.pulse_train_numba  GOOFER.py:473-554
    onepole(...)        dynamic_butter_filter SillySampler.py:95-174
    analyse_envelope(y) envelope half of gf.extract_features + gf.compress_env_to_knots  GOOFER.py:940-946, 97-147
torch is used for device memory and the stream only.
"""
from __future__ import annotations

import ctypes as C

from . import capi

N_FFT, HOP, N_BINS = 1024, 256, 513


def _stream(t):
    import torch
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _as2d(x):
    return x.reshape(1, -1) if x.dim() == 1 else x


def stft(x):
    """x: (n,) or (B, n) float32 CUDA tensor -> (513, T) / (B, 513, T) complex64, T = 1 + n // 256."""
    import torch
    lib = capi.load()
    x2 = _as2d(x).contiguous().float()
    B, n = x2.shape
    T = 1 + n // HOP
    S = torch.empty((B, N_BINS, T), dtype=torch.complex64, device=x.device)
    with torch.cuda.device(x.device):
        capi.check(lib.goofer_stft_batch(x2.data_ptr(), B, n, S.data_ptr(), _stream(x2)))
    return S[0] if x.dim() == 1 else S


def istft(S, length: int):
    """S: (513, T) or (B, 513, T) complex64 CUDA tensor -> (length,) / (B, length) float32."""
    import torch
    lib = capi.load()
    S3 = (S.unsqueeze(0) if S.dim() == 2 else S).contiguous().to(torch.complex64)
    B, nb, T = S3.shape
    if nb != N_BINS:
        raise ValueError("istft needs 513 frequency bins (n_fft = 1024)")
    y = torch.empty((B, int(length)), dtype=torch.float32, device=S.device)
    with torch.cuda.device(S.device):
        capi.check(lib.goofer_istft_batch(S3.data_ptr(), B, T, int(length), y.data_ptr(), _stream(S3)))
    return y[0] if S.dim() == 2 else y


def pulse_train(f0, sr: int = 44100):
    """f0: (n,) or (B, n) float32 CUDA tensor -> LF glottal pulse train of the same shape."""
    import torch
    lib = capi.load()
    f2 = _as2d(f0).contiguous().float()
    B, n = f2.shape
    out = torch.empty_like(f2)
    work = torch.empty(int(lib.goofer_pulse_work_bytes(B, n)), dtype=torch.uint8, device=f0.device)
    with torch.cuda.device(f0.device):
        capi.check(lib.goofer_pulse_train_batch(f2.data_ptr(), B, n, int(sr), out.data_ptr(), work.data_ptr(), _stream(f2)))
    return out[0] if f0.dim() == 1 else out


def onepole(x, f0, sr: int, cutoff_factor: float, order: int = 4, btype: str = "lowpass"):
    """dynamic_butter_filter: x, f0 (n,) or (B, n) float32 CUDA tensors."""
    import torch
    lib = capi.load()
    x2, f2 = _as2d(x).contiguous().float(), _as2d(f0).contiguous().float()
    if x2.shape != f2.shape:
        raise ValueError("x and f0 must have the same shape")
    B, n = x2.shape
    y = torch.empty_like(x2)
    with torch.cuda.device(x.device):
        capi.check(lib.goofer_onepole_batch(x2.data_ptr(), f2.data_ptr(), B, n, int(sr), float(cutoff_factor), int(order),
                                            0 if btype == "lowpass" else 1, y.data_ptr(), _stream(x2)))
    return y[0] if x.dim() == 1 else y


def analyse_envelope(y, sr: int = 44100):
    """y: (n,) or (B, n) float32 CUDA tensor of source waveforms -> list of knot packs
    {"mode", "knot_vals_log" (K, T) float16, "hz_knots" (K,) float32, "n_bins", "n_fft", "sr"} -- what
    gf.compress_env_to_knots returns and gf.save_features stores (numpy arrays on the host)."""
    import numpy as np
    import torch
    lib = capi.load()
    y2 = _as2d(y).contiguous().float()
    B, n = y2.shape
    T = 1 + n // HOP
    knots = torch.empty((B, 192, T), dtype=torch.float16, device=y.device)
    hz = torch.empty((B, 192), dtype=torch.float32, device=y.device)
    K = torch.empty((B,), dtype=torch.int32, device=y.device)
    work = torch.empty(int(lib.goofer_analyse_work_bytes(B, n)), dtype=torch.uint8, device=y.device)
    with torch.cuda.device(y.device):
        capi.check(lib.goofer_analyse_batch(y2.data_ptr(), B, n, int(sr), knots.data_ptr(), hz.data_ptr(), K.data_ptr(),
                                            work.data_ptr(), _stream(y2)))
    Kh = K.cpu().numpy()
    knots_h, hz_h = knots.cpu().numpy(), hz.cpu().numpy()
    return [{"mode": "knots", "knot_vals_log": np.ascontiguousarray(knots_h[b, :Kh[b]]), "hz_knots": np.ascontiguousarray(hz_h[b, :Kh[b]]),
             "n_bins": N_BINS, "n_fft": N_FFT, "sr": int(sr)} for b in range(B)]
