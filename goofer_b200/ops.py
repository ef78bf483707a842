"""Stage-level mirrors of the reference's DSP primitives, running as CUDA kernels on torch tensors.

    stft(x)             gf.stft               GOOFER.py:355-370   (n_fft 1024, hop 256, sqrt-Hann)
    istft(S, length)    gf.istft              GOOFER.py:392-413
    pulse_train(f0, sr) gf.pulse_train_numba  GOOFER.py:473-554
    onepole(...)        dynamic_butter_filter SillySampler.py:95-174
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
