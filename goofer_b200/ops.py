"""Stage-level mirrors of the reference's DSP primitives, running as CUDA kernels on torch tensors.

    stft(x)             gf.stft               GOOFER.py:355-370   (n_fft 1024, hop 256, sqrt-Hann)
    istft(S, length)    gf.istft              GOOFER.py:392-413
    pulse_train(f0, sr) gf.pulse_train_numba  GOOFER.py:473-554
    onepole(...)        dynamic_butter_filter SillySampler.py:95-174
    analyse_envelope(y) envelope half of gf.extract_features + gf.compress_env_to_knots  GOOFER.py:940-946, 97-147
    synthesize(...)     gf.synthesize         GOOFER.py:971-1220  (the call SillyEditor.py:227, 559 and test.py:38 make)
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


def synthesize(env_spec, f0_interp, voicing_mask, y=None, sr: int = 44100, n_fft: int = N_FFT, hop_length: int = HOP, *,
               normalize: float = 1.0, uv_strength: float = 0.75, breath_strength: float = 0.1, pitch_shift: float = 1.0,
               formant_shift: float = 1.0, F1_shift: float = 1.0, F2_shift: float = 1.0,
               F3_shift: float = 1.0, F4_shift: float = 1.0, formants=None, f0_jitter: bool = False,
               f0_jitter_strength: float = 1.5, volume_jitter: bool = False, volume_jitter_strength_harm: float = 50,
               volume_jitter_strength_breath: float = 100, noise=None, device: str = "cuda:0", **unsupported):
    """gf.synthesize (GOOFER.py:971-1220) on the GPU: the direct call the editor preview (SillyEditor.py:227, 559) and
    test.py:38 make, same positional arguments, same return tuple (reconstruct, harmonic, aper_uv, aper_bre), each
    (len(voicing_mask),) float32.  `env_spec` is a (513, T) array or a knots dict; `y` is only looked at for its length
    in the reference and is ignored here (len(voicing_mask) plays that role).  One note rendered through
    goofer_render_batch with GooferNote.f0_off (include/goofer_b200.h).  The keyword arguments take the reference's
    continuous values: they travel as GooferNote.override_val (normalize, uv_strength, breath_strength, formant_shift,
    F1..F4_shift, f0_jitter_strength, volume_jitter_strength_harm / _breath); pitch_shift scales the f0 curve on the
    host exactly like GOOFER.py:995 (`f0_interp *= pitch_shift`, float32).  Keyword arguments whose defaults no caller
    in the reference changes for this path (stretch_factor, glottal_smoothing, roughness_*, subharm_*) stay unsupported.
    `noise`: a host.SeededNoise-like provider (default: fresh noise like the reference)."""
    import numpy as np
    from . import host
    defaults = {"stretch_factor": 1.0, "start_sec": None, "end_sec": None, "glottal_smoothing": False, "apply_brightness": True,
                "noise_transition_smoothness": 100, "f0_jitter_speed": 100, "volume_jitter_speed": 150, "volume_vibrato": False,
                "add_subharm": False, "roughness_on": False}
    bad = sorted(k for k, v in unsupported.items() if not (k in defaults and v == defaults[k]))
    if bad:
        raise NotImplementedError("ops.synthesize: keyword arguments no resampler flag reaches are not implemented: " + ", ".join(bad))
    if int(n_fft) != N_FFT or int(hop_length) != HOP:
        raise NotImplementedError("ops.synthesize: n_fft = 1024, hop_length = 256 only")
    mask = np.ascontiguousarray(voicing_mask, dtype=np.float32).reshape(-1)
    f0 = np.ascontiguousarray(f0_interp, dtype=np.float32).reshape(-1)
    if f0.size != mask.size:
        raise ValueError("len(f0_interp) must equal len(voicing_mask)")
    if float(pitch_shift) != 1.0:
        f0 = f0 * np.float32(pitch_shift)                     # to_compute'd f32 array times a Python float stays f32 (GOOFER.py:995)

    # the flag columns decide which stages run; the overrides carry the exact values
    flags = ""
    ovr = {"normalize": float(normalize), "breath_strength": float(breath_strength), "uv_strength": float(uv_strength)}
    if float(formant_shift) != 1.0:
        flags += "g1"
        ovr["formant_shift"] = float(formant_shift)
    for nm, key, val in (("fa", "F1_shift", F1_shift), ("fb", "F2_shift", F2_shift), ("fc", "F3_shift", F3_shift), ("fd", "F4_shift", F4_shift)):
        if float(val) != 1.0:
            flags += f"{nm}1"
            ovr[key] = float(val)
    if f0_jitter:
        flags += "sh1"
        ovr["f0_jitter_strength"] = float(f0_jitter_strength)
    if volume_jitter:
        flags += "sr1"
        ovr["volume_jitter_strength_harm"] = float(volume_jitter_strength_harm)
        ovr["volume_jitter_strength_breath"] = float(volume_jitter_strength_breath)
    forms = {}
    if isinstance(formants, dict):
        for k, v in formants.items():
            if isinstance(k, str) and k.upper().startswith("F"):
                try:
                    k = int(k[1:])
                except Exception:
                    continue
            if isinstance(k, int) and 1 <= k <= 4:
                forms[k] = np.asarray(v, dtype=np.float64).reshape(-1)
    if isinstance(env_spec, dict):
        T = int(np.asarray(env_spec["knot_vals_log"]).shape[1])
    else:
        env_spec = np.ascontiguousarray(env_spec, dtype=np.float32)
        T = int(env_spec.shape[1])
    for i in (1, 2, 3, 4):                                   # pad_trim_to_len (GOOFER.py:64-70, 999-1002)
        x = forms.get(i, np.zeros(1))
        x = np.zeros(T) if x.size == 0 else (np.pad(x, (0, T - x.size), mode="edge") if x.size < T else x[:T])
        forms[i] = np.ascontiguousarray(x, dtype=np.float64)
    if isinstance(env_spec, dict):
        src = host.SourceFeatures.from_knot_pack(env_spec, mask, forms, int(sr), int(mask.size))
    else:
        src = host.SourceFeatures.from_dense(env_spec, mask, forms, int(sr), int(mask.size))
    b = host.Batch()
    b.add_source(src)
    b.add_note(host.NoteArgs(source=0, pitch="C4", flags=flags, f0_curve=f0, overrides=ovr))
    ab = b.assemble(noise or host.FreshNoise(), taps=True)
    db = ab.to_device(device)
    db.render()
    out = db.outputs()[0]
    harm, uv, bre = (t[:ab.out_total].cpu().numpy() for t in db.tap)
    return out, harm, uv, bre
