"""Drop-in for the UTAU resampler command line of SillySampler.py (/root/reference/SillySampler.py:1226-1275):

    python -m goofer_b200.cli in.wav out.wav pitch velocity flags offset length consonant cutoff volume
                              modulation !tempo pitch_string

Same 13 positional arguments, same defaults, same log messages and exit codes (0 / 1 with the usage text on
a TypeError).  The cached features `<in stem>_features.goofy` (GOOFER.py:287-339) must exist: extracting them
needs Praat (third-party, out of scope -- DESIGN.md section 7).  The note is rendered on cuda:0 through the C ABI
and written as 16-bit PCM like soundfile's default for .wav (SillySampler.py:1185).
`render_notes()` is the batch form of the same call: many argument lists, one GPU launch sequence.
"""
from __future__ import annotations

import logging
import sys
import wave
from pathlib import Path
from typing import List, Sequence

import numpy as np

from . import host

VERSION = "goofer_b200 (SillySampler v2.6.1 CLI surface)"
HELP = ("Usage:\n"
        "  python -m goofer_b200.cli in.wav out.wav pitch velocity flags\n"
        "           offset(ms) length(ms) consonant(ms) cutoff(ms)\n"
        "           volume(%) modulation(%) !tempo pitch_string\n\n"
        "Example:\n"
        "  python -m goofer_b200.cli in.wav out.wav C4 100 g0 0 1000 0 700 100 0 !120 AA")


def feature_path(in_wav) -> Path:
    p = Path(in_wav)
    return p.with_name(f"{p.stem}_features.goofy")


def pcm16_like_soundfile(samples: np.ndarray) -> np.ndarray:
    """float -> int16 the way `sf.write(path.wav, x, sr)` does it (SillySampler.py:1185): python-soundfile turns
    libsndfile's clipping on, and pcm.c d2s_clip_array then computes saturate(lrint(x * 2^31)) >> 16.  Host-side
    restatement of the device encoder (k_tail.cu gf_pcm16); the CLI itself receives PCM from the device."""
    s = np.asarray(samples, dtype=np.float64) * 2147483648.0
    v = np.rint(np.clip(s, -2147483648.0, 2147483647.0)).astype(np.int64) >> 16
    v = np.where(s >= 2147483647.0, 0x7FFF, v)
    return np.clip(v, -0x8000, 0x7FFF).astype("<i2")


def write_wav_pcm16(path, samples: np.ndarray, sr: int) -> None:
    """Write a mono 16-bit wav; `samples` is int16 PCM (from the device) or float (encoded like soundfile would)."""
    a = np.asarray(samples)
    pcm = a.astype("<i2") if a.dtype == np.int16 else pcm16_like_soundfile(a)
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(sr))
        w.writeframes(pcm.tobytes())


# feature files already read / already resident in HBM, shared by every render_notes call of the process (the server)
_HOST_FEATURES: dict = {}
SOURCE_CACHE = host.DeviceSourceCache()


def _features(path: Path) -> "host.SourceFeatures":
    st = path.stat()
    key = (str(path.resolve()), st.st_mtime_ns, st.st_size)
    f = _HOST_FEATURES.get(key)
    if f is None:
        if len(_HOST_FEATURES) >= 4096:
            _HOST_FEATURES.pop(next(iter(_HOST_FEATURES)))
        f = _HOST_FEATURES[key] = host.load_goofy(path)
    return f


_MULTI = {}                                                  # tuple(devices) -> MultiGpuRenderer (persistent worker threads)


def default_devices() -> List[str]:
    """GOOFER_DEVICES="0,1,2,3" (or "all") selects the GPUs render_notes / the server shard a batch over; default cuda:0."""
    import os
    env = os.environ.get("GOOFER_DEVICES", "").strip()
    if not env:
        return ["cuda:0"]
    if env == "all":
        import torch
        return [f"cuda:{i}" for i in range(torch.cuda.device_count())]
    return [f"cuda:{int(x)}" for x in env.split(",") if x.strip()]


def render_notes(arg_lists: Sequence[Sequence[str]], noise=None, device: str = "cuda:0", pcm16: bool = False,
                 devices: "Sequence[str] | None" = None) -> List[np.ndarray]:
    """Render many resampler invocations (each the 13 CLI strings) as ONE batch; returns the sample arrays
    (float32, or int16 PCM encoded on the device with pcm16=True).  devices=[...] (more than one) shards the batch by
    note over those GPUs, one host thread per GPU (goofer_b200.multi) -- same samples either way."""
    batch = host.Batch()
    src_index = {}
    for args in arg_lists:
        if len(args) < 13:
            raise TypeError(f"Expected 13 arguments but got {len(args)}")
        feat = feature_path(args[0])
        if not feat.exists():
            raise FileNotFoundError(f"{feat} not found: feature extraction needs Praat and is out of scope of goofer_b200")
        key = str(feat.resolve())
        if key not in src_index:
            src_index[key] = batch.add_source(_features(feat))
        batch.add_note(host.NoteArgs.from_cli(src_index[key], list(args[2:13])))
    if devices is not None and len(devices) > 1 and len(batch.notes) > 1:
        from . import multi
        key = tuple(str(d) for d in devices)
        if key not in _MULTI:
            _MULTI[key] = multi.MultiGpuRenderer(key, source_cache=SOURCE_CACHE)
        return _MULTI[key].render(batch, noise or host.FreshDeviceNoise(), pcm16=pcm16)
    if devices is not None and len(devices) == 1:
        device = devices[0]
    ab = batch.assemble(noise or host.FreshDeviceNoise())      # phases drawn on the device (GooferNote.phi_rng)
    db = ab.to_device(device, source_cache=SOURCE_CACHE)
    if pcm16:
        db.enable_pcm16()
    db.render()
    host.capi.check(db.status())
    return db.outputs_pcm16() if pcm16 else db.outputs()


def main(argv: Sequence[str]) -> int:
    logging.basicConfig(format="%(message)s", level=logging.INFO)
    logging.info(VERSION)
    args = list(argv)
    if not args:
        # no arguments: server mode, like the reference (SillySampler.py:1238-1240 -> run(), :1220-1224)
        from . import server
        try:
            server.run()
        except TypeError:
            logging.info(HELP)
        return 0
    logging.info(f"Args: {args} (count={len(args)})")
    try:
        if len(args) < 13:
            raise TypeError(f"Expected 13 arguments but got {len(args)}")
        logging.info("Loading cached features")
        logging.info("Synthesizing")
        out = render_notes([args[:13]], pcm16=True, devices=default_devices()[:1])[0]
        sr = _features(feature_path(args[0])).sr
        logging.info(f"Writing {args[1]}")
        write_wav_pcm16(args[1], out, sr)
    except TypeError as e:
        logging.error("Argument parsing failed: %s", str(e))
        logging.error(HELP)
        return 1
    except Exception:
        logging.exception("Failed to render")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
