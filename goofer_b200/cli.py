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


def write_wav_pcm16(path, samples: np.ndarray, sr: int) -> None:
    pcm = np.clip(np.rint(np.asarray(samples, dtype=np.float64) * 32767.0), -32768, 32767).astype("<i2")
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(sr))
        w.writeframes(pcm.tobytes())


def render_notes(arg_lists: Sequence[Sequence[str]], noise=None, device: str = "cuda:0") -> List[np.ndarray]:
    """Render many resampler invocations (each the 13 CLI strings) as ONE batch; returns the sample arrays."""
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
            src_index[key] = batch.add_source(host.load_goofy(feat))
        batch.add_note(host.NoteArgs.from_cli(src_index[key], list(args[2:13])))
    ab = batch.assemble(noise or host.FreshNoise())
    db = ab.to_device(device)
    db.render()
    return db.outputs()


def main(argv: Sequence[str]) -> int:
    logging.basicConfig(format="%(message)s", level=logging.INFO)
    logging.info(VERSION)
    args = list(argv)
    logging.info(f"Args: {args} (count={len(args)})")
    try:
        if len(args) < 13:
            raise TypeError(f"Expected 13 arguments but got {len(args)}")
        logging.info("Loading cached features")
        logging.info("Synthesizing")
        out = render_notes([args[:13]])[0]
        sr = host.load_goofy(feature_path(args[0])).sr
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
