"""goofer_b200 -- B200-native (sm_100a) render path of GOOFER / SillySampler behind a C ABI.

Host code (this package) parses the UTAU resampler arguments and packs batches; every numeric array is
produced by hand-written CUDA kernels in goofer_b200/csrc (libgoofer_b200.so, see include/goofer_b200.h).
There is no CPU fallback: importing works without a GPU, rendering does not.
"""
from . import capi, host, shard                                    # noqa: F401
from .host import (Batch, NoteArgs, SourceFeatures, SeededNoise, FreshNoise, load_goofy, parse_flags,  # noqa: F401
                   note_to_midi, pitch_string_to_cents)

__all__ = ["capi", "host", "shard", "Batch", "NoteArgs", "SourceFeatures", "SeededNoise", "FreshNoise", "load_goofy",
           "parse_flags", "note_to_midi", "pitch_string_to_cents"]
