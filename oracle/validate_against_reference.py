"""ORACLE pinning run (build container only; needs /root/reference).

Renders the same notes through the UNMODIFIED reference (oracle/ref_harness.py) and through the
oracle restatement with identical features and identical noise buffers; prints the max-abs
difference per case.  ``python -m oracle.validate_against_reference [--quick]``.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

from . import ref_harness, resampler, sources

CASES = [
    # (name, source idx, seconds, cli args after the two paths)
    ("default_C4", 0, 1.0, ["C4", "100", "", "0", "1000", "0", "0", "100", "0", "!120", "AA"]),
    ("default_A3_tie", 0, 1.0, ["A3", "100", "", "0", "1000", "0", "0", "100", "0", "!120", "AA"]),
    ("formant", 1, 1.0, ["A3", "100", "g-20fa10fb-10fc5fd-5fw20fst30br20es40", "0", "1000", "0", "0", "100", "0", "!120", "AA"]),
    ("formant2", 2, 1.0, ["D4", "100", "g35fa-9fb8fc-7fd9fw-60fst-50fsta20fstd-30br-70es-60", "0", "800", "0", "0", "90", "0", "!120", "AA"]),
    ("fric_cons", 3, 1.0, ["E4", "100", "", "20", "700", "150", "100", "100", "0", "!120", "AA"]),
    ("velocity", 3, 1.0, ["E4", "60", "B20U-20V90", "20", "700", "150", "100", "100", "0", "!120", "AA"]),
    ("velocity150", 3, 1.0, ["F4", "150", "P50", "10", "600", "180", "-700", "80", "0", "!150", "AA"]),
    ("bend", 0, 1.0, ["A4", "100", "t15", "0", "1000", "0", "0", "100", "0", "!120", "AAAB#3#ACADAFAIALAOAQASATATASAQAOALAIAFADACAB#20#"]),
    ("full", 4, 1.0, ["A3", "100", "B20U-20V90sh30sr30sg40sd20sj30sa20su40vf30st30pd50", "0", "1000", "0", "0", "100", "0", "!120", "AAABACAEAGAIAKAMAOAQ#40#AOAKAGACAA#30#"]),
    ("full_neg", 5, 1.0, ["G3", "100", "sh60sr50sg80sd70sj50sa60su30vf-40vh35vl40st-60pd-80FV1", "0", "900", "0", "0", "100", "0", "!120", "AA///+/9/7/5/3#30#/5/9AA#40#"]),
    ("L0_long", 0, 1.0, ["C4", "100", "L0", "100", "2500", "150", "300", "100", "0", "!120", "AA"]),
    ("L1_long", 1, 1.0, ["C4", "100", "L1", "100", "2500", "150", "300", "100", "0", "!120", "AA"]),
    ("L2_long", 2, 1.0, ["C4", "100", "L2", "100", "2500", "150", "300", "100", "0", "!120", "AA"]),
    ("L0_R1", 3, 1.0, ["C4", "100", "L0R1", "100", "2500", "150", "300", "100", "0", "!120", "AA"]),
    ("L2_R1_vel", 3, 1.0, ["C4", "130", "L2R1", "100", "2000", "150", "300", "100", "0", "!120", "AA"]),
    ("short", 0, 1.0, ["B3", "100", "", "0", "400", "0", "600", "100", "0", "!120", "AA"]),
]


def metrics(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b))) if a.size else 0.0


def run_case(tmp, name, src_idx, seconds, cli, seed_base=20000, legacy=777):
    feat, pack, y, tr = sources.source_features(src_idx, seconds)
    wav = os.path.join(tmp, f"src{src_idx}_{int(seconds * 1000)}.wav")
    goofy = wav[:-4] + "_features.goofy"
    if not os.path.exists(goofy):
        ref_harness.write_goofy(goofy, pack, tr["f0"], tr["mask"], {i + 1: np.full(feat.env.shape[1], tr["F"][i]) for i in range(4)}, feat.sr, len(y))
    out_wav = os.path.join(tmp, name + "_out.wav")
    ref_out, sr, cap = ref_harness.render_note(goofy, [wav, out_wav] + cli, seed_base, legacy)
    spec = resampler.NoteSpec.from_cli(*cli)
    orc = resampler.resample(feat, spec, lambda n, T: resampler.noise_for_note(spec, n, T, seed_base, legacy))
    return ref_out, orc


def main():
    quick = "--quick" in sys.argv
    worst = 0.0
    with tempfile.TemporaryDirectory() as tmp:
        for name, si, secs, cli in CASES[: (4 if quick else None)]:
            ref_out, orc = run_case(tmp, name, si, secs, cli)
            ok = ref_out.shape == orc.shape
            d = metrics(ref_out, orc) if ok else float("inf")
            worst = max(worst, d)
            print(f"{name:16s} n={len(ref_out):7d} peak={np.max(np.abs(ref_out)):.4f} shapes_ok={ok} max_abs_diff={d:.3e}")
    print("worst", worst)
    return 0 if worst < 1e-6 else 1


if __name__ == "__main__":
    sys.exit(main())
