"""OpenUtau resampler manifest of goofer_b200: the expression list that exposes the flag surface in the editor.

`python -m goofer_b200.manifest > goofer_b200.yaml` writes the file OpenUtau reads next to the launcher
(integration/goofer_b200.sh).  It declares the same expressions as the reference's SillySampler.yaml (keys, abbreviations,
ranges, defaults, flags -- a project saved with one resampler opens with the other), generated from the table below
rather than shipped as a copy.  `g`, `B`, `P` have no expression in the reference's manifest either (OpenUtau's built-in
GEN / BRE expressions and the raw flag field carry them).
"""
from __future__ import annotations

TAG = " (SillySampler)"
# (key, name, abbr, min, max, default, flag)
NUMERICAL = [
    ("cent", "Pitch Offset", "foff", -100, 100, 0, "t"),
    ("fmwd", "Formant Width" + TAG, "S_FW", -100, 100, 0, "fw"),
    ("fmst", "Formant Strength Global" + TAG, "S_FT", -100, 100, 0, "fst"),
    *[(f"SF{k}", f"Scale Formant (F{k})" + TAG, f"S_F{k}", -100, 100, 0, "f" + "abcd"[k - 1]) for k in (1, 2, 3, 4)],
    *[(f"STF{k}", f"Strength Formant (F{k})" + TAG, f"STF{k}", -100, 100, 0, "fst" + "abcd"[k - 1]) for k in (1, 2, 3, 4)],
    ("Hvoi", "Voiced Harmonics" + TAG, "S_V", 0, 100, 100, "V"),
    ("cons", "Unvoiced Consonant Gain" + TAG, "S_C", -100, 100, 0, "U"),
    ("grit", "Grittiness" + TAG, "S_G", 0, 100, 0, "sh"),
    ("dist", "Distortion" + TAG, "S_D", 0, 100, 0, "sr"),
    ("tens", "Tension" + TAG, "S_T", -100, 100, 0, "st"),
    ("grwl", "Growl" + TAG, "S_GW", 0, 100, 0, "sg"),
    ("vfry", "Vocal Fry" + TAG, "S_VF", -100, 100, 0, "vf"),
    ("vfhz", "Vocal Fry Base Hz" + TAG, "S_VZ", 0, 100, 50, "vh"),
    ("vfsl", "Vocal Fry Slide Amount" + TAG, "S_VL", 0, 100, 15, "vl"),
    ("thdr", "Dryness" + TAG, "S_DR", 0, 100, 0, "sd"),
    ("rasp", "Rasp" + TAG, "S_SJ", 0, 100, 0, "sj"),
    ("wgwl", "Whisper Growl" + TAG, "S_WG", 0, 100, 0, "sa"),
    ("subh", "Subharmonics" + TAG, "S_SH", 0, 100, 0, "su"),
    ("brig", "Brightness", "BRI", -100, 100, 0, "br"),
    ("evsh", "Envelope Shaping" + TAG, "EVSH", -100, 100, 0, "es"),
    ("pdyn", "Dynamic from Pitch" + TAG, "PDYN", -100, 100, 0, "pd"),
]
# (key, name, abbr, options)
OPTIONS = [
    ("sust", "Sustain Behavior" + TAG, "S_SS", ["L0", "L1", "L2"]),
    ("fvoi", "Force Voicing" + TAG, "FVOI", ["FV0", "FV1"]),
    ("rev", "Reverse", "REV", ["R0", "R1"]),
    ("edit", "SillyEditor", "SEDI", ["SE0", "SE1"]),       # SE1 is rejected by goofer_b200 (GOOFER_NOTE_EDITOR): the Tk editor is out of scope
]


def flags() -> set:
    """Every flag letter group the manifest can emit."""
    out = {f for *_, f in NUMERICAL}
    for *_, opts in OPTIONS:
        out |= {o.rstrip("0123456789") for o in opts}
    return out


def render() -> str:
    lines = ["expressions:"]
    for key, name, abbr, lo, hi, dflt, flag in NUMERICAL:
        lines += [f"  {key}:", f"    name: {name}", f"    abbr: {abbr}", "    type: Numerical", f"    min: {lo}", f"    max: {hi}",
                  f"    default_value: {dflt}", "    is_flag: true", f"    flag: {flag}"]
    for key, name, abbr, opts in OPTIONS:
        lines += [f"  {key}:", f"    name: {name}", f"    abbr: {abbr}", "    type: Options", "    min: 0", "    max: 1",
                  "    default_value: 0", "    is_flag: true", "    options:"] + [f"    - {o}" for o in opts]
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    import sys
    sys.stdout.write(render())
