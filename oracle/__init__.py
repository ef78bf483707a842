"""ORACLE -- CPU restatement of GOOFER's render path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import
this package; the product package ``goofer_b200`` never does (tests/test_boundary.py enforces it).
Parity status: pinned by execution against the unmodified reference + committed golden vectors
(the reference ships no tests of its own -- SURVEY.md section 4).
"""
from . import dsp, synth, resampler, sources  # noqa: F401
