"""GPU bring-up report: per case, max-abs error of the render and of the main-pass taps against the
oracle.  `python tests/gpu_debug.py [case ...]` on the GPU box; prints a table."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tests import cases
from goofer_b200 import host, capi

only = set(sys.argv[1:])
for name, si, secs, cli in cases.CASES:
    if only and name not in only:
        continue
    feat, sf = cases.source_for(si, secs)
    taps = {}
    ref = cases.oracle_render(feat, cli, taps=taps)
    b = host.Batch()
    b.add_source(sf)
    b.add_note(host.NoteArgs.from_cli(0, cli))
    try:
        ab = b.assemble(host.SeededNoise(cases.SEED_BASE, cases.SEED_LEGACY), taps=True)
        t0 = time.time()
        outs, tp = ab.render_host()
        dt = time.time() - t0
    except Exception as e:
        print(f"{name:16s} FAILED: {e}")
        continue
    got = outs[0].astype(np.float64)
    err = np.max(np.abs(got - ref))
    pos = int(np.argmax(np.abs(got - ref)))
    print(f"{name:16s} n={len(ref):7d} max_abs={err:.3e} at {pos} peak={np.max(np.abs(ref)):.3f} lsd={cases.lsd_db(ref, got):.4f} dB  t={dt*1e3:.1f} ms "
          f"launches={capi.last_stats()['kernel_launches']}")
