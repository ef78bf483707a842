"""Marginal cost of each formant flag in the envelope kernel: renders the c2 batch with one flag family at a time
(python tools/probe_env_flags.py on a B200)."""
import os, sys, re, argparse
sys.path.insert(0, os.getcwd())
import torch, bench, bench_data
from goofer_b200 import capi
orig = bench_data.formant_flags
def keep_only(names):
    def f(i):
        s = orig(i)
        toks = re.findall(r"([A-Za-z]+)(-?\d+)", s)
        return "".join(k + v for k, v in toks if k in names)
    return f
sets = {"none": [], "g": ["g"], "fa-fd": ["fa", "fb", "fc", "fd"], "fw": ["fw"], "fst": ["fst"], "br": ["br"], "es": ["es"],
        "all": ["g", "fa", "fb", "fc", "fd", "fw", "fst", "br", "es"]}
print(orig(1), orig(2))
for name, ks in sets.items():
    bench_data.formant_flags = keep_only(ks)
    ab, _ = bench.build_batch(argparse.Namespace(workload="c2", notes=1024), 0)
    db = ab.to_device("cuda:0")
    for _ in range(3): db.render()
    torch.cuda.synchronize()
    capi.profile(True)
    for _ in range(5): db.render()
    torch.cuda.synchronize()
    p = capi.profile_summary(); capi.profile(False)
    print(f"{name:6s} env={p['env'][1]/5:.3f} tracks={p['tracks'][1]/5:.3f} total={sum(v[1] for v in p.values())/5:.3f}", flush=True)
