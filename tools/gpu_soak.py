#!/usr/bin/env python
"""Determinism soak on a B200: render the same batch many times and compare every output bit for bit with the first
render (catches races: named barriers, atomics, stream ordering, workspace reuse).

    python tools/gpu_soak.py [workload=c2] [notes=512] [steps=200]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

ap = argparse.ArgumentParser()
ap.add_argument("workload", nargs="?", default="c2")
ap.add_argument("notes", nargs="?", type=int, default=512)
ap.add_argument("steps", nargs="?", type=int, default=200)
a = ap.parse_args()
ab, _ = bench.build_batch(argparse.Namespace(workload=a.workload, notes=a.notes), 0)
db = ab.to_device("cuda:0")
ref = db.render().clone()
torch.cuda.synchronize()
assert torch.isfinite(ref).all()
bad = 0
for k in range(a.steps):
    out = db.render()
    if not torch.equal(out, ref):
        bad += 1
        print(f"step {k}: {int((out != ref).sum())} samples differ, max {float((out - ref).abs().max()):.3e}")
ab.pin()
host_ref = None
for k in range(max(4, a.steps // 20)):
    flat = torch.from_numpy(__import__("numpy").concatenate(ab.render_host()))
    if host_ref is None:
        host_ref = flat.clone()
        if not torch.equal(host_ref, ref.cpu()[:host_ref.numel()]):
            bad += 1
            print("host entry point differs from the device-resident render")
    elif not torch.equal(flat, host_ref):
        bad += 1
        print(f"host step {k} differs")
print(f"soak {a.workload} x{a.notes}: {a.steps} device renders + host renders, {bad} mismatches")
sys.exit(1 if bad else 0)
