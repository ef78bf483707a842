"""GOOFER_HOST_TRACE timeline of the headline end-to-end variant (device-drawn phases, PCM16) on the c2 batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from goofer_b200 import capi

class A: workload = "c2"; notes = 1024
ab = bench.build_batch(A, 0, device_noise=True)[0]
ab.pin()
for pcm in (True, False):
    for _ in range(3):
        ab.render_host(pcm16=pcm)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); ab.render_host(pcm16=pcm); ts.append(1e3 * (time.perf_counter() - t0))
    print("pcm16" if pcm else "f32", "device phases: ms per call", [round(t, 3) for t in ts], capi.last_stats(), file=sys.stderr)
