import os, sys, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from goofer_b200 import capi
args = argparse.Namespace(workload="c2", notes=1024)
ab, _ = bench.build_batch(args, 0)
ab.pin()
os.environ["GOOFER_HOST_TRACE"] = "1"
for chunk in (1024, 512):
    os.environ["GOOFER_HOST_CHUNK"] = str(chunk)
    for _ in range(3):
        t = time.perf_counter(); ab.render_host(); print("call ms", 1e3 * (time.perf_counter() - t), file=sys.stderr)
