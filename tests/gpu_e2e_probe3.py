import os, sys, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
args = argparse.Namespace(workload="c2", notes=1024)
ab, _ = bench.build_batch(args, 0)
ab.pin()
for chunk in (512, 342, 256, 1024):
    os.environ["GOOFER_HOST_CHUNK"] = str(chunk)
    for _ in range(3):
        ab.render_host()
    t = time.perf_counter()
    for _ in range(5):
        ab.render_host()
    print("chunk", chunk, "ms/step", 1e3 * (time.perf_counter() - t) / 5)
