"""GPU bring-up probe: e2e time of goofer_render_batch_host vs chunk size, and raw pinned copy bandwidth."""
import os, sys, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from goofer_b200 import capi

args = argparse.Namespace(workload="c2", notes=1024)
ab, _ = bench.build_batch(args, 0)
ab.pin()
capi.load()
x = torch.empty(380_000_000 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(2):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize()
print("H2D 380MB ms", 1e3 * (time.perf_counter() - t))
t = time.perf_counter(); x[:45_000_000].copy_(d[:45_000_000], non_blocking=True); torch.cuda.synchronize()
print("D2H 180MB ms", 1e3 * (time.perf_counter() - t))
for chunk in (1024, 512, 342, 256, 128):
    os.environ["GOOFER_HOST_CHUNK"] = str(chunk)
    for _ in range(2):
        ab.render_host()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(4):
        ab.render_host()
    torch.cuda.synchronize()
    print("chunk", chunk, "ms/step", 1e3 * (time.perf_counter() - t) / 4)
db = ab.to_device("cuda:0")
for _ in range(3):
    db.render()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); db.render(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("render(): host return ms", 1e3 * (t1 - t0), "total ms", 1e3 * (t2 - t0))
import ctypes as C
lib = capi.load()
t0 = time.perf_counter(); ws = lib.goofer_workspace_bytes(C.byref(db.desc), 0); t1 = time.perf_counter()
print("workspace_bytes ms", 1e3 * (t1 - t0))
info = (capi.GooferNotePlanInfo * 1024)()
t0 = time.perf_counter(); lib.goofer_plan_batch(C.byref(db.desc), info); t1 = time.perf_counter()
print("plan_batch ms", 1e3 * (t1 - t0))
