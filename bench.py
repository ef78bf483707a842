#!/usr/bin/env python
"""bench.py -- notes/s of the GOOFER render path on B200 (BASELINE.json metric) with the roofline of the
dominant kernel and the CPU reference path timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--notes 1024]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...        # the reference algorithm (oracle port) on the host cores

A "step" is one goofer_render_batch call over one batch of synthetic notes (default: configs[1] of
BASELINE.json = 1,024 one-second notes with formant flags).  Notes are independent, so with N ranks every
rank renders its own 1,024-note shard (weak scaling, no collective on the data path); `value` is the total
notes of all ranks divided by the slowest rank's device time.
  value     inputs resident in HBM, CUDA events on the launching stream
  e2e       goofer_render_batch_host: pinned HOST buffers in, H2D + kernels + D2H inside the timed region
  roofline  dominant kernel (named in the line), algorithmic bytes per launch over its
            CUDA-event duration inside the timed steps, against MEASURED_PEAKS.json's copy bandwidth
  cpu_baseline  oracle port (numpy + C restatement of the reference), 1 core, bounded sample, rank 0 only
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC, UNIT = "notes_per_sec", "notes/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--notes", type=int, default=1024, help="notes per rank per step")
    ap.add_argument("--cpu-sample", type=int, default=96, help="notes of the workload timed on the CPU (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_config(args, n_gpus):
    import bench_data
    return {"workload": f"{args.workload}: {bench_data.WORKLOADS[args.workload]}", "notes_per_gpu": args.notes,
            "global_notes": args.notes * n_gpus, "note_seconds": 16.0 if args.workload == "c4" else 1.0, "sample_rate": 44100, "n_sources": 64,
            "parallelism": f"notes sharded over {n_gpus} rank(s), no collective on the data path",
            "l2_policy": "inputs larger than L2 (noise phases alone are 4*513*173 B per note)"}


# ------------------------------------------------------------------------------------------------------
# CPU legs (the ONLY place bench.py touches oracle/): cpu_baseline of the native arm, and --impl reference
# ------------------------------------------------------------------------------------------------------
_ORC = {}
_ORC_SECONDS = [1.0]


def _oracle_feat(src_idx):
    import bench_data
    from oracle import dsp
    from oracle.resampler import Features
    f = _ORC.get(src_idx)
    if f is None:
        s = bench_data.make_source(src_idx, _ORC_SECONDS[0])
        env = dsp.decode_knots({"knot_vals_log": s["knot_vals_log"], "hz_knots": s["hz_knots"], "n_fft": 1024,
                                "sr": s["sr"], "n_bins": 513})
        f = Features(env=env, mask=s["mask"], formants=s["formants"], sr=s["sr"], ylen=s["ylen"])
        _ORC[src_idx] = f
    return f


def _oracle_note(job):
    """Render global note index `i` of `workload` with the oracle; returns the number of output samples."""
    workload, i = job
    import bench_data
    from oracle import resampler
    _ORC_SECONDS[0] = bench_data.SOURCE_SECONDS.get(workload, 1.0)
    src, cli = bench_data.note_cli(i, workload, n_sources=8 if workload == "c4" else 64)
    spec = resampler.NoteSpec.from_cli(*cli)
    out = resampler.resample(_oracle_feat(src), spec,
                             lambda n, T: resampler.noise_for_note(spec, n, T, 20000 + 16 * i, 777 + i))
    return len(out)


def cpu_baseline_1core(workload, n_sample):
    for i in range(2):
        _oracle_note((workload, i))                      # warm caches / tables
    t0 = time.perf_counter()
    for i in range(n_sample):
        _oracle_note((workload, i))
    dt = time.perf_counter() - t0
    return {"value": n_sample / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {n_sample} notes of the workload, oracle port (numpy + C), in-process, 1 thread",
            "seconds": round(dt, 3)}


def run_reference(args):
    """--impl reference: the reference algorithm (oracle port: the reference is numpy/numba Python and cannot
    travel to the GPU box) on every host core, one process per core, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_step = max(cores * 2, 16)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        jobs = lambda k: [(args.workload, (k * per_step + j) % max(args.notes, 1)) for j in range(per_step)]  # noqa: E731
        for k in range(args.warmup):
            pool.map(_oracle_note, jobs(k), chunksize=1)
        t0 = time.perf_counter()
        samples = 0
        for k in range(args.steps):
            samples += sum(pool.map(_oracle_note, jobs(args.warmup + k), chunksize=1))
        dt = time.perf_counter() - t0
    n = per_step * args.steps
    val = n / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.gpus),
        "audio_sec_per_sec": samples / 44100.0 / dt,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} notes of the workload per step, oracle port, {cores} processes"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz, self.ok, self.live = None, False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                if self.live:
                    self.samples.append(mhz)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
def build_batch(args, rank, device_noise=False):
    import bench_data
    from goofer_b200 import host
    b = host.Batch()
    n_src = 8 if args.workload == "c4" else 64
    feats = [bench_data.make_source(s, bench_data.SOURCE_SECONDS.get(args.workload, 1.0)) for s in range(n_src)]
    for f in feats:
        b.add_source(host.SourceFeatures.from_knot_pack(f, f["mask"], f["formants"], f["sr"], f["ylen"]))
    first = rank * args.notes
    for j in range(args.notes):
        src, cli = bench_data.note_cli(first + j, args.workload, n_sources=n_src)
        b.add_note(host.NoteArgs.from_cli(src, cli))
    noise_cls = host.DeviceNoise if device_noise else host.SeededNoise
    noise = noise_cls(base_seed=lambda j: 20000 + 16 * (first + j), legacy_seed=lambda j: 777 + first + j)
    ab = b.assemble(noise)
    algo = sum(bench_data.algorithmic_bytes(inf, feats[0]["knot_vals_log"].shape[1] - 1, feats[0]["ylen"]) for inf in ab.infos)
    return ab, algo


def run_native(args):
    import torch
    import torch.distributed as dist
    from goofer_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: goofer_b200 has no CPU fallback")
    from goofer_b200 import shard
    numa = shard.bind_rank_to_gpu_numa(local)                 # before any pinned allocation: host buffers on the GPU's node
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner to stdout at the VERSION and WARN levels
        if "GOOFER_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["GOOFER_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)
        dist.init_process_group("nccl", device_id=dev)
    capi.load()

    ab, algo_bytes = build_batch(args, rank)
    n_notes = len(ab.infos)
    samples = sum(inf["n_total"] for inf in ab.infos)
    db = ab.to_device(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        db.render()
    barrier()
    capi.profile(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.live = True
    ev0.record()
    for _ in range(args.steps):
        db.render()
    ev1.record()
    barrier()
    clocks.live = False
    dev_ms = ev0.elapsed_time(ev1)
    launches_step = capi.last_stats()["kernel_launches"]
    prof = capi.profile_summary()
    capi.profile(False)
    # The excitation chain runs on a side stream beside the envelope kernel (GOOFER_OVERLAP, default on): the spans of
    # those kernels overlap in the timed region, while the frame kernel and everything after it run alone.  A second,
    # untimed pass with both chains on one stream gives per-kernel durations that add up (the table of the line).
    prof_serial = prof
    if os.environ.get("GOOFER_OVERLAP", "1") != "0":
        capi.profile(True, serial=True)
        for _ in range(args.steps):
            db.render()
        torch.cuda.synchronize()
        prof_serial = capi.profile_summary()
        capi.profile(False)

    # ---- end to end through the host-buffer C-ABI call (pinned host memory) ----
    e2e = None
    if not args.no_e2e:
        ab.pin()
        for _ in range(2):
            ab.render_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ab.render_host()
        torch.cuda.synchronize()
        e2e_ms = 1e3 * (time.perf_counter() - t0)
        st = capi.last_stats()
        e2e = (e2e_ms, st["h2d_bytes"], st["d2h_bytes"])
        # the same call returning the notes as 16-bit PCM encoded on the device (what the reference's CLI writes to
        # the .wav): half the download; reported beside the f32 number, which stays the headline
        for _ in range(2):
            ab.render_host(pcm16=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ab.render_host(pcm16=True)
        torch.cuda.synchronize()
        e2e_pcm = (1e3 * (time.perf_counter() - t0), capi.last_stats()["d2h_bytes"])
        # the same batch with the noise phases drawn on the device from the host's PCG64 states (GooferNote.phi_rng:
        # bit-identical phases, tests/test_gpu_parity.py) -- what the CLI / server do; only seeds cross PCIe for them
        ab_dn = build_batch(args, rank, device_noise=True)[0]
        ab_dn.pin()
        for _ in range(2):
            ab_dn.render_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ab_dn.render_host()
        torch.cuda.synchronize()
        e2e_dn = (1e3 * (time.perf_counter() - t0), capi.last_stats()["h2d_bytes"], capi.last_stats()["d2h_bytes"])
    clocks.stop_flag = True

    t = torch.tensor([dev_ms, e2e[0] if e2e else 0.0], dtype=torch.float64, device=dev)
    e2e_by_rank = [float(t[1])]
    if world > 1:
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        e2e_by_rank = [float(x[1]) for x in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        total_notes = n_notes * world * args.steps
        value = total_notes / (dev_ms * 1e-3)
        top = max(prof_serial.items(), key=lambda kv: kv[1][1]) if prof_serial else (None, (0, 0.0))
        forked = ("mask", "fir", "f0", "walk", "onset", "pulse", "tracks", "env", "phi")     # overlapping spans in the timed region
        if top[0] in prof and not (top[0] in forked or top[0].startswith("sg")):
            top = (top[0], prof[top[0]])                      # measured live in the timed region (runs alone there)
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s"
        roof = None
        if top[0]:
            per_launch_ms = top[1][1] / max(1, top[1][0])
            # a step may launch the kernel several times (waves of 2,048 notes, host parts): bytes per LAUNCH
            launches_per_step = max(1.0, top[1][0] / max(1, args.steps))
            algo_per_launch = algo_bytes / launches_per_step
            ach = algo_per_launch / (per_launch_ms * 1e-3) / 1e9
            traffic = None                                    # ncu DRAM bytes per launch of this kernel, when a capture of this workload is committed
            try:
                with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
                    traffic = json.load(fh).get(f"{args.workload}:{args.notes}", {}).get(top[0])
            except Exception:
                pass
            roof = {"bound": "hbm", "kernel": top[0], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "peak_source": peak_src, "kernel_ms_per_launch": per_launch_ms,
                    "algorithmic_bytes_per_launch": int(algo_per_launch), "launches_per_step": launches_per_step,
                    "kernel_share_of_step": top[1][1] / dev_ms,
                    "step_algorithmic_gbs": algo_bytes * args.steps / (dev_ms * 1e-3) / 1e9,
                    "kernels_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in prof_serial.items()},
                    "kernels_note": "table: untimed pass with both preparation chains on one stream (durations add up); "
                                    "in the timed region the excitation chain overlaps the envelope kernel on a side stream, "
                                    "the roofline kernel runs alone and is timed there"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "audio_sec_per_sec": samples * world * args.steps / 44100.0 / (dev_ms * 1e-3),
            "gpu_launches": int(launches_step) * args.steps * world, "clocks": clocks.result(), "roofline": roof,
        }
        if e2e:
            line["e2e"] = {"value": n_notes * world * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                           "h2d_bytes_per_step": int(e2e[1]), "d2h_bytes_per_step": int(e2e[2]),
                           "ms_per_step": e2e_ms / args.steps,
                           "ms_per_step_by_rank": [round(x / args.steps, 3) for x in e2e_by_rank],
                           "host_numa_binding_rank0": numa,
                           "api": "goofer_render_batch_host (C ABI, pinned host buffers, per-rank wall clock, max over ranks)",
                           "pcm16_output": {"value": n_notes * args.steps / (e2e_pcm[0] * 1e-3), "unit": UNIT + " (rank 0)",
                                            "ms_per_step": e2e_pcm[0] / args.steps, "d2h_bytes_per_step": int(e2e_pcm[1])},
                           "device_drawn_phases": {"value": n_notes * args.steps / (e2e_dn[0] * 1e-3), "unit": UNIT + " (rank 0)",
                                                   "ms_per_step": e2e_dn[0] / args.steps, "h2d_bytes_per_step": int(e2e_dn[1]),
                                                   "d2h_bytes_per_step": int(e2e_dn[2]),
                                                   "note": "same notes, same noise: PCG64 states instead of phase buffers"}}
        if args.cpu_sample > 0 and world == 1:               # the CPU leg is timed at N = 1 only (rank 0)
            line["cpu_baseline"] = cpu_baseline_1core(args.workload, args.cpu_sample)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
