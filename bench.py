#!/usr/bin/env python
"""bench.py -- notes/s of the GOOFER render path on B200 (BASELINE.json metric) with the roofline of the
dominant kernel and the CPU reference path timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--notes 1024] [--verify 8]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...        # the reference algorithm on the host cores (see run_reference)

A "step" is one render call over one batch of synthetic notes (default: configs[1] of BASELINE.json = 1,024
one-second notes with formant flags).  Notes are independent, so with N ranks every rank renders its own
shard (weak scaling, no collective on the data path); `value` is the total notes of all ranks divided by the
slowest rank's device time.
  value     inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e       goofer_render_batch_host with pinned HOST buffers: H2D + kernels + D2H inside the timed region, max
            over ranks.  Headline variant = the production transfer set of the drop-in (what cli / server use):
            the host hands over the PCG64 STATES of the noise generators the reference would draw from
            (GOOFER.py:1151: the reference draws its noise itself; the phases are bit-identical to the uploaded
            ones, tests/test_gpu_parity.py) and reads back 16-bit PCM (what SillySampler.py:1185 writes to the
            .wav).  The other three combinations (host-supplied float phases, float output) are timed the same
            way and reported beside it in e2e.variants.
  roofline  dominant kernel (named in the line), algorithmic bytes per launch over its
            CUDA-event duration inside the timed steps, against MEASURED_PEAKS.json's copy bandwidth
  verify    K notes of the batch that was just timed, compared with the oracle (max-abs, log-spectral distance)
  cpu_baseline  oracle port (numpy + C restatement of the reference) on the host cores, bounded sample, N = 1 only
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
E2E_VARIANTS = ("device_phases_pcm16", "device_phases_f32", "host_phases_pcm16", "host_phases_f32")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--notes", type=int, default=1024, help="notes per rank per step")
    ap.add_argument("--cpu-sample", type=int, default=96, help="notes of the workload timed on the CPU (0 = skip)")
    ap.add_argument("--verify", type=int, default=8, help="notes of the timed batch compared with the oracle (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-variants", default="auto", help="'all', 'prod' (headline only) or 'auto' (all up to 2,048 notes per rank)")
    ap.add_argument("--noise", default="auto", choices=["auto", "host", "device"],
                    help="noise phases of the device-resident leg: uploaded once (host) or drawn by gf_phi_kernel inside "
                         "every step (device); auto = host up to 8,192 notes per rank")
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "reference", "port"])
    return ap.parse_args()


def workload_config(args, n_gpus):
    import bench_data
    return {"workload": f"{args.workload}: {bench_data.WORKLOADS[args.workload]}", "notes_per_gpu": args.notes,
            "global_notes": args.notes * n_gpus, "note_seconds": 16.0 if args.workload == "c4" else 1.0, "sample_rate": 44100,
            "n_sources": 8 if args.workload == "c4" else 64,
            "parallelism": f"notes sharded over {n_gpus} rank(s), no collective on the data path",
            "l2_policy": "inputs larger than L2 (noise phases alone are 4*513*173 B per note)"}


# ------------------------------------------------------------------------------------------------------
# CPU legs (the ONLY place bench.py touches oracle/): verify, cpu_baseline of the native arm, --impl reference
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


def oracle_render(workload, i):
    """Global note index `i` of `workload` rendered by the oracle port (float64 samples)."""
    import bench_data
    from oracle import resampler
    _ORC_SECONDS[0] = bench_data.SOURCE_SECONDS.get(workload, 1.0)
    src, cli = bench_data.note_cli(i, workload, n_sources=8 if workload == "c4" else 64)
    spec = resampler.NoteSpec.from_cli(*cli)
    return resampler.resample(_oracle_feat(src), spec,
                              lambda n, T: resampler.noise_for_note(spec, n, T, 20000 + 16 * i, 777 + i))


def _oracle_note(job):
    """Pool worker: renders one note with the oracle port; returns the number of output samples."""
    return len(oracle_render(*job))


# ---- the UNMODIFIED reference, when its sources are on this machine (build container; a pod with baseline/_ref) ----
_REF = {"dir": None, "tmp": None}


def reference_dir():
    for d in (os.environ.get("GOOFER_REFERENCE_DIR"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if d and os.path.isfile(os.path.join(d, "GOOFER.py")) and os.path.isfile(os.path.join(d, "SillySampler.py")):
            return d
    return None


def reference_usable():
    d = reference_dir()
    if not d:
        return None
    try:
        import numba  # noqa: F401
        import scipy  # noqa: F401
    except Exception:
        return None
    return d


def _reference_note(job):
    """Pool worker: GooferResampler(*13 CLI strings) of the unmodified reference through oracle/ref_harness.py (stub
    soundfile / Praat / Tk modules, seeded noise) -- CLI-equivalent in-process: load_features + resample + every
    gf.synthesize call + the sf.write hand-over; no interpreter start, JIT warm."""
    workload, i = job
    import tempfile
    import numpy as np
    import bench_data
    if _REF["dir"] is None:
        os.environ["GOOFER_REFERENCE_DIR"] = reference_dir()
        _REF["dir"] = os.environ["GOOFER_REFERENCE_DIR"]
        _REF["tmp"] = tempfile.mkdtemp(prefix="goofer_ref_bench_")
    from oracle import ref_harness
    n_src = 8 if workload == "c4" else 64
    src, cli = bench_data.note_cli(i, workload, n_sources=n_src)
    wav = os.path.join(_REF["tmp"], f"src{src}.wav")
    goofy = wav[:-4] + "_features.goofy"
    if not os.path.exists(goofy):
        s = bench_data.make_source(src, bench_data.SOURCE_SECONDS.get(workload, 1.0))
        pack = {"mode": "knots", "knot_vals_log": s["knot_vals_log"], "hz_knots": s["hz_knots"], "n_bins": 513, "n_fft": 1024, "sr": s["sr"]}
        ref_harness.write_goofy(goofy, pack, (220.0 * s["mask"]).astype(np.float64), s["mask"].astype(np.float64), s["formants"], s["sr"], s["ylen"])
    out, _, _ = ref_harness.render_note(goofy, [wav, os.path.join(_REF["tmp"], "out.wav")] + list(cli), 20000 + 16 * i, 777 + i)
    return len(out)


def cpu_pool_rate(worker, workload, n_sample, cores, first=0):
    """notes/s of `worker` over `n_sample` notes of the workload on `cores` processes (1 = in-process)."""
    if cores <= 1:
        for i in range(2):
            worker((workload, i))                            # warm caches / tables / JIT
        t0 = time.perf_counter()
        for i in range(n_sample):
            worker((workload, first + i))
        return n_sample / (time.perf_counter() - t0), time.perf_counter() - t0
    import multiprocessing as mp
    # spawn, not fork: the native arm calls this with a live CUDA context, NVML and pinned host mappings in the parent
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(worker, [(workload, j) for j in range(2 * cores)], chunksize=1)     # warm every process
        t0 = time.perf_counter()
        pool.map(worker, [(workload, first + j) for j in range(n_sample)], chunksize=1)
        dt = time.perf_counter() - t0
    return n_sample / dt, dt


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_baseline(workload, n_sample):
    """The native arm's CPU leg (rank 0, N = 1): oracle port on ALL host cores and on one core, bounded sample."""
    cores = host_cores()
    n_all = max(n_sample, 4 * cores)
    v_all, dt_all = cpu_pool_rate(_oracle_note, workload, n_all, cores)
    n_one = max(8, min(n_sample, 32))
    v_one, dt_one = cpu_pool_rate(_oracle_note, workload, n_one, 1)
    return {"value": v_all, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n_all} notes of the workload, oracle port (numpy + C), {cores} processes; resample + synthesize "
                      f"only (features decoded once per source)",
            "seconds": round(dt_all, 3),
            "one_core": {"value": v_one, "unit": UNIT, "cores": 1, "sample": f"first {n_one} notes, in-process, 1 thread", "seconds": round(dt_one, 3)}}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on every host core, one process per core,
    bounded sample per step.  kind "reference" = the UNMODIFIED GOOFER.py / SillySampler.py driven through
    oracle/ref_harness.py when their sources are on this machine (baseline/_ref, /root/reference) and numba imports;
    else kind "port" = the oracle restatement (bit-identical to the reference on the pinned cases).  A reference
    "step" is `per_step` notes, not the native arm's batch: notes/s compare, ms_per_step does not."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = host_cores()
    use_ref = args.ref_kind != "port" and reference_usable() is not None
    if args.ref_kind == "reference" and not use_ref:
        print(json.dumps({"impl": "reference", "unavailable": "reference sources (baseline/_ref, /root/reference) or numba absent on this machine"}), flush=True)
        return
    worker = _reference_note if use_ref else _oracle_note
    per_step = max(cores * 2, 16)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        jobs = lambda k: [(args.workload, (k * per_step + j) % max(args.notes, 1)) for j in range(per_step)]  # noqa: E731
        for k in range(max(args.warmup, 2)):
            pool.map(worker, jobs(k), chunksize=1)
        t0 = time.perf_counter()
        samples = 0
        for k in range(args.steps):
            samples += sum(pool.map(worker, jobs(args.warmup + k), chunksize=1))
        dt = time.perf_counter() - t0
    n = per_step * args.steps
    val = n / dt
    kind = "reference" if use_ref else "port"
    scope = ("CLI-equivalent in-process: load_features + resample + synthesize + write hand-over (no interpreter start, numba JIT warm)"
             if use_ref else "resample + synthesize only (features decoded once per source)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.gpus),
        "audio_sec_per_sec": samples / 44100.0 / dt,
        "reference_step": f"{per_step} notes per step (bounded sample of the workload); compare notes/s, not ms_per_step",
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "scope": scope,
                         "sample": f"{per_step} notes of the workload per step, {'unmodified reference (' + reference_dir() + ')' if use_ref else 'oracle port'}, {cores} processes"},
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
            time.sleep(0.005)

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


def lsd_db(ref, got, floor_db=-100.0):
    """Log-spectral distance on the reference STFT grid (1024 / 256, sqrt-Hann), floor -100 dB re the reference peak
     -- the checker's own numpy STFT, not the product's."""
    import numpy as np
    from oracle import dsp
    A = np.abs(dsp.stft(ref.astype(np.float32)))
    B = np.abs(dsp.stft(got.astype(np.float32)))
    fl = max(A.max(), 1e-30) * 10 ** (floor_db / 20)
    return float(np.sqrt(np.mean((20 * np.log10(np.maximum(A, fl)) - 20 * np.log10(np.maximum(B, fl))) ** 2)))


def verify_indices(args):
    K = min(args.verify, args.notes)
    return sorted({int(round(k * (args.notes - 1) / max(1, K - 1))) for k in range(K)}) if K > 0 else []


def verify_batch(args, rank, outs_f32, outs_pcm):
    """Compare K notes, spread evenly over this rank's batch, of the outputs the timed calls produced with the oracle."""
    import numpy as np
    from goofer_b200 import cli as gcli
    idx = verify_indices(args)
    first = rank * args.notes
    worst_abs, worst_lsd, worst_lsb, worst_note = 0.0, 0.0, 0, -1
    for j in idx:
        ref = oracle_render(args.workload, first + j)
        got = np.asarray(outs_f32[j], dtype=np.float64)
        if got.shape != ref.shape:
            return {"checked": len(idx), "error": f"note {first + j}: {got.shape} vs oracle {ref.shape}"}
        d = float(np.max(np.abs(got - ref))) if ref.size else 0.0
        if d > worst_abs:
            worst_abs, worst_note = d, first + j
        worst_lsd = max(worst_lsd, lsd_db(ref, got))
        if outs_pcm is not None:
            want = gcli.pcm16_like_soundfile(ref).astype(np.int64)
            worst_lsb = max(worst_lsb, int(np.max(np.abs(outs_pcm[j].astype(np.int64) - want))) if ref.size else 0)
    out = {"checked": len(idx), "notes": [first + j for j in idx], "worst_max_abs": worst_abs, "worst_lsd_db": worst_lsd,
           "worst_note": worst_note, "tolerance_max_abs": 1e-4, "tolerance_lsd_db": 0.05,
           "what": "device-resident f32 output of the timed batch vs the oracle port, same seeded noise"}
    if outs_pcm is not None:
        out["e2e_pcm16_worst_lsb_diff"] = worst_lsb
        out["e2e_pcm16_what"] = ("headline e2e output (device-drawn phases, PCM16) vs libsndfile-style PCM16 of the oracle; "
                                 "1e-4 of full scale = 3.3 LSB")
    out["ok"] = bool(worst_abs <= 1e-4 and worst_lsd <= 0.05 and worst_lsb <= 4)
    return out


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
    numa = shard.bind_rank_to_gpu_numa(local, world)          # before any pinned allocation: host buffers on the GPU's node
    if world > 2:
        # several ranks share one host memory path: render in sub-batches so that the downloads of all ranks spread over the
        # step instead of bursting at its end (goofer_render_batch_host, GOOFER_HOST_SUBBATCHES).  Measured on an 8 x B200 box
        # (c2, e2e ms per step, max over ranks; S = 1 / 2 / 3 / 4): 4 ranks 6.72 / 6.36 / 6.57 / 6.55, 8 ranks 10.23 / 9.48 /
        # 8.24 / 8.40; one or two ranks render fastest in one piece (5.6 / 5.8 ms)
        os.environ.setdefault("GOOFER_HOST_SUBBATCHES", "2" if world <= 4 else "3")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL logs to stdout by default; stdout must stay the one JSON line, so the log goes to stderr (where the
        # driver reads the rank count from) unless the caller already chose a file
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=dev)
    capi.load()

    dev_noise = args.noise == "device" or (args.noise == "auto" and args.notes > 8192)
    ab, algo_bytes = build_batch(args, rank, device_noise=dev_noise)
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
    clocks.stop_flag = True                                   # NVML polling ends with the region it documents: eight ranks polling the
    clocks.join(timeout=1.0)                                  # driver during the end-to-end legs would perturb what they measure
    launches_step = capi.last_stats()["kernel_launches"]
    prof = capi.profile_summary()
    capi.profile(False)
    capi.check(db.status())                                   # no truncated pulse lists in the timed batch
    vidx = verify_indices(args)
    offs = [0]
    for inf in ab.infos:
        offs.append(offs[-1] + inf["n_total"])
    outs_f32 = {j: db.out[offs[j]:offs[j + 1]].cpu().numpy() for j in vidx} if args.verify > 0 else None   # only the checked notes cross PCIe
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
    del db
    torch.cuda.empty_cache()

    # ---- end to end through the host-buffer C-ABI call (pinned host memory), every variant on every rank ----
    e2e_ms = {}
    e2e_calls = {}
    e2e_bytes = {}
    outs_pcm = None
    if not args.no_e2e:
        which = args.e2e_variants
        if which == "auto":
            which = "all" if args.notes <= 2048 else "prod"
        variants = E2E_VARIANTS if which == "all" else E2E_VARIANTS[:1]
        ab_dn = ab if dev_noise else build_batch(args, rank, device_noise=True)[0]
        ab_host = None
        if any(v.startswith("host") for v in variants):
            ab_host = ab if not dev_noise else build_batch(args, rank, device_noise=False)[0]
            ab_host.pin()
        ab_dn.pin()
        for v in variants:
            a = ab_dn if v.startswith("device") else ab_host
            pcm = v.endswith("pcm16")
            for _ in range(3):
                res = a.render_host(pcm16=pcm)
            barrier()
            calls = []
            t0 = time.perf_counter()
            for _ in range(args.steps):
                tc = time.perf_counter()
                res = a.render_host(pcm16=pcm)
                calls.append(1e3 * (time.perf_counter() - tc))
            torch.cuda.synchronize()
            e2e_ms[v] = 1e3 * (time.perf_counter() - t0)
            e2e_calls[v] = calls
            st = capi.last_stats()
            e2e_bytes[v] = (st["h2d_bytes"], st["d2h_bytes"])
            if v == E2E_VARIANTS[0] and args.verify > 0:
                outs_pcm = {j: res[j].copy() for j in vidx}
    keys = [v for v in E2E_VARIANTS if v in e2e_ms]
    t = torch.tensor([dev_ms] + [e2e_ms[v] for v in keys], dtype=torch.float64, device=dev)
    by_rank = [[float(x) for x in t]]
    if world > 1:
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        by_rank = [[float(x) for x in e] for e in every]
    worst = [max(r[c] for r in by_rank) for c in range(len(by_rank[0]))]
    dev_ms = worst[0]

    verify = None
    if args.verify > 0:
        verify = verify_batch(args, rank, outs_f32, outs_pcm)          # every rank checks its own shard
        if world > 1:
            allv = [None] * world
            dist.all_gather_object(allv, verify)
            if rank == 0:
                verify = {"checked": sum(v.get("checked", 0) for v in allv), "ranks": world,
                          "worst_max_abs": max(v.get("worst_max_abs", float("inf")) for v in allv),
                          "worst_lsd_db": max(v.get("worst_lsd_db", float("inf")) for v in allv),
                          "e2e_pcm16_worst_lsb_diff": max(v.get("e2e_pcm16_worst_lsb_diff", 0) for v in allv),
                          "tolerance_max_abs": 1e-4, "tolerance_lsd_db": 0.05, "ok": all(v.get("ok", False) for v in allv),
                          "what": allv[0].get("what"), "e2e_pcm16_what": allv[0].get("e2e_pcm16_what")}

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
            traffic, traffic_src, step_traffic = None, None, None     # ncu DRAM bytes per launch, from tools/ncu_traffic.py's output
            try:
                with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
                    tj = json.load(fh)
                ent = tj.get(f"{args.workload}:{args.notes}", {})
                traffic = ent.get(top[0])
                traffic_src = tj.get("_source")
                step_traffic = ent.get("_step_total")         # DRAM bytes of one whole captured step
            except Exception:
                pass
            roof = {"bound": "hbm", "kernel": top[0], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "step_dram_traffic": step_traffic,
                    "peak_source": peak_src, "kernel_ms_per_launch": per_launch_ms,
                    "algorithmic_bytes_per_launch": int(algo_per_launch), "launches_per_step": launches_per_step,
                    "kernel_share_of_step": top[1][1] / dev_ms,
                    "step_algorithmic_gbs": algo_bytes * args.steps / (dev_ms * 1e-3) / 1e9,
                    "step_frac": algo_bytes * args.steps / (dev_ms * 1e-3) / 1e9 / peak,
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
            "value_noise": "phases drawn on the device inside every step (gf_phi_kernel)" if dev_noise else "host-supplied phases resident in HBM",
        }
        if keys:
            def leg(c, v):
                ms = worst[c]
                return {"value": n_notes * world * args.steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / args.steps,
                        "h2d_bytes_per_step": int(e2e_bytes[v][0]), "d2h_bytes_per_step": int(e2e_bytes[v][1]),
                        "ms_per_step_by_rank": [round(r[c] / args.steps, 3) for r in by_rank],
                        "rank0_call_ms": {"median": round(statistics.median(e2e_calls[v]), 3), "min": round(min(e2e_calls[v]), 3),
                                          "max": round(max(e2e_calls[v]), 3)}}
            head = leg(1, keys[0])
            head.update({
                "variant": keys[0],
                "api": "goofer_render_batch_host (C ABI, pinned host buffers): noise phases drawn on the device from the host's PCG64 "
                       "states (GooferNote.phi_rng, bit-identical to the uploaded phases), 16-bit PCM read back (SillySampler.py:1185); "
                       "per-rank wall clock over the timed calls, max over ranks",
                "host_numa_binding_rank0": numa,
                "host_subbatches": int(os.environ.get("GOOFER_HOST_SUBBATCHES", "1")),
                "variants": {v: leg(1 + c, v) for c, v in enumerate(keys)},
                "variants_note": "every variant: same notes, same noise, same call, max over ranks; host_phases_f32 is the "
                                 "north star's parity transfer set (float phases up, float samples down)"})
            line["e2e"] = head
        if verify is not None:
            line["verify"] = verify
        if args.cpu_sample > 0 and world == 1:               # the CPU leg is timed at N = 1 only (rank 0)
            line["cpu_baseline"] = cpu_baseline(args.workload, args.cpu_sample)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
