"""TEST INFRASTRUCTURE: fans oracle renders over the host cores so that whole BASELINE.json configs (1,024 notes of c2,
256 of c3, the 16 s grid of c4) can be compared note by note with the CUDA output in seconds.

The parent (a pytest process with a live CUDA context) writes the GPU output to a .npy file; spawned workers memory-map
it, render their notes with the oracle port and return (note, max-abs error, log-spectral distance)."""
from __future__ import annotations

import multiprocessing as mp
import os
import tempfile

import numpy as np

_W = {"out": None, "feat": {}}


def _init(npy_path):
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    _W["out"] = np.load(npy_path, mmap_mode="r")


def _feature(src_idx, seconds):
    import bench_data
    from oracle import dsp
    from oracle.resampler import Features
    key = (src_idx, seconds)
    f = _W["feat"].get(key)
    if f is None:
        s = bench_data.make_source(src_idx, seconds)
        env = dsp.decode_knots({"knot_vals_log": s["knot_vals_log"], "hz_knots": s["hz_knots"], "n_fft": 1024, "sr": s["sr"], "n_bins": 513})
        f = _W["feat"][key] = Features(env=env, mask=s["mask"], formants=s["formants"], sr=s["sr"], ylen=s["ylen"])
    return f


def _job(job):
    """job = (tag, source index, source seconds, 11 CLI strings, base seed, legacy seed, output offset, expected length)"""
    from oracle import resampler
    from tests import cases
    tag, src, seconds, cli, base, legacy, off, n = job
    spec = resampler.NoteSpec.from_cli(*cli)
    ref = resampler.resample(_feature(src, seconds), spec, lambda m, T: resampler.noise_for_note(spec, m, T, base, legacy))
    if len(ref) != n:
        return tag, float("inf"), float("inf"), f"length {len(ref)} vs {n}"
    got = np.asarray(_W["out"][off:off + n], dtype=np.float64)
    err = float(np.max(np.abs(got - ref))) if n else 0.0
    return tag, err, cases.lsd_db(ref, got), ""


def compare(jobs, flat_out: np.ndarray, processes: int | None = None):
    """[(tag, max_abs, lsd_db, message)] for every job, in job order."""
    cores = processes or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        path = os.path.join(tmp, "gpu_out.npy")
        np.save(path, np.ascontiguousarray(flat_out))
        ctx = mp.get_context("spawn")                       # the parent holds a CUDA context: no fork
        with ctx.Pool(min(cores, max(1, len(jobs))), initializer=_init, initargs=(path,)) as pool:
            return pool.map(_job, jobs, chunksize=max(1, len(jobs) // (8 * cores)))


def workload_jobs(workload, idx, infos, n_sources=64):
    """Jobs of bench_data notes `idx` (global note numbers); seeds as bench.py / host.SeededNoise use them."""
    import bench_data
    seconds = bench_data.SOURCE_SECONDS.get(workload, 1.0)
    jobs, off = [], 0
    for j, i in enumerate(idx):
        src, cli = bench_data.note_cli(i, workload, n_sources=n_sources)
        n = infos[j]["n_total"]
        jobs.append((i, src, seconds, cli, 20000 + 16 * i, 777 + i, off, n))
        off += n
    return jobs
