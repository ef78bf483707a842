"""GPU (-m gpu), two or more devices in ONE process: per-device library state (the dynamic shared-memory
limits, the frame-slot count and the thread-local caches were per process), one host thread per GPU through
goofer_render_batch_host, and the multi-GPU product entry (goofer_b200.multi / cli.render_notes(devices=...)).
Skipped on a single-GPU box; the driver's 8-GPU tier runs them."""
import threading

import numpy as np
import pytest

import bench_data
from goofer_b200 import capi, host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_multi(lib):
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device; goofer_b200 has no CPU fallback"
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    return torch


def _batch(idx, workload="c2", noise_cls=host.SeededNoise):
    b = host.Batch()
    for s in range(16):
        f = bench_data.make_source(s)
        b.add_source(host.SourceFeatures.from_knot_pack(f, f["mask"], f["formants"], f["sr"], f["ylen"]))
    for i in idx:
        src, cli = bench_data.note_cli(i, workload, n_sources=16)
        b.add_note(host.NoteArgs.from_cli(src, cli))
    return b, noise_cls(base_seed=lambda j: 20000 + 16 * idx[j], legacy_seed=lambda j: 777 + idx[j])


def test_second_device_in_the_same_process(torch_multi):
    """cuda:0 first, then cuda:1 from the same thread: same bits (the frame / envelope kernels need their shared-memory
    limit raised on EVERY device)."""
    idx = list(range(24))
    b, noise = _batch(idx, "c3")
    ab = b.assemble(noise)
    outs = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        db = ab.to_device(dev)
        db.render()
        capi.check(db.status())
        outs.append(db.out[:ab.out_total].cpu().numpy())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    # the host entry point from one thread that moves between devices: its cache follows
    ab.pin()
    for d in (0, 1, 0):
        torch_multi.cuda.set_device(d)
        got = np.concatenate(ab.render_host())
        assert np.array_equal(got, outs[0])
    torch_multi.cuda.set_device(0)


def test_one_host_thread_per_gpu(torch_multi):
    """goofer_render_batch_host concurrently from one thread per GPU (the ABI promises re-entrancy per (device, stream)
    and thread-local state): every thread gets the single-GPU result."""
    torch = torch_multi
    n_dev = min(torch.cuda.device_count(), 8)
    idx = list(range(64))
    b, noise = _batch(idx)
    ref_ab = b.assemble(noise)
    db = ref_ab.to_device("cuda:0")
    db.render()
    ref = db.out[:ref_ab.out_total].cpu().numpy()
    results, errors = [None] * n_dev, []

    def work(d):
        try:
            torch.cuda.set_device(d)
            ab = b.assemble(noise)
            ab.pin()
            for _ in range(3):
                out = np.concatenate(ab.render_host())
            results[d] = out
            capi.load().goofer_host_release()
        except Exception as e:  # noqa: BLE001
            errors.append((d, repr(e)))

    ts = [threading.Thread(target=work, args=(d,)) for d in range(n_dev)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    for d in range(n_dev):
        assert np.array_equal(results[d], ref), d


def test_multi_gpu_product_entry(torch_multi):
    """goofer_b200.multi.MultiGpuRenderer: the batch sharded by cost over every GPU of the box equals the single-GPU
    render note for note (float and device-encoded PCM16), with seeded and with device-drawn noise."""
    from goofer_b200 import multi
    torch = torch_multi
    devs = [f"cuda:{i}" for i in range(min(torch.cuda.device_count(), 8))]
    idx = list(range(200))
    b, noise = _batch(idx, "c5", host.DeviceNoise)
    ab = b.assemble(noise)
    db = ab.to_device("cuda:0").enable_pcm16()
    db.render()
    capi.check(db.status())
    ref, ref_pcm = db.outputs(), db.outputs_pcm16()
    with multi.MultiGpuRenderer(devs) as mr:
        parts = mr.partition(b)
        assert sorted(i for p in parts for i in p) == list(range(len(idx))) and all(len(p) > 0 for p in parts)
        got = mr.render(b, noise)
        got_pcm = mr.render(b, noise, pcm16=True)
    for r, g, rp, gp in zip(ref, got, ref_pcm, got_pcm):
        assert np.array_equal(r, g) and np.array_equal(rp, gp)
