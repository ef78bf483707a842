"""One process, several GPUs: a render batch split by note over the devices of a box.

The reference's fan-out is one OS process per note (SillySampler.sh:9), a thread per POST in server mode
(SillySampler.py:1196-1224) and a pool over files in folder mode (:235-238).  Here a batch is partitioned by cost
(`shard.balanced_partition`: output samples x synth passes), every GPU gets its own host thread -- the C ABI keeps its
caches per (thread, device) and is re-entrant per (device, stream) -- renders its shard through the same call a
single-GPU caller makes, and downloads into its own page-locked buffer.  No collective, no inter-GPU traffic: notes are
independent.  A note renders to the same bits whichever GPU and whichever shard it lands in (tests/test_gpu_multi.py).
"""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import capi, host, shard


class _IndexedNoise:
    """Noise provider of a shard: local note j draws what global note idx[j] draws in the unsharded batch."""

    def __init__(self, inner: Callable[[int, dict], dict], idx: Sequence[int]):
        self.inner, self.idx = inner, list(idx)

    def __call__(self, j: int, info: dict) -> dict:
        return self.inner(self.idx[j], info)


class MultiGpuRenderer:
    """render(batch, noise) over `devices` (e.g. ["cuda:0", .. "cuda:7"]); one persistent worker thread per device.

    close() joins the workers after each released what the C library cached for it (goofer_host_release)."""

    def __init__(self, devices: Sequence, source_cache: Optional[host.DeviceSourceCache] = None):
        import torch
        self.torch = torch
        self.devices = [torch.device(d) for d in devices]
        if not self.devices:
            raise ValueError("no devices")
        self.source_cache = source_cache
        self.pools = [ThreadPoolExecutor(max_workers=1, thread_name_prefix=f"goofer-{d}") for d in self.devices]
        self._staging: Dict[int, "object"] = {}                  # per worker: page-locked download buffer (torch tensor)
        self._lock = threading.Lock()
        for k, p in enumerate(self.pools):
            p.submit(self._init_worker, k).result()

    def _init_worker(self, k: int) -> None:
        self.torch.cuda.set_device(self.devices[k])
        capi.load()

    def partition(self, batch: host.Batch) -> List[List[int]]:
        infos = batch.plan()
        return shard.balanced_partition([shard.note_cost(i) for i in infos], len(self.devices))

    def _render_shard(self, k: int, batch: host.Batch, idx: List[int], noise, pcm16: bool):
        torch = self.torch
        if not idx:
            return []
        dev = self.devices[k]
        torch.cuda.set_device(dev)
        sub = host.Batch()
        remap: Dict[int, int] = {}
        for i in idx:                                            # only the sources this shard uses
            nt = batch.notes[i]
            if nt.source not in remap:
                remap[nt.source] = sub.add_source(batch.sources[nt.source])
            sub.add_note(host.NoteArgs(**{**nt.__dict__, "source": remap[nt.source]}))
        ab = sub.assemble(_IndexedNoise(noise, idx))
        db = ab.to_device(dev, source_cache=self.source_cache)
        if pcm16:
            db.enable_pcm16()
        db.render()
        src = db.pcm if pcm16 else db.out
        n = ab.out_total
        stage = self._staging.get(k)
        if stage is None or stage.numel() * stage.element_size() < n * src.element_size():
            stage = self._staging[k] = torch.empty(max(n * src.element_size(), 1 << 20), dtype=torch.uint8).pin_memory()
        view = stage[: n * src.element_size()].view(src.dtype)
        view.copy_(src[:n], non_blocking=True)                   # this GPU's own D2H into its own pinned buffer
        capi.check(db.status())                                  # waits for the stream; names a note whose pulse list overflowed
        torch.cuda.current_stream(dev).synchronize()
        return [a.copy() for a in ab.split(view.numpy())]

    def render(self, batch: host.Batch, noise=None, pcm16: bool = False) -> List[np.ndarray]:
        noise = noise or host.FreshDeviceNoise()
        parts = self.partition(batch)
        futs = [p.submit(self._render_shard, k, batch, parts[k], noise, pcm16) for k, p in enumerate(self.pools)]
        outs: List[Optional[np.ndarray]] = [None] * len(batch.notes)
        for k, f in enumerate(futs):
            for i, o in zip(parts[k], f.result()):
                outs[i] = o
        return outs                                              # type: ignore[return-value]

    def close(self) -> None:
        for p in self.pools:
            p.submit(capi.load().goofer_host_release).result()
            p.shutdown(wait=True)
        self.pools = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
