"""Note sharding across ranks.

Notes are independent units -- the reference renders one note per process (SillySampler.sh:9) -- so a
render batch is partitioned by note over the GPUs of a box with NO collective on the data path.  The only
communication is the optional gather of per-rank results (object / padded tensor all_gather through
torch.distributed: NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence


def contiguous_range(n_notes: int, rank: int, world: int) -> range:
    """Contiguous, balanced-by-count slice of note indices of `rank` (first `n % world` ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(int(n_notes), world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def note_cost(info: dict) -> int:
    """Relative cost of a note: output samples x synth passes (a full-flag note runs gf.synthesize up to
    four times, SillySampler.py:1006,1041,1067,1156)."""
    return int(info["n_total"]) * int(info["n_passes"])


def balanced_partition(costs: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-processing-time partition of note indices into `world` bins of near-equal cost.
    Deterministic: ties go to the lowest rank; indices inside a bin stay in ascending order."""
    if world <= 0:
        raise ValueError("world must be positive")
    bins: List[List[int]] = [[] for _ in range(world)]
    load = [0] * world
    for i in sorted(range(len(costs)), key=lambda k: (-int(costs[k]), k)):
        r = min(range(world), key=lambda q: (load[q], q))
        bins[r].append(i)
        load[r] += int(costs[i])
    for b in bins:
        b.sort()
    return bins


def gather_outputs(local_indices: Sequence[int], local_outputs: Sequence, group=None, device=None) -> dict:
    """All-gather {note index: output array} over the process group (no-op without torch.distributed).

    One fixed-stride `all_gather_into_tensor` of the concatenated samples (`ncclAllGather` of
    (notes_per_gpu x N_out) over NVLink / NVSwitch when a single buffer is required) plus one of the small (index, length)
    table -- no pickling of sample data.  `device`: where the collective runs ("cuda:k" for NCCL; CPU tensors for gloo).
    Every rank returns every note.  The data path of a render needs none of this: ranks write their own outputs."""
    import numpy as np
    import torch
    import torch.distributed as dist
    mine = {int(i): o for i, o in zip(local_indices, local_outputs)}
    if not (dist.is_available() and dist.is_initialized()):
        return mine
    world = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    dtype = np.asarray(local_outputs[0]).dtype if len(local_outputs) else np.dtype(np.float32)
    tdtype = torch.from_numpy(np.zeros(1, dtype=dtype)).dtype
    n_local = len(local_indices)
    total = int(sum(len(o) for o in local_outputs))
    # 1. sizes: (note count, sample count) of every rank
    sizes = torch.tensor([n_local, total], dtype=torch.int64, device=device)
    all_sizes = torch.empty(2 * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_sizes, sizes, group=group)
    all_sizes = all_sizes.cpu().view(world, 2)
    max_notes, max_total = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())
    # 2. the (index, length) table, fixed stride
    table = torch.full((max(1, max_notes), 2), -1, dtype=torch.int64)
    for k, (i, o) in enumerate(zip(local_indices, local_outputs)):
        table[k, 0], table[k, 1] = int(i), len(o)
    all_tables = torch.empty((world * table.shape[0], 2), dtype=torch.int64, device=device)      # concatenated along dim 0
    dist.all_gather_into_tensor(all_tables, table.to(device), group=group)
    # 3. the samples, fixed stride = the largest shard
    flat = torch.zeros(max(1, max_total), dtype=tdtype)
    if total:
        flat[:total] = torch.from_numpy(np.concatenate([np.asarray(o) for o in local_outputs]))
    all_flat = torch.empty(world * flat.numel(), dtype=tdtype, device=device)
    dist.all_gather_into_tensor(all_flat, flat.to(device), group=group)
    all_tables = all_tables.cpu().numpy().reshape(world, table.shape[0], 2)
    all_flat = all_flat.cpu().numpy().reshape(world, flat.numel())
    out = {}
    for r in range(world):
        off = 0
        for k in range(int(all_sizes[r, 0])):
            i, n = int(all_tables[r, k, 0]), int(all_tables[r, k, 1])
            out[i] = all_flat[r, off:off + n]
            off += n
    return out


def _parse_cpulist(txt: str) -> set:
    cpus = set()
    for part in txt.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def _bind_slice(local_rank: int, world: int, out: dict) -> dict:
    """Fallback when the platform reports no NUMA node for the GPU (a VM with a flat topology): give every rank of the box its
    own contiguous slice of the allowed CPUs, so that the ranks' host threads do not migrate onto each other."""
    import os
    if world <= 1:
        return out
    cpus = sorted(os.sched_getaffinity(0))
    k = len(cpus) // world
    if k < 1:
        return out
    mine = set(cpus[local_rank * k:(local_rank + 1) * k])
    os.sched_setaffinity(0, mine)
    out.update(bound=True, cpus=len(mine), how="contiguous slice of the allowed CPUs (no NUMA information)")
    return out


def bind_rank_to_gpu_numa(local_rank: int, world: int = 1) -> dict:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates page-locked buffers.

    One process per GPU renders its own shard and, through `goofer_render_batch_host`, streams hundreds of MB per call
    between pinned host memory and its GPU.  Linux places pinned pages on the node of the allocating thread, so an
    unbound rank may end up with its buffers on the other socket and every DMA crossing the inter-socket link -- with
    eight ranks that link, not PCIe, bounds the end-to-end rate.  Returns what was done (for the bench line); never
    raises: no NVML / sysfs / permission => {"bound": False, ...}.  `GOOFER_NUMA_BIND=0` disables it."""
    import os
    out = {"bound": False}
    if os.environ.get("GOOFER_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        out["why"] = "disabled"
        return out
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.lower().split(":", 1)
        sysfs = f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/numa_node"
        with open(sysfs) as fh:
            node = int(fh.read().strip())
        out["node"] = node
        if node < 0:
            out["why"] = "no NUMA node reported for the GPU"
            return _bind_slice(local_rank, world, out)
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = _parse_cpulist(fh.read())
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            out["why"] = "none of the node's CPUs is in this process's cpuset"
            return out
        os.sched_setaffinity(0, allowed)
        out.update(bound=True, cpus=len(allowed))
    except Exception as ex:  # noqa: BLE001 -- a launch nicety, never fatal
        out["why"] = f"{type(ex).__name__}: {ex}"
    return out
