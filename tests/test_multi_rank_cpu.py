"""CPU, world_size 2 over gloo: the N > 1 path is 'shard notes by rank, no collective on the data path, gather
results' (SURVEY.md section 8e).  Two processes plan the same batch, take their shards, and the gathered
result must cover every note exactly once with the planner's lengths."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import bench_data
    from goofer_b200 import host, shard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = host.Batch()
        for s in range(4):
            f = bench_data.make_source(s)
            b.add_source(host.SourceFeatures.from_knot_pack(f, f["mask"], f["formants"], f["sr"], f["ylen"]))
        n = 11
        for i in range(n):
            src, cli = bench_data.note_cli(i, "c3" if i % 3 == 0 else "c2", n_sources=4)
            b.add_note(host.NoteArgs.from_cli(src, cli))
        infos = b.assemble(host.SeededNoise()).infos            # planning only (CPU)
        bins = shard.balanced_partition([shard.note_cost(x) for x in infos], world)
        mine = bins[rank]
        # stand-in for the render: an array of the planned length tagged with the note index
        outs = [np.full(infos[i]["n_total"], float(i), dtype=np.float32) for i in mine]
        allout = shard.gather_outputs(mine, outs)
        ok = sorted(allout) == list(range(n)) and all(len(allout[i]) == infos[i]["n_total"] and allout[i][0] == i for i in range(n))
        loads = [sum(shard.note_cost(infos[i]) for i in bn) for bn in bins]
        q.put((rank, ok, loads, list(shard.contiguous_range(n, rank, world))))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    loads = res[0][2]
    assert max(loads) / max(1, min(loads)) < 1.5            # cost-balanced: full-flag notes cost up to 4x
    cover = sorted(i for r in res for i in r[3])
    assert cover == list(range(11))


def test_numa_binding_helper_never_raises(monkeypatch):
    """shard.bind_rank_to_gpu_numa is a launch nicety: without NVML / a GPU it reports why and leaves the affinity alone."""
    import os
    from goofer_b200 import shard
    assert shard._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert shard._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    r = shard.bind_rank_to_gpu_numa(0)
    assert r["bound"] is False and "why" in r
    assert os.sched_getaffinity(0) == before
    monkeypatch.setenv("GOOFER_NUMA_BIND", "0")
    assert shard.bind_rank_to_gpu_numa(0) == {"bound": False, "why": "disabled"}
