"""CPU: host-side logic of the one-process-per-GPU deployment — shard geometry, result merge and the
timing / counter reductions of bench.py — with a world_size-2 gloo group.  The GPU side of the same path
(ltg_scan_shard) is covered by tests/test_gpu_parity.py::test_shards_reproduce_the_whole_record."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fasim_b200 as fb


def reference_segments(n, cut=5000, overlap=100):
    """cutSequence (fastsim.h:71-90): segment k starts at k*(cut-overlap) while that is < n."""
    out, pos = [], 0
    while pos < n:
        out.append((pos, min(cut, n - pos)))
        pos += cut - overlap
    return out


@pytest.mark.parametrize("n", [0, 1, 99, 100, 4899, 4900, 4901, 5000, 5001, 9800, 123457, 1_000_000])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shards_partition_the_segment_list(n, world):
    segs = reference_segments(n)
    seen = []
    for r in range(world):
        first, count, first_byte, n_bytes = fb.shard_segments(n, world, r)
        mine = segs[first:first + count]
        seen += mine
        if count:
            assert first_byte == mine[0][0]
            assert first_byte + n_bytes == mine[-1][0] + mine[-1][1]          # the byte range covers its segments exactly
        else:
            assert n_bytes == 0
    assert seen == segs                                                      # every segment once, in order
    counts = [fb.shard_segments(n, world, r)[1] for r in range(world)]
    assert max(counts) - min(counts) <= 1                                    # balanced


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count, first_byte, n_bytes = fb.shard_segments(n, world, rank)
    # stand-in for the per-rank scan: one "row" per owned segment carrying its global coordinates
    rows = [{"seg": first + k, "start": (first + k) * 4900} for k in range(count)]
    gathered = [None] * world
    dist.all_gather_object(gathered, rows)
    merged = fb.merge_shard_rows(gathered)
    # reductions exactly as bench.py does them: time = MAX over ranks, work counters = SUM
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cells = torch.tensor([float(sum(min(5000, n - r["start"]) for r in rows))], dtype=torch.float64)
    dist.all_reduce(cells, op=dist.ReduceOp.SUM)
    dist.barrier()
    if rank == 0:
        q.put((merged, float(t[0]), float(cells[0])))
    dist.destroy_process_group()


def test_two_rank_gloo_gather_and_reductions():
    n, world = 123457, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged, tmax, cells = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    segs = reference_segments(n)
    assert [r["start"] for r in merged] == [s for s, _ in segs]              # rank order == segment order
    assert tmax == 11.0
    assert cells == float(sum(l for _, l in segs))


def _queue_worker(rank, world, port, n_jobs, q):
    """bench.py's multi-query path on CPU: every rank pulls (lncRNA, DNA part) jobs from ONE atomic counter (the c10d
    store's fetch-and-add) and ships its result blobs to rank 0 through gather_object on a gloo group."""
    import pickle
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    queue = bench.JobQueue(dist)
    seen = []
    for epoch in range(2):                      # two "steps": the counter restarts under a new key
        queue.reset()
        mine = []
        while True:
            j = queue.next()
            if j >= n_jobs:
                break
            mine.append(j)
        payload = pickle.dumps([(b"x" * bench.TRI_BYTES * len(mine), bytes(mine))])
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(payload, gathered, dst=0)
        if rank == 0:
            jobs = sorted(j for pl in gathered for _, text in pickle.loads(pl) for j in text)
            seen.append(jobs)
        dist.barrier()
    if rank == 0:
        q.put(seen)
    dist.destroy_process_group()


def test_two_rank_atomic_job_queue():
    n_jobs, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_queue_worker, args=(r, world, port, n_jobs, q)) for r in range(world)]
    for p in procs:
        p.start()
    seen = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert seen == [list(range(n_jobs))] * 2                                 # every job exactly once per step


def test_bench_workload_definitions():
    """The synthetic workloads are the ones SURVEY.md 8(d) pins: generator, seeds, lncRNA lengths, cell counts."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from _harness import splitmix_bases
    assert bench.splitmix_bases(1001, 64).tobytes().decode() == splitmix_bases(1001, 64)
    assert bench.splitmix_bases(1001, 10, 54).tobytes().decode() == splitmix_bases(1001, 64)[54:]
    qs = bench.synthetic_queries(64)
    assert len(qs) == 64 and all(1000 <= len(s) <= 10000 for _, s in qs)
    assert abs(sum(len(s) for _, s in qs) - 3.5e5) < 0.5e5                   # SURVEY: sum of m ~ 3.5e5
    assert bench.region_cells(100_000_000, 3000) == 102_040_800 * 3000 * 48  # 20 408 full segments + one of 800 bp
    for name in ("demo", "meg3", "h19", "malat1", "neat1"):
        wl, rn, rna, recs, flags, params = bench.real_config(name)
        assert len(rna) > 1000 and len(recs) in (1, 532)
