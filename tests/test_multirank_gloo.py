"""The N > 1 path on CPU (gloo, world_size 2): station sharding, the barrier / max-over-ranks / sum-of-units plumbing
bench.py uses, and that the union of the ranks' results equals the single-rank result (checked with the oracle as the
per-station chain, since there is no GPU here; the GPU chain is pinned to the same oracle by the -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import pytest

from fmrx import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_every_station_once():
    for n, w in [(4096, 1), (4096, 2), (4096, 8), (4097, 8), (5, 8), (0, 3), (7, 2)]:
        seen = []
        for r in range(w):
            rg = shard.shard_range(n, w, r)
            seen.extend(rg)
            assert all(shard.owner_of(s, n, w) == r for s in rg)
        assert seen == list(range(n))
        sizes = [len(shard.shard_range(n, w, r)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_shard_range_rejects_bad_requests():
    for args in [(10, 0, 0), (10, 2, 2), (10, 2, -1), (-1, 2, 0)]:
        with pytest.raises(ValueError):
            shard.shard_range(*args)
    assert list(shard.weak_range(4, 2)) == [8, 9, 10, 11]
    for args in [(4096, 4096, 8), (-1, 4096, 8), (5, 5, 8), (0, 0, 3), (3, 10, 0)]:  # a station outside the job has no owner
        with pytest.raises(ValueError):
            shard.owner_of(*args)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "real-time-software-defined-radio_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    from fmrx import shard as sh
    from fmrx import synth
    from oracle import Chain

    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sh.shard_range(n_total, world, rank)
    res = {}
    for s in mine:  # one block per station: audio + RDS bits of station s depend on nothing but station s
        raw = synth.synth_station(s, 1, 0)
        audio, _, bits, _, _ = Chain(0, 1).run(raw)
        res[s] = (audio, np.concatenate(bits) if len(bits) else np.zeros(0, np.uint8))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), stations=np.array(list(mine)), **{f"a{s}": v[0] for s, v in res.items()},
             **{f"b{s}": v[1] for s, v in res.items()})
    dist.barrier()
    # rank r pretends its timed region took (r + 1) seconds over len(mine) units
    units, tmax, rate = sh.job_throughput(float(len(mine)), float(rank + 1), dist)
    assert units == n_total and tmax == float(world) and abs(rate - n_total / world) < 1e-12
    dist.destroy_process_group()


def test_two_ranks_cover_the_batch_and_match_one_rank(tmp_path):
    import torch.multiprocessing as mp

    from fmrx import synth
    from oracle import Chain

    n_total, world = 3, 2  # ragged on purpose: ranks own 2 and 1 stations
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    got = {}
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert list(z["stations"]) == list(shard.shard_range(n_total, world, r))
        for s in z["stations"]:
            assert int(s) not in got
            got[int(s)] = (z[f"a{s}"], z[f"b{s}"])
    assert sorted(got) == list(range(n_total))
    for s in range(n_total):
        audio, _, bits, _, _ = Chain(0, 1).run(synth.synth_station(s, 1, 0))
        assert np.array_equal(got[s][0], audio)
        assert np.array_equal(got[s][1], np.concatenate(bits) if len(bits) else np.zeros(0, np.uint8))


def _bar_worker(rank, world, key, out):
    import time

    os.environ["MASTER_PORT"] = "45678"
    sys.path.insert(0, ROOT)
    import bench

    bench.os.getppid = lambda: key          # the launcher's pid keys the segment: the same for every rank of one launch
    if rank:
        time.sleep(0.1 * rank)              # late attachers wait for rank 0's segment
    bar = bench.ShmBarrier(rank, world)
    order = []
    for step in range(50):
        if step % world == rank:
            time.sleep(0.002)               # one straggler per step: nobody may pass the barrier before it arrives
        bar.wait()
        order.append(int(bar.slots.min()))
    out.put((rank, order))
    bar.wait()
    bar.close()


def test_shared_memory_barrier_between_ranks():
    """bench.ShmBarrier (the host barrier of the ingest arbitration): after wait() number k every rank's slot has reached k."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 3
    ps = [ctx.Process(target=_bar_worker, args=(r, world, os.getpid(), q)) for r in range(world)]
    for p in ps:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert all(m >= k + 1 for k, m in enumerate(got[r])), f"rank {r} passed a barrier early"


def test_gpu_map_is_identity_when_there_is_nothing_to_choose():
    """bench.pick_gpu: one rank, or as many ranks as visible GPUs (none here): LOCAL_RANK, no probe, no file."""
    sys.path.insert(0, ROOT)
    import bench

    for rank, world, local in ((0, 1, 0), (3, 8, 3)):
        idx, info = bench.pick_gpu(rank, world, local)
        assert idx == local and info["policy"] == "identity"
