"""CPU, world_size 2 over gloo: the N>1 host logic (shard planning + final padded all-gather of mel/audio)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from f5_tts_b200.sharding import gather_padded, plan_shards, shard_bounds


def test_shard_bounds_cover_everything():
    for n in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_plan_shards_sorted_by_duration():
    dur = [469, 1875, 670, 1674, 871, 1473, 1072, 1272]
    shards = plan_shards(dur, 2)
    assert sorted(i for s in shards for i in s) == list(range(8))
    assert min(dur[i] for i in shards[0]) >= max(dur[i] for i in shards[1])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        durations = [12, 30, 18, 25, 9]
        mine = plan_shards(durations, world)[rank]
        L = max(durations[i] for i in mine)
        mel = torch.zeros(len(mine), L, 4)
        for j, i in enumerate(mine):
            mel[j, : durations[i]] = float(i + 1)
        lens = torch.tensor([durations[i] for i in mine])
        got, glens = gather_padded(mel, lens)
        ok = True
        for r in range(world):
            idx = plan_shards(durations, world)[r]
            ok &= got[r].shape[0] == len(idx) and [int(v) for v in glens[r]] == [durations[i] for i in idx]
            for j, i in enumerate(idx):
                ok &= bool((got[r][j, : durations[i]] == float(i + 1)).all()) and bool((got[r][j, durations[i]:] == 0).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gather_padded_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == [(0, True), (1, True)]
