"""N > 1 host logic on CPU: world_size-2 gloo processes (127.0.0.1 rendezvous)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dart_env_b200.parallel import gather_batch, max_over_ranks, shard_worlds


def test_shard_worlds_partitions_exactly():
    for total in (1, 7, 4096, 131072, 100003):
        for ws in (1, 2, 3, 4, 8):
            seen = []
            for r in range(ws):
                off, cnt = shard_worlds(total, r, ws)
                seen += list(range(off, off + cnt)) if total < 10000 else [off, off + cnt]
                assert cnt in (total // ws, total // ws + 1)
            if total < 10000:
                assert seen == list(range(total))
            else:
                assert seen[0] == 0 and seen[-1] == total and all(seen[2 * i + 1] == seen[2 * i + 2] for i in range(ws - 1))
    assert shard_worlds(131072, 3, 8) == (49152, 16384)  # BASELINE config 4: 16384 worlds / GPU
    with pytest.raises(ValueError):
        shard_worlds(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        total, nobs = 10, 3
        off, cnt = shard_worlds(total, rank, ws)
        assert cnt == 5
        # obs row w of the global batch = [w, w, w]: the gathered batch must be ordered by world id
        local = (torch.arange(off, off + cnt, dtype=torch.float32)[:, None]).repeat(1, nobs)
        full = gather_batch(local)
        ok = full.shape == (total, nobs) and torch.equal(full[:, 0], torch.arange(total, dtype=torch.float32))
        t = max_over_ranks(1.0 + rank)
        q.put((rank, bool(ok), t))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gather_and_max_over_ranks_gloo_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, 2.0), (1, True, 2.0)]
