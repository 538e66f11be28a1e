"""N > 1 host logic on CPU: the game-index list shards across ranks like selfplay_main.py:44-47 and the throughput
counters reduce to (sum of moves, max of seconds).  world_size 2 over gloo, no GPU."""
import os
import socket

import torch.multiprocessing as tmp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from tamago_b200.selfplay.shard import shard_for_rank, reduce_counters
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_for_rank(11)
    moves, secs = reduce_counters(100 * (rank + 1), 1.0 + rank)
    import numpy as np
    from tamago_b200.selfplay.shard import gather_training_arrays
    n = 3 + 2 * rank
    got = gather_training_arrays({"value": np.full(n, rank, np.int32), "policy": np.full((n, 4), rank + 0.5)})
    if rank == 0:
        assert list(got["value"]) == [0, 0, 0, 1, 1, 1, 1, 1] and got["policy"].shape == (8, 4) and got["policy"][5, 0] == 1.5
    else:
        assert got is None
    out.put((rank, mine, moves, secs))
    dist.barrier()
    dist.destroy_process_group()


def test_shards_and_counters_world2():
    ctx = tmp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == [1, 2, 3, 4, 5, 6] and res[1][1] == [7, 8, 9, 10, 11]       # ceil(11/2) = 6, like the reference
    for _, _, moves, secs in res:
        assert moves == 300.0 and secs == 2.0


def test_split_matches_reference_rule():
    from tamago_b200.selfplay.shard import split_indices
    assert split_indices(10, 4) == [[1, 2, 3], [4, 5, 6], [7, 8, 9], [10]]
    assert split_indices(3, 8) == [[1], [2], [3]]
    assert sum(split_indices(10000, 8), []) == list(range(1, 10001))
