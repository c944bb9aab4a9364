"""World-size-2 gloo test (CPU) of the only inference collective: the per-clip gather of track states."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    from siammot_b200.parallel import gather_track_states, shard_streams, unpack_track_states
    from siammot_b200.structures import BoxList
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n = 3 + rank
    b = BoxList(torch.arange(n * 4, dtype=torch.float32).view(n, 4) + 100 * rank, (1280, 704))
    b.add_field("scores", torch.linspace(0.5, 0.9, n))
    b.add_field("ids", torch.tensor([-1] + list(range(10 * rank, 10 * rank + n - 1))))
    b.add_field("labels", torch.ones(n, dtype=torch.int64))
    rec = gather_track_states(b, max_tracks=8)
    states = unpack_track_states(rec)
    ok = (rec.shape == (world, 8, 8) and [len(s["ids"]) for s in states] == [2, 3]
          and states[1]["ids"].tolist() == [10, 11, 12] and float(states[1]["boxes"][0, 0]) == 104.0
          and shard_streams(5, rank, world) == ([0, 2, 4] if rank == 0 else [1, 3]))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gather_track_states_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
