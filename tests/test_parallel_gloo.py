"""The N>1 path on CPU: two gloo ranks shard a round of candidates, evaluate them independently and exchange the
16-byte records with the single all-gather; every rank must end up with the same table in candidate order."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:  # spawned workers re-import this module without conftest.py
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_eval(c):
    # deterministic stand-in for train_task0 + validate of candidate c
    return (c * 0.01 + 0.1, c * 0.02, c * 0.03, c * 0.04)


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import nas_segm_b200  # noqa: F401
    from nas_segm_b200 import parallel
    r, w, dev = parallel.init("gloo")
    assert (r, w) == (rank, world) and dev.type == "cpu"
    cands = list(range(10, 16))                      # one round of 6 candidates on 2 ranks
    assert parallel.shard(len(cands), rank, world) == list(range(rank, 6, 2))
    seen = []

    def ev(c):
        seen.append(c)
        return _fake_eval(c) if c != 13 else 0      # candidate 13 "failed" (engine returned 0)
    table = parallel.evaluate_round(cands, ev)
    q.put((rank, table.tolist(), seen))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_round_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, t0, seen0), (_, t1, seen1) = res
    assert t0 == t1                                  # identical on every rank
    assert seen0 == [10, 12, 14] and seen1 == [11, 13, 15]   # each rank evaluated only its own candidates
    for i, c in enumerate(range(10, 16)):
        want = list(_fake_eval(c)) if c != 13 else [0.0, 0.0, 0.0, 0.0]
        assert all(abs(a - b) < 1e-6 for a, b in zip(t0[i], want))


def test_single_rank_degenerates():
    from nas_segm_b200 import parallel
    t = parallel.evaluate_round([1, 2, 3], _fake_eval, rank=0, world=1)
    assert t.shape == (3, 4) and abs(float(t[2, 0]) - 0.13) < 1e-6
