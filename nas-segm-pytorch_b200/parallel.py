"""Multi-GPU plumbing of the search loop: ONE sampled candidate per GPU, one process per GPU, no data-path collective.

The reference evaluates candidates strictly one after another on a 2-GPU ``nn.DataParallel`` (src/main_search.py:507,548-680).
On an 8 x B200 node every rank trains / validates its own candidate (own encoder copy, decoder, optimiser state and cached
task0 tensors -- a few GB of 180); the only exchange is one all-gather of a 16-byte record per rank per round:
``(reward, miou, macc, fwiou)``.  Every rank then holds the records of the whole round in candidate order, so the
controller update (src/main_search.py:657-660) can be replayed identically everywhere (or on rank 0 only).
"""
import os

import torch
import torch.distributed as dist

RECORD = 4  # floats per candidate record: reward, mean IoU, mean accuracy, frequency-weighted IoU


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment.  Returns (rank, world, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cuda = torch.cuda.is_available()
    device = torch.device("cuda", local) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if cuda else "gloo"), rank=rank, world_size=world)
    return rank, world, device


def shard(n_candidates, rank, world):
    """Indices of the candidates rank `rank` evaluates in a round of `n_candidates` (round-robin: candidate i -> rank i % world)."""
    return list(range(rank, n_candidates, world))


def gather_records(records, device=None):
    """All-gather this rank's records ([k, RECORD] floats, k identical on all ranks) -> tensor [world, k, RECORD] on the host.
    The single collective of the path."""
    t = torch.as_tensor(records, dtype=torch.float32).reshape(-1, RECORD)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t[None].clone()
    world = dist.get_world_size()
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    send = t.to(dev).contiguous().view(-1)
    recv = torch.empty(world * send.numel(), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(recv, send)
    return recv.cpu().view(world, -1, RECORD)


def evaluate_round(candidates, evaluate_fn, rank=None, world=None, device=None):
    """Evaluate one round of candidates, `len(candidates)` being a multiple of the world size.

    evaluate_fn(candidate) -> (reward, miou, macc, fwiou) (an engine ``validate`` that swallowed a RuntimeError returns 0:
    record (0, 0, 0, 0)).  Returns a [len(candidates), RECORD] tensor in candidate order, identical on every rank."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    n = len(candidates)
    if n % world:
        raise ValueError("a round must hold a multiple of world_size candidates (%d vs %d)" % (n, world))
    mine = []
    for i in shard(n, rank, world):
        r = evaluate_fn(candidates[i])
        if not isinstance(r, (tuple, list)):
            r = (float(r), 0.0, 0.0, 0.0)
        mine.append([float(v) for v in r])
    allr = gather_records(mine, device)          # [world, n/world, RECORD]
    out = torch.empty((n, RECORD), dtype=torch.float32)
    for r in range(world):
        for j, i in enumerate(shard(n, r, world)):
            out[i] = allr[r, j]
    return out
