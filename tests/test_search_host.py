"""Host logic of the outer search loop (nas-segm-pytorch_b200/engine/search.py) on CPU: the
early-stopping judge against decisions recorded from the reference's own class, the genotypes.out line format, the
per-candidate recipe with a fake engine (call order, Polyak, validation points, interruption) and a two-rank gloo round
(identical sampling on every rank, one candidate per rank, one all-gather, controller update in slot order, rank-0 log)."""
import io
import os
import re
import socket
import sys
import types

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_task_performer_matches_reference_decisions():
    """Decisions, running maximum, tolerance and step counter recorded from src/helpers/utils.py:207-243 (np.random.seed
    123 / 5) -- the judge uses numpy's global RNG exactly like the reference."""
    from nas_segm_b200.engine.search import TaskPerformer
    vals = np.round(np.abs(np.sin(np.arange(1, 41) * 0.7)) * 0.05, 5)
    np.random.seed(123)
    tp = TaskPerformer(maxval=0.01, delta=0.9)
    dec = [int(tp.step(float(v))) for v in vals]
    want = [1] * 40
    for i in (8, 17, 26, 35):
        want[i] = 0
    assert dec == want
    assert abs(tp.maxval - 0.017253561571911324) < 1e-15 and tp.delta == 0.9 and tp.n_steps == 74
    np.random.seed(5)
    tp = TaskPerformer(maxval=0.02, delta=0.9)
    tp.n_steps = 98                                    # crosses the 100-step point of the tolerance schedule
    assert [int(tp.step(0.019)) for _ in range(6)] == [1] * 6
    assert abs(tp.maxval - 0.019941480149400996) < 1e-15 and abs(tp.delta - 0.81) < 1e-12 and tp.n_steps == 104


def test_genotype_log_line_format():
    from nas_segm_b200.engine.search import GenotypeLog
    buf = io.StringIO()
    log = GenotypeLog(buf)
    geno = [[8, [0, 0, 5, 2], [0, 2, 8, 8], [0, 5, 1, 4]], [[3, 3], [3, 2], [3, 0]]]
    log.write(0.123456, 7, 2847123, 12.34567, geno)
    log.write(0, 8, 10, 1, [[0]])
    log.close()
    lines = buf.getvalue().splitlines()
    assert lines[0] == "reward: 0.1235, epoch: 7, params: 2847123, epoch_time: 12.3457, genotype: " + str(geno)
    assert re.fullmatch(r"reward: \d+\.\d{4}, epoch: \d+, params: \d+, epoch_time: \d+\.\d{4}, genotype: .+", lines[1])


class _FakeEngine:
    def __init__(self, rewards):
        self.calls, self.rewards = [], list(rewards)

    def train_task0(self, Xy, seg, optim_dec, epoch, *a, **kw):
        assert isinstance(optim_dec, torch.optim.Adam) and kw["polyak_decay"] == 0.9
        self.calls.append(("t0", epoch))
        with torch.no_grad():
            for p in seg.module.decoder.parameters():
                p.add_(1.0)                             # "training" moves the weights away from the Polyak average
            for a_, p in zip(kw["avg_param"], seg.module.decoder.parameters()):
                a_.mul_(0.5).add_(p, alpha=0.5)

    def train_segmenter(self, seg, loader, optim_enc, optim_dec, epoch, *a, **kw):
        assert isinstance(optim_enc, torch.optim.SGD) and kw["polyak_decay"] == 0.99
        self.calls.append(("t1", epoch))

    def validate(self, seg, loader, epoch, epoch_segm, **kw):
        self.calls.append(("val", epoch_segm, kw["num_classes"]))
        return self.rewards.pop(0)


def _args(**over):
    a = types.SimpleNamespace(
        num_tasks=2, enc_optim="sgd", dec_optim="adam", enc_lr=[1e-3, 1e-3], dec_lr=[3e-3, 3e-3], enc_mom=[0.9, 0.9],
        dec_mom=[0.9, 0.9], enc_wd=[1e-5, 1e-5], dec_wd=[1e-5, 1e-5], do_polyak=True, num_segm_epochs=[4, 2], val_every=[2, 2],
        batch_size=[64, 32], freeze_bn=[False, False], do_kd=False, kd_coeff=0.3, dec_grad_clip=3.0, enc_grad_clip=3.0,
        dec_aux_weight=0.15, print_every=20, num_classes=[21, 21], val_omit_classes=[0], segm_crit=None, kd_crit=None)
    for k, v in over.items():
        setattr(a, k, v)
    return a


def _segmenter():
    m = types.SimpleNamespace(encoder=torch.nn.Linear(2, 2), decoder=torch.nn.Linear(2, 2))
    seg = torch.nn.Module()
    seg.module = torch.nn.Module()
    seg.module.encoder, seg.module.decoder = m.encoder, m.decoder
    return seg


def test_evaluate_candidate_recipe_and_interruption():
    from nas_segm_b200.engine.search import evaluate_candidate, make_task_performers
    args = _args()
    np.random.seed(0)
    eng = _FakeEngine([0.5, 0.6, 0.7])
    seg = _segmenter()
    w0 = seg.module.decoder.weight.detach().clone()
    tasks = []
    reward, n_ep = evaluate_candidate(seg, {}, [], [], args, make_task_performers(args.num_segm_epochs, args.val_every), engine=eng,
                                      set_task=tasks.append)
    assert eng.calls == [("t0", 0), ("t0", 1), ("val", 1, 21), ("t0", 2), ("t0", 3), ("val", 3, 21), ("t1", 0), ("t1", 1), ("val", 1, 21)]
    assert tasks == [0, 1] and n_ep == 6 and abs(reward - 0.7) < 1e-12
    # after every task0 epoch the Polyak average was copied into the decoder: the weights are NOT w0 + 4
    assert not torch.allclose(seg.module.decoder.weight, w0 + 4.0)
    # a candidate far below the running maximum is interrupted at its first validation point
    ps = make_task_performers(args.num_segm_epochs, args.val_every)
    ps[0][0].maxval, ps[0][0].delta = 0.9, 0.0
    eng = _FakeEngine([0.01])
    reward, n_ep = evaluate_candidate(_segmenter(), {}, [], [], args, ps, engine=eng)
    assert eng.calls == [("t0", 0), ("t0", 1), ("val", 1, 21)] and n_ep == 2 and abs(reward - 0.01) < 1e-12
    # an engine that swallowed a RuntimeError returns 0: the candidate's reward is 0
    eng = _FakeEngine([0, 0, 0])
    np.random.seed(1)
    reward, _ = evaluate_candidate(_segmenter(), {}, [], [], _args(num_tasks=1), make_task_performers([4], [2]), engine=eng)
    assert reward == 0.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, logpath, sync_every=1, rounds=2):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from nas_segm_b200 import parallel
    from nas_segm_b200.engine.search import GenotypeLog, search_rounds
    parallel.init("gloo")
    rng = np.random.RandomState(7)                     # identical controller state on every rank
    built, updates = [], []

    def sample(rnd, slot):
        return ([[rnd, slot, int(rng.randint(0, 9))]], 0.5 + slot, -1.0 - rnd)

    def build(cfg):
        built.append(cfg)
        return torch.nn.Linear(3, 5)                   # 20 parameters

    def evaluate(seg, cfg):
        return (0.1 * (cfg[0][0] + 1) + 0.01 * cfg[0][1], 3) if cfg[0][1] == 0 else 0   # slot 1 "fails" -> reward 0

    log = GenotypeLog(logpath) if rank == 0 else None
    hist = search_rounds(rounds, sample, build, evaluate, update_fn=updates.append, log=log, first_epoch=10, sync_every=sync_every)
    if log is not None:
        log.close()
    q.put((rank, [h[:, 0].tolist() for h in hist], built, updates))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_search_rounds_gloo(tmp_path):
    world, port = 2, _free_port()
    logpath = str(tmp_path / "genotypes.out")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, logpath)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, rew0, built0, upd0), (_, rew1, built1, upd1) = res
    assert rew0 == rew1 and upd0 == upd1                               # every rank ends the round with the same information
    assert all(abs(a - b) < 1e-6 for a, b in zip(rew0[0], [0.1, 0.0])) and all(abs(a - b) < 1e-6 for a, b in zip(rew0[1], [0.2, 0.0]))
    assert [c[0][:2] for c in built0] == [[0, 0], [1, 0]] and [c[0][:2] for c in built1] == [[0, 1], [1, 1]]   # own slot only
    assert [[u[0][0][:2] for u in rnd] for rnd in upd0] == [[[0, 0], [0, 1]], [[1, 0], [1, 1]]]               # slot order
    assert [u[2] for u in upd0[0]] == [0.5, 1.5] and [u[3] for u in upd0[1]] == [-2.0, -2.0]
    lines = open(logpath).read().splitlines()
    assert len(lines) == 4
    assert [int(re.search(r"epoch: (\d+)", ln).group(1)) for ln in lines] == [10, 11, 12, 13]
    assert all("params: 20," in ln for ln in lines) and lines[0].startswith("reward: 0.1000,") and lines[1].startswith("reward: 0.0000,")


def test_two_rank_blocked_exchange_gloo(tmp_path):
    """sync_every = k: k rounds per all-gather (a rank whose candidate stopped early moves on instead of waiting); history,
    controller updates and log lines are the same as with one exchange per round, in (round, slot) order."""
    ctx = mp.get_context("spawn")
    out = {}
    for sync_every in (1, 3):
        world, port = 2, _free_port()
        logpath = str(tmp_path / ("genotypes_%d.out" % sync_every))
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, world, port, q, logpath, sync_every, 3)) for r in range(world)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=120) for _ in range(world))
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert res[0][1] == res[1][1] and res[0][3] == res[1][3]
        lines = [re.sub(r"epoch_time: [0-9.]+", "epoch_time: T", ln) for ln in open(logpath).read().splitlines()]
        out[sync_every] = (res[0][1], res[0][3], lines)
    assert out[1][0] == out[3][0] and out[1][1] == out[3][1] and out[1][2] == out[3][2] and len(out[3][2]) == 6


def test_uniform_sampler_covers_the_reference_search_space():
    """engine.search.uniform_sampler: ranges of src/rl/micro_controllers.py:94-120,180-262; deterministic in (seed, round, slot)."""
    from nas_segm_b200.engine.search import uniform_sampler
    s = uniform_sampler(seed=3)
    seen_ops, seen_pos = set(), set()
    for rnd in range(40):
        for slot in range(4):
            (ctx, conns), ent, logp = s(rnd, slot)
            assert s(rnd, slot)[0] == [ctx, conns] and abs(ent + logp) < 1e-12
            assert len(conns) == 3 and all(len(c) == 2 and 0 <= c[0] < 4 + i and 0 <= c[1] < 4 + i for i, c in enumerate(conns))
            assert 0 <= ctx[0] < 11 and len(ctx) == 4
            for layer, cfg in enumerate(ctx[1:], start=1):
                assert len(cfg) == 4 and all(0 <= p < 1 + 3 * (layer - 1) for p in cfg[:2]) and all(0 <= o < 11 for o in cfg[2:])
                seen_ops.update(cfg[2:])
                seen_pos.add((layer, cfg[0]))
    assert seen_ops == set(range(11)) and (3, 6) in seen_pos
    assert s(0, 0)[0] != s(0, 1)[0] or s(0, 0)[0] != s(1, 0)[0]
