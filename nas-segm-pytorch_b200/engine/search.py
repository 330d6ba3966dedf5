"""Outer search loop, one sampled candidate per GPU (reference: src/main_search.py:523-529,548-680).

The reference evaluates candidates strictly one after another: sample a decoder, train it on the cached task0 features,
(optionally) train end to end, validate, hand the reward to the controller, write one line to ``genotypes.out``, repeat.
Here a *round* evaluates ``world_size`` candidates at once -- rank r owns candidate r of the round (its own encoder copy,
decoder, optimiser state and task0 cache) -- and the only exchange is the all-gather of one 16-byte reward record per
rank (``parallel.gather_records``).  Every rank then sees the round's records in candidate order, so the controller
update can be replayed identically everywhere.

What stays outside (reused from the reference as-is): the RL controller and its PPO/REINFORCE update,
the datasets, the checkpoint saver.  They enter through plain callables:

    sample_fn(round_idx, slot)                  -> (decoder_config, entropy, log_prob)     agent.controller.sample()
    build_fn(decoder_config)                    -> segmenter                              create_segmenter()
    update_fn([(config, reward, entropy, log_prob), ...])                                 train_agent(), in slot order

``evaluate_candidate`` is the per-candidate recipe of main_search.py:551-656 on top of this package's engine functions
(``train_task0`` / ``train_segmenter`` / ``validate`` replay CUDA graphs when ``config().cuda_graphs`` is on).
"""
import gc
import queue
import threading
import time

import numpy as np
import torch

from .. import parallel
from ..helpers.utils import apply_polyak, compute_params, init_polyak

GENOTYPE_LINE = "reward: {:.4f}, epoch: {}, params: {}, epoch_time: {:.4f}, genotype: {}\n"  # main_search.py:669-673


class TaskPerformer(object):
    """Early-stopping judge of a partially trained candidate (reference: src/helpers/utils.py:207-243): keeps a running
    estimate of the reward reached at this validation point and lets a candidate continue if it is above it, or -- with a
    tolerance `delta` that shrinks after 100/200/300/400 judged candidates -- not too far below.  Same arithmetic and the
    same use of numpy's global RNG as the reference, so a seeded search takes the same decisions."""

    SCHEDULE = {100: 0.9, 200: 0.8, 300: 0.7, 400: 0.6}

    def __init__(self, maxval, delta=0.3):
        self.maxval, self.delta, self.n_steps, self.decay = maxval, delta, 0, 0.99

    def step(self, newval):
        self.delta *= self.SCHEDULE.get(self.n_steps, 1.0)
        self.n_steps += 1
        self.maxval = self.decay * self.maxval + (1.0 - self.decay) * newval
        if newval > self.maxval:
            self.n_steps += 1
            return True
        return bool(newval > self.maxval * (1.0 - np.random.uniform(0.0, high=self.delta)))


def make_task_performers(num_segm_epochs, val_every):
    """One judge per (task, validation point) (main_search.py:523-529)."""
    return [[TaskPerformer(maxval=0.01, delta=0.9) for _ in range(n // v)] for n, v in zip(num_segm_epochs, val_every)]


def create_optimisers(optim_enc, optim_dec, lr_enc, lr_dec, mom_enc, mom_dec, wd_enc, wd_dec, param_enc, param_dec):
    """'sgd' | 'adam' for each half (reference: src/utils/solvers.py:6-52)."""
    def make(kind, params, lr, mom, wd, what):
        if kind == "sgd":
            return torch.optim.SGD(params, lr=lr, momentum=mom, weight_decay=wd)
        if kind == "adam":
            return torch.optim.Adam(params, lr=lr, weight_decay=wd)
        raise ValueError("Unknown {} Optimiser: {}".format(what, kind))
    return make(optim_enc, param_enc, lr_enc, mom_enc, wd_enc, "Encoder"), make(optim_dec, param_dec, lr_dec, mom_dec, wd_dec, "Decoder")


class GenotypeLog(object):
    """``genotypes.out`` writer in the reference's line format.  Lines are formatted and flushed by a background thread, so
    the rank that logs never waits on the file system between two rounds; ``close()`` drains the queue."""

    def __init__(self, path_or_file):
        self._own = isinstance(path_or_file, str)
        self._f = open(path_or_file, "a") if self._own else path_or_file
        self._q = queue.Queue()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            self._f.write(GENOTYPE_LINE.format(*item))
            self._f.flush()

    def write(self, reward, epoch, params, epoch_time, genotype):
        self._q.put((float(reward), int(epoch), params, float(epoch_time), genotype))

    def close(self):
        self._q.put(None)
        self._t.join()
        if self._own:
            self._f.close()


def evaluate_candidate(segmenter, Xy_train, train_loader, val_loader, args, task_ps, engine=None, epoch=0, set_task=None):
    """Train and validate ONE candidate: for each task (0 = decoder only on the cached features, 1 = end to end) create the
    optimisers, run ``num_segm_epochs[task]`` epochs, copy the Polyak average in after each, validate every
    ``val_every[task]`` epochs and let ``task_ps[task][k].step(reward)`` decide whether the candidate goes on.
    Returns (reward, n_epochs_run).  `args` carries the reference's option names (utils/default_args.py); `engine` defaults
    to this package's trainer / inference modules; `set_task(task_idx)` re-configures the data loaders (crop sizes, batch
    size) where the caller has real datasets."""
    if engine is None:
        from . import inference, trainer

        class engine:  # noqa: N801
            train_task0, train_segmenter, validate = trainer.train_task0, trainer.train_segmenter, inference.validate
    reward, epochs_run = 0.0, 0
    for task_idx in range(args.num_tasks):
        if set_task is not None:
            set_task(task_idx)
        optim_enc, optim_dec = create_optimisers(
            args.enc_optim, args.dec_optim, args.enc_lr[task_idx], args.dec_lr[task_idx], args.enc_mom[task_idx],
            args.dec_mom[task_idx], args.enc_wd[task_idx], args.dec_wd[task_idx], segmenter.module.encoder.parameters(),
            segmenter.module.decoder.parameters())
        averaged = segmenter.module.decoder if task_idx == 0 else segmenter
        avg_param = init_polyak(args.do_polyak, averaged)
        task_reward = None
        for epoch_segm in range(args.num_segm_epochs[task_idx]):
            if task_idx == 0:
                engine.train_task0(Xy_train, segmenter, optim_dec, epoch_segm, args.segm_crit, args.kd_crit, args.batch_size[0],
                                   args.freeze_bn[0], args.do_kd, args.kd_coeff, args.dec_grad_clip, args.do_polyak,
                                   avg_param=avg_param, polyak_decay=0.9, aux_weight=args.dec_aux_weight)
            else:
                engine.train_segmenter(segmenter, train_loader, optim_enc, optim_dec, epoch_segm, args.segm_crit,
                                       args.freeze_bn[1], args.enc_grad_clip, args.dec_grad_clip, args.do_polyak,
                                       args.print_every, aux_weight=args.dec_aux_weight, avg_param=avg_param, polyak_decay=0.99)
            epochs_run += 1
            apply_polyak(args.do_polyak, averaged, avg_param)
            if (epoch_segm + 1) % args.val_every[task_idx] == 0:
                task_reward = engine.validate(segmenter, val_loader, epoch, epoch_segm, num_classes=args.num_classes[task_idx],
                                              print_every=args.print_every, omit_classes=args.val_omit_classes)
                judge = task_ps[task_idx][(epoch_segm + 1) // args.val_every[task_idx] - 1]
                if not judge.step(task_reward):
                    return float(task_reward), epochs_run  # interrupted: the reward so far is what the controller sees
        if task_reward is not None:  # the reference leaves `reward` unbound when a task never validates (Appendix A.14)
            reward = float(task_reward)
    return reward, epochs_run


def uniform_sampler(seed=9314, enc_num_layers=4, dec_num_cells=3, cell_num_layers=4, num_ops=11):
    """`sample_fn` over the CVPR search space of the reference controller (src/rl/micro_controllers.py:94-120,180-262:
    decoder block i draws two inputs from enc_num_layers + i candidates; cell layer 0 draws one op, layer l >= 1 draws two
    positions from 1 + 3(l-1) candidates and two ops), uniform instead of LSTM-parametrised -- the controller itself is
    outside the hot path and its freshly initialised policy (weights in +-0.1) is within a few percent
    of uniform.  Deterministic in (seed, round, slot), identical on every rank.  Entropy / log-prob of the uniform policy
    are returned so that `update_fn` sees the controller's tuple shape."""
    def sample(rnd, slot):
        rs = np.random.RandomState([seed & 0x7fffffff, int(rnd), int(slot)])
        logp = 0.0
        conns = []
        for i in range(dec_num_cells):
            n = enc_num_layers + i
            conns.append([int(rs.randint(n)), int(rs.randint(n))])
            logp -= 2 * np.log(n)
        ctx = [int(rs.randint(num_ops))]
        logp -= np.log(num_ops)
        for layer in range(1, cell_num_layers):
            n = 1 + 3 * (layer - 1)
            ctx.append([int(rs.randint(n)), int(rs.randint(n)), int(rs.randint(num_ops)), int(rs.randint(num_ops))])
            logp -= 2 * np.log(n) + 2 * np.log(num_ops)
        return [ctx, conns], float(-logp), float(logp)
    return sample


def search_rounds(n_rounds, sample_fn, build_fn, evaluate_fn, update_fn=None, log=None, first_epoch=0, sync_every=1):
    """Rank-parallel search: every round, slot s of the round (s = 0 .. world-1) is sampled by `sample_fn(round, s)` -- on
    EVERY rank, so that all ranks hold identical controller inputs -- rank r builds and evaluates slot r only, one
    all-gather exchanges the rewards, and `update_fn` receives the round's samples in slot order on every rank.  Rank 0
    appends one genotype line per candidate to `log`.  Returns the list of per-round reward tensors [world, RECORD].

    `sync_every` = k > 1 exchanges the records of k consecutive rounds with ONE all-gather (k samples per rank from one
    policy snapshot): a rank whose candidate was stopped early by its TaskPerformer moves on to its next candidate instead
    of idling until the slowest rank of the round is done (SURVEY 8e "stragglers").  Controller updates and log lines are
    then replayed in (round, slot) order after the exchange; k = 1 is the synchronous schedule.

    `evaluate_fn(segmenter, decoder_config) -> reward | (reward, miou, macc, fwiou) | (reward, n_epochs)` as returned by
    `evaluate_candidate`; an engine function that swallowed a RuntimeError returns 0, which is recorded as reward 0
    (src/helpers/utils.py:172-187).  The logged epoch_time is the candidate's wall time per training epoch, as in the
    reference (main_search.py:666-668)."""
    rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
    world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
    history, pending = [], []
    sync_every = max(int(sync_every), 1)

    def exchange():
        if not pending:
            return
        tables = parallel.gather_records([m for _, _, m in pending])  # [world, k, RECORD]: the single collective of the block
        for j, (rnd, samples, _) in enumerate(pending):
            table = tables[:, j, :]
            history.append(table)
            if update_fn is not None:
                update_fn([(samples[s][0], float(table[s, 0]), samples[s][1], samples[s][2]) for s in range(world)])
            if log is not None and rank == 0:
                for s in range(world):
                    log.write(float(table[s, 0]), first_epoch + rnd * world + s, int(table[s, 3]), float(table[s, 2]), samples[s][0])
        del pending[:]

    for rnd in range(n_rounds):
        samples = [sample_fn(rnd, s) for s in range(world)]
        t0 = time.time()
        segmenter = build_fn(samples[rank][0])
        n_params = compute_params(segmenter)[1] if hasattr(segmenter, "named_parameters") else 0
        out = evaluate_fn(segmenter, samples[rank][0])
        del segmenter
        # a candidate's captured graphs hang off its modules and their closures point back at the modules: collect the cycle
        # now, so the graphs' memory returns to the shared capture pool before the next candidate captures (graphs.capture
        # skips the gc.collect() + empty_cache() torch's own context manager would do at every capture)
        gc.collect()
        n_epochs = 1
        if isinstance(out, (tuple, list)) and len(out) == 2:  # (reward, epochs run) from evaluate_candidate
            out, n_epochs = float(out[0]), max(int(out[1]), 1)
        mine = list(out) if isinstance(out, (tuple, list)) else [float(out), 0.0, 0.0, 0.0]
        # record = (reward, miou, time per epoch, #params): the last two fields feed the genotype log
        mine = [float(mine[0]), float(mine[1]) if len(mine) > 1 else 0.0, (time.time() - t0) / n_epochs, float(n_params)]
        pending.append((rnd, samples, mine))
        if len(pending) == sync_every:
            exchange()
    exchange()
    return history
