"""Discrete model of the mbarrier protocol of pw_tc_ws_kernel (nas-segm-pytorch_b200/csrc/pw_tcgen05.cu; the opt-in
warp-specialised pointwise kernel was written without GPU access, so its producer / MMA / epilogue hand-shake is checked
here on the CPU).  The three roles run as Python threads under a randomised scheduler and use the SAME slot / parity
expressions as the kernel; an mbarrier is modelled with its phase bit and pending-arrival count (a parity wait passes
when the barrier's current phase differs from the awaited parity bit's phase, exactly like mbarrier.try_wait.parity).
Checked: termination (no deadlock), no A stage refilled before the MMAs that read it committed, no TMEM accumulator
overwritten before all 128 epilogue threads arrived, every tile stored exactly once and in order per buffer."""
import random
import threading

import pytest

WS_SA = 2


class MBar:
    def __init__(self, count, cv):
        self.count, self.pending, self.phase, self.cv = count, count, 0, cv

    def arrive(self, n=1):
        with self.cv:
            self.pending -= n
            assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
            if self.pending == 0:
                self.phase ^= 1
                self.pending = self.count
            self.cv.notify_all()

    def wait(self, parity, deadline):
        # try_wait.parity(P) succeeds once the phase with parity P has completed, i.e. the current phase bit != P
        with self.cv:
            while self.phase == parity:
                if not self.cv.wait(timeout=deadline):
                    raise TimeoutError("deadlock: parity %d never completed" % parity)


def run_model(my_n, seed, n_epi=8):
    rnd = random.Random(seed)
    cv = threading.Condition()
    a_full = [MBar(1, cv) for _ in range(WS_SA)]
    a_empty = [MBar(1, cv) for _ in range(WS_SA)]
    acc_full = [MBar(1, cv) for _ in range(2)]
    acc_empty = [MBar(n_epi, cv) for _ in range(2)]
    stage_owner = [None] * WS_SA      # tile whose data sits in the A stage (None = free / consumed)
    acc_owner = [None] * 2            # tile whose result sits in the accumulator
    acc_readers = [0, 0]
    stored, errors = [], []
    lock = threading.Lock()

    def jitter():
        if rnd.random() < 0.5:
            threading.Event().wait(rnd.random() * 0.002)

    def producer():
        try:
            for i in range(my_n):
                s = i % WS_SA
                if i >= WS_SA:
                    a_empty[s].wait(((i // WS_SA) - 1) & 1, 5.0)
                jitter()
                with lock:
                    assert stage_owner[s] is None, "A stage %d refilled while tile %s is still in it" % (s, stage_owner[s])
                    stage_owner[s] = i
                a_full[s].arrive()        # the TMA's complete_tx
        except Exception as e:  # noqa: BLE001
            errors.append(("producer", e))

    def mma():
        try:
            for i in range(my_n):
                s, a = i % WS_SA, i & 1
                if i >= 2:
                    acc_empty[a].wait(((i >> 1) - 1) & 1, 5.0)
                a_full[s].wait((i // WS_SA) & 1, 5.0)
                jitter()
                with lock:
                    assert stage_owner[s] == i, "MMA %d read stage %d holding %s" % (i, s, stage_owner[s])
                    assert acc_owner[a] is None and acc_readers[a] == 0, "accumulator %d overwritten (tile %s)" % (a, acc_owner[a])
                    acc_owner[a] = i
                    stage_owner[s] = None
                a_empty[s].arrive()       # tcgen05.commit
                acc_full[a].arrive()      # tcgen05.commit
        except Exception as e:  # noqa: BLE001
            errors.append(("mma", e))

    def epilogue(t):
        try:
            for i in range(my_n):
                a = i & 1
                acc_full[a].wait((i >> 1) & 1, 5.0)
                jitter()
                with lock:
                    assert acc_owner[a] == i, "epilogue thread %d of tile %d found tile %s" % (t, i, acc_owner[a])
                    acc_readers[a] += 1
                    if acc_readers[a] == n_epi:          # last reader: the accumulator is free again
                        acc_owner[a], acc_readers[a] = None, 0
                        stored.append(i)
                acc_empty[a].arrive()
        except Exception as e:  # noqa: BLE001
            errors.append(("epilogue%d" % t, e))

    threads = [threading.Thread(target=producer), threading.Thread(target=mma)] + [
        threading.Thread(target=epilogue, args=(t,)) for t in range(n_epi)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=30)
        assert not th.is_alive(), "model did not terminate"
    assert not errors, errors
    assert stored == list(range(my_n))


@pytest.mark.parametrize("my_n", [1, 2, 3, 4, 5, 8, 17])
def test_ws_barrier_protocol_model(my_n):
    for seed in range(6):
        run_model(my_n, seed)


# ------------------------------------------------------------------------------------------------------------------------
# The same kind of model for the pipelines that DID run on the B200 (they are validated by the GPU parity tests; the models
# document why their parity arithmetic cannot dead-lock under any interleaving).

def run_dw_double_buffer(total, grid, n_threads, seed):
    """dw_tile_kernel / dw_wgrad_tile_kernel / dw_dgrad_s2k3_kernel (dw_tma.cu): persistent CTA, two shared buffers, thread 0
    issues the TMA of patch t + grid into buffer buf^1 at the top of iteration `it`, every thread waits bar[buf] with parity
    (it >> 1) & 1, consumes, __syncthreads."""
    rnd = random.Random(seed)
    cv = threading.Condition()
    bar = [MBar(1, cv), MBar(1, cv)]
    content = [None, None]
    sync = threading.Barrier(n_threads)
    errors, lock = [], threading.Lock()
    tiles = list(range(0, total, grid))     # the patches of CTA 0

    def thread(tid):
        try:
            if tid == 0 and tiles:
                with lock:
                    content[0] = tiles[0]
                bar[0].arrive()
            for it, t in enumerate(tiles):
                buf = it & 1
                if tid == 0 and it + 1 < len(tiles):
                    with lock:
                        content[buf ^ 1] = tiles[it + 1]      # released by the sync that ended iteration it-1
                    bar[buf ^ 1].arrive()
                bar[buf].wait((it >> 1) & 1, 5.0)
                if rnd.random() < 0.3:
                    threading.Event().wait(rnd.random() * 0.001)
                with lock:
                    assert content[buf] == t, "thread %d iteration %d read patch %s instead of %s" % (tid, it, content[buf], t)
                sync.wait(timeout=10)
        except Exception as e:  # noqa: BLE001
            errors.append((tid, e))
            sync.abort()

    ths = [threading.Thread(target=thread, args=(i,)) for i in range(n_threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join(timeout=30)
        assert not th.is_alive()
    assert not errors, errors


def run_c3_ring(n_items, S, seed):
    """c3_tc_kernel's single-thread TMA/MMA pipeline (conv3_tcgen05.cu): S stages, item g goes to slot g % S after waiting
    done[slot] with parity ((g / S) - 1) & 1; item c is consumed after waiting full[c % S] with parity (c / S) & 1; the slot
    freed one item ago is refilled before item c is consumed.  The asynchronous agents (TMA completing, tcgen05.commit
    arriving) are separate threads with random latency."""
    rnd = random.Random(seed)
    cv = threading.Condition()
    full = [MBar(1, cv) for _ in range(S)]
    done = [MBar(1, cv) for _ in range(S)]
    slot = [None] * S
    pending, errors, lock = [], [], threading.Lock()

    def later(fn):
        th = threading.Timer(rnd.random() * 0.002, fn)
        pending.append(th)
        th.start()

    def issue(g):
        s = g % S
        if g >= S:
            done[s].wait(((g // S) - 1) & 1, 5.0)

        def land():
            with lock:
                if slot[s] is not None:
                    errors.append("slot %d overwritten while item %s unread" % (s, slot[s]))
                slot[s] = g
            full[s].arrive()
        later(land)

    try:
        g_issued = 0
        while g_issued < n_items and g_issued < S:
            issue(g_issued)
            g_issued += 1
        for c in range(n_items):
            if c >= 1 and g_issued < n_items:
                issue(g_issued)
                g_issued += 1
            s = c % S
            full[s].wait((c // S) & 1, 5.0)
            with lock:
                assert slot[s] == c, "item %d found %s in its slot" % (c, slot[s])

            def commit(s=s):
                with lock:
                    slot[s] = None
                done[s].arrive()
            later(commit)
    finally:
        for th in pending:
            th.join()
    assert not errors, errors


def test_validated_pipelines_cannot_deadlock():
    for seed in range(3):
        for total, grid in [(1, 4), (5, 2), (9, 1), (16, 3)]:
            run_dw_double_buffer(total, grid, n_threads=6, seed=seed)
        for n_items, S in [(1, 3), (2, 3), (9, 3), (20, 7), (27, 2)]:
            run_c3_ring(n_items, S, seed)
