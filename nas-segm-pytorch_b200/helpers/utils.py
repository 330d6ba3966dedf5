"""The few host utilities the engine functions depend on (reference: src/helpers/utils.py)."""
import copy
import logging
import os

logger = logging.getLogger(__name__)


class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def try_except(func):
    """Engine error convention (utils.py:172-187): a RuntimeError (OOM, CUDA failure, a genotype whose shapes do not
    fit) makes the call return 0 -> the candidate gets reward 0; every other exception propagates."""

    def wrapper_func(*args, **kwargs):
        try:
            return func(*args, **kwargs)
        except RuntimeError as e:
            if os.environ.get("NASB_RAISE"):  # debugging aid: see the failure instead of "reward 0"
                raise
            logger.warning(" %s failed and returns 0 (reference convention): %s", getattr(func, "__name__", "call"),
                           str(e).splitlines()[0] if str(e) else type(e).__name__)
            return 0

    wrapper_func.__wrapped__ = func
    wrapper_func.__name__ = getattr(func, "__name__", "wrapper_func")
    wrapper_func.__doc__ = func.__doc__
    return wrapper_func


def compute_params(model):
    n_total = n_aux = 0
    for name, p in model.named_parameters():
        n_total += p.numel()
        if "aux" in name:
            n_aux += p.numel()
    return n_total, n_total - n_aux


def init_polyak(do_polyak, module):
    if not do_polyak:
        return None
    try:
        return copy.deepcopy(list(p.data for p in module.parameters()))
    except RuntimeError:
        return None


def apply_polyak(do_polyak, module, avg_param):
    if do_polyak:
        try:
            for p, avg_p in zip(module.parameters(), avg_param):
                p.data.copy_(avg_p)
        except RuntimeError:
            return None
