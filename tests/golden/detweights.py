"""Deterministic, platform-stable weights for parity fixtures.

Fixtures store only (key, shape) lists plus inputs/outputs; the tensors themselves are regenerated
from numpy's PCG64 keyed by crc32(key), so the same values are produced in the container that
ran the real reference (make_golden.py) and wherever the tests run."""
import zlib

import numpy as np
import torch


def det_tensor(key, shape, seed=0):
    rng = np.random.default_rng([zlib.crc32(key.encode()), seed])
    shape = tuple(int(s) for s in shape)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros((), dtype=torch.int64)
    if leaf == "running_var":
        a = rng.uniform(0.5, 1.5, shape)
    elif leaf == "running_mean":
        a = rng.normal(0.0, 0.1, shape)
    elif leaf in ("a", "b"):
        a = rng.uniform(0.5, 1.5, shape)
    elif len(shape) == 4:
        fan_in = shape[1] * shape[2] * shape[3]
        a = rng.normal(0.0, (1.5 / fan_in) ** 0.5, shape)
    elif leaf == "weight":  # BN gamma
        a = rng.uniform(0.5, 1.5, shape)
    else:  # BN beta / conv bias
        a = rng.normal(0.0, 0.1, shape)
    return torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(shape).copy())


def det_state_dict(keys_shapes, seed=0):
    """keys_shapes: iterable of (key, shape)."""
    return {k: det_tensor(k, s, seed) for k, s in keys_shapes}


def det_array(tag, shape, seed=0, kind="normal", lo=0, hi=1):
    rng = np.random.default_rng([zlib.crc32(tag.encode()), seed])
    if kind == "normal":
        return rng.normal(0.0, 1.0, shape).astype(np.float32)
    if kind == "int":
        return rng.integers(lo, hi, shape)
    if kind == "uniform":
        return rng.uniform(lo, hi, shape).astype(np.float32)
    raise KeyError(kind)


def keys_shapes_of(module):
    return [(k, tuple(v.shape)) for k, v in module.state_dict().items()]
