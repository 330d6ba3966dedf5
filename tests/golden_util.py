"""Helpers shared by the golden-fixture tests."""
import numpy as np
import torch

from detweights import det_array, det_state_dict  # noqa: F401  (tests/golden on sys.path via conftest)

W0 = [[[3, 0, 1], [4, 1, 1], [3, 1, 1]],
      [[0, 1, 0, 0, 1], [2, 1, 2, 1, 0], [3, 1, 1, 1, 0], [1, 1, 2, 0, 0], [3, 0, 2, 0, 0], [5, 3, 2, 1, 0],
       [0, 5, 0, 1, 0]]]
W1 = [[[1, 1, 0], [1, 3, 0], [3, 4, 0]],
      [[1, 1, 0, 0, 0], [0, 1, 1, 1, 1], [3, 1, 2, 3, 0], [3, 0, 2, 2, 0], [0, 1, 2, 0, 0], [2, 1, 1, 3, 0],
       [4, 0, 2, 2, 0]]]
C0 = [[8, [0, 0, 5, 2], [0, 2, 8, 8], [0, 5, 1, 4]], [[3, 3], [3, 2], [3, 0]]]
C1 = [[2, [1, 0, 3, 6], [0, 1, 2, 8], [2, 0, 6, 1]], [[2, 3], [3, 1], [4, 4]]]
C2 = [[5, [0, 0, 4, 1], [3, 2, 0, 1], [5, 6, 5, 0]], [[1, 3], [4, 3], [2, 2]]]

# tag -> (paper, config, classes, agg_size, repeats, aux_cell)   [mirrors make_golden.gen_nets]
NETS = {
    "W0": ("wacv", W0, 19, 64, 2, False),
    "W1": ("wacv", W1, 19, 64, 2, False),
    "W0cv": ("wacv", W0, 11, 64, 2, False),
    "C0search": ("cvpr", C0, 21, 48, 1, True),
    "C1search": ("cvpr", C1, 21, 48, 1, True),
    "C2final": ("cvpr", C2, 21, 64, 2, False),
    "D0depth": ("cvpr", C0, 1, 64, 2, False),
}


def keys_shapes(fx, prefix=""):
    keys = [str(k) for k in fx[prefix + "keys"]]
    shapes = [tuple(int(v) for v in str(s).split(",") if v != "") for s in fx[prefix + "shapes"]]
    return list(zip(keys, shapes))


def sub_state(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))
