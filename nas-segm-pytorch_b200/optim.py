"""Gradient clipping + optimiser step + Polyak average of an engine iteration as two multi-tensor launches (scope row f2).

The reference runs ``clip_grad_norm_`` / ``optim.step()`` / a per-parameter ``mul_().add_()`` loop over ~490 tiny tensors
(src/engine/trainer.py:163-169,258-272; optimisers from src/utils/solvers.py:35-52) -- ~1500 launches.  Here the caller's
``torch.optim.SGD`` / ``torch.optim.Adam`` objects stay the owners of hyper-parameters and state (``momentum_buffer``,
``exp_avg``, ``exp_avg_sq``, ``step`` live in ``optim.state`` exactly as torch would create them, so ``state_dict()`` and a
later plain ``optim.step()`` keep working); only the arithmetic moves into ``nasb_mt_grad_sumsq`` + ``nasb_mt_optim_step``.

Optimisers the kernels do not cover (anything but plain SGD / Adam: amsgrad, maximize, decoupled decay, tensor lr, other
classes) make ``FusedStep.supported`` return False and the engine keeps the caller's own ``step()`` for them.
"""
import ctypes as C

import numpy as np
import torch

from . import lib

_TENSOR = np.dtype([("param", np.uint64), ("grad", np.uint64), ("state1", np.uint64), ("state2", np.uint64),
                    ("avg", np.uint64), ("step", np.uint64), ("numel", np.int64), ("group", np.int32), ("clip", np.int32)])
_GROUP = np.dtype([("kind", np.int32), ("first", np.int32), ("nesterov", np.int32), ("lr", np.float32),
                   ("beta1", np.float32), ("beta2", np.float32), ("eps", np.float32), ("weight_decay", np.float32)])
NONE, SGD, ADAM = 0, 1, 2


def _plain(optim):
    """Is this optimiser one whose update the kernel reproduces?"""
    if type(optim) is torch.optim.SGD:
        # dampening != 0 needs torch's per-tensor "first step" rule (buf = grad); with dampening 0 a zero-initialised
        # buffer gives exactly that, which is what the kernel relies on
        return all(not g.get("maximize", False) and not g.get("differentiable", False) and g.get("dampening", 0) == 0
                   and isinstance(g["lr"], (int, float)) for g in optim.param_groups)
    if type(optim) is torch.optim.Adam:
        return all(not g.get("amsgrad", False) and not g.get("maximize", False) and not g.get("differentiable", False)
                   and not g.get("decoupled_weight_decay", False) and isinstance(g["lr"], (int, float))
                   and all(isinstance(b, (int, float)) for b in g["betas"]) for g in optim.param_groups)
    return False


class FusedStep:
    """``FusedStep([(optim, clip_params, max_norm), ...], polyak_params, avg_param)``; ``step(polyak_decay)`` performs, for
    every entry, ``clip_grad_norm_(clip_params, max_norm)`` (skipped when max_norm <= 0) and ``optim.step()``, then the
    Polyak update of ``avg_param`` against ``polyak_params`` -- with torch's semantics, in two launches per 320 tensors."""

    def __init__(self, entries, polyak_params=None, avg_param=None):
        self.entries = [(o, list(cp) if cp is not None else [], float(mn)) for o, cp, mn in entries]
        self.polyak = list(polyak_params) if (polyak_params is not None and avg_param is not None) else []
        self.avg = list(avg_param) if avg_param is not None else []
        # strong references: id()s of optimisers / parameters in cache keys must not be recycled while this object lives
        self._built = False
        self.cells = None

    @staticmethod
    def supported(entries):
        return all(_plain(o) for o, _, _ in entries) and len(entries) <= 8

    def hyper_key(self):
        """Everything a captured graph bakes in: hyper-parameters of every group and the clip norms."""
        key = []
        for o, _, mn in self.entries:
            key.append(mn)
            for g in o.param_groups:
                key.append((g["lr"], g.get("momentum"), g.get("dampening"), g.get("nesterov"), g.get("betas"), g.get("eps"),
                            g["weight_decay"], len(g["params"])))
        return tuple(key)

    # ---- table construction
    def _build(self):
        rows, self.params, self.groups, self.group_of = [], [], [], []
        index = {}
        dev = None

        def row(p):
            i = index.get(id(p))
            if i is None:
                if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                    raise RuntimeError("fused optimiser step needs contiguous fp32 CUDA parameters")
                i = index[id(p)] = len(rows)
                rows.append([p.data_ptr(), 0, 0, 0, 0, 0, p.numel(), -1, -1])
                self.params.append(p)
            return i

        for ci, (o, clip_params, mn) in enumerate(self.entries):
            for g in o.param_groups:
                gi = len(self.groups)
                self.groups.append((o, g))
                for p in g["params"]:
                    if not p.requires_grad:
                        continue
                    dev = p.device
                    r = rows[row(p)]
                    r[7] = gi
            for p in clip_params:
                if p.requires_grad:
                    rows[row(p)][8] = ci
        if len(self.groups) > 8:
            raise RuntimeError("fused optimiser step supports at most 8 parameter groups")
        for p, a in zip(self.polyak, self.avg):
            r = rows[row(p)]
            if a.dtype != torch.float32 or not a.is_contiguous() or a.numel() != p.numel():
                raise RuntimeError("Polyak averages must be contiguous fp32 tensors shaped like their parameters")
            r[4] = a.data_ptr()
            dev = p.device
        self.device = dev
        self.table = np.zeros(len(rows), dtype=_TENSOR)
        for i, r in enumerate(rows):
            self.table[i] = tuple(r)
        self.grad_rows = [i for i, r in enumerate(rows) if r[7] >= 0 or r[8] >= 0]
        self.cells = torch.zeros(8, dtype=torch.float64, device=dev)
        self.max_norm = np.zeros(8, dtype=np.float32)
        for ci, (_, _, mn) in enumerate(self.entries):
            self.max_norm[ci] = mn
        self.garr = np.zeros(max(len(self.groups), 1), dtype=_GROUP)
        self._built, self._state_done, self._ids = True, False, self._state_ids()

    def _ensure_state(self):
        """Create optimiser state the way torch's first step() would, and point the table at it.  Returns the per-group
        `first` flags (SGD: this step creates the momentum buffers, so buf = grad)."""
        first = [0] * len(self.groups)
        if self._state_done:
            return first
        done = True
        tb = self.table
        for i, p in enumerate(self.params):
            gi = int(tb["group"][i])
            if gi < 0:
                continue
            o, g = self.groups[gi]
            st = o.state[p]
            if type(o) is torch.optim.SGD:
                if g["momentum"] != 0:
                    buf = st.get("momentum_buffer")
                    if buf is None:
                        if p.grad is None:
                            done = False
                            continue
                        buf = st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    tb["state1"][i] = buf.data_ptr()
            else:
                if len(st) == 0:
                    if p.grad is None:
                        done = False
                        continue
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                elif not st["step"].is_cuda:
                    st["step"] = st["step"].to(device=p.device, dtype=torch.float32)
                tb["state1"][i] = st["exp_avg"].data_ptr()
                tb["state2"][i] = st["exp_avg_sq"].data_ptr()
                tb["step"][i] = st["step"].data_ptr()
        # momentum buffers start at zero: buf = mom*0 + g = g is torch's first-step rule for dampening 0 (see _plain)
        self._state_done = done
        return first

    def _state_ids(self):
        return tuple((id(o.state), id(o.param_groups)) for o, _, _ in self.entries)

    def step(self, polyak_decay=0.0):
        if not self._built or self._ids != self._state_ids():  # load_state_dict() replaces state and groups
            self._build()
        first = self._ensure_state()
        tb = self.table
        # gradients: contiguous fp32; a parameter without .grad is skipped by torch's clip and step alike
        gcol = tb["grad"]
        for i in self.grad_rows:
            p = self.params[i]
            g = p.grad
            if g is None:
                gcol[i] = 0
                continue
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = p.grad = g.to(torch.float32).contiguous()
            gcol[i] = g.data_ptr()
        for gi, (o, g) in enumerate(self.groups):
            if type(o) is torch.optim.SGD:
                self.garr[gi] = (SGD, first[gi], int(bool(g["nesterov"])), g["lr"], g["momentum"], g["dampening"], 0.0,
                                 g["weight_decay"])
            else:
                self.garr[gi] = (ADAM, 0, 0, g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"])
        n = len(tb)
        with torch.cuda.device(self.device):
            lib.call("nasb_mt_grad_sumsq", C.c_void_p(tb.ctypes.data), n, lib.ptr(self.cells), 8)
            lib.call("nasb_mt_optim_step", C.c_void_p(tb.ctypes.data), n, C.c_void_p(self.garr.ctypes.data), len(self.groups),
                     C.c_void_p(self.max_norm.ctypes.data), 8, lib.ptr(self.cells), float(polyak_decay))
        for o, _, _ in self.entries:  # what torch.optim.Optimizer.step's wrapper records for LR schedulers
            o._opt_called = True

    def grad_norms(self):
        """Total gradient norms of the last step(), one per entry (what clip_grad_norm_ returns)."""
        return self.cells[:len(self.entries)].sqrt().float()
