"""Persistent bf16 tensor-core operands of a model's convolution weights.

The tensor-core kernels consume the fp32 master weights as bf16 packs (pointwise: [R][Kp], forward and transposed; dense
3x3: [9][Nr][Kp], forward and flipped-transposed).  Packing per call cost 135 launches per arch0 iteration (1.6 ms of
25.7).  Inside a *scope* -- one engine iteration, or one validate() loop -- the weights cannot change between the scope's
begin() and the optimiser step at its end, so begin() re-packs EVERY operand of the model with one multi-tensor launch
(``nasb_mt_pack_bf16``) into buffers that live as long as the model, and the conv units pick them up by weight address.
Outside of a scope (a user calling modules directly) the conv units pack per call as before, which is always correct.

A captured CUDA graph of an engine iteration starts with that one pack launch, so replays see the weights the optimiser
kernel of the previous replay wrote.
"""
import contextlib
import ctypes as C

import numpy as np
import torch
from torch import nn

from . import lib

PW, PW_T, C3, C3_T = 0, 1, 2, 3
_JOB = np.dtype([("src", np.uint64), ("dst", np.uint64), ("c_out", np.int32), ("c_in", np.int32), ("kind", np.int32),
                 ("reserved", np.int32)])

_active = None  # {(weight.data_ptr(), kind): packed tensor} of the open scope

# ---- inference units (functional._conv_unit_infer): BN fold + operand pack of EVERY unit of the model in one launch
_scope_model = None   # the model of the open scope
_prepared = None      # {scratch address} of the units the open scope has prepared (None: no scope)
_pending = None       # units met inside the scope that were not in the model's plan yet: [(unit_state, c_in)]


class _ModelPacks:
    def __init__(self, model):
        lb = lib.load()
        self.weights, jobs, self.table = [], [], {}
        dev = None
        for m in model.modules():
            if not isinstance(m, nn.Conv2d) or m.groups != 1 or not m.weight.is_cuda or m.weight.dtype != torch.float32:
                continue
            w = m.weight
            co, ci, kh, kw = w.shape
            kinds = ()
            if (kh, kw) == (1, 1):
                kinds = ((PW, co, ci), (PW_T, co, ci))
            elif (kh, kw) == (3, 3) and ci == 3:  # encoder stem as a K = 32 GEMM on the im2col patch matrix
                kinds = ((PW, co, 27),)
            elif (kh, kw) == (3, 3) and lb.nasb_conv3_tc_supported(ci, co):
                kinds = ((C3, co, ci), (C3_T, co, ci))
            for kind, r, c in kinds:
                n = int(lb.nasb_pack_elems(kind, r, c))
                buf = torch.empty(n, dtype=torch.bfloat16, device=w.device)
                self.table[(w.data_ptr(), kind)] = buf
                jobs.append((w.data_ptr(), buf.data_ptr(), r, c, kind, 0))
                self.weights.append(w)
                dev = w.device
        self.jobs = np.array(jobs, dtype=_JOB) if jobs else np.zeros(0, dtype=_JOB)
        self.device = dev

    def valid(self):
        src = self.jobs["src"]
        return all(int(src[i]) == w.data_ptr() for i, w in enumerate(self.weights))

    def refresh(self):
        if len(self.jobs):
            with torch.cuda.device(self.device):
                lib.call("nasb_mt_pack_bf16", C.c_void_p(self.jobs.ctypes.data), len(self.jobs))


def _prepare_infer(model):
    """One launch that folds BatchNorm and packs the operand of every inference unit the model has run before."""
    global _scope_model, _prepared, _pending
    from . import config
    _scope_model, _prepared, _pending = model, set(), []
    plan = model.__dict__.get("_nasb_infer_plan")
    if not plan or torch.is_grad_enabled():
        return
    units = [(st, cin) for st, cin in plan if st[0].weight == st[4].data_ptr() and st[1].is_cuda]
    if len(units) != len(plan):  # parameters moved (model.to(...)): the stale states are rebuilt at their next call
        object.__setattr__(model, "_nasb_infer_plan", units)
    if not units:
        return
    n = len(units)
    arr_u = (C.c_void_p * n)(*[C.addressof(st[0]) for st, _ in units])
    arr_c = (C.c_int * n)(*[cin for _, cin in units])
    arr_s = (C.c_void_p * n)(*[st[1].data_ptr() for st, _ in units])
    cfg = config()
    with torch.cuda.device(units[0][0][1].device):
        lib.call("nasb_conv_units_prepare", arr_u, arr_c, arr_s, n, (1 if cfg.use_tcgen05 else 0) | (2 if cfg.use_tma_tiles else 0))
    _prepared = {st[1].data_ptr() for st, _ in units}


def infer_prepared(state, c_in):
    """Is this inference unit's scratch already filled by the open scope?  Unknown units are noted for the next scope."""
    if _prepared is None:
        return False
    if state[1].data_ptr() in _prepared:
        return True
    _pending.append((state, c_in))
    return False


def begin(model):
    """Open a scope: re-pack every tensor-core operand of `model` (one launch) and serve them to the conv units."""
    global _active
    from . import config
    _prepare_infer(model)
    if config().act_dtype != torch.bfloat16 or not config().use_tcgen05:
        _active = None
        return
    mp = getattr(model, "_nasb_packs", None)
    if mp is None or not mp.valid():
        if torch.cuda.is_current_stream_capturing():  # never allocate the long-lived buffers from a graph's private pool
            _active = None
            return
        mp = _ModelPacks(model)
        object.__setattr__(model, "_nasb_packs", mp)
    mp.refresh()
    _active = mp.table


def end():
    global _active, _scope_model, _prepared, _pending
    _active = None
    if _scope_model is not None and _pending:
        plan = _scope_model.__dict__.get("_nasb_infer_plan") or []
        known = {st[1].data_ptr() for st, _ in plan}
        plan = plan + [(st, cin) for st, cin in _pending if st[1].data_ptr() not in known and not known.add(st[1].data_ptr())]
        object.__setattr__(_scope_model, "_nasb_infer_plan", plan)
    _scope_model = _prepared = _pending = None


@contextlib.contextmanager
def scope(model):
    """``with packs.scope(model):`` -- begin() / end() around a region in which the model's weights do not change."""
    begin(model)
    try:
        yield
    finally:
        end()


def get(weight, kind):
    """The packed operand of `weight` if a scope is open and covers it, else None (the caller packs per call)."""
    if _active is None:
        return None
    return _active.get((weight.data_ptr(), kind))
