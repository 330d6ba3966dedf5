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


def begin(model):
    """Open a scope: re-pack every tensor-core operand of `model` (one launch) and serve them to the conv units."""
    global _active
    from . import config
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
    global _active
    _active = None


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
