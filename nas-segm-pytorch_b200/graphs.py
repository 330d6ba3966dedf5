"""CUDA-graph capture of a fixed-shape forward pass.

At inference sizes the network is launch-bound on the host (~300 C-ABI calls per image, a few microseconds each), so a
replayed graph is what a latency-sensitive user runs.  The kernels take explicit stream arguments and allocate nothing,
tensor maps are passed by value, so the whole forward is capturable; torch's allocator serves the activations from the
graph's private pool."""
import contextlib
import gc

import torch

_pools = {}  # device index -> (pool handle, keep-alive graph)


def shared_pool(device=None):
    """ONE allocator pool for every capture of the process (per device).  torch gives each captured graph a private pool
    that starts empty, so every activation of a capture is a fresh cudaMalloc (measured in the search loop, where each
    candidate captures four graphs: 12.8 k torch.empty calls cost 1.07 s of a 10 s candidate), and returns the memory to
    the driver when the graph dies.  A shared pool kept alive by a one-node graph hands the blocks of the previous
    candidate's graphs to the next capture.  Sharing is safe here because graphs are replayed one at a time on one stream
    and nothing but their static inputs / outputs (alive, hence never reused) is read across a replay boundary."""
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    if dev not in _pools:
        handle = torch.cuda.graph_pool_handle()
        g, s = torch.cuda.CUDAGraph(), torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            g.capture_begin(pool=handle)
            try:
                keep = torch.zeros(8, device="cuda:%d" % dev)
            finally:
                g.capture_end()
        torch.cuda.current_stream(dev).wait_stream(s)
        _pools[dev] = (handle, g, keep)
    return _pools[dev][0]


@contextlib.contextmanager
def capture(graph, stream=None):
    """`with capture(g):` -- like `torch.cuda.graph(g)`, but into the shared pool and without the gc.collect() /
    empty_cache() of torch's context manager (which hands every cached block back to the driver, so the eager iterations
    of the next candidate start with cudaMalloc again).  The cyclic collector is suspended for the duration instead."""
    torch.cuda.synchronize()
    pool = shared_pool()
    stream = torch.cuda.Stream() if stream is None else stream
    stream.wait_stream(torch.cuda.current_stream())
    # The cyclic collector must not run while the stream is capturing: an unreachable graph of an earlier capture (modules and
    # graph closures reference each other) would be destroyed from inside the capture, and cudaGraphExecDestroy / the pool
    # release it triggers invalidate a global-mode capture ("operation failed due to a previous error during capture"; seen
    # once the test suite's allocation pattern moved a generation-2 collection into a re-capture).  torch.cuda.graph avoids
    # this with a full gc.collect() before every capture; suspending the collector costs nothing.
    gc_was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.stream(stream):
            graph.capture_begin(pool=pool)
            try:
                yield
            finally:
                graph.capture_end()
    finally:
        if gc_was_enabled:
            gc.enable()
    torch.cuda.current_stream().wait_stream(stream)


class GraphedForward:
    """``g = GraphedForward(model, example_input); out = g(x)`` -- x must have example_input's shape / dtype."""

    def __init__(self, model, example, warmup=1):
        assert example.is_cuda
        self.model = model
        self.static_in = example.clone()
        was_training = model.training
        model.eval()
        from . import packs
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(warmup):  # first calls configure kernel attributes, the per-device workspace and the unit plan
                with packs.scope(model):
                    model(self.static_in)
        torch.cuda.current_stream().wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        # The captured forward starts with ONE launch that folds BatchNorm and packs the tensor-core operand of every unit
        # (packs.scope -> nasb_conv_units_prepare), so a replay always sees the current weights without the two small
        # per-unit launches (200 of ~360 for arch0).
        with torch.no_grad(), capture(self.graph, s), packs.scope(model):
            self.static_out = model(self.static_in)
        model.train(was_training)

    def __call__(self, x):
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out


def make_capturable(optim):
    """Put a torch optimiser into graph-capturable form (Adam/AdamW keep their step counters on the device)."""
    if optim is None:
        return
    for g in optim.param_groups:
        if "capturable" in g:
            g["capturable"] = True
    for st in optim.state.values():
        if "step" in st and torch.is_tensor(st["step"]) and not st["step"].is_cuda:
            p = next(iter(optim.param_groups[0]["params"]))
            st["step"] = st["step"].to(p.device)


class StepGraph:
    """A whole training iteration (forward, loss, backward, gradient clipping, optimiser steps, Polyak) as one CUDA graph.

    ``fn(*static_inputs) -> loss`` must read only from the static input tensors.  The first ``warmup`` calls run eagerly
    (they are real iterations on the caller's data; they also initialise optimiser state and per-kernel attributes), the
    next call captures the iteration and replays it, later calls copy the inputs and replay.  At the NAS-loop sizes
    (64x64 cached features, ~1300 launches per iteration) the host cannot issue launches as fast as the GPU retires
    them; a replayed graph removes that bound."""

    def __init__(self, fn, static_inputs, warmup=None):
        from . import config
        warmup = config().graph_warmup if warmup is None else warmup
        self.fn, self.static_inputs, self.warmup = fn, list(static_inputs), warmup
        self.calls, self.graph, self.static_loss = 0, None, None
        self.launches_per_replay = 0
        self.side = torch.cuda.Stream()

    def matches(self, inputs):
        return len(inputs) == len(self.static_inputs) and all(
            tuple(a.shape) == tuple(b.shape) and a.dtype == b.dtype for a, b in zip(inputs, self.static_inputs))

    def __call__(self, *inputs):
        from . import lib
        for dst, src in zip(self.static_inputs, inputs):
            if dst is not src:
                dst.copy_(src, non_blocking=True)
        self.calls += 1
        if self.graph is not None:
            self.graph.replay()
            lib.launches += self.launches_per_replay
            return self.static_loss
        if self.calls <= self.warmup:
            self.side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.side):
                loss = self.fn(*self.static_inputs)
            torch.cuda.current_stream().wait_stream(self.side)
            return loss.detach()
        torch.cuda.synchronize()
        l0 = lib.launches
        self.graph = torch.cuda.CUDAGraph()
        with capture(self.graph, self.side):
            self.static_loss = self.fn(*self.static_inputs).detach()
        self.launches_per_replay = lib.launches - l0
        lib.launches = l0  # capture launched nothing
        self.graph.replay()
        lib.launches += self.launches_per_replay
        return self.static_loss
