"""CUDA-graph capture of a fixed-shape forward pass.

At inference sizes the network is launch-bound on the host (~300 C-ABI calls per image, a few microseconds each), so a
replayed graph is what a latency-sensitive user runs.  The kernels take explicit stream arguments and allocate nothing,
tensor maps are passed by value, so the whole forward is capturable; torch's allocator serves the activations from the
graph's private pool."""
import torch


class GraphedForward:
    """``g = GraphedForward(model, example_input); out = g(x)`` -- x must have example_input's shape / dtype."""

    def __init__(self, model, example, warmup=3):
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
        with torch.no_grad(), torch.cuda.graph(self.graph), packs.scope(model):
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
        with torch.cuda.graph(self.graph):
            self.static_loss = self.fn(*self.static_inputs).detach()
        self.launches_per_replay = lib.launches - l0
        lib.launches = l0  # capture launched nothing
        self.graph.replay()
        lib.launches += self.launches_per_replay
        return self.static_loss
