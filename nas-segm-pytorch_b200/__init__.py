"""nas-segm-pytorch_b200 -- B200-native (sm_100a) implementation of the NAS inner-loop hot path of
DrSleep/nas-segm-pytorch behind the reference's own registry / module / engine API.

Host code is Python/PyTorch (device memory, streams, autograd glue); every arithmetic step on the path runs in the
hand-written CUDA kernels of ``libnasb200.so`` (C ABI: ``include/nasb200.h``).  There is no CPU or eager fallback:
importing the compute modules without the built library, or running them without a CUDA device, raises.

Sub-packages mirror the reference's ``src/`` layout: ``nn`` (op/cell registry, decoders, encoder), ``rl.genotypes``
(name tables), ``engine`` (populate_task0 / train_task0 / train_segmenter / validate), ``helpers.miou_utils``
(fast_cm / compute_iu / compute_ius_accs).  ``dropin()`` registers them under the reference's top-level module names.
"""
import os as _os
import sys as _sys

__version__ = "0.1.0"


class _Config:
    """Process-wide numeric mode.  ``act_dtype`` is the storage type of activations inside the network:
    torch.float32 = parity mode (matches the reference within 1e-3 on logits), torch.bfloat16 = speed mode."""

    def __init__(self):
        import torch
        self.act_dtype = torch.float32
        self.use_tcgen05 = True  # bf16 pointwise convolutions on the tensor cores when shapes allow
        self.graph_warmup = 1  # eager iterations before an engine loop captures its CUDA graph
        self.cuda_graphs = False  # engine loops replay one captured CUDA graph per iteration (graphs.StepGraph)
        # training-BN statistics inside the depthwise tile kernel (per-thread register partials over the persistent loop,
        # one shared-memory merge per CTA): +15-25 % on the kernel, cheaper than the separate pass that re-reads z
        self.fuse_dw_stats = True
        self.use_tma_tiles = True  # bf16 depthwise convolutions on TMA-staged shared-memory tiles when shapes allow
        # clip + optimiser step + Polyak of an engine iteration as two multi-tensor launches (optim.FusedStep) when the
        # caller's optimisers are plain torch.optim.SGD / Adam; off = the caller's own optim.step() (torch kernels)
        self.fused_optim = True
        # training-mode BN backward folded into the neighbouring kernels where the graph guarantees a sole consumer
        # (functional._BnRec): gated data-gradient epilogues, dz pass from their reductions, the pointwise unit's backward
        # without dz.  Validated (tests/test_gpu_tcgen05.py::test_bn_backward_fusion_matches_unfused_path) and MEASURED on a
        # B200: the third operand costs the issue-bound producers what the separate reduction pass costs (step 24.5 vs
        # 24.4 ms, profiles/r2_bn_fusion_kbench.txt), so it is off by default.
        self.fuse_bn_bwd = False
        # inference: depthwise -> pointwise pairs as one kernel (csrc/sep_tcgen05.cu) when the shapes allow
        self.fused_sep = _os.environ.get("NASB_FUSED_SEP", "1") != "0"
        # weight-gradient kernels of an engine iteration on a second stream, joined before the optimiser step (lib._WgradStream)
        self.async_wgrad = _os.environ.get("NASB_ASYNC_WGRAD", "1") != "0"
        # independent decoder branches (the two inputs of every aggregation) on concurrent streams (lib._BranchStreams)
        self.branch_streams = _os.environ.get("NASB_BRANCH_STREAMS", "1") != "0"


_config = None


def config():
    global _config
    if _config is None:
        _config = _Config()
    return _config


def set_act_dtype(dtype):
    config().act_dtype = dtype


def dropin():
    """Register this package's sub-packages under the reference's import names (``nn``, ``rl``, ``engine``,
    ``helpers``) so reference-side code such as ``from nn.micro_decoders import MicroDecoder`` or
    ``from helpers.miou_utils import fast_cm`` binds to the B200 implementation unchanged."""
    import importlib
    me = __name__
    for sub in ("rl", "rl.genotypes", "nn", "nn.layer_factory", "nn.micro_decoders", "nn.encoders", "helpers",
                "helpers.utils", "helpers.miou_utils", "engine", "engine.trainer", "engine.inference"):
        _sys.modules[sub] = importlib.import_module(me + "." + sub)
