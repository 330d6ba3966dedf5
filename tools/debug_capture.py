"""Debug aid: which part of an fp32 MicroDecoder train_segmenter iteration invalidates a CUDA-graph capture?"""
import sys, os, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch
from torch import nn
import nas_segm_b200
from nas_segm_b200 import lib
import nas_segm_b200.functional as Fn
from nas_segm_b200.engine import trainer
from nas_segm_b200.graphs import StepGraph, make_capturable
from nas_segm_b200.nn.encoders import mbv2
from nas_segm_b200.nn.micro_decoders import MicroDecoder
from golden_util import C0


class Seg(nn.Module):
    def __init__(self, enc, dec):
        super().__init__()
        self.encoder, self.decoder = enc, dec

    def forward(self, x):
        return self.decoder(self.encoder(x))


class Wrap(nn.Module):
    def __init__(self, m):
        super().__init__()
        self.module = m

    def forward(self, x):
        return self.module(x)


orig_call, orig_try = lib.call, lib.try_call
names = []
def traced(name, *a):
    names.append(name)
    return orig_call(name, *a)
def traced_try(name, *a):
    names.append(name + "?")
    return orig_try(name, *a)

nas_segm_b200.set_act_dtype(torch.float32)
cfg = nas_segm_b200.config()
for variant in ("full", "no_polyak", "no_aux", "torch_optim", "dec_only_optim"):
    torch.manual_seed(0)
    enc = mbv2()
    dec = MicroDecoder(list(enc.out_sizes), 21, C0, agg_size=16, aux_cell=True, repeats=1)
    seg = Wrap(Seg(enc, dec)).cuda().train()
    oe = torch.optim.SGD(enc.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-5)
    od = torch.optim.Adam(dec.parameters(), lr=3e-3, weight_decay=1e-5)
    make_capturable(oe), make_capturable(od)
    avg = [p.data.clone() for p in seg.parameters()]
    x = torch.randn(4, 3, 192, 192, device="cuda")
    y = torch.randint(0, 21, (4, 192, 192), device="cuda").to(torch.uint8)
    crit = nn.NLLLoss(ignore_index=255)
    cfg.fused_optim = variant != "torch_optim"
    polyak = variant != "no_polyak"
    auxw = -1 if variant == "no_aux" else 0.15

    def step(im, tg):
        return trainer.segmenter_step(seg, im, tg, oe, od, crit, 0.0 if variant == "dec_only_optim" else 3.0, 3.0, polyak, auxw,
                                      avg if polyak else None, 0.99)
    sg = StepGraph(step, [x.clone(), y.clone()], warmup=1)
    sg(x, y)
    torch.cuda.synchronize()
    names.clear()
    lib.call = traced; Fn.call = traced; lib.try_call = traced_try; Fn.try_call = traced_try
    try:
        sg(x, y)
        torch.cuda.synchronize()
        print(variant, "capture OK,", len(names), "calls; loss", float(sg(x, y)))
    except Exception as e:  # noqa: BLE001
        print(variant, "capture FAILED after", len(names), "calls; last:", names[-6:], "|", str(e).splitlines()[0][:160])
    finally:
        lib.call = orig_call; Fn.call = orig_call; lib.try_call = orig_try; Fn.try_call = orig_try
    torch.cuda.synchronize()

# ---- the engine function itself, as the failing test drives it
class Loader(list):
    class _DS:
        def set_stage(self, s):
            pass
    dataset = _DS()

for variant in ("engine_pageable_f64", "engine_pinned_f32", "engine_pageable_f32"):
    torch.manual_seed(0)
    enc = mbv2()
    dec = MicroDecoder(list(enc.out_sizes), 21, C0, agg_size=16, aux_cell=True, repeats=1)
    seg = Wrap(Seg(enc, dec)).cuda()
    oe = torch.optim.SGD(enc.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-5)
    od = torch.optim.Adam(dec.parameters(), lr=3e-3, weight_decay=1e-5)
    avg = [p.data.clone() for p in seg.parameters()]
    loader = Loader()
    for i in range(2):
        img = torch.randn(4, 3, 192, 192)
        msk = torch.randint(0, 21, (4, 192, 192)).to(torch.uint8)
        if variant == "engine_pageable_f64":
            img = img.double()
        if variant == "engine_pinned_f32":
            img, msk = img.pin_memory(), msk.pin_memory()
        loader.append({"image": img, "mask": msk})
    cfg.cuda_graphs, cfg.graph_warmup, cfg.fused_optim = True, 1, True
    names.clear()
    lib.call = traced; Fn.call = traced; lib.try_call = traced_try; Fn.try_call = traced_try
    try:
        for ep in range(2):
            trainer.train_segmenter.__wrapped__(seg, loader, oe, od, ep, nn.NLLLoss(ignore_index=255), False, 3.0, 3.0, True,
                                                print_every=1, aux_weight=0.15, avg_param=avg, polyak_decay=0.99)
        torch.cuda.synchronize()
        print(variant, "OK,", len(names), "calls")
    except Exception as e:  # noqa: BLE001
        print(variant, "FAILED after", len(names), "calls; last:", names[-6:], "|", str(e).splitlines()[0][:160])
    finally:
        lib.call = orig_call; Fn.call = orig_call; lib.try_call = orig_try; Fn.try_call = orig_try
        cfg.cuda_graphs, cfg.graph_warmup = False, 3
    torch.cuda.synchronize()
