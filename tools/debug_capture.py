"""Debug aid: which part of an fp32 MicroDecoder training forward invalidates a CUDA-graph capture?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch
from torch import nn
import nas_segm_b200
from nas_segm_b200 import lib
from nas_segm_b200.nn.encoders import mbv2
from nas_segm_b200.nn.micro_decoders import MicroDecoder
from golden_util import C0

nas_segm_b200.set_act_dtype(torch.float32)
torch.manual_seed(0)
enc = mbv2().cuda().train()
dec = MicroDecoder(list(enc.out_sizes), 21, C0, agg_size=16, aux_cell=True, repeats=1).cuda().train()
x = torch.randn(4, 3, 192, 192, device="cuda")
orig_call = lib.call
names = []
def traced(name, *a):
    names.append(name)
    return orig_call(name, *a)
import nas_segm_b200.functional as Fn
for mode in ("global", "thread_local", "relaxed"):
    for what in ("enc", "enc+dec"):
        for _ in range(2):
            f = enc(x)
            if what != "enc":
                dec(f)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        names.clear()
        lib.call = traced; Fn.call = traced
        try:
            with torch.cuda.graph(g, capture_error_mode=mode):
                f = enc(x)
                if what != "enc":
                    o = dec(f)
            print(mode, what, "capture OK,", len(names), "calls")
        except Exception as e:  # noqa: BLE001
            print(mode, what, "capture FAILED after", len(names), "calls; last:", names[-4:], "|", str(e).splitlines()[0][:150])
        finally:
            lib.call = orig_call; Fn.call = orig_call
        torch.cuda.synchronize()
