"""Diagnostic (GPU box): per-parameter gradient error of a bf16 InvertedResidual block against torch fp32."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch  # noqa: E402

import nas_segm_b200  # noqa: E402
from nas_segm_b200 import lib  # noqa: E402
from nas_segm_b200.nn.layer_factory import InvertedResidual  # noqa: E402

torch.manual_seed(11)
for (inp, oup, stride, n, h, w) in [(32, 32, 1, 2, 40, 56), (32, 64, 2, 2, 36, 52)]:
    m = InvertedResidual(inp, oup, stride, 6).cuda().train()
    ref = copy.deepcopy(m.conv).float()
    x = torch.randn(n, inp, h, w, device="cuda").to(torch.bfloat16)
    nas_segm_b200.set_act_dtype(torch.bfloat16)
    xi = lib.to_nhwc(x.clone()).requires_grad_(True)
    y = m(xi)
    gy = torch.randn(y.shape, device="cuda").to(torch.bfloat16)
    (y.float() * gy.float()).sum().backward()
    nas_segm_b200.set_act_dtype(torch.float32)
    xr = x.float().requires_grad_(True)
    yr = ref(xr) + (xr if m.use_res_connect else 0)
    (yr * gy.float()).sum().backward()
    print("block", inp, oup, stride, "y", float((y.float() - yr).abs().mean() / yr.abs().mean()),
          "dx", float((xi.grad.float() - xr.grad).abs().mean() / xr.grad.abs().mean()))
    for (k, a), b in zip(m.conv.named_parameters(), ref.parameters()):
        print("   %-10s %-18s err %.4f   |ours| %.4e |ref| %.4e" % (k, tuple(a.shape), float((a.grad - b.grad).abs().mean() / b.grad.abs().mean().clamp_min(1e-9)),
                                                                    float(a.grad.abs().mean()), float(b.grad.abs().mean())))
