"""tcgen05 implicit-GEMM 3x3 convolution (forward / data gradient / weight gradient) vs fp32 torch on the same bf16
operands: channel counts of the classifier heads and the CVPR cell ops, dilations 1/3/12, odd and tiny images, fp32 and bf16
outputs, many patches per CTA (pipeline phase logic), odd channel counts with a padded pixel pitch."""
import pytest
import torch
import torch.nn.functional as F

from nas_segm_b200 import functional as Fn
from nas_segm_b200 import lib

pytestmark = pytest.mark.gpu


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


CASES = [  # Cin, Cout, dil, N, H, W
    (64, 19, 1, 2, 37, 53), (48, 48, 1, 2, 81, 81), (48, 48, 3, 2, 21, 21), (48, 48, 12, 2, 11, 11), (64, 64, 1, 1, 64, 128),
    (24, 40, 1, 3, 5, 200), (128, 64, 1, 1, 33, 47), (64, 21, 1, 4, 128, 256), (16, 8, 3, 1, 9, 9), (64, 1, 1, 2, 30, 40),
    (48, 48, 3, 1, 20, 130), (128, 64, 1, 1, 12, 100), (64, 19, 12, 1, 30, 300),  # wide rows: the row-halo staging mode
]


def test_conv3_tc_forward_dgrad_wgrad():
    g = torch.Generator(device="cuda").manual_seed(0)
    bad = []
    for (ci, co, d, n, H, W) in CASES:
        assert lib.load().nasb_conv3_tc_supported(ci, co) == 1
        x = _nhwc(torch.randn(n, ci, H, W, generator=g, device="cuda").to(torch.bfloat16))
        w = torch.randn(co, ci, 3, 3, generator=g, device="cuda") / (3 * ci ** 0.5)
        bias = torch.randn(co, generator=g, device="cuda")
        wq = w.to(torch.bfloat16).float()
        ref = F.conv2d(x.float(), wq, bias, 1, d, d)
        for odt in (torch.float32, torch.bfloat16):
            if odt == torch.bfloat16 and co % 8:
                continue
            out = lib.new_act(n, co, H, W, odt, "cuda")
            lib.call("nasb_conv3_tc_fwd", lib.ref(lib.desc(x)), lib.ptr(Fn._pack_conv3(w, 0)), co, d, d, None, lib.ptr(bias),
                     lib.ACT_NONE, lib.ref(lib.desc(out)), None)
            torch.cuda.synchronize()
            e = float((out.float() - ref).abs().max() / ref.abs().max())
            if not e < (1e-2 if odt == torch.bfloat16 else 2e-3):
                bad.append(("fwd", ci, co, d, H, W, str(odt), e))
        # gradients
        dzf = torch.randn(n, co, H, W, generator=g, device="cuda")
        dz = Fn._bf16_padded_copy(_nhwc(dzf))
        xr = x.float().requires_grad_(True)
        wr = wq.clone().requires_grad_(True)
        F.conv2d(xr, wr, None, 1, d, d).backward(dz.float())
        dx = lib.new_act(n, ci, H, W, torch.bfloat16, "cuda")
        lib.call("nasb_conv3_tc_fwd", lib.ref(lib.desc(dz)), lib.ptr(Fn._pack_conv3(w, 1)), ci, d, 2 * d - d, None, None,
                 lib.ACT_NONE, lib.ref(lib.desc(dx)), None)
        dw = torch.zeros(co, ci, 3, 3, device="cuda")
        lib.call("nasb_conv3_tc_wgrad", lib.ref(lib.desc(x)), lib.ref(lib.desc(dz)), d, d, lib.ptr(dw))
        torch.cuda.synchronize()
        e = float((dx.float() - xr.grad).abs().max() / xr.grad.abs().max())
        if not e < 1.5e-2:
            bad.append(("dgrad", ci, co, d, H, W, e))
        e = float((dw - wr.grad).abs().max() / wr.grad.abs().max())
        if not e < 3e-3:
            bad.append(("wgrad", ci, co, d, H, W, e))
    assert not bad, bad


def test_conv3_unit_through_autograd():
    """The fused conv unit picks the tensor-core 3x3 path for bf16 (classifier head: fp32 logits, 19 classes)."""
    from nas_segm_b200.nn.layer_factory import conv3x3
    from nas_segm_b200.nn.micro_decoders import clf3x3
    torch.manual_seed(0)
    conv = conv3x3(64, 19, stride=1, bias=True).cuda()
    x = _nhwc(torch.randn(2, 64, 40, 56, device="cuda").to(torch.bfloat16)).requires_grad_(True)
    l0 = lib.launches
    y = clf3x3(conv, x)
    assert y.dtype == torch.float32
    ct = torch.randn_like(y)
    (y * ct).sum().backward()
    xr = x.detach().float().requires_grad_(True)
    wr = conv.weight.detach().to(torch.bfloat16).float().requires_grad_(True)
    br = conv.bias.detach().clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, 1, 1, 1)
    (yr * ct.to(torch.bfloat16).float()).sum().backward()
    assert float((y - yr).abs().max() / yr.abs().max()) < 2e-3
    assert float((x.grad.float() - xr.grad).abs().max() / xr.grad.abs().max()) < 1.5e-2
    assert float((conv.weight.grad - wr.grad).abs().max() / wr.grad.abs().max()) < 5e-3
    assert float((conv.bias.grad - br.grad).abs().max() / br.grad.abs().max()) < 5e-3


def test_pw_tc_wgrad_many_chunks_per_cta():
    """Regression: the end-of-kernel wait must not be satisfiable by an earlier phase of a cycling mbarrier."""
    g = torch.Generator(device="cuda").manual_seed(5)
    M, Co, Ci = 128 * 2400 + 77, 96, 64
    dz = torch.randn(1, 1, M, Co, generator=g, device="cuda").to(torch.bfloat16)
    x = torch.randn(1, 1, M, Ci, generator=g, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        dw = torch.zeros(Co, Ci, device="cuda")
        lib.call("nasb_pw_tc_wgrad", lib.ref(lib.desc(x.permute(0, 3, 1, 2))), lib.ref(lib.desc(dz.permute(0, 3, 1, 2))), lib.ptr(dw))
        torch.cuda.synchronize()
        ref = dz.float().reshape(M, Co).t() @ x.float().reshape(M, Ci)
        assert float((dw - ref).abs().max() / ref.abs().max()) < 2e-3
