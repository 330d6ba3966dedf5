"""TMA-tiled depthwise kernels (bf16) vs the gather kernels and vs an fp32 torch reference: forward (+ folded BN/act),
stride-1 data gradient, weight gradient; channel counts, kernel sizes, strides, dilations and odd image sizes of the
real networks, plus channel-slice views."""
import pytest
import torch
import torch.nn.functional as F

from nas_segm_b200 import lib

pytestmark = pytest.mark.gpu

CASES = [  # C, k, stride, dil, pad, H, W
    (32, 3, 1, 1, 1, 37, 53), (96, 3, 2, 1, 1, 64, 96), (144, 3, 1, 1, 1, 33, 65), (144, 3, 2, 1, 1, 31, 45),
    (192, 3, 1, 1, 1, 16, 32), (24, 5, 1, 1, 2, 40, 72), (32, 5, 1, 1, 2, 19, 23), (64, 5, 1, 1, 2, 8, 9),
    (48, 5, 2, 1, 2, 37, 53), (32, 5, 1, 6, 12, 33, 47), (64, 3, 1, 3, 3, 21, 34), (16, 3, 1, 1, 1, 5, 7),
    (64, 3, 2, 1, 1, 1, 1),
    # the search space's dilated separable ops at the search loop's map sizes (large patches of the dilated plan)
    (48, 5, 1, 6, 12, 64, 64), (48, 5, 1, 6, 12, 88, 88), (48, 3, 1, 3, 3, 88, 50), (64, 5, 1, 6, 12, 11, 11), (40, 3, 1, 3, 3, 9, 70),
]


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


def test_dw_tile_forward_dgrad_wgrad():
    g = torch.Generator(device="cuda").manual_seed(0)
    bad, covered = [], 0
    for (C, k, s, d, pad, H, W) in CASES:
        x = _nhwc(torch.randn(2, C, H, W, generator=g, device="cuda").to(torch.bfloat16))
        w = torch.randn(C, 1, k, k, generator=g, device="cuda") / k
        scale = torch.rand(C, generator=g, device="cuda") + 0.5
        shift = torch.randn(C, generator=g, device="cuda") * 0.1
        ref = F.conv2d(x.float(), w, None, s, pad, d, groups=C)
        OH, OW = ref.shape[2:]
        out = lib.new_act(2, C, OH, OW, torch.bfloat16, "cuda")
        ok = lib.try_call("nasb_dwconv_tile", lib.ref(lib.desc(x)), lib.ptr(w), k, s, d, pad, 0, lib.ptr(scale), lib.ptr(shift),
                          lib.ACT_RELU6, lib.ref(lib.desc(out)), None)
        if not ok:
            continue
        covered += 1
        torch.cuda.synchronize()
        want = (ref * scale[None, :, None, None] + shift[None, :, None, None]).clamp(0, 6)
        e = float((out.float() - want).abs().max() / want.abs().max().clamp_min(1e-6))
        if e > 1e-2:
            bad.append(("fwd", C, k, s, d, H, W, e))
        # weight gradient
        dz = _nhwc(torch.randn(2, C, OH, OW, generator=g, device="cuda").to(torch.bfloat16))
        dw = torch.zeros(C, 1, k, k, device="cuda")
        if lib.try_call("nasb_dwconv_wgrad_tile", lib.ref(lib.desc(x)), lib.ref(lib.desc(dz)), k, s, d, pad, lib.ptr(dw)):
            torch.cuda.synchronize()
            xr = x.float().requires_grad_(True)
            wr = w.clone().requires_grad_(True)
            F.conv2d(xr, wr, None, s, pad, d, groups=C).backward(dz.float())
            e = float((dw - wr.grad).abs().max() / wr.grad.abs().max().clamp_min(1e-6))
            if e > 5e-3:
                bad.append(("wgrad", C, k, s, d, H, W, e))
            dx = lib.new_act(2, C, H, W, torch.bfloat16, "cuda")
            dx.fill_(float("nan"))
            if s == 1:
                done = lib.try_call("nasb_dwconv_tile", lib.ref(lib.desc(dz)), lib.ptr(w), k, s, d, pad, 1, None, None,
                                    lib.ACT_NONE, lib.ref(lib.desc(dx)), None)
            else:
                done = lib.try_call("nasb_dwconv_dgrad_strided_tile", lib.ref(lib.desc(dz)), lib.ptr(w), k, s, d, pad,
                                    lib.ref(lib.desc(dx)))
            if done:
                torch.cuda.synchronize()
                e = float((dx.float() - xr.grad).abs().max() / xr.grad.abs().max().clamp_min(1e-6))
                if e > 1e-2:
                    bad.append(("dgrad", C, k, s, d, H, W, e))
    assert covered >= 10, covered
    assert not bad, bad


def test_dw_tile_channel_slices():
    g = torch.Generator(device="cuda").manual_seed(1)
    wide = torch.randn(2, 21, 30, 96, generator=g, device="cuda").to(torch.bfloat16)  # NHWC storage
    x = wide.permute(0, 3, 1, 2)[:, 16:64]                                            # 48-channel slice, pitch 96
    w = torch.randn(48, 1, 3, 3, generator=g, device="cuda") / 3
    outw = torch.zeros(2, 21, 30, 80, device="cuda", dtype=torch.bfloat16)
    out = outw.permute(0, 3, 1, 2)[:, 8:56]
    st = torch.zeros(2 * 48, dtype=torch.float64, device="cuda")
    assert lib.try_call("nasb_dwconv_tile", lib.ref(lib.desc(x)), lib.ptr(w), 3, 1, 1, 1, 0, None, None, lib.ACT_NONE,
                        lib.ref(lib.desc(out)), lib.ptr(st))
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), w, None, 1, 1, 1, groups=48)
    assert float((out.float() - ref).abs().max() / ref.abs().max()) < 1e-2
    assert float(outw[..., :8].abs().max()) == 0 and float(outw[..., 56:].abs().max()) == 0
    od = out.double().permute(0, 2, 3, 1).reshape(-1, 48)  # fused statistics == sums of the stored (bf16) output
    assert torch.allclose(st[:48], od.sum(0), rtol=1e-4, atol=1e-2) and torch.allclose(st[48:], (od * od).sum(0), rtol=1e-4, atol=1e-2)
