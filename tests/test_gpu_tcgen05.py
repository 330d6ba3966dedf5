"""tcgen05 pointwise-GEMM kernel vs an fp32 torch matmul of the same bf16 operands and vs the CUDA-core path, over
channel counts of the real networks, M tails, channel-slice views, every epilogue variant and the fused BN statistics."""
import pytest
import torch

import nas_segm_b200
from nas_segm_b200 import functional as Fn
from nas_segm_b200 import lib

pytestmark = pytest.mark.gpu


def _ref(x, w, scale, shift, act, res):
    y = x.float().reshape(-1, x.shape[-1]) @ w.to(torch.bfloat16).float().t()
    if scale is not None:
        y = y * scale + shift
    if act == lib.ACT_RELU:
        y = y.clamp_min(0)
    elif act == lib.ACT_RELU6:
        y = y.clamp(0, 6)
    if res is not None:
        y = y + res.float().reshape(-1, res.shape[-1])
    return y


def _run_tc(x_nhwc, w, scale, shift, act, res_nhwc, stats=False, out_buf=None):
    """x_nhwc: [n,h,w,K] bf16 (possibly a channel slice of a wider buffer)."""
    n, h, ww, K = x_nhwc.shape
    N = w.shape[0]
    x = x_nhwc.permute(0, 3, 1, 2)
    out = lib.new_act(n, N, h, ww, torch.bfloat16, x.device) if out_buf is None else out_buf
    wp = Fn._pack_weight(w.reshape(N, K, 1, 1).contiguous(), False)
    st = torch.zeros(2 * N, dtype=torch.float64, device=x.device) if stats else None
    lib.call("nasb_pw_tc_fwd", lib.ref(lib.desc(x)), lib.ptr(wp), N, lib.ptr(scale), lib.ptr(shift), act,
             lib.ref(lib.desc(res_nhwc.permute(0, 3, 1, 2))) if res_nhwc is not None else None, lib.ref(lib.desc(out)),
             lib.ptr(st))
    torch.cuda.synchronize()
    return out.permute(0, 2, 3, 1), st


def test_tc_gemm_shapes_and_tails():
    g = torch.Generator(device="cuda").manual_seed(0)
    bad = []
    for K, N in [(16, 96), (96, 24), (24, 144), (144, 24), (144, 32), (32, 192), (192, 32), (24, 48), (32, 64), (64, 64),
                 (128, 64), (224, 64), (8, 8), (48, 48), (64, 16), (16, 256), (256, 16), (200, 72), (96, 576), (448, 40),
                 (64, 136),
                 # K-ring mode of the warp-specialised kernel (more than 7 K blocks): MobileNet-v2's 960-channel layers
                 (960, 160), (960, 320), (576, 96), (520, 24), (1280, 40)]:
        assert lib.load().nasb_pw_tc_supported(K, N) == 1, (K, N)
        for M in (1, 100, 128, 129, 4099):
            x = torch.randn(1, 1, M, K, generator=g, device="cuda").to(torch.bfloat16)
            w = torch.randn(N, K, generator=g, device="cuda") / K ** 0.5
            y, _ = _run_tc(x, w, None, None, lib.ACT_NONE, None)
            ref = _ref(x, w, None, None, lib.ACT_NONE, None)
            err = float((y.float().reshape(-1, N) - ref).abs().max() / ref.abs().max().clamp_min(1e-6))
            if not err < 1e-2:
                bad.append((K, N, M, err))
    assert not bad, bad


def test_tc_epilogues_slices_and_stats():
    g = torch.Generator(device="cuda").manual_seed(1)
    n, h, w_, K, N = 3, 37, 41, 64, 96
    wide = torch.randn(n, h, w_, K + 32, generator=g, device="cuda").to(torch.bfloat16)
    x = wide[..., 16:16 + K]                       # channel slice: pixel pitch 96, 32-byte offset
    wt = torch.randn(N, K, generator=g, device="cuda") / 8
    scale = torch.rand(N, generator=g, device="cuda") + 0.5
    shift = torch.randn(N, generator=g, device="cuda")
    res = torch.randn(n, h, w_, N, generator=g, device="cuda").to(torch.bfloat16)
    for act in (lib.ACT_NONE, lib.ACT_RELU, lib.ACT_RELU6):
        for r in (None, res):
            y, _ = _run_tc(x, wt, scale, shift, act, r)
            ref = _ref(x, wt, scale, shift, act, r)
            err = float((y.float().reshape(-1, N) - ref).abs().max() / ref.abs().max())
            assert err < 1e-2, (act, r is not None, err)
    # output into a channel slice of a wider concat buffer; untouched channels stay untouched
    buf = torch.full((n, h, w_, N + 64), 7.0, device="cuda", dtype=torch.bfloat16)
    out_view = buf.permute(0, 3, 1, 2)[:, 32:32 + N]
    y, _ = _run_tc(x, wt, None, None, lib.ACT_NONE, None, out_buf=out_view)
    ref = _ref(x, wt, None, None, lib.ACT_NONE, None)
    assert float((buf[..., 32:32 + N].float().reshape(-1, N) - ref).abs().max() / ref.abs().max()) < 1e-2
    assert float((buf[..., :32] - 7).abs().max()) == 0 and float((buf[..., 32 + N:] - 7).abs().max()) == 0
    # fused statistics == sums of the stored (bf16) output
    y, st = _run_tc(x, wt, None, None, lib.ACT_NONE, None, stats=True)
    yd = y.double().reshape(-1, N)
    assert torch.allclose(st[:N], yd.sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(st[N:], (yd * yd).sum(0), rtol=1e-4, atol=1e-2)


def test_tc_k_ring_epilogues_stats_and_many_tiles():
    """Large C_in (K-ring mode): every epilogue variant, the fused statistics, and enough tiles per CTA for the ring and the
    accumulator barriers to cycle through many phases."""
    g = torch.Generator(device="cuda").manual_seed(11)
    K, N = 960, 160
    for M in (128 * 700 + 5, 3872):
        x = torch.randn(1, 1, M, K, generator=g, device="cuda").to(torch.bfloat16)
        wt = torch.randn(N, K, generator=g, device="cuda") / K ** 0.5
        scale = torch.rand(N, generator=g, device="cuda") + 0.5
        shift = torch.randn(N, generator=g, device="cuda")
        res = torch.randn(1, 1, M, N, generator=g, device="cuda").to(torch.bfloat16)
        for act, r in ((lib.ACT_NONE, None), (lib.ACT_RELU6, res), (lib.ACT_RELU, None)):
            y, _ = _run_tc(x, wt, scale, shift, act, r)
            ref = _ref(x, wt, scale, shift, act, r)
            err = float((y.float().reshape(-1, N) - ref).abs().max() / ref.abs().max())
            assert err < 1e-2, (M, act, r is not None, err)
        y, st = _run_tc(x, wt, None, None, lib.ACT_NONE, None, stats=True)
        yd = y.float().reshape(-1, N).double()
        assert torch.allclose(st[:N], yd.sum(0), rtol=1e-4, atol=5e-2)
        assert torch.allclose(st[N:], (yd * yd).sum(0), rtol=1e-4, atol=5e-2)


def test_conv_unit_tc_vs_cuda_core_path():
    """The fused conv unit (train-mode BN, backward) on the tensor-core path vs the CUDA-core path, same bf16 inputs."""
    from nas_segm_b200.nn.layer_factory import conv_bn_relu
    torch.manual_seed(0)
    m = conv_bn_relu(64, 96, 1, 1, 0).cuda().train()
    x = torch.randn(4, 64, 45, 53, device="cuda").to(torch.bfloat16)
    outs = {}
    for tc in (True, False):
        nas_segm_b200.config().use_tcgen05 = tc
        try:
            xi = x.clone().requires_grad_(True)
            m.zero_grad()
            m[1].reset_running_stats()
            y = m(xi)
            (y.float() * torch.linspace(-1, 1, y.numel(), device="cuda").reshape(y.shape)).sum().backward()
            outs[tc] = (y.detach().float(), xi.grad.float(), m[0].weight.grad.clone(), m[1].weight.grad.clone(),
                        m[1].running_var.clone())
        finally:
            nas_segm_b200.config().use_tcgen05 = True
    # The two paths round differently (bf16 weights on the tensor cores), so a few ReLU masks flip for |y| ~ 0 and single
    # elements of the input gradient legitimately move by a whole term: compare in the mean, not in the max.
    for a, b, tol in zip(outs[True], outs[False], (1e-2, 5e-2, 5e-2, 5e-2, 1e-3)):
        assert float((a - b).abs().mean() / b.abs().mean()) < tol


def test_tc_dgrad_is_the_same_kernel_with_transposed_pack():
    g = torch.Generator(device="cuda").manual_seed(2)
    cout, cin, M = 96, 64, 3001
    dz = torch.randn(1, 1, M, cout, generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randn(cout, cin, 1, 1, generator=g, device="cuda") / 10
    dx = lib.new_act(1, cin, 1, M, torch.bfloat16, "cuda")
    lib.call("nasb_pw_tc_fwd", lib.ref(lib.desc(dz.permute(0, 3, 1, 2))), lib.ptr(Fn._pack_weight(w, True)), cin, None, None,
             lib.ACT_NONE, None, lib.ref(lib.desc(dx)), None)
    ref = dz.float().reshape(M, cout) @ w.reshape(cout, cin).to(torch.bfloat16).float()
    got = dx.permute(0, 2, 3, 1).float().reshape(M, cin)
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-2
    # expand convolution of an inverted-residual block with 960 hidden channels: the data gradient has K = 960
    cout, cin, M = 960, 160, 3872
    dz = torch.randn(1, 1, M, cout, generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randn(cout, cin, 1, 1, generator=g, device="cuda") / 30
    dx = lib.new_act(1, cin, 1, M, torch.bfloat16, "cuda")
    lib.call("nasb_pw_tc_fwd", lib.ref(lib.desc(dz.permute(0, 3, 1, 2))), lib.ptr(Fn._pack_weight(w, True)), cin, None, None,
             lib.ACT_NONE, None, lib.ref(lib.desc(dx)), None)
    ref = dz.float().reshape(M, cout) @ w.reshape(cout, cin).to(torch.bfloat16).float()
    got = dx.permute(0, 2, 3, 1).float().reshape(M, cin)
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-2


def test_tc_wgrad_mn_major():
    """dW = dz^T x on the tensor cores (MN-major descriptors) vs an fp32 matmul of the same bf16 operands."""
    g = torch.Generator(device="cuda").manual_seed(3)
    bad = []
    for Co, Ci in [(64, 64), (96, 16), (24, 96), (144, 24), (32, 144), (64, 128), (64, 224), (192, 32), (48, 24), (8, 8),
                   (256, 64), (16, 256), (960, 160), (160, 960), (320, 960), (40, 520)]:  # > 256 input channels: x channel blocks
        assert lib.load().nasb_pw_tc_wgrad_supported(Co, Ci) == 1, (Co, Ci)
        for M in (1, 127, 128, 300, 20011):
            dz = torch.randn(1, 1, M, Co, generator=g, device="cuda").to(torch.bfloat16)
            x = torch.randn(1, 1, M, Ci, generator=g, device="cuda").to(torch.bfloat16)
            dw = torch.zeros(Co, Ci, device="cuda")
            lib.call("nasb_pw_tc_wgrad", lib.ref(lib.desc(x.permute(0, 3, 1, 2))), lib.ref(lib.desc(dz.permute(0, 3, 1, 2))),
                     lib.ptr(dw))
            torch.cuda.synchronize()
            ref = dz.float().reshape(M, Co).t() @ x.float().reshape(M, Ci)
            err = float((dw - ref).abs().max() / ref.abs().max().clamp_min(1e-6))
            if not err < 2e-3:
                bad.append((Co, Ci, M, err))
    assert not bad, bad
    # accumulates (+=) and honours channel-slice views
    wide = torch.randn(2, 9, 11, 96, generator=g, device="cuda").to(torch.bfloat16)
    x, dz = wide[..., 8:40], wide[..., 48:96]
    dw = torch.ones(48, 32, device="cuda")
    lib.call("nasb_pw_tc_wgrad", lib.ref(lib.desc(x.permute(0, 3, 1, 2))), lib.ref(lib.desc(dz.permute(0, 3, 1, 2))), lib.ptr(dw))
    ref = 1 + dz.float().reshape(-1, 48).t() @ x.float().reshape(-1, 32)
    assert float((dw - ref).abs().max() / ref.abs().max()) < 2e-3


def test_stem_as_tensor_core_gemm():
    """Speed-mode encoder stem (planar fp32 image -> bf16 patch matrix -> tcgen05 GEMM, K = 32) against torch fp32:
    conv 3->32, stride 2, training-mode BN, ReLU6; output, weight / BN gradients, odd image sizes."""
    import torch.nn.functional as F
    from nas_segm_b200.nn.layer_factory import conv_bn_relu6
    torch.manual_seed(3)
    for (n, h, w) in [(2, 64, 96), (3, 37, 53), (1, 18, 130)]:
        m = conv_bn_relu6(3, 32, 2).cuda().train()
        x = torch.randn(n, 3, h, w, device="cuda")
        nas_segm_b200.set_act_dtype(torch.bfloat16)
        try:
            y = m(x)
            assert y.dtype == torch.bfloat16
            gy = torch.randn(y.shape, device="cuda")
            (y.float() * gy).sum().backward()
        finally:
            nas_segm_b200.set_act_dtype(torch.float32)
        wr = m[0].weight.detach().clone().requires_grad_(True)
        gr, br = m[1].weight.detach().clone().requires_grad_(True), m[1].bias.detach().clone().requires_grad_(True)
        z = F.conv2d(x, wr, None, 2, 1)
        yr = torch.clamp(F.batch_norm(z, None, None, gr, br, True, 0.1, 1e-5), 0, 6)
        (yr * gy.to(torch.bfloat16).float()).sum().backward()
        assert float((y.float() - yr).abs().max() / yr.abs().max()) < 2e-2, (n, h, w)
        for a, b, tol in ((m[0].weight.grad, wr.grad, 5e-2), (m[1].weight.grad, gr.grad, 5e-2), (m[1].bias.grad, br.grad, 5e-2)):
            assert float((a - b).abs().mean() / b.abs().mean()) < tol, (n, h, w, float((a - b).abs().mean() / b.abs().mean()))


def test_inverted_residual_block_bf16_training_vs_torch():
    """A MobileNet-v2 residual block in speed mode, training-mode BN -- pointwise tcgen05 GEMMs with fused statistics,
    depthwise TMA tile kernel with fused statistics, finalise+apply+residual in one pass, packed bf16 BN backward --
    against the same block evaluated by torch in fp32: output, input gradient, weight / BN gradients."""
    import copy
    from nas_segm_b200.nn.layer_factory import InvertedResidual
    torch.manual_seed(11)
    for (inp, oup, stride, n, h, w) in [(32, 32, 1, 2, 40, 56), (24, 24, 1, 3, 33, 47), (32, 64, 2, 2, 36, 52)]:
        m = InvertedResidual(inp, oup, stride, 6).cuda().train()
        ref = copy.deepcopy(m.conv).float()
        x = torch.randn(n, inp, h, w, device="cuda").to(torch.bfloat16)
        nas_segm_b200.set_act_dtype(torch.bfloat16)
        try:
            xi = lib.to_nhwc(x.clone()).requires_grad_(True)
            y = m(xi)
            gy = torch.randn(y.shape, device="cuda").to(torch.bfloat16)
            (y.float() * gy.float()).sum().backward()
        finally:
            nas_segm_b200.set_act_dtype(torch.float32)
        xr = x.float().requires_grad_(True)
        yr = ref(xr) + (xr if m.use_res_connect else 0)
        (yr * gy.float()).sum().backward()
        tag = (inp, oup, stride)
        assert float((y.float() - yr).abs().mean() / yr.abs().mean()) < 2e-2, tag
        assert float((xi.grad.float() - xr.grad).abs().mean() / xr.grad.abs().mean()) < 6e-2, tag
        # The gradient of the FIRST BatchNorm's gamma is analytically ~0 (a per-channel scale in front of a depthwise conv
        # is normalised away by the next BatchNorm; only the ReLU6 kinks break the invariance): 1e-4 of the bias gradient
        # in fp32, pure cancellation noise in any bf16 evaluation.  BatchNorm gradients are therefore judged on the scale of
        # the (gamma, beta) pair, not of each vector alone.
        named, refp = dict(m.conv.named_parameters()), dict(ref.named_parameters())
        for k, a in named.items():
            b = refp[k].grad
            scale = b.abs().mean()
            if a.dim() == 1:
                idx = k.split(".")[0]
                scale = torch.maximum(refp[idx + ".weight"].grad.abs().mean(), refp[idx + ".bias"].grad.abs().mean())
            e = float((a.grad - b).abs().mean() / scale.clamp_min(1e-6))
            assert e < 8e-2, (tag, k, tuple(a.shape), e)
        for a, b in ((m.conv[1].running_var, ref[1].running_var), (m.conv[7].running_mean, ref[7].running_mean)):
            assert float((a - b).abs().max() / b.abs().max()) < 2e-2, tag
        assert int(m.conv[4].num_batches_tracked) == 1


def test_bn_backward_fusion_matches_unfused_path():
    """InvertedResidual in speed mode with the BN-backward fusion links on (gated depthwise data gradient + the expansion
    unit's backward without dz: functional._BnRec, include/nasb200.h NasbGate / nasb_pw_bn_bwd_prepare) against the same
    block with them off (separate sums + dz passes): every gradient agrees to bf16 rounding, and the fused entry points
    really ran."""
    import copy
    from nas_segm_b200.nn.layer_factory import InvertedResidual
    cfg = nas_segm_b200.config()
    torch.manual_seed(5)
    seen = []
    orig, orig_try = lib.call, lib.try_call

    def spy(name, *a):
        seen.append(name)
        return orig(name, *a)

    def spy_try(name, *a):
        ok = orig_try(name, *a)
        if ok:
            seen.append(name)
        return ok
    for (inp, oup, stride, t, n, h, w) in [(16, 24, 2, 6, 2, 64, 96), (24, 24, 1, 6, 3, 33, 47), (32, 16, 1, 1, 2, 40, 56),
                                           (32, 64, 2, 6, 2, 37, 53)]:
        m = InvertedResidual(inp, oup, stride, t).cuda().train()
        m2 = copy.deepcopy(m)
        x = torch.randn(n, inp, h, w, device="cuda").to(torch.bfloat16)
        nas_segm_b200.set_act_dtype(torch.bfloat16)
        try:
            res = []
            for mod, fuse in ((m, True), (m2, False)):
                cfg.fuse_bn_bwd = fuse
                xi = lib.to_nhwc(x.clone()).requires_grad_(True)
                del seen[:]
                lib.call = Fn.call = spy
                lib.try_call = Fn.try_call = spy_try
                try:
                    y = mod(xi)
                    gy = torch.randn(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)).to(torch.bfloat16)
                    (y.float() * gy.float()).sum().backward()
                finally:
                    lib.call = Fn.call = orig
                    lib.try_call = Fn.try_call = orig_try
                res.append((y.detach().float(), xi.grad.float(), {k: p.grad.clone() for k, p in mod.conv.named_parameters()}, list(seen)))
        finally:
            cfg.fuse_bn_bwd = False
            nas_segm_b200.set_act_dtype(torch.float32)
        (y1, gx1, g1, calls1), (y2, gx2, g2, calls2) = res
        tag = (inp, oup, stride, t)
        # expansion: no BN pass at all (dz never formed); depthwise: the dz pass only, from the projection's gated epilogue
        nodz = t >= 2  # the algebraic path needs a narrow input (2 * C_in <= C_out); t = 1 keeps the dz pass, from the sums
        assert ("nasb_pw_bn_bwd_prepare" in calls1) == nodz and "nasb_pw_bn_bwd_prepare" not in calls2, tag
        assert "nasb_dwconv_dgrad_gated" in calls1 and "nasb_pw_tc_dgrad_gated" in calls1 and "nasb_bn_bwd_from_sums" in calls1, tag
        assert calls1.count("nasb_bn_bwd_from_sums") == (1 if nodz else 2), tag
        assert not any(c in calls2 for c in ("nasb_dwconv_dgrad_gated", "nasb_pw_tc_dgrad_gated", "nasb_bn_bwd_from_sums")), tag
        assert calls1.count("nasb_bn_act_bwd") == calls2.count("nasb_bn_act_bwd") - 2, tag
        # same forward kernels; the fp64 statistics are reduced with atomics, so a rare last-bit difference of a batch mean can
        # flip single bf16 roundings of y
        assert float((y1 - y2).abs().max()) <= 2.0 ** -7 * float(y2.abs().max()), tag
        assert float((gx1 - gx2).norm() / gx2.norm()) < 2e-2, (tag, float((gx1 - gx2).norm() / gx2.norm()))
        for k in g1:
            scale = g2[k].norm()
            if g1[k].dim() == 1:  # BN vectors: judged on the scale of the (gamma, beta) pair (see the test above)
                idx = k.split(".")[0]
                scale = torch.maximum(g2[idx + ".weight"].norm(), g2[idx + ".bias"].norm())
            e = float((g1[k] - g2[k]).norm() / scale.clamp_min(1e-6))
            assert e < 3e-2, (tag, k, e)


def test_fused_separable_kernel_matches_two_unit_path():
    """csrc/sep_tcgen05.cu (depthwise -> folded BN / activation -> tcgen05 pointwise -> epilogue in ONE kernel, inference)
    against the same modules evaluated as two units (TMA depthwise tile kernel + tcgen05 pointwise kernel) and against torch
    fp32: InvertedResidual blocks with and without residual, SepConv 3x3 / 5x5 with two repeats, ragged sizes, a channel
    count that is not a multiple of 64 (144) and more than one 64-channel block (192)."""
    import copy
    from nas_segm_b200.nn.layer_factory import OPS, InvertedResidual
    cfg = nas_segm_b200.config()
    torch.manual_seed(9)
    seen = []
    orig_try = lib.try_call

    def spy_try(name, *a):
        ok = orig_try(name, *a)
        if ok:
            seen.append(name)
        return ok
    # (the kernel takes channel counts with <= 15 % padding to whole 64-channel blocks: 56, 64, 120, 128, 192, ...)
    cases = [("ir", dict(inp=32, oup=32, stride=1, expand_ratio=6), (1, 32, 40, 64)),     # 192 channels, residual
             ("ir", dict(inp=32, oup=16, stride=1, expand_ratio=2), (2, 32, 21, 50)),     # 64 channels, no residual
             ("ir", dict(inp=40, oup=40, stride=1, expand_ratio=3), (2, 40, 33, 47)),     # 120 channels: partial last block
             ("sep_conv_3x3", 56, (2, 56, 37, 29)), ("sep_conv_5x5", 64, (1, 64, 40, 72)), ("sep_conv_5x5", 64, (3, 64, 9, 17))]
    for kind, arg, shape in cases:
        m = (InvertedResidual(**arg) if kind == "ir" else OPS[kind](arg, arg, 1, True, repeats=2)).cuda().eval()
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.2)
                mod.running_var.uniform_(0.5, 1.5)
                mod.weight.data.uniform_(0.5, 1.5)
                mod.bias.data.normal_(0, 0.2)
        x = torch.randn(*shape, device="cuda").to(torch.bfloat16)
        nas_segm_b200.set_act_dtype(torch.bfloat16)
        try:
            outs = []
            for fused in (True, False):
                cfg.fused_sep = fused
                del seen[:]
                lib.try_call = Fn.try_call = spy_try
                try:
                    with torch.no_grad():
                        outs.append((m(lib.to_nhwc(x.clone())).float(), list(seen)))
                finally:
                    lib.try_call = Fn.try_call = orig_try
        finally:
            cfg.fused_sep = True
            nas_segm_b200.set_act_dtype(torch.float32)
        (yf, calls_f), (yu, calls_u) = outs
        assert "nasb_sep_unit_infer" in calls_f and "nasb_sep_unit_infer" not in calls_u, (kind, arg)
        assert calls_f.count("nasb_sep_unit_infer") == (1 if kind == "ir" else 2), (kind, arg)
        scale = float(yu.abs().max())
        assert float((yf - yu).abs().max()) <= 2.0 ** -6 * scale, (kind, arg, float((yf - yu).abs().max()) / scale)
        ref = copy.deepcopy(m)
        # torch fp32 of the same modules (their nn.Conv2d / BatchNorm2d children are real parameter holders)
        with torch.no_grad():
            if kind == "ir":
                yr = ref.conv(x.float()) + (x.float() if ref.use_res_connect else 0)
            else:
                yr = ref.op(x.float())
        assert float((yf - yr).abs().max()) <= 3e-2 * float(yr.abs().max()), (kind, arg)
