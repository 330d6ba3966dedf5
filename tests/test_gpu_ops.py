"""GPU parity of every registry primitive: CUDA kernels (through the C ABI) vs (a) fixtures produced by the real
reference and (b) the CPU oracle, forward (eval + train BN) and backward.  Tolerances: fp32 mode 1e-4 relative on
outputs, 1e-3 on gradients (north_star: 1e-3 relative on float logits); bf16 mode checked against fp32 loosely."""
import numpy as np
import pytest
import torch

import nas_segm_b200
from detweights import det_array
from golden_util import det_state_dict, keys_shapes, rel_err, t

pytestmark = pytest.mark.gpu


def _tags(fx, kind):
    return sorted({k.split("/")[0] for k in fx.files if k.startswith(kind)})


def _np(x):
    return x.detach().float().cpu().numpy()


def _build_op(fx, tag, idx):
    from nas_segm_b200.nn.layer_factory import OPS
    name, cin, cout, stride, repeats = [str(v) for v in fx[tag + "/meta"]]
    m = OPS[name](int(cin), int(cout), int(stride), True, int(repeats))
    m.load_state_dict(det_state_dict(keys_shapes(fx, tag + "/"), seed=idx), strict=True)
    return name, int(cin), m.cuda()


def test_registry_ops_forward_eval(golden):
    nas_segm_b200.set_act_dtype(torch.float32)
    fx = golden("ops")
    bad = []
    for idx, tag in enumerate(_tags(fx, "op")):
        name, cin, m = _build_op(fx, tag, idx)
        x = t(det_array(tag + "/x", (2, cin, 13, 17))).cuda()
        m.eval()
        with torch.no_grad():
            y = m(x)
        ref = fx[tag + "/y_eval"]
        if tuple(y.shape) != ref.shape or rel_err(_np(y), ref) > 1e-4:
            bad.append((tag, tuple(y.shape), ref.shape, rel_err(_np(y), ref) if tuple(y.shape) == ref.shape else None))
    assert not bad, bad


def test_registry_ops_train_forward_backward(golden):
    nas_segm_b200.set_act_dtype(torch.float32)
    fx = golden("ops")
    bad = []
    for idx, tag in enumerate(_tags(fx, "op")):
        name, cin, m = _build_op(fx, tag, idx)
        x = t(det_array(tag + "/x", (2, cin, 13, 17))).cuda().requires_grad_(True)
        m.train()
        y = m(x)
        e = rel_err(_np(y), fx[tag + "/y_train"])
        if e > 1e-4:
            bad.append((tag, "y_train", e))
        ct = t(det_array(tag + "/ct", tuple(y.shape))).cuda()
        params = dict(m.named_parameters())
        pk = [k for k in fx.files if k.startswith(tag + "/g/")]
        plist = [params[k.split("/g/")[1]] for k in pk]
        grads = torch.autograd.grad((y * ct).sum(), [x] + plist, allow_unused=True)
        gx = grads[0] if grads[0] is not None else torch.zeros_like(x)
        ref_gx = fx[tag + "/gx"]
        if np.abs(ref_gx).max() > 0:
            e = rel_err(_np(gx), ref_gx)
            if e > 1e-3:
                bad.append((tag, "gx", e))
        elif float(gx.abs().max()) != 0:
            bad.append((tag, "gx nonzero", float(gx.abs().max())))
        for k, g in zip(pk, grads[1:]):
            ref = fx[k]
            scale = max(np.abs(ref).max(), 1e-3)
            e = float(np.abs(_np(g) - ref).max() / scale)
            if e > 2e-3:
                bad.append((k, e))
        sd = m.state_dict()
        for k in [k for k in fx.files if k.startswith(tag + "/post/")]:
            e = rel_err(_np(sd[k.split("/post/")[1]]), fx[k])
            if e > 1e-4:
                bad.append((k, e))
    assert not bad, bad


def test_agg_ops_forward_backward(golden):
    from nas_segm_b200.nn.layer_factory import AGG_OPS
    nas_segm_b200.set_act_dtype(torch.float32)
    fx = golden("ops")
    bad = []
    for idx, tag in enumerate(_tags(fx, "agg")):
        meta = [str(v) for v in fx[tag + "/meta"]]
        name, (c0, c1, cout, larger, h0, w0, h1, w1) = meta[0], [int(v) for v in meta[1:]]
        m = AGG_OPS[name](c0, c1, cout, True, repeats=1, larger=bool(larger))
        m.load_state_dict(det_state_dict(keys_shapes(fx, tag + "/"), seed=100 + idx), strict=True)
        m = m.cuda()
        x = t(det_array(tag + "/x", (2, c0, h0, w0))).cuda().requires_grad_(True)
        y = t(det_array(tag + "/y", (2, c1, h1, w1))).cuda().requires_grad_(True)
        m.eval()
        with torch.no_grad():
            z = m(x, y)
        e = rel_err(_np(z), fx[tag + "/z_eval"])
        if e > 1e-4:
            bad.append((tag, "z_eval", e))
        m.train()
        z = m(x, y)
        e = rel_err(_np(z), fx[tag + "/z_train"])
        if e > 1e-4:
            bad.append((tag, "z_train", e))
        ct = t(det_array(tag + "/ct", tuple(z.shape))).cuda()
        params = dict(m.named_parameters())
        pk = [k for k in fx.files if k.startswith(tag + "/g/")]
        grads = torch.autograd.grad((z * ct).sum(), [x, y] + [params[k.split("/g/")[1]] for k in pk])
        for nm, g, ref in (("gx", grads[0], fx[tag + "/gx"]), ("gy", grads[1], fx[tag + "/gy"])):
            e = rel_err(_np(g), ref)
            if e > 1e-3:
                bad.append((tag, nm, e))
        for k, g in zip(pk, grads[2:]):
            ref = fx[k]
            e = float(np.abs(_np(g) - ref).max() / max(np.abs(ref).max(), 1e-3))
            if e > 2e-3:
                bad.append((k, e))
    assert not bad, bad


def test_registry_ops_bf16_close_to_fp32(golden):
    fx = golden("ops")
    bad = []
    try:
        for idx, tag in enumerate(_tags(fx, "op")):
            name, cin, m = _build_op(fx, tag, idx)
            x = t(det_array(tag + "/x", (2, cin, 13, 17))).cuda()
            m.eval()
            with torch.no_grad():
                y = m(x.to(torch.bfloat16))
            assert y.dtype == torch.bfloat16
            ref = fx[tag + "/y_eval"]
            e = rel_err(_np(y), ref)
            if e > 4e-2:
                bad.append((tag, e))
    finally:
        nas_segm_b200.set_act_dtype(torch.float32)
    assert not bad, bad


def test_ops_vs_oracle_odd_shapes():
    """Seeded comparison against the CPU oracle at shapes the fixtures do not cover (odd sizes, batch 3, C=24/40)."""
    from nas_segm_b200.nn.layer_factory import OPS
    from oracle import nas_oracle as O
    nas_segm_b200.set_act_dtype(torch.float32)
    bad = []
    for idx, (name, cin, cout, stride, rep, hw) in enumerate([
            ("sep_conv_5x5_dil6", 24, 24, 1, 2, (19, 31)), ("sep_conv_3x3", 24, 48, 2, 2, (33, 21)),
            ("max_pool_3x3", 40, 40, 2, 1, (17, 9)), ("conv3x3_dil12", 16, 16, 1, 1, (11, 11)),
            ("global_average_pool", 24, 24, 1, 1, (7, 5)), ("conv3x3", 8, 24, 2, 1, (9, 14)),
            ("sep_conv_7x7", 8, 8, 1, 1, (15, 15)), ("dil_conv_5x5", 16, 16, 2, 1, (21, 13)),
            ("avg_pool_3x3", 8, 16, 2, 1, (10, 10)), ("conv1x1", 8, 40, 2, 1, (9, 9))]):
        m = OPS[name](cin, cout, stride, True, rep)
        ks = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
        sd = det_state_dict(ks, seed=500 + idx)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().train()
        xs = det_array("odd%d" % idx, (3, cin) + hw)
        x = t(xs).cuda().requires_grad_(True)
        y = m(x)
        xo = t(xs).requires_grad_(True)
        P = O.Params({k: v.clone() for k, v in sd.items()}).requires_grad_()
        yo = O.op_forward(name, xo, P, "", cin, cout, stride, rep, training=True)
        if tuple(y.shape) != tuple(yo.shape):
            bad.append((name, "shape", tuple(y.shape), tuple(yo.shape)))
            continue
        e = rel_err(_np(y), yo.detach().numpy())
        if e > 1e-4:
            bad.append((name, "y", e))
        ct = t(det_array("oddct%d" % idx, tuple(yo.shape)))
        (gxo,) = torch.autograd.grad((yo * ct).sum(), [xo])
        (gx,) = torch.autograd.grad((y * ct.cuda()).sum(), [x])
        e = rel_err(_np(gx), gxo.numpy())
        if e > 1e-3:
            bad.append((name, "gx", e))
    assert not bad, bad


@pytest.mark.parametrize("act", [0, 1, 2])
def test_bf16_bn_affine_and_backward_vs_torch(act):
    """Speed-mode (bf16) training BatchNorm pieces through the C ABI -- nasb_bn_stats, nasb_affine_act, nasb_bn_act_bwd (the
    4-pixel-in-flight packed kernels) -- against torch autograd in fp32 on the same bf16-rounded inputs; channel counts
    whose vector count does not divide the block (24, 144), ragged pixel counts, a channel-slice view."""
    from nas_segm_b200 import lib
    from nas_segm_b200.lib import call, desc, ptr, ref
    g = torch.Generator(device="cuda").manual_seed(5 + act)
    for (n, c, h, w, sliced) in [(2, 24, 37, 53, False), (3, 144, 19, 23, False), (2, 64, 64, 96, True), (1, 32, 7, 5, False),
                                 (4, 192, 40, 64, False)]:
        cs = c + 16 if sliced else c
        zbuf = (torch.randn(n, h, w, cs, generator=g, device="cuda") * 1.5 + 0.7).to(torch.bfloat16)
        z = zbuf.permute(0, 3, 1, 2)[:, :c]
        dy = lib.new_act(n, c, h, w, torch.bfloat16, "cuda")
        dy.copy_(torch.randn(n, c, h, w, generator=g, device="cuda"))
        gamma = torch.rand(c, generator=g, device="cuda") + 0.5
        beta = torch.randn(c, generator=g, device="cuda") * 0.3
        rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
        sm, sr, sc, sh = (torch.empty(c, device="cuda") for _ in range(4))
        ws = lib.workspace(torch.device("cuda"), 1 << 20)
        call("nasb_bn_stats", ref(desc(z)), ptr(gamma), ptr(beta), 1e-5, 0.1, ptr(rm), ptr(rv), ptr(sm), ptr(sr), ptr(sc), ptr(sh), None, ptr(ws))
        y = lib.new_act(n, c, h, w, torch.bfloat16, "cuda")
        call("nasb_affine_act", ref(desc(z)), ptr(sc), ptr(sh), act, ref(desc(y)))
        dz = lib.new_act(n, c, h, w, torch.bfloat16, "cuda")
        dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
        call("nasb_bn_act_bwd", ref(desc(dy)), None, ref(desc(z)), act, ptr(gamma), ptr(beta), ptr(sc), ptr(sh), ptr(sm), ptr(sr), 1,
             ptr(dg), ptr(db), ref(desc(dz)), ptr(ws))
        torch.cuda.synchronize()
        zr = z.float().clone().requires_grad_(True)
        gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        yr = torch.nn.functional.batch_norm(zr, None, None, gr, br, True, 0.1, 1e-5)
        yr = torch.relu(yr) if act == 1 else (torch.clamp(yr, 0, 6) if act == 2 else yr)
        yr.backward(dy.float())
        tag = (n, c, h, w, act)
        assert float((y.float() - yr).abs().max() / yr.abs().max()) < 1e-2, tag
        # elements whose pre-activation is within bf16 rounding of a kink may gate differently: compare in the mean
        assert float((dz.float() - zr.grad).abs().mean() / zr.grad.abs().mean()) < 1.5e-2, tag
        assert float((dg - gr.grad).abs().max() / gr.grad.abs().max()) < 1e-2, tag
        assert float((db - br.grad).abs().max() / br.grad.abs().max()) < 1e-2, tag
