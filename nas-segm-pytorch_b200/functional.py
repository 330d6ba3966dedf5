"""Autograd glue between torch tensors and the C-ABI kernels (include/nasb200.h).

Every function here is a thin ``torch.autograd.Function``: ``forward`` / ``backward`` only allocate outputs and call
the library; there is no torch arithmetic on activations.  Tensors are logical NCHW views over dense NHWC storage
(``lib.new_act``); parameters stay fp32 in the reference's layouts.
"""
import contextlib
import ctypes as C

import torch

from . import lib, packs
from .lib import ACT_NONE, ACT_RELU, ACT_RELU6, POOL_AVG, POOL_MAX, call, desc, ptr, ref, try_call  # noqa: F401


def act_dtype():
    return _cfg().act_dtype


class _NoGradCtx:
    """Stand-in for the autograd context when gradients are off: the forward of every Function below then runs as a plain
    call (torch.autograd.Function.apply costs ~15 us per op on the host, 135 ops per arch0 forward)."""
    needs_input_grad = (False,) * 32

    def save_for_backward(self, *tensors):
        pass


def _apply(fn, *args):
    if torch.is_grad_enabled():
        return fn.apply(*args)
    return fn.forward(_NoGradCtx(), *args)


_config_obj = None


def _cfg():
    global _config_obj
    if _config_obj is None:
        from . import config
        _config_obj = config()
    return _config_obj


def conv_out_hw(h, w, ks, stride, dil, pad):
    return ((h + 2 * pad - dil * (ks - 1) - 1) // stride + 1, (w + 2 * pad - dil * (ks - 1) - 1) // stride + 1)


def _ws(dev, c=0):
    return lib.workspace(dev, max(1 << 20, 24 * c + 256))


def _grad_in(dy, like_dtype):
    """Incoming gradient -> NHWC storage of the expected dtype (a no-op when it comes from our own kernels)."""
    return lib.to_nhwc(dy, like_dtype)


def _dw_stats_on():
    return _cfg().fuse_dw_stats


def _bn_momentum(bn):
    """nn.BatchNorm2d's update factor: `momentum`, or 1 / num_batches_tracked (cumulative moving average) when momentum is
    None -- that reads the device counter (a host sync, and an error under CUDA-graph capture, which cannot bake it in)."""
    if bn.momentum is not None:
        return float(bn.momentum)
    return 1.0 / (int(bn.num_batches_tracked.item()) + 1)


def _tiles_on():
    return _cfg().use_tma_tiles


def _tc_ok(x, cin, cout, out_dtype):
    """Can this pointwise GEMM (K=cin -> N=cout) run on the tcgen05 kernel?  bf16 in/out, 8-aligned channels/pitches."""
    if not _cfg().use_tcgen05 or x.dtype != torch.bfloat16 or out_dtype != torch.bfloat16:
        return False
    if x.data_ptr() % 16 or lib.desc(x).cstride % 8:
        return False
    return bool(lib.load().nasb_pw_tc_supported(int(cin), int(cout)))


def _tc_wgrad_ok(x, dz, cout):
    if not _cfg().use_tcgen05 or x.dtype != torch.bfloat16 or dz.dtype != torch.bfloat16:
        return False
    if x.data_ptr() % 16 or dz.data_ptr() % 16 or lib.desc(x).cstride % 8 or lib.desc(dz).cstride % 8:
        return False
    return bool(lib.load().nasb_pw_tc_wgrad_supported(int(cout), int(x.shape[1])))


def _stem_tc_ok(cout):
    lb = lib.load()
    return bool(_cfg().use_tcgen05 and cout % 8 == 0 and lb.nasb_pw_tc_supported(32, int(cout))
                and lb.nasb_pw_tc_wgrad_supported(int(cout), 32))


def _c3_ok(x, cin, cout, ks, stride, dil, pad):
    """3x3 implicit GEMM on the tensor cores: bf16 input with a 16-byte pixel pitch, stride 1, 'same' geometry."""
    if not _cfg().use_tcgen05 or ks != 3 or stride != 1 or pad != dil or x.dtype != torch.bfloat16:
        return False
    if x.data_ptr() % 16 or lib.desc(x).cstride % 8:
        return False
    return bool(lib.load().nasb_conv3_tc_supported(int(cin), int(cout)))


def _pack_conv3(weight, mode):
    wp = packs.get(weight, packs.C3_T if mode else packs.C3)
    if wp is not None:
        return wp
    cout, cin = weight.shape[0], weight.shape[1]
    n = int(lib.load().nasb_pack_conv3_elems(cout, cin, mode))
    wp = torch.empty(n, dtype=torch.bfloat16, device=weight.device)
    call("nasb_pack_conv3_bf16", ptr(weight), cout, cin, mode, ptr(wp))
    return wp


def _bf16_padded_copy(t):
    """fp32 / bf16 NHWC tensor -> bf16 copy whose pixel pitch is a multiple of 8 elements (TMA needs a 16-byte pitch; the
    channel count itself may be odd, e.g. the 19 / 21 logit channels)."""
    n, c, h, w = t.shape
    pitch = (c + 7) // 8 * 8
    buf = torch.empty((n, h, w, pitch), dtype=torch.bfloat16, device=t.device)
    view = buf[..., :c].permute(0, 3, 1, 2)
    call("nasb_scale_copy", ref(desc(t)), None, 0, ref(desc(view)))
    return view


def _pack_weight(weight, transpose):
    """fp32 [C_out, C_in, 1, 1] -> bf16 [R][Kp] operand of the tensor-core kernel.  Inside an engine iteration the operand
    comes from the model's persistent packs (packs.begin: one launch for the whole model); otherwise it is packed here."""
    wp = packs.get(weight, packs.PW_T if transpose else packs.PW)
    if wp is not None:
        return wp
    cout, cin = weight.shape[0], weight.shape[1]
    r, k = (cin, cout) if transpose else (cout, cin)
    if not torch.is_grad_enabled():  # inference: one buffer per weight, rewritten (in stream order) by every call
        wp = getattr(weight, "_nasb_wp", None)
        if wp is None or wp.device != weight.device:
            wp = weight._nasb_wp = torch.empty(r * ((k + 7) // 8 * 8), dtype=torch.bfloat16, device=weight.device)
    else:
        wp = torch.empty(r * ((k + 7) // 8 * 8), dtype=torch.bfloat16, device=weight.device)
    call("nasb_pack_weight_bf16", ptr(weight), cout, cin, 1 if transpose else 0, ptr(wp))
    return wp


class _BnRec:
    """Link between a conv -> BN(train) -> act unit and the SOLE consumer of its output (set by InvertedResidual / SepConv,
    where the graph guarantees it): the consumer's data-gradient kernel gates its result with this unit's activation mask
    and accumulates this unit's two BatchNorm-backward reductions in its epilogue (NasbGate), so that the unit's own backward
    needs no pass over (dy, z) for them -- and, for a pointwise unit, no dz at all (include/nasb200.h)."""
    __slots__ = ("z", "ss", "act", "g", "sums")

    def __init__(self, z, ss, act):
        self.z, self.ss, self.act, self.g, self.sums = z, ss, act, None, None


def _pw_unit_bwd_nodz(g, x0, weight, has_g, has_b, ss, sv, sums, need_dx):
    """Backward of a pointwise conv -> BN(train) -> act unit from the gated gradient g and the reductions S1, S2 its consumer
    left behind: dW, dgamma, dbeta and dx without forming dz (nasb_pw_bn_bwd_prepare)."""
    dev = g.device
    cout, cin = weight.shape[0], weight.shape[1]
    n, _, h, w = g.shape
    gx = lib.zeros((cout, cin), torch.float32, dev)
    call("nasb_pw_tc_wgrad", ref(desc(x0)), ref(desc(g)), ptr(gx))
    xx = lib.zeros((cin, cin), torch.float32, dev)
    call("nasb_pw_tc_wgrad", ref(desc(x0)), ref(desc(x0)), ptr(xx))
    sx = lib.zeros(cin, torch.float32, dev)
    call("nasb_channel_sum", ref(desc(x0)), ptr(sx), None)
    dweight = lib.zeros(tuple(weight.shape), torch.float32, dev)
    dgamma = lib.zeros(cout, torch.float32, dev) if has_g else None
    dbeta = lib.zeros(cout, torch.float32, dev) if has_b else None
    pack_g = torch.empty(cin * ((cout + 7) // 8 * 8), dtype=torch.bfloat16, device=dev)
    pack_x = torch.empty(cin * ((cin + 7) // 8 * 8), dtype=torch.bfloat16, device=dev)
    bias_row = torch.empty(cin + 3 * cout, dtype=torch.float32, device=dev)  # + the kernel's [3][cout] coefficient scratch
    call("nasb_pw_bn_bwd_prepare", ptr(weight), cout, cin, ptr(ss[0]), ptr(sv[0]), ptr(sv[1]), ptr(sums), C.c_longlong(n * h * w),
         ptr(gx), ptr(xx), ptr(sx), ptr(dweight), ptr(dgamma), ptr(dbeta), ptr(pack_g), ptr(pack_x), ptr(bias_row),
         bias_row.data_ptr() + 4 * cin)
    dx = None
    if need_dx:
        tmp = lib.new_act(*x0.shape, x0.dtype, dev)   # x.(W^T diag(A) W) + B^T W : the z terms of dz, through x
        call("nasb_pw_tc_fwd", ref(desc(x0)), ptr(pack_x), cin, None, ptr(bias_row), ACT_NONE, None, ref(desc(tmp)), None)
        dx = lib.new_act(*x0.shape, x0.dtype, dev)
        call("nasb_pw_tc_fwd", ref(desc(g)), ptr(pack_g), cin, None, None, ACT_NONE, ref(desc(tmp)), ref(desc(dx)), None)
    return dx, dweight, dgamma, dbeta


class _ConvUnit(torch.autograd.Function):
    """conv (dense 1x1/3x3 or depthwise kxk) [+ BatchNorm2d train/eval] [+ bias] [+ ReLU/ReLU6] [+ residual].

    One fused unit of the reference graph: nn.Conv2d -> nn.BatchNorm2d -> nn.ReLU(6) chains
    (layer_factory.py:56-75,94-114,125-158,225-265; micro_decoders.py adapt/pre_clf/conv_clf)."""

    @staticmethod
    def forward(ctx, x0, x1, weight, gamma, beta, bias, res, bn, cfg):
        lib.require_cuda(x0)
        dev = x0.device
        ks, stride, dil, pad = cfg["ks"], cfg["stride"], cfg["dil"], cfg["pad"]
        act, dw, in_relu, image = cfg["act"], cfg.get("dw", False), cfg.get("in_relu", 0), cfg.get("image", False)
        n, _, h, w = x0.shape
        cout = weight.shape[0]
        oh, ow = conv_out_hw(h, w, ks, stride, dil, pad)
        out_dtype = cfg.get("out_dtype") or (act_dtype() if image else x0.dtype)
        training = bn is not None and bn.training
        dx0 = lib.desc_nchw_f32(x0) if image else desc(x0)
        dx1 = desc(x1) if x1 is not None else None
        if bn is not None and bn.running_mean is None:
            raise RuntimeError("BatchNorm2d without running statistics is not supported")

        use_tc = (not dw and ks == 1 and stride == 1 and pad == 0 and x1 is None and not in_relu and not image
                  and _tc_ok(x0, x0.shape[1], cout, out_dtype))
        wpack = _pack_weight(weight, False) if use_tc else None
        # speed mode stem: image -> bf16 patch matrix [n, 32, oh, ow] (k = ci*9 + ky*3 + kx), then the pointwise tensor-core
        # kernel with K = 32; the patch matrix is what the backward pass keeps (weight gradient = nasb_pw_tc_wgrad)
        stem_tc = (image and not dw and ks == 3 and x1 is None and not in_relu and res is None and x0.shape[1] == 3
                   and out_dtype == torch.bfloat16 and not x0.requires_grad and _stem_tc_ok(cout))
        if stem_tc:
            xp = lib.new_act(n, 32, oh, ow, torch.bfloat16, dev)
            call("nasb_stem_im2col", ref(dx0), ks, stride, dil, pad, ref(desc(xp)))
            wpack = packs.get(weight, packs.PW)
            if wpack is None:
                wpack = torch.empty(cout * 32, dtype=torch.bfloat16, device=dev)
                call("nasb_pack_weight_bf16", ptr(weight), cout, 27, 0, ptr(wpack))
            x0, dx0, use_tc = xp, desc(xp), True
        use_c3 = (not dw and x1 is None and not in_relu and not image and res is None
                  and out_dtype in (torch.bfloat16, torch.float32) and _c3_ok(x0, x0.shape[1], cout, ks, stride, dil, pad))
        wpack3 = _pack_conv3(weight, 0) if use_c3 else None

        def run_conv(out, scale, shift, a, r, stats=None):
            if use_c3:
                call("nasb_conv3_tc_fwd", ref(dx0), ptr(wpack3), cout, dil, pad, ptr(scale), ptr(shift), a, ref(desc(out)),
                     ptr(stats))
                return
            if use_tc:
                call("nasb_pw_tc_fwd", ref(dx0), ptr(wpack), cout, ptr(scale), ptr(shift), a,
                     ref(desc(r)) if r is not None else None, ref(desc(out)), ptr(stats))
                return
            if dw:
                assert r is None
                if not in_relu and x0.dtype == torch.bfloat16 and _tiles_on() and try_call(
                        "nasb_dwconv_tile", ref(dx0), ptr(weight), ks, stride, dil, pad, 0, ptr(scale), ptr(shift), a,
                        ref(desc(out)), ptr(stats)):
                    return True
                call("nasb_dwconv_fwd", ref(dx0), ptr(weight), ks, stride, dil, pad, in_relu, ptr(scale), ptr(shift), a,
                     ref(desc(out)))
            else:
                if image and r is None and try_call("nasb_stem_fwd", ref(dx0), ptr(weight), ks, stride, dil, pad, ptr(scale),
                                                    ptr(shift), a, ref(desc(out))):
                    return
                call("nasb_conv_fwd", ref(dx0), ref(dx1), ptr(weight), ks, stride, dil, pad, None, None, in_relu,
                     ptr(scale), ptr(shift), a, ref(desc(r)) if r is not None else None, ref(desc(out)))

        z = ss = sv = None
        res_done = False
        y = lib.new_act(n, cout, oh, ow, out_dtype, dev)
        # The residual add is fused into the conv epilogue only when no gradient is needed: the backward pass
        # rebuilds xhat / the activation mask from the unit's own output, which must then exclude the residual.
        late_res = res is not None and any(ctx.needs_input_grad)
        fused_res = None if late_res else res
        if bn is None:
            run_conv(y, None, bias, act, fused_res)
        elif not training:
            if torch.is_grad_enabled():
                ss = torch.empty((2, cout), dtype=torch.float32, device=dev)
            else:  # inference: the folded constants live in one buffer per BN module, rewritten (in stream order) per call
                ss = bn.__dict__.get("_nasb_ss")
                if ss is None or ss.device != dev:
                    ss = torch.empty((2, cout), dtype=torch.float32, device=dev)
                    object.__setattr__(bn, "_nasb_ss", ss)
            call("nasb_bn_fold", ptr(gamma), ptr(beta), ptr(bn.running_mean), ptr(bn.running_var), float(bn.eps), cout,
                 ptr(ss[0]), ptr(ss[1]))
            run_conv(y, ss[0], ss[1], act, fused_res)
        else:
            if n * oh * ow == 1:  # same guard (and exception type) as torch.nn.functional.batch_norm
                raise ValueError("Expected more than 1 value per channel when training, got input size {}".format(
                    [n, cout, oh, ow]))
            z = lib.new_act(n, cout, oh, ow, out_dtype, dev)
            ss = torch.empty((2, cout), dtype=torch.float32, device=dev)
            sv = torch.empty((2, cout), dtype=torch.float32, device=dev)
            mom = _bn_momentum(bn)
            fused_stats = use_tc or use_c3
            sums = None
            if fused_stats or (dw and not in_relu and x0.dtype == torch.bfloat16 and _tiles_on() and _dw_stats_on()):
                sums = lib.zeros(2 * cout, torch.float64, dev)
                fused_stats = bool(run_conv(z, None, None, ACT_NONE, None, sums)) or fused_stats
            applied = False
            if fused_stats:  # batch statistics were accumulated by the conv kernel's epilogue: finalise + apply in one launch
                # (+ the residual: training-mode backward rebuilds everything from z, so y itself is never needed)
                call("nasb_bn_finalize_affine_act", ptr(sums), C.c_longlong(n * oh * ow), ref(desc(z)), ptr(gamma), ptr(beta),
                     float(bn.eps), mom, ptr(bn.running_mean), ptr(bn.running_var), ptr(sv[0]), ptr(sv[1]), ptr(ss[0]),
                     ptr(ss[1]), ptr(bn.num_batches_tracked), act, ref(desc(res)) if res is not None else None, ref(desc(y)))
                applied = True
                res_done = res is not None
            else:
                if sums is None:
                    run_conv(z, None, None, ACT_NONE, None)
                call("nasb_bn_stats", ref(desc(z)), ptr(gamma), ptr(beta), float(bn.eps), mom, ptr(bn.running_mean),
                     ptr(bn.running_var), ptr(sv[0]), ptr(sv[1]), ptr(ss[0]), ptr(ss[1]), ptr(bn.num_batches_tracked),
                     ptr(_ws(dev, cout)))
            if not applied:
                call("nasb_affine_act", ref(desc(z)), ptr(ss[0]), ptr(ss[1]), act, ref(desc(y)))
            if res is not None and not late_res and not res_done:
                call("nasb_resize_axpby", ref(desc(y)), None, ref(desc(res)), None, 0, ref(desc(y)))
        ctx.stem_tc = stem_tc
        ctx.cfg, ctx.bn_mode = cfg, (0 if bn is None else (2 if training else 1))
        ctx.has = (x1 is not None, gamma is not None, beta is not None, bias is not None, res is not None)
        ctx.save_for_backward(x0, x1, weight, gamma, beta, y, z, ss, sv)
        # BN-backward fusion links (see _BnRec): what this unit's input producer left on x0, what this unit leaves on y
        fuse = _cfg().fuse_bn_bwd and out_dtype == torch.bfloat16
        ctx.prod = getattr(x0, "_nasb_rec", None) if (fuse and cfg.get("sole") and any(ctx.needs_input_grad)) else None
        ctx.rec = None
        if late_res and not res_done:
            out = lib.new_act(n, cout, oh, ow, out_dtype, dev)
            call("nasb_resize_axpby", ref(desc(y)), None, ref(desc(res)), None, 0, ref(desc(out)))
            return out
        if fuse and training and res is None and any(ctx.needs_input_grad):
            ctx.rec = y._nasb_rec = _BnRec(z, ss, act)
        return y

    @staticmethod
    def backward(ctx, dy):
        x0, x1, weight, gamma, beta, y, z, ss, sv = ctx.saved_tensors
        cfg = ctx.cfg
        ks, stride, dil, pad = cfg["ks"], cfg["stride"], cfg["dil"], cfg["pad"]
        act, dw, in_relu, image = cfg["act"], cfg.get("dw", False), cfg.get("in_relu", 0), cfg.get("image", False)
        has_x1, has_g, has_b, has_bias, has_res = ctx.has
        dev = y.device
        cout = weight.shape[0]
        dy = _grad_in(dy, y.dtype)
        need = ctx.needs_input_grad
        rec, gsums = ctx.rec, None
        if rec is not None and rec.g is not None:
            g, sums = rec.g, rec.sums
            rec.g = rec.sums = rec.z = None  # one-shot; drop the references
            # the consumer's gated gradient IS the incoming gradient (nothing else was accumulated into it)
            if g.data_ptr() == dy.data_ptr() and tuple(g.shape) == tuple(dy.shape) and ctx.bn_mode == 2:
                cin = x0.shape[1]
                if (not dw and ks == 1 and stride == 1 and pad == 0 and not has_x1 and not in_relu and not image and not has_bias
                        and not has_res and not ctx.stem_tc and 2 * cin <= cout and cin <= 64 and _tc_wgrad_ok(x0, dy, cout)
                        and lib.load().nasb_pw_tc_wgrad_supported(int(cin), int(cin))
                        and _tc_ok(dy, cout, cin, x0.dtype) and _tc_ok(x0, cin, cin, x0.dtype)):
                    # pointwise unit with a narrow input (MobileNet-v2 expansion): no dz at all
                    dx0, dweight, dgamma, dbeta = _pw_unit_bwd_nodz(dy, x0, weight, has_g, has_b, ss, sv, sums, need[0])
                    return dx0, None, dweight if need[2] else None, dgamma, dbeta, None, None, None, None
                gsums = sums  # any other unit: the reductions exist, only the dz pass remains
        dgamma = lib.zeros(cout, torch.float32, dev) if has_g else None
        dbeta = lib.zeros(cout, torch.float32, dev) if has_b else None
        if ctx.bn_mode or act != ACT_NONE:
            dz = lib.new_act(*y.shape, y.dtype, dev)
            if not (gsums is not None and try_call(
                    "nasb_bn_bwd_from_sums", ref(desc(dy)), ref(desc(z)), act, ptr(ss[0]), ptr(ss[1]), ptr(sv[0]), ptr(sv[1]),
                    ptr(gsums), ptr(dgamma), ptr(dbeta), ref(desc(dz)), ptr(_ws(dev, cout)))):
                # training: the mask is recomputed from z, y is neither read nor passed
                call("nasb_bn_act_bwd", ref(desc(dy)), ref(desc(y)) if z is None else None,
                     ref(desc(z)) if z is not None else None, act, ptr(gamma), ptr(beta),
                     ptr(ss[0]) if ss is not None else None, ptr(ss[1]) if ss is not None else None,
                     ptr(sv[0]) if sv is not None else None, ptr(sv[1]) if sv is not None else None,
                     1 if ctx.bn_mode == 2 else 0, ptr(dgamma), ptr(dbeta), ref(desc(dz)), ptr(_ws(dev, cout)))
        else:
            dz = dy
        dbias = None
        if has_bias:
            dbias = lib.zeros(cout, torch.float32, dev)
            call("nasb_channel_sum", ref(desc(dz)), ptr(dbias), None)
        dweight = None
        c3 = (not dw and not has_x1 and not in_relu and not image
              and _c3_ok(x0, x0.shape[1], cout, ks, stride, dil, pad))
        if c3 and (dz.dtype != torch.bfloat16 or dz.data_ptr() % 16 or desc(dz).cstride % 8):
            dz = _bf16_padded_copy(dz)  # e.g. fp32 logit gradients with 19 channels
        ddz = desc(dz)
        if need[2] and ctx.stem_tc:  # x0 is the saved bf16 patch matrix; columns 27..31 of the product are zero padding
            dw32 = lib.zeros((cout, 32), torch.float32, dev)
            dweight = torch.empty(tuple(weight.shape), dtype=torch.float32, device=dev)
            with (lib.wgrad_stream.fork((x0, dz)) or contextlib.nullcontext()):
                call("nasb_pw_tc_wgrad", ref(desc(x0)), ref(ddz), ptr(dw32))
                dweight.view(cout, 27).copy_(dw32[:, :27])  # on the stream that produced dw32
        elif need[2] and c3:
            dweight = lib.zeros(tuple(weight.shape), torch.float32, dev)
            with (lib.wgrad_stream.fork((x0, dz)) or contextlib.nullcontext()):
                call("nasb_conv3_tc_wgrad", ref(desc(x0)), ref(ddz), dil, pad, ptr(dweight))
        elif need[2]:
            dweight = lib.zeros(tuple(weight.shape), torch.float32, dev)
            # the weight gradient is not needed before the optimiser step: inside an engine iteration it runs on the side
            # stream (lib._WgradStream) while this stream goes on with the data gradient
            with (lib.wgrad_stream.fork((x0, x1, dz)) or contextlib.nullcontext()):
                if (not dw and ks == 1 and stride == 1 and pad == 0 and not has_x1 and not image and not in_relu
                        and _tc_wgrad_ok(x0, dz, cout)):
                    call("nasb_pw_tc_wgrad", ref(desc(x0)), ref(ddz), ptr(dweight))
                elif dw:
                    if not (not in_relu and x0.dtype == torch.bfloat16 and _tiles_on() and try_call(
                            "nasb_dwconv_wgrad_tile", ref(desc(x0)), ref(ddz), ks, stride, dil, pad, ptr(dweight))):
                        call("nasb_dwconv_wgrad", ref(desc(x0)), in_relu, ref(ddz), ks, stride, dil, pad, ptr(dweight))
                elif image and try_call("nasb_stem_wgrad", ref(lib.desc_nchw_f32(x0)), ref(ddz), ks, stride, dil, pad,
                                        ptr(dweight)):
                    pass
                else:
                    dsrc0 = lib.desc_nchw_f32(x0) if image else desc(x0)
                    call("nasb_conv_wgrad", ref(dsrc0), ref(desc(x1)) if has_x1 else None, None, None, in_relu, ref(ddz), ks,
                         stride, dil, pad, ptr(dweight))
        dx0 = dx1 = None
        if (need[0] or (has_x1 and need[1])) and not ctx.stem_tc:
            dx0 = lib.new_act(*x0.shape, x0.dtype, dev)
            if has_x1:
                dx1 = lib.new_act(*x1.shape, x1.dtype, dev)
            if c3:
                call("nasb_conv3_tc_fwd", ref(ddz), ptr(_pack_conv3(weight, 1)), x0.shape[1], dil, 2 * dil - pad, None, None,
                     ACT_NONE, ref(desc(dx0)), None)
            elif (not dw and ks == 1 and stride == 1 and pad == 0 and not has_x1 and not image
                    and _tc_ok(dz, cout, x0.shape[1], x0.dtype)):
                prod, gated = ctx.prod, False
                if prod is not None and prod.z is not None:  # sole consumer of a BN(train) unit: see the depthwise branch
                    sums = lib.zeros(2 * x0.shape[1], torch.float64, dev)
                    gate = lib.NasbGate(C.pointer(desc(prod.z)), ptr(prod.ss[0]), ptr(prod.ss[1]), prod.act, 0, ptr(sums))
                    gated = try_call("nasb_pw_tc_dgrad_gated", ref(ddz), ptr(_pack_weight(weight, True)), x0.shape[1],
                                     C.byref(gate), ref(desc(dx0)))
                    if gated:
                        prod.g, prod.sums = dx0, sums
                if not gated:
                    call("nasb_pw_tc_fwd", ref(ddz), ptr(_pack_weight(weight, True)), x0.shape[1], None, None, ACT_NONE, None,
                         ref(desc(dx0)), None)
            elif dw:
                tiled = False
                prod = ctx.prod
                if prod is not None and prod.z is not None and dz.dtype == torch.bfloat16 and _tiles_on():
                    # sole consumer of a conv -> BN(train) -> act unit: gate dx with its mask and leave its reductions behind
                    sums = lib.zeros(2 * x0.shape[1], torch.float64, dev)
                    gate = lib.NasbGate(C.pointer(desc(prod.z)), ptr(prod.ss[0]), ptr(prod.ss[1]), prod.act, 0, ptr(sums))
                    tiled = try_call("nasb_dwconv_dgrad_gated", ref(ddz), ptr(weight), ks, stride, dil, pad, C.byref(gate),
                                     ref(desc(dx0)))
                    if tiled:
                        prod.g, prod.sums = dx0, sums
                if not tiled and dz.dtype == torch.bfloat16 and _tiles_on():
                    if stride == 1:
                        tiled = try_call("nasb_dwconv_tile", ref(ddz), ptr(weight), ks, stride, dil, pad, 1, None, None, ACT_NONE,
                                         ref(desc(dx0)), None)
                    else:
                        tiled = try_call("nasb_dwconv_dgrad_strided_tile", ref(ddz), ptr(weight), ks, stride, dil, pad,
                                         ref(desc(dx0)))
                if not tiled:
                    call("nasb_dwconv_dgrad", ref(ddz), ptr(weight), ks, stride, dil, pad, ref(desc(dx0)))
            else:
                call("nasb_conv_dgrad", ref(ddz), ptr(weight), ks, stride, dil, pad, ref(desc(dx0)),
                     ref(desc(dx1)) if has_x1 else None)
            if in_relu:
                call("nasb_relu_bwd", ref(desc(dx0)), ref(desc(x0)), ref(desc(dx0)))
                if has_x1:
                    call("nasb_relu_bwd", ref(desc(dx1)), ref(desc(x1)), ref(desc(dx1)))
        dres = dy if has_res else None
        return dx0, dx1, dweight, dgamma, dbeta, dbias, dres, None, None


def _unit_state(cin, weight, bn, ks, stride, dil, pad, act, dw, in_relu, bias):
    """The NasbConvUnit parameter block + scratch buffer of an inference unit; they live on the weight tensor and are rebuilt
    when any address or setting changes.  -> (struct, scratch, settings, scratch bytes, weight)"""
    cout = weight.shape[0]
    if bn is not None and bn.running_mean is None:
        raise RuntimeError("BatchNorm2d without running statistics is not supported")
    settings = (ks, stride, dil, pad, act, dw, in_relu, id(bn))
    st = weight.__dict__.get("_nasb_unit")
    wp = weight.data_ptr()
    if (st is None or st[2] != settings or st[0].weight != wp or st[0].bias != (bias.data_ptr() if bias is not None else None)
            or (bn is not None and st[0].running_mean != bn.running_mean.data_ptr())):
        u = lib.NasbConvUnit(wp, ptr(bn.weight) if bn is not None else None, ptr(bn.bias) if bn is not None else None,
                             ptr(bn.running_mean) if bn is not None else None, ptr(bn.running_var) if bn is not None else None,
                             ptr(bias), float(bn.eps) if bn is not None else 0.0, cout, ks, stride, dil, pad, int(bool(dw)),
                             int(in_relu), act)
        nbytes = int(lib.load().nasb_conv_unit_scratch(cout, int(cin)))
        st = (u, torch.empty(nbytes, dtype=torch.uint8, device=weight.device), settings, nbytes, weight)
        weight._nasb_unit = st
    return st


def _conv_unit_infer(x0, weight, bn, ks, stride, dil, pad, act, bias, res, dw, in_relu, out_dtype):
    """Inference path of a conv unit: BN fold + operand pack + kernel choice behind ONE C-ABI call (csrc/unit.cu)."""
    lib.require_cuda(x0)
    n, cin, h, w = x0.shape
    cout = weight.shape[0]
    st = _unit_state(cin, weight, bn, ks, stride, dil, pad, act, dw, in_relu, bias)
    oh, ow = conv_out_hw(h, w, ks, stride, dil, pad)
    y = lib.new_act(n, cout, oh, ow, out_dtype or x0.dtype, x0.device)
    cfg = _cfg()
    flags = (1 if cfg.use_tcgen05 else 0) | (2 if cfg.use_tma_tiles else 0)
    if packs.infer_prepared(st, int(cin)):  # a scope (validate(), GraphedForward) folded / packed every unit at once
        flags |= 4
    call("nasb_conv_unit_infer", ref(desc(x0)), C.byref(st[0]), ref(desc(res)) if res is not None else None, ref(desc(y)),
         st[1].data_ptr(), st[3], flags)
    return y


def sep_unit(x, dw_weight, dw_bn, dw_act, pw_weight, pw_bn, pw_act, *, ks, stride, dil, pad, res=None, sole=(False, False)):
    """Depthwise conv unit followed by a pointwise conv unit (a SepConv repeat, layer_factory.py:241-256; the depthwise +
    projection half of InvertedResidual, :141-158).  At inference, when the shapes allow, ONE kernel (csrc/sep_tcgen05.cu):
    the depthwise output goes from registers into the tensor core's shared-memory operand and never touches HBM.  Otherwise
    (training, or shapes outside the fused kernel) the two units run one after the other.  `sole`: see conv_unit."""
    cfg = _cfg()
    if (cfg.fused_sep and cfg.use_tcgen05 and not torch.is_grad_enabled() and x.dtype == torch.bfloat16
            and (dw_bn is None or not dw_bn.training) and (pw_bn is None or not pw_bn.training)
            and lib.load().nasb_sepconv_tc_supported(int(x.shape[1]), int(pw_weight.shape[0]), ks, stride, dil, pad)):
        lib.require_cuda(x)
        n, c, h, w = x.shape
        st_dw = _unit_state(c, dw_weight, dw_bn, ks, stride, dil, pad, dw_act, True, 0, None)
        st_pw = _unit_state(c, pw_weight, pw_bn, 1, 1, 1, 0, pw_act, False, 0, None)
        y = lib.new_act(n, pw_weight.shape[0], h, w, x.dtype, x.device)
        flags = 1 | (2 if cfg.use_tma_tiles else 0)
        if packs.infer_prepared(st_dw, int(c)) & packs.infer_prepared(st_pw, int(c)):
            flags |= 4
        if try_call("nasb_sep_unit_infer", ref(desc(x)), C.byref(st_dw[0]), st_dw[1].data_ptr(), st_dw[3], C.byref(st_pw[0]),
                    st_pw[1].data_ptr(), st_pw[3], ref(desc(res)) if res is not None else None, ref(desc(y)), flags):
            return y
    t = conv_unit(x, dw_weight, dw_bn, ks=ks, stride=stride, dil=dil, pad=pad, act=dw_act, dw=True, sole=sole[0])
    return conv_unit(t, pw_weight, pw_bn, ks=1, act=pw_act, res=res, sole=sole[1])


def conv_unit(x0, weight, bn=None, *, ks, stride=1, dil=1, pad=0, act=ACT_NONE, x1=None, bias=None, res=None, dw=False,
              in_relu=0, image=False, out_dtype=None, sole=False):
    """Fused conv unit.  ``bn`` is the nn.BatchNorm2d module holding gamma/beta/running stats (or None).  ``sole``: the
    caller guarantees that this unit is the ONLY consumer of x0 (enables the BN-backward fusion link, see _BnRec)."""
    if x1 is None and not image and not torch.is_grad_enabled() and (bn is None or not bn.training):
        return _conv_unit_infer(x0, weight, bn, ks, stride, dil, pad, act, bias, res, dw, in_relu, out_dtype)
    cfg = dict(ks=ks, stride=stride, dil=dil, pad=pad, act=act, dw=dw, in_relu=in_relu, image=image, out_dtype=out_dtype,
               sole=sole)
    gamma = bn.weight if bn is not None else None
    beta = bn.bias if bn is not None else None
    return _apply(_ConvUnit, x0, x1, weight, gamma, beta, bias, res, bn, cfg)


class _BnAct(torch.autograd.Function):
    """Stand-alone BatchNorm2d (+ReLU): the head of ConcatReduce (layer_factory.py:372-376)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, bn, act):
        lib.require_cuda(x)
        dev, c = x.device, x.shape[1]
        training = bn.training
        ss = torch.empty((2, c), dtype=torch.float32, device=dev)
        sv = None
        if training:
            if x.shape[0] * x.shape[2] * x.shape[3] == 1:
                raise ValueError("Expected more than 1 value per channel when training, got input size {}".format(
                    list(x.shape)))
            sv = torch.empty((2, c), dtype=torch.float32, device=dev)
            mom = _bn_momentum(bn)
            call("nasb_bn_stats", ref(desc(x)), ptr(gamma), ptr(beta), float(bn.eps), mom, ptr(bn.running_mean),
                 ptr(bn.running_var), ptr(sv[0]), ptr(sv[1]), ptr(ss[0]), ptr(ss[1]), ptr(bn.num_batches_tracked),
                 ptr(_ws(dev, c)))
        else:
            call("nasb_bn_fold", ptr(gamma), ptr(beta), ptr(bn.running_mean), ptr(bn.running_var), float(bn.eps), c,
                 ptr(ss[0]), ptr(ss[1]))
        y = lib.new_act(*x.shape, x.dtype, dev)
        call("nasb_affine_act", ref(desc(x)), ptr(ss[0]), ptr(ss[1]), act, ref(desc(y)))
        ctx.act, ctx.training = act, training
        ctx.has = (gamma is not None, beta is not None)
        ctx.save_for_backward(x, gamma, beta, y, ss, sv)
        ctx.rec = None
        if training and _cfg().fuse_bn_bwd and x.dtype == torch.bfloat16 and ctx.needs_input_grad[0]:
            ctx.rec = y._nasb_rec = _BnRec(x, ss, act)  # the BN input plays the part of z (see _BnRec)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, y, ss, sv = ctx.saved_tensors
        dev, c = y.device, y.shape[1]
        dy = _grad_in(dy, y.dtype)
        dgamma = lib.zeros(c, torch.float32, dev) if ctx.has[0] else None
        dbeta = lib.zeros(c, torch.float32, dev) if ctx.has[1] else None
        dx = lib.new_act(*y.shape, y.dtype, dev)
        rec = ctx.rec
        if rec is not None and rec.g is not None:
            g, sums = rec.g, rec.sums
            rec.g = rec.sums = rec.z = None
            if g.data_ptr() == dy.data_ptr() and tuple(g.shape) == tuple(dy.shape) and try_call(
                    "nasb_bn_bwd_from_sums", ref(desc(dy)), ref(desc(x)), ctx.act, ptr(ss[0]), ptr(ss[1]), ptr(sv[0]), ptr(sv[1]),
                    ptr(sums), ptr(dgamma), ptr(dbeta), ref(desc(dx)), ptr(_ws(dev, c))):
                return dx, dgamma, dbeta, None, None
        call("nasb_bn_act_bwd", ref(desc(dy)), None if ctx.training else ref(desc(y)), ref(desc(x)) if ctx.training else None,
             ctx.act, ptr(gamma),
             ptr(beta), ptr(ss[0]), ptr(ss[1]), ptr(sv[0]) if sv is not None else None, ptr(sv[1]) if sv is not None else None,
             1 if ctx.training else 0, ptr(dgamma), ptr(dbeta), ref(desc(dx)), ptr(_ws(dev, c)))
        return dx, dgamma, dbeta, None, None


def bn_act(x, bn, act=ACT_RELU):
    return _apply(_BnAct, x, bn.weight, bn.bias, bn, act)


class _Pool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mode, stride):
        lib.require_cuda(x)
        n, c, h, w = x.shape
        oh, ow = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
        y = lib.new_act(n, c, oh, ow, x.dtype, x.device)
        arg = None
        if mode == POOL_MAX and ctx.needs_input_grad[0]:
            arg = torch.empty((n, oh, ow, c), dtype=torch.uint8, device=x.device)
        call("nasb_pool3x3_fwd", ref(desc(x)), mode, stride, ref(desc(y)), ptr(arg))
        ctx.mode, ctx.stride, ctx.xshape = mode, stride, tuple(x.shape)
        ctx.save_for_backward(arg)
        return y

    @staticmethod
    def backward(ctx, dy):
        (arg,) = ctx.saved_tensors
        dy = _grad_in(dy, dy.dtype)
        dx = lib.new_act(*ctx.xshape, dy.dtype, dy.device)
        call("nasb_pool3x3_bwd", ref(desc(dy)), ctx.mode, ctx.stride, ptr(arg), ref(desc(dx)))
        return dx, None, None


def pool3x3(x, mode, stride):
    return _apply(_Pool, x, mode, stride)


class _ResizeAxpby(torch.autograd.Function):
    """out = sa*resize(x) + sb*y (bilinear, align_corners=False; identity when sizes agree)."""

    @staticmethod
    def forward(ctx, x, y, sa, sb, size):
        lib.require_cuda(x)
        n, c = x.shape[:2]
        oh, ow = size
        out = lib.new_act(n, c, oh, ow, x.dtype, x.device)
        call("nasb_resize_axpby", ref(desc(x)), ptr(sa), ref(desc(y)) if y is not None else None, ptr(sb), 0, ref(desc(out)))
        ctx.save_for_backward(x, y, sa, sb)
        return out

    @staticmethod
    def backward(ctx, dz):
        x, y, sa, sb = ctx.saved_tensors
        dev = x.device
        dz = _grad_in(dz, x.dtype)
        need = ctx.needs_input_grad
        dx = dyy = dsa = dsb = None
        if need[0]:
            if sa is None and tuple(dz.shape) == tuple(x.shape):
                dx = dz  # same size, no per-channel scale: the adjoint of the identity (gradients are values, aliasing is safe)
            else:
                dx = lib.new_act(*x.shape, x.dtype, dev)
                call("nasb_resize_bwd", ref(desc(dz)), ptr(sa), ref(desc(dx)))
        if y is not None and need[1]:
            if sb is None:
                dyy = dz
            else:
                dyy = lib.new_act(*y.shape, y.dtype, dev)
                call("nasb_scale_copy", ref(desc(dz)), ptr(sb), 0, ref(desc(dyy)))
        if (sa is not None and need[2]) or (sb is not None and need[3]):
            c = x.shape[1]
            dsa = lib.zeros(c, torch.float32, dev) if sa is not None else None
            dsb = lib.zeros(c, torch.float32, dev) if sb is not None else None
            call("nasb_axpby_bwd_params", ref(desc(dz)), ref(desc(x)), ref(desc(y)) if y is not None else None, ptr(dsa),
                 ptr(dsb), None)
        return dx, dyy, dsa, dsb, None


def resize(x, size):
    size = (int(size[0]), int(size[1]))
    if tuple(x.shape[2:]) == size:
        return x
    return _apply(_ResizeAxpby, x, None, None, None, size)


def resize_add(x, y, sa=None, sb=None):
    """sa*resize(x -> y's size) + sb*y."""
    return _apply(_ResizeAxpby, x, y, sa, sb, (y.shape[2], y.shape[3]))


class _ConcatResize(torch.autograd.Function):
    """[relu](torch.cat([resize(t, size) for t in tensors], 1)): every producer-side resize writes straight into its channel
    slice of the concat buffer (collect_all, micro_decoders.py:11-25; ConcatReduce's cat, layer_factory.py:380); the
    optional ReLU is the F.relu the decoders apply to the collected map (micro_decoders.py:251,395)."""

    @staticmethod
    def forward(ctx, size, relu, *tensors):
        lib.require_cuda(tensors[0])
        n = tensors[0].shape[0]
        ctot = sum(t.shape[1] for t in tensors)
        out = lib.new_act(n, ctot, size[0], size[1], tensors[0].dtype, tensors[0].device)
        c0 = 0
        for t in tensors:
            c = t.shape[1]
            call("nasb_resize_axpby", ref(desc(t)), None, None, None, 1 if relu else 0, ref(desc(out[:, c0:c0 + c])))
            c0 += c
        ctx.shapes = [tuple(t.shape) for t in tensors]
        ctx.relu = relu
        if relu:
            ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _grad_in(dout, dout.dtype)
        if ctx.relu:
            (out,) = ctx.saved_tensors
            masked = lib.new_act(*out.shape, dout.dtype, dout.device)
            call("nasb_relu_bwd", ref(desc(dout)), ref(desc(out)), ref(desc(masked)))
            dout = masked
        grads, c0 = [], 0
        for i, shp in enumerate(ctx.shapes):
            c = shp[1]
            if ctx.needs_input_grad[i + 2]:
                dx = lib.new_act(*shp, dout.dtype, dout.device)
                call("nasb_resize_bwd", ref(desc(dout[:, c0:c0 + c])), None, ref(desc(dx)))
                grads.append(dx)
            else:
                grads.append(None)
            c0 += c
        return (None, None, *grads)


def concat_resize(tensors, size, relu=False):
    return _apply(_ConcatResize, (int(size[0]), int(size[1])), bool(relu), *tensors)


class _ChannelTile(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, reps, stride, scale):
        lib.require_cuda(x)
        n, c, h, w = x.shape
        oh, ow = (h + stride - 1) // stride, (w + stride - 1) // stride
        out = lib.new_act(n, c * reps, oh, ow, x.dtype, x.device)
        call("nasb_channel_tile", ref(desc(x)), stride, float(scale), ref(desc(out)))
        ctx.args = (reps, stride, scale, tuple(x.shape))
        return out

    @staticmethod
    def backward(ctx, dz):
        reps, stride, scale, xshape = ctx.args
        dz = _grad_in(dz, dz.dtype)
        dx = lib.new_act(*xshape, dz.dtype, dz.device)
        call("nasb_channel_tile_bwd", ref(desc(dz)), stride, float(scale), ref(desc(dx)))
        return dx, None, None, None


def channel_tile(x, reps, stride=1, scale=1.0):
    return _apply(_ChannelTile, x, reps, stride, scale)


class _SpatialMean(torch.autograd.Function):
    """x.mean(2).mean(3) -> fp32 [N,C,1,1] (GAPConv1x1, layer_factory.py:189)."""

    @staticmethod
    def forward(ctx, x):
        lib.require_cuda(x)
        n, c, h, w = x.shape
        out = lib.new_act(n, c, 1, 1, torch.float32, x.device)
        call("nasb_spatial_mean", ref(desc(x)), ptr(out))
        ctx.xshape, ctx.xdtype = tuple(x.shape), x.dtype
        return out

    @staticmethod
    def backward(ctx, dv):
        dv = lib.to_nhwc(dv, torch.float32)
        n, c, h, w = ctx.xshape
        dx = lib.new_act(n, c, h, w, ctx.xdtype, dv.device)
        call("nasb_spatial_bcast", ref(desc(dv)), 1.0 / float(h * w), ref(desc(dx)))
        return dx


class _SpatialBcast(torch.autograd.Function):
    """Bilinear interpolation of a 1x1 map to HxW == broadcast (layer_factory.py:191-194)."""

    @staticmethod
    def forward(ctx, v, h, w, dtype):
        lib.require_cuda(v)
        n, c = v.shape[:2]
        out = lib.new_act(n, c, h, w, dtype, v.device)
        call("nasb_spatial_bcast", ref(desc(v)), 1.0, ref(desc(out)))
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _grad_in(dout, dout.dtype)
        n, c = dout.shape[:2]
        dv = lib.new_act(n, c, 1, 1, torch.float32, dout.device)
        call("nasb_spatial_sum", ref(desc(dout)), ptr(dv))
        return dv, None, None, None


def spatial_mean(x):
    return _apply(_SpatialMean, x)


def spatial_bcast(v, h, w, dtype):
    return _apply(_SpatialBcast, v, h, w, dtype)


class _Cast(torch.autograd.Function):
    """dtype conversion at the fp32 <-> bf16 boundary of the network (nasb_scale_copy)."""

    @staticmethod
    def forward(ctx, x, dtype):
        out = lib.new_act(*x.shape, dtype, x.device)
        call("nasb_scale_copy", ref(desc(x)), None, 0, ref(desc(out)))
        ctx.src = x.dtype
        return out

    @staticmethod
    def backward(ctx, dy):
        dy = _grad_in(dy, dy.dtype)
        dx = lib.new_act(*dy.shape, ctx.src, dy.device)
        call("nasb_scale_copy", ref(desc(dy)), None, 0, ref(desc(dx)))
        return dx, None


def as_act(x, dtype=None):
    """User tensor -> NHWC storage in the activation dtype."""
    dtype = dtype or act_dtype()
    lib.require_cuda(x)
    if x.dtype == dtype:
        return lib.to_nhwc(x)
    x = lib.to_nhwc(x, None if x.dtype in (torch.float32, torch.bfloat16) else torch.float32)
    return _apply(_Cast, x, dtype)


# ------------------------------------------------------------------------------------------------------- losses
class _CrossEntropy2d(torch.autograd.Function):
    """nn.NLLLoss2d(ignore_index)(nn.LogSoftmax()(logits), target), trainer.py:144-146 / main_search.py:435."""

    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        lib.require_cuda(logits)
        if target.dtype != torch.int64:
            target = target.long()
        target = target.contiguous()
        assert tuple(target.shape) == (logits.shape[0], logits.shape[2], logits.shape[3])
        out2 = torch.empty(2, dtype=torch.float32, device=logits.device)
        call("nasb_ce_fwd", ref(desc(logits)), ptr(target), int(ignore_index), ptr(out2), ptr(_ws(logits.device)))
        ctx.ignore = int(ignore_index)
        ctx.save_for_backward(logits, target, out2)
        return out2[0].clone()

    @staticmethod
    def backward(ctx, g):
        logits, target, out2 = ctx.saved_tensors
        g = g.to(torch.float32).contiguous()
        dl = lib.new_act(*logits.shape, logits.dtype, logits.device)
        call("nasb_ce_bwd", ref(desc(logits)), ptr(target), ctx.ignore, ptr(out2), ptr(g), ref(desc(dl)))
        return dl, None, None


def cross_entropy2d(logits, target, ignore_index=255):
    return _CrossEntropy2d.apply(lib.to_nhwc(logits), target, ignore_index)


class _MSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        lib.require_cuda(x)
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        call("nasb_mse_fwd", ref(desc(x)), ref(desc(y)), ptr(out), ptr(_ws(x.device)))
        ctx.save_for_backward(x, y)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        g = g.to(torch.float32).contiguous()
        dx = lib.new_act(*x.shape, x.dtype, x.device)
        call("nasb_mse_bwd", ref(desc(x)), ref(desc(y)), ptr(g), ref(desc(dx)))
        return dx, None


def mse_loss(x, y):
    """nn.MSELoss()(x, y) (knowledge-distillation term, trainer.py:147-149)."""
    return _MSE.apply(lib.to_nhwc(x), lib.to_nhwc(y))


class _BerHu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, valid_min):
        lib.require_cuda(x)
        out3 = torch.empty(3, dtype=torch.float32, device=x.device)
        call("nasb_berhu_fwd", ref(desc(x)), ref(desc(y)), float(valid_min), ptr(out3), ptr(_ws(x.device)))
        ctx.vmin = float(valid_min)
        ctx.save_for_backward(x, y, out3)
        return out3[0].clone()

    @staticmethod
    def backward(ctx, g):
        x, y, out3 = ctx.saved_tensors
        g = g.to(torch.float32).contiguous()
        dx = lib.new_act(*x.shape, x.dtype, x.device)
        call("nasb_berhu_bwd", ref(desc(x)), ref(desc(y)), ctx.vmin, ptr(out3), ptr(g), ref(desc(dx)))
        return dx, None, None


def berhu_loss(pred, target, valid_min=0.0):
    """Reverse Huber depth loss (defined by this repo; the reference ships depth inference only)."""
    return _BerHu.apply(lib.to_nhwc(pred), lib.to_nhwc(target), valid_min)


# ------------------------------------------------------------------------------------------------------- metric
def confmat_labels(pred_u8, gt_u8, n_classes, cm=None):
    lib.require_cuda(pred_u8)
    assert pred_u8.dtype == torch.uint8 and gt_u8.dtype == torch.uint8
    pred_u8, gt_u8 = pred_u8.contiguous().view(-1), gt_u8.contiguous().view(-1)
    assert pred_u8.numel() == gt_u8.numel()
    if cm is None:
        cm = torch.zeros((n_classes, n_classes), dtype=torch.int64, device=pred_u8.device)
    call("nasb_confmat_labels", ptr(pred_u8), ptr(gt_u8), C.c_longlong(pred_u8.numel()), int(n_classes), ptr(cm))
    return cm


def confmat_logits(logits, gt_u8, n_classes, cm=None):
    """Fused upsample(bilinear, align_corners=False) + argmax + (gt < C) mask + histogram; accumulates into cm."""
    lib.require_cuda(logits)
    logits = lib.to_nhwc(logits.detach())
    assert gt_u8.dtype == torch.uint8 and gt_u8.dim() == 3 and gt_u8.shape[0] == logits.shape[0]
    gt_u8 = gt_u8.contiguous()
    if cm is None:
        cm = torch.zeros((n_classes, n_classes), dtype=torch.int64, device=logits.device)
    call("nasb_confmat_logits", ref(desc(logits)), ptr(gt_u8), int(gt_u8.shape[1]), int(gt_u8.shape[2]), int(n_classes),
         ptr(cm))
    return cm


def ius_accs(cm):
    lib.require_cuda(cm)
    c = cm.shape[0]
    iu = torch.empty(c, dtype=torch.float64, device=cm.device)
    acc = torch.empty(c, dtype=torch.float64, device=cm.device)
    npx = torch.empty(c, dtype=torch.int64, device=cm.device)
    call("nasb_ius_accs", ptr(cm.contiguous()), int(c), ptr(iu), ptr(npx), ptr(acc))
    return iu, npx, acc
