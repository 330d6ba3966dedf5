"""GPU parity of whole networks (encoder + decoder) against fixtures from the real reference: logits within 1e-3
relative (north_star) in fp32 mode -- in practice ~1e-5; one training step (loss + gradients); bf16 mode sanity."""
import os

import numpy as np
import pytest
import torch

import nas_segm_b200
from golden_util import NETS, det_state_dict, keys_shapes, rel_err, sub_state, t

pytestmark = pytest.mark.gpu


def _np(x):
    return x.detach().float().cpu().numpy()


def build(tag, fx):
    from nas_segm_b200.nn.encoders import mbv2
    from nas_segm_b200.nn.micro_decoders import MicroDecoder, TemplateDecoder
    paper, cfg, ncls, agg, rep, aux = NETS[tag]
    if paper == "wacv":
        enc = mbv2(return_layers=[1, 2])
        dec = TemplateDecoder(list(enc.out_sizes), ncls, cfg, agg_size=agg, repeats=rep)
    else:
        enc = mbv2()
        dec = MicroDecoder(list(enc.out_sizes), ncls, cfg, agg_size=agg, aux_cell=aux, repeats=rep)
    sd = det_state_dict(keys_shapes(fx), seed=7)
    enc.load_state_dict(sub_state(sd, "encoder."), strict=True)
    dec.load_state_dict(sub_state(sd, "decoder."), strict=True)
    return enc.cuda(), dec.cuda()


@pytest.mark.parametrize("tag", sorted(NETS))
def test_network_eval_logits(golden, tag):
    nas_segm_b200.set_act_dtype(torch.float32)
    fx = golden("net_" + tag)
    enc, dec = build(tag, fx)
    enc.eval(), dec.eval()
    x = t(fx["x"]).cuda()
    with torch.no_grad():
        feats = enc(x)
        out = dec(feats)
    for i, f in enumerate(feats):
        assert rel_err(_np(f), fx["feat%d" % i]) < 1e-4, "feat%d" % i
    auxs = []
    if isinstance(out, tuple):
        out, auxs = out
    assert out.dtype == torch.float32
    assert rel_err(_np(out), fx["out"]) < 1e-3
    for i, a in enumerate(auxs):
        assert rel_err(_np(a), fx["aux%d" % i]) < 1e-3
    if NETS[tag][2] > 1:  # label maps: the reference's own end-to-end criterion (tests/test_inference.py:172-174)
        agree = (_np(out).argmax(1) == fx["out"].argmax(1)).mean()
        assert agree > 0.999


def _oracle_train(tag, fx, dtype, x=None):
    """fp64 / fp32 evaluation of the same training step in the CPU oracle (optionally on a perturbed input)."""
    from oracle import nas_oracle as O
    paper, cfg, ncls, agg, rep, aux = NETS[tag]
    sd = det_state_dict(keys_shapes(fx), seed=7)
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
    Pe, Pd = O.Params(sub_state(sd, "encoder."), dtype=dtype), O.Params(sub_state(sd, "decoder."), dtype=dtype)
    rl = (1, 2) if paper == "wacv" else (1, 2, 4, 6)
    feats = O.mbv2_encoder((t(fx["x"]) if x is None else x).to(dtype), Pe, rl, True)
    if paper == "wacv":
        out, auxs = O.template_decoder(feats, Pd, cfg, O.encoder_out_sizes(rl), ncls, agg, rep, training=True), []
    else:
        out, auxs = O.micro_decoder(feats, Pd, cfg, O.encoder_out_sizes(rl), ncls, agg, aux, rep, training=True)
    y = t(fx["train_y"])
    loss = O.segm_loss(out, y)
    for a in auxs:
        loss = loss + 0.15 * O.segm_loss(a, y, y.shape[1:])
    loss.backward()
    return sd


def _fp32_band(tag, fx, trials=4):
    """How far the reference algorithm's OWN fp32 gradients sit from fp64 on this fixture: the largest (median relative
    error, global relative L2 error) over `trials` evaluations whose inputs differ by 1e-6 relative.  On the 65x65 CVPR
    fixtures that figure moves between 2e-3 and 3e-2 from one such evaluation to the next (tools/diag_grads.py, DESIGN.md
    section 2): a ReLU or a 2-sample BatchNorm decides differently and every gradient upstream moves with it.  A single fp32
    evaluation is therefore not a band; the maximum over a few is."""
    g = torch.Generator().manual_seed(1)
    x0 = t(fx["x"])
    med, l2 = 0.0, 0.0
    for i in range(trials):
        x = x0 if i == 0 else (x0.double() * (1 + (torch.rand(x0.shape, generator=g, dtype=torch.float64) - 0.5) * 2e-6)).float()
        s64, s32 = _oracle_train(tag, fx, torch.float64, x), _oracle_train(tag, fx, torch.float32, x)
        es, num, den = [], 0.0, 0.0
        for k, v in s64.items():
            if v.grad is None or float(v.grad.abs().max()) < 1e-4:
                continue
            es.append(rel_err(s32[k].grad.numpy(), v.grad.numpy()))
            num += float(((s32[k].grad.double() - v.grad) ** 2).sum())
            den += float((v.grad ** 2).sum())
        med, l2 = max(med, float(np.median(es))), max(l2, (num / den) ** 0.5)
    return med, l2


@pytest.mark.parametrize("tag", ["W0cv", "C0search", "C1search"])
def test_network_train_step(golden, tag):
    """One training step (BN in training mode): logits and loss against the real reference's fixture; gradients against
    an fp64 evaluation of the oracle.  The 65x65 CVPR fixtures put 18 samples per channel through the deepest
    BatchNorms, which makes the encoder gradients ill-conditioned: the reference's own fp32 gradients are 1-2 % away from
    fp64 there (tools/diag_grads.py), so each parameter is held to max(5e-3, 3 x the fp32 CPU error)."""
    from nas_segm_b200 import functional as Fn
    nas_segm_b200.set_act_dtype(torch.float32)
    fx = golden("net_" + tag)
    enc, dec = build(tag, fx)
    enc.train(), dec.train()
    x = t(fx["x"]).cuda()
    out = dec(enc(x))
    auxs = []
    if isinstance(out, tuple):
        out, auxs = out
    y = t(fx["train_y"]).cuda()
    loss = Fn.cross_entropy2d(out, y, 255)
    for a in auxs:
        loss = loss + 0.15 * Fn.cross_entropy2d(Fn.resize(a, tuple(y.shape[1:])), y, 255)
    assert rel_err(_np(out), fx["train_out"]) < 1e-3
    assert abs(float(loss.detach()) - float(fx["train_loss"])) < 1e-4 * max(1.0, abs(float(fx["train_loss"])))
    loss.backward()
    named = {("encoder." + k): p for k, p in enc.named_parameters()}
    named.update({("decoder." + k): p for k, p in dec.named_parameters()})
    sd64, sd32 = _oracle_train(tag, fx, torch.float64), _oracle_train(tag, fx, torch.float32)
    bad, checked, e_all, e_cpu_all = [], 0, [], []
    num = den = num_cpu = 0.0
    for k, p in named.items():
        g64 = sd64[k].grad
        if g64 is None:
            if p.grad is not None and float(p.grad.abs().max()) != 0.0:
                bad.append((k, "expected no gradient"))
            continue
        scale = float(g64.abs().max())
        if scale < 1e-4:  # analytically-zero gradients (e.g. a bias in front of a training-mode BatchNorm): noise only
            if float(p.grad.abs().max()) > 1e-3:
                bad.append((k, "expected ~0", float(p.grad.abs().max())))
            continue
        e_cuda = rel_err(_np(p.grad), g64.numpy())
        e_cpu = rel_err(sd32[k].grad.numpy(), g64.numpy())
        checked += 1
        e_all.append(e_cuda)
        e_cpu_all.append(e_cpu)
        num += float(((p.grad.double().cpu() - g64) ** 2).sum())
        num_cpu += float(((sd32[k].grad.double() - g64) ** 2).sum())
        den += float((g64 ** 2).sum())
        # A ReLU whose pre-activation is ~0 can flip between two fp32 evaluations and move the gradient of everything
        # upstream by a whole term; on the 3x3 / 5x5 maps of these fixtures (18-50 samples per channel) one flip is worth
        # several per cent of a parameter's gradient.  Per parameter only gross errors are rejected; the aggregate
        # (median, global relative L2) must stay in the fp32 band.
        if e_cuda > max(0.35, 3.0 * e_cpu):
            bad.append((k, e_cuda, e_cpu))
    assert checked > 100 and not bad, bad[:20]
    band_med, band_l2 = _fp32_band(tag, fx)
    assert np.median(e_all) <= max(1e-3, 3.0 * max(band_med, np.median(e_cpu_all))), (np.median(e_all), band_med)
    assert (num / den) ** 0.5 <= max(1e-2, 3.0 * max(band_l2, (num_cpu / den) ** 0.5)), ((num / den) ** 0.5, band_l2)
    # the real reference's fixture gradients (fp32) must sit in the same error band around fp64
    for k in [k for k in fx.files if k.startswith("grad/")]:
        g64 = sd64[k[5:]].grad.numpy()
        assert rel_err(_np(named[k[5:]].grad), g64) <= max(5e-2, 3.0 * rel_err(fx[k], g64)), k


@pytest.mark.parametrize("tag", ["W0", "C0search"])
def test_network_bf16_mode(golden, tag):
    """Speed mode: bf16 activations; judged against fp32 (the reference's own bf16-autocast error is ~6e-3 of max|logit|,
    BASELINE.md)."""
    fx = golden("net_" + tag)
    enc, dec = build(tag, fx)
    enc.eval(), dec.eval()
    nas_segm_b200.set_act_dtype(torch.bfloat16)
    try:
        with torch.no_grad():
            out = dec(enc(t(fx["x"]).cuda()))
    finally:
        nas_segm_b200.set_act_dtype(torch.float32)
    if isinstance(out, tuple):
        out = out[0]
    assert out.dtype == torch.float32
    assert rel_err(_np(out), fx["out"]) < 5e-2


def test_cuda_graph_forward_matches_eager(golden):
    """A captured CUDA graph of the eval forward replays to the same logits as the eager call (fp32 and bf16)."""
    from nas_segm_b200.graphs import GraphedForward
    fx = golden("net_W0")
    enc, dec = build("W0", fx)
    model = torch.nn.Sequential(enc, dec).eval()
    for dtype in (torch.float32, torch.bfloat16):
        nas_segm_b200.set_act_dtype(dtype)
        try:
            x = t(fx["x"]).cuda()
            with torch.no_grad():
                ref = model(x).clone()
            g = GraphedForward(model, x)
            out = g(x).clone()
            assert torch.equal(out, ref)
            x2 = x * 0.5 + 0.1
            with torch.no_grad():
                ref2 = model(x2)
            assert torch.equal(g(x2), ref2)
        finally:
            nas_segm_b200.set_act_dtype(torch.float32)


@pytest.mark.parametrize("bn_training", [False, True])
def test_bf16_training_step_at_well_conditioned_size(bn_training):
    """The mode bench.py times (bf16 activations, tcgen05 / TMA kernels, fused statistics) held to asserted bounds on a full
    forward + backward of WACV arch0 at batch 2 @512x256 (every BatchNorm sees >= 4096 samples per channel): logits and the
    gradient of EVERY parameter against the fp32 oracle, next to what stock PyTorch's bf16 autocast loses on the same step,
    plus run-to-run reproducibility (weight-gradient / statistics reductions are atomic: bit-equality is not promised).

    bn_training=False (the reference's FREEZE_BN mode): absolute bounds.  bn_training=True: a randomly initialised
    BatchNorm network in training mode amplifies ANY perturbation by ~1.1x per layer (gradient explosion of BN networks at
    initialisation), so stock bf16 autocast itself is ~16 % from fp32 on these logits; the bound there is "no worse than
    stock bf16", and fp32 parity mode must still meet the north star's 1e-3."""
    from golden_util import W0
    from nas_segm_b200 import functional as Fn
    from nas_segm_b200.nn.encoders import mbv2
    from nas_segm_b200.nn.micro_decoders import TemplateDecoder
    from oracle import nas_oracle as O
    torch.manual_seed(0)
    enc = mbv2(return_layers=[1, 2])
    dec = TemplateDecoder(list(enc.out_sizes), 19, W0, agg_size=64, repeats=2)
    ks = [("encoder." + k, tuple(v.shape)) for k, v in enc.state_dict().items()]
    ks += [("decoder." + k, tuple(v.shape)) for k, v in dec.state_dict().items()]
    sd = det_state_dict(ks, seed=5)
    enc, dec = enc.cuda().train(bn_training), dec.cuda().train(bn_training)
    g = torch.Generator().manual_seed(9314)
    x = torch.randn(2, 3, 256, 512, generator=g)
    y = torch.randint(0, 19, (2, 64, 128), generator=g)
    y[torch.rand(2, 64, 128, generator=g) < 0.05] = 255

    def ours(dtype):
        nas_segm_b200.set_act_dtype(dtype)
        try:
            enc.load_state_dict(sub_state(sd, "encoder."))  # same parameters AND running statistics for every run
            dec.load_state_dict(sub_state(sd, "decoder."))
            enc.zero_grad(), dec.zero_grad()
            out = dec(enc(x.cuda()))
            loss = Fn.cross_entropy2d(out, y.cuda(), 255)
            loss.backward()
            grads = {"encoder." + k: p.grad.double().cpu() for k, p in enc.named_parameters() if p.grad is not None}
            grads.update({"decoder." + k: p.grad.double().cpu() for k, p in dec.named_parameters() if p.grad is not None})
            return out.detach().double().cpu(), float(loss.detach()), grads
        finally:
            nas_segm_b200.set_act_dtype(torch.float32)

    def oracle(device, autocast):
        Pe = O.Params({k[8:]: v.clone().to(device) for k, v in sd.items() if k.startswith("encoder.")}).requires_grad_()
        Pd = O.Params({k[8:]: v.clone().to(device) for k, v in sd.items() if k.startswith("decoder.")}).requires_grad_()
        with torch.autocast(device, dtype=torch.bfloat16, enabled=autocast):
            out = O.template_decoder(O.mbv2_encoder(x.to(device), Pe, (1, 2), bn_training), Pd, W0, [24, 32], 19, 64, 2,
                                     training=bn_training)
        loss = O.segm_loss(out.float(), y.to(device))
        loss.backward()
        grads = {"encoder." + k: v.grad.double().cpu() for k, v in Pe.sd.items() if v.grad is not None}
        grads.update({"decoder." + k: v.grad.double().cpu() for k, v in Pd.sd.items() if v.grad is not None})
        return out.detach().double().cpu(), float(loss.detach()), grads

    def errs(out, grads, ref_out, ref_grads):
        e_out = float((out - ref_out).abs().max() / ref_out.abs().max())
        num = sum(float(((grads[k] - ref_grads[k]) ** 2).sum()) for k in ref_grads)
        den = sum(float((ref_grads[k] ** 2).sum()) for k in ref_grads)
        return e_out, (num / den) ** 0.5

    ref_out, ref_loss, ref_g = oracle("cpu", False)                    # fp32, torch CPU kernels
    o32, l32, g32 = ours(torch.float32)
    o16, l16, g16 = ours(torch.bfloat16)
    o16b, l16b, g16b = ours(torch.bfloat16)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        oac, lac, gac = oracle("cuda", True)                            # stock PyTorch, bf16 autocast
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert set(g16) == set(ref_g) and len(ref_g) > 200, (sorted(set(g16) ^ set(ref_g))[:20], len(g16), len(ref_g))
    e32, e16, eac, rep = errs(o32, g32, ref_out, ref_g), errs(o16, g16, ref_out, ref_g), errs(oac, gac, ref_out, ref_g), errs(o16b, g16b, o16, g16)
    print("\nbf16 step @512x256 b2, BN %s: logits rel-err ours %.2e (fp32 mode %.1e, torch autocast %.2e); gradient global "
          "rel-L2 ours %.2e (fp32 mode %.1e, torch autocast %.2e); run-to-run logits %.1e grads %.1e; loss %.5f / fp32 mode %.5f / "
          "oracle %.5f" % ("training" if bn_training else "frozen", e16[0], e32[0], eac[0], e16[1], e32[1], eac[1], rep[0], rep[1],
                           l16, l32, ref_loss))
    assert all(torch.isfinite(v).all() for v in g16.values()) and torch.isfinite(o16).all()
    # parity mode: the north star's tolerance on logits; gradients in the fp32 band
    assert e32[0] < 1e-3 and e32[1] < 5e-3 and abs(l32 - ref_loss) < 1e-4 * max(1.0, abs(ref_loss))
    if not bn_training:
        assert e16[0] < 3e-2 and e16[1] < 8e-2 and abs(l16 - ref_loss) < 3e-2 * max(1.0, abs(ref_loss))
        assert rep[0] < 1e-3 and rep[1] < 1e-2  # only the atomic weight-gradient sums reorder between runs
    # speed mode is as accurate as stock PyTorch's bf16 on the same step (both against fp32)
    assert e16[0] < 2.0 * eac[0] + 1e-2 and e16[1] < 1.5 * eac[1] + 3e-2
    assert rep[0] < 0.5 and rep[1] < 1.0  # a race in the fused-statistics / tile / tcgen05 paths would not be this quiet


def test_full_size_properties_2048x1024():
    """BASELINE size (2048x1024, WACV arch0, speed mode) through size-independent properties: images of a batch do not
    interact in eval mode (batch of 2 == each image alone: exercises every tile / tail / persistent-loop path at full
    size), the fused CE equals torch's CE on the same logits, and the fused upsample+argmax+histogram reward path counts
    every valid pixel exactly once with the ground-truth marginals."""
    from nas_segm_b200 import functional as Fn
    torch.manual_seed(0)
    from nas_segm_b200.nn.encoders import mbv2
    from nas_segm_b200.nn.micro_decoders import TemplateDecoder
    from golden_util import W0
    nas_segm_b200.set_act_dtype(torch.bfloat16)
    try:
        enc = mbv2(return_layers=[1, 2])
        dec = TemplateDecoder(list(enc.out_sizes), 19, W0, agg_size=64, repeats=2)
        for m in list(enc.modules()) + list(dec.modules()):
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
        enc, dec = enc.cuda().eval(), dec.cuda().eval()
        g = torch.Generator().manual_seed(9314)
        x = torch.randn(2, 3, 1024, 2048, generator=g).cuda()
        with torch.no_grad():
            out = dec(enc(x))
            o0, o1 = dec(enc(x[0:1].contiguous())), dec(enc(x[1:2].contiguous()))
        assert out.shape == (2, 19, 256, 512) and out.dtype == torch.float32
        assert torch.isfinite(out).all()
        scale = float(out.abs().max())
        assert float((out[0:1] - o0).abs().max()) <= 1e-6 * scale and float((out[1:2] - o1).abs().max()) <= 1e-6 * scale
        y = torch.randint(0, 19, (2, 256, 512), generator=g)
        y[torch.rand(2, 256, 512, generator=g) < 0.05] = 255
        y = y.cuda()
        loss = Fn.cross_entropy2d(out, y, 255)
        ref = torch.nn.functional.cross_entropy(out, y, ignore_index=255)
        assert abs(float(loss) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))
        gt = torch.randint(0, 21, (2, 1024, 2048), generator=g).to(torch.uint8).cuda()  # 19, 20 = out of range -> skipped
        cm = Fn.confmat_logits(out, gt, 19).cpu().numpy()
        gtn = gt.cpu().numpy()
        assert cm.sum() == int((gtn < 19).sum())
        assert (cm.sum(1) == np.bincount(gtn[gtn < 19].ravel(), minlength=19)).all()
    finally:
        nas_segm_b200.set_act_dtype(torch.float32)
