"""GPU parity of whole networks (encoder + decoder) against fixtures from the real reference: logits within 1e-3
relative (north_star) in fp32 mode -- in practice ~1e-5; one training step (loss + gradients); bf16 mode sanity."""
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


@pytest.mark.parametrize("tag", ["W0cv", "C0search", "C1search"])
def test_network_train_step(golden, tag):
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
    assert abs(float(loss) - float(fx["train_loss"])) < 1e-4 * max(1.0, abs(float(fx["train_loss"])))
    loss.backward()
    named = {("encoder." + k): p for k, p in enc.named_parameters()}
    named.update({("decoder." + k): p for k, p in dec.named_parameters()})
    bad = []
    for k in [k for k in fx.files if k.startswith("grad/")]:
        e = rel_err(_np(named[k[5:]].grad), fx[k])
        if e > 5e-3:
            bad.append((k, e))
    for k in [k for k in fx.files if k.startswith("gnorm/")]:
        ref = float(fx[k])
        g = named[k[6:]].grad
        if ref < 0:
            if g is not None and float(g.abs().max()) != 0.0:
                bad.append((k, "expected no grad"))
        elif abs(float(g.norm()) - ref) > 5e-3 * ref + 2e-4:
            bad.append((k, float(g.norm()), ref))
    assert not bad, bad[:20]


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
