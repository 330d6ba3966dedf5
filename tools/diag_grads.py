"""Diagnostic (GPU box): per-parameter gradient error of the CUDA path and of the fp32 CPU oracle, both measured against
an fp64 evaluation of the oracle -- separates implementation error from the conditioning of tiny-batch BatchNorm."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import nas_segm_b200  # noqa: E402
from golden_util import NETS, det_state_dict, keys_shapes, rel_err, sub_state, t  # noqa: E402
from nas_segm_b200 import functional as Fn  # noqa: E402
from oracle import nas_oracle as O  # noqa: E402
from test_gpu_nets import build  # noqa: E402


def oracle_run(tag, fx, dtype):
    paper, cfg, ncls, agg, rep, aux = NETS[tag]
    sd = det_state_dict(keys_shapes(fx), seed=7)
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
    Pe, Pd = O.Params(sub_state(sd, "encoder."), dtype=dtype), O.Params(sub_state(sd, "decoder."), dtype=dtype)
    x = t(fx["x"]).to(dtype)
    rl = (1, 2) if paper == "wacv" else (1, 2, 4, 6)
    feats = O.mbv2_encoder(x, Pe, rl, True)
    for f in feats:
        f.retain_grad()
    if paper == "wacv":
        out, auxs = O.template_decoder(feats, Pd, cfg, O.encoder_out_sizes(rl), ncls, agg, rep, training=True), []
    else:
        out, auxs = O.micro_decoder(feats, Pd, cfg, O.encoder_out_sizes(rl), ncls, agg, aux, rep, training=True)
    y = t(fx["train_y"])
    loss = O.segm_loss(out, y)
    for a in auxs:
        loss = loss + 0.15 * O.segm_loss(a, y, y.shape[1:])
    loss.backward()
    return sd, feats, out, loss


for tag in sys.argv[1:] or ["C0search", "C1search", "W0cv"]:
    fx = np.load(os.path.join(ROOT, "tests", "golden", "net_%s.npz" % tag))
    sd64, f64, o64, l64 = oracle_run(tag, fx, torch.float64)
    sd32, f32, o32, l32 = oracle_run(tag, fx, torch.float32)
    nas_segm_b200.set_act_dtype(torch.float32)
    enc, dec = build(tag, fx)
    enc.train(), dec.train()
    feats = enc(t(fx["x"]).cuda())
    for f in feats:
        f.retain_grad()
    out = dec(feats)
    auxs = []
    if isinstance(out, tuple):
        out, auxs = out
    y = t(fx["train_y"]).cuda()
    loss = Fn.cross_entropy2d(out, y, 255)
    for a in auxs:
        loss = loss + 0.15 * Fn.cross_entropy2d(Fn.resize(a, tuple(y.shape[1:])), y, 255)
    loss.backward()
    print("== %s  loss cuda %.7f  cpu32 %.7f  cpu64 %.7f" % (tag, float(loss), float(l32), float(l64)))
    for i in range(len(feats)):
        print("  feat%d      cuda %.2e   cpu32 %.2e    | dfeat cuda %.2e  cpu32 %.2e" % (
            i, rel_err(feats[i].detach().cpu().numpy(), f64[i].detach().numpy()),
            rel_err(f32[i].detach().numpy(), f64[i].detach().numpy()),
            rel_err(feats[i].grad.cpu().numpy(), f64[i].grad.numpy()), rel_err(f32[i].grad.numpy(), f64[i].grad.numpy())))
    print("  out        cuda %.2e   cpu32 %.2e" % (rel_err(out.detach().cpu().numpy(), o64.detach().numpy()),
                                                   rel_err(o32.detach().numpy(), o64.detach().numpy())))
    named = {("encoder." + k): p for k, p in enc.named_parameters()}
    named.update({("decoder." + k): p for k, p in dec.named_parameters()})
    rows = []
    for k, p in named.items():
        g64 = sd64[k].grad
        if g64 is None or p.grad is None:
            continue
        rows.append((rel_err(p.grad.cpu().numpy(), g64.numpy()), rel_err(sd32[k].grad.numpy(), g64.numpy()), k))
    rows.sort(reverse=True)
    print("  worst parameters (cuda-vs-fp64, cpu32-vs-fp64):")
    for r in rows[:12]:
        print("    %.2e  %.2e  %s" % r)
    if os.environ.get("DIAG_ALL"):
        print("  decoder parameters in module order:")
        for k, p in named.items():
            if k.startswith("decoder.") and sd64[k].grad is not None and p.grad is not None and float(sd64[k].grad.abs().max()) > 1e-4:
                print("    %.2e  %.2e  %s" % (rel_err(p.grad.cpu().numpy(), sd64[k].grad.numpy()),
                                              rel_err(sd32[k].grad.numpy(), sd64[k].grad.numpy()), k))
    enc_rows = [r for r in rows if r[2].startswith("encoder.")]
    print("  median over encoder params: cuda %.2e cpu32 %.2e ; decoder: cuda %.2e cpu32 %.2e" % (
        np.median([r[0] for r in enc_rows]), np.median([r[1] for r in enc_rows]),
        np.median([r[0] for r in rows if r[2].startswith("decoder.")]),
        np.median([r[1] for r in rows if r[2].startswith("decoder.")])))
