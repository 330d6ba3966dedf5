"""Pins oracle/ against fixtures produced by the REAL reference (tests/golden/make_golden.py).
CPU only.  fp32 torch CPU kernels on both sides, so tolerances are tight."""
import numpy as np
import pytest
import torch

from golden_util import NETS, det_state_dict, keys_shapes, rel_err, sub_state, t
from oracle import miou_oracle as MO
from oracle import nas_oracle as O


def _op_tags(fx, kind):
    return sorted({k.split("/")[0] for k in fx.files if k.startswith(kind)})


def test_registry_ops_eval_train_and_grads(golden):
    fx = golden("ops")
    tags = _op_tags(fx, "op")
    assert len(tags) == 32  # 16 registry names x 2 shapes
    for idx, tag in enumerate(tags):
        name, cin, cout, stride, repeats = [str(v) for v in fx[tag + "/meta"]]
        cin, cout, stride, repeats = int(cin), int(cout), int(stride), int(repeats)
        ks = keys_shapes(fx, tag + "/")
        sd = det_state_dict(ks, seed=idx)
        from detweights import det_array
        x = t(det_array(tag + "/x", (2, cin, 13, 17))).requires_grad_(True)
        P = O.Params(dict(sd))
        y = O.op_forward(name, x, P, "", cin, cout, stride, repeats, training=False)
        # the oracle must touch exactly the reference's keys
        assert rel_err(y.detach().numpy(), fx[tag + "/y_eval"]) < 1e-5, tag
        P = O.Params({k: v.clone() for k, v in sd.items()}).requires_grad_()
        y = O.op_forward(name, x, P, "", cin, cout, stride, repeats, training=True)
        assert rel_err(y.detach().numpy(), fx[tag + "/y_train"]) < 1e-5, tag
        ct = t(det_array(tag + "/ct", tuple(y.shape)))
        pkeys = [k for k in fx.files if k.startswith(tag + "/g/")]
        params = [P.sd[k[len(tag) + 4 - 1:].lstrip("/")] for k in pkeys]
        grads = torch.autograd.grad((y * ct).sum(), [x] + params, allow_unused=True)
        gx = grads[0] if grads[0] is not None else torch.zeros_like(x)
        assert rel_err(gx.numpy(), fx[tag + "/gx"]) < 2e-5 or np.abs(fx[tag + "/gx"]).max() == 0, tag
        for k, g in zip(pkeys, grads[1:]):
            assert rel_err(g.numpy(), fx[k]) < 5e-5, k
        for k in [k for k in fx.files if k.startswith(tag + "/post/")]:
            assert rel_err(P.sd[k.split("/post/")[1]].detach().numpy(), fx[k]) < 1e-5, k


def test_registry_keys_match_reference(golden):
    fx = golden("ops")
    for idx, tag in enumerate(_op_tags(fx, "op")):
        name, cin, cout, stride, repeats = [str(v) for v in fx[tag + "/meta"]]
        P = O.Params()
        O.op_forward(name, torch.zeros(2, int(cin), 13, 17), P, "", int(cin), int(cout), int(stride), int(repeats),
                     training=False)
        assert {(k, tuple(v.shape)) for k, v in P.sd.items()} == set(keys_shapes(fx, tag + "/")), tag


def test_agg_ops(golden):
    from detweights import det_array
    fx = golden("ops")
    tags = _op_tags(fx, "agg")
    assert len(tags) == 8
    for idx, tag in enumerate(tags):
        m = [str(v) for v in fx[tag + "/meta"]]
        name, (c0, c1, cout, larger, h0, w0, h1, w1) = m[0], [int(v) for v in m[1:]]
        sd = det_state_dict(keys_shapes(fx, tag + "/"), seed=100 + idx)
        x = t(det_array(tag + "/x", (2, c0, h0, w0))).requires_grad_(True)
        y = t(det_array(tag + "/y", (2, c1, h1, w1))).requires_grad_(True)
        z = O.agg_forward(name, x, y, O.Params(dict(sd)), "", c0, c1, cout, bool(larger), training=False)
        assert rel_err(z.detach().numpy(), fx[tag + "/z_eval"]) < 1e-5, tag
        P = O.Params({k: v.clone() for k, v in sd.items()}).requires_grad_()
        z = O.agg_forward(name, x, y, P, "", c0, c1, cout, bool(larger), training=True)
        assert rel_err(z.detach().numpy(), fx[tag + "/z_train"]) < 1e-5, tag
        ct = t(det_array(tag + "/ct", tuple(z.shape)))
        pkeys = [k for k in fx.files if k.startswith(tag + "/g/")]
        params = [P.sd[k.split("/g/")[1]] for k in pkeys]
        grads = torch.autograd.grad((z * ct).sum(), [x, y] + params)
        assert rel_err(grads[0].numpy(), fx[tag + "/gx"]) < 2e-5, tag
        assert rel_err(grads[1].numpy(), fx[tag + "/gy"]) < 2e-5, tag
        for k, g in zip(pkeys, grads[2:]):
            assert rel_err(g.numpy(), fx[k]) < 5e-5, k


def _run_net(tag, fx, training=False, sd=None):
    paper, cfg, ncls, agg, rep, aux = NETS[tag]
    sd = sd if sd is not None else det_state_dict(keys_shapes(fx), seed=7)
    Pe = O.Params(sub_state(sd, "encoder."))
    Pd = O.Params(sub_state(sd, "decoder."))
    x = t(fx["x"])
    if paper == "wacv":
        rl = (1, 2)
        feats = O.mbv2_encoder(x, Pe, rl, training)
        out = O.template_decoder(feats, Pd, cfg, O.encoder_out_sizes(rl), ncls, agg, rep, training=training)
        return feats, out, [], Pe, Pd
    rl = (1, 2, 4, 6)
    feats = O.mbv2_encoder(x, Pe, rl, training)
    out, auxs = O.micro_decoder(feats, Pd, cfg, O.encoder_out_sizes(rl), ncls, agg, aux, rep, training=training)
    return feats, out, auxs, Pe, Pd


@pytest.mark.parametrize("tag", sorted(NETS))
def test_networks_eval(golden, tag):
    fx = golden("net_" + tag)
    with torch.no_grad():
        feats, out, auxs, _, _ = _run_net(tag, fx)
    for i, f in enumerate(feats):
        assert rel_err(f.numpy(), fx["feat%d" % i]) < 1e-5
    assert rel_err(out.numpy(), fx["out"]) < 2e-5
    for i, a in enumerate(auxs):
        assert rel_err(a.numpy(), fx["aux%d" % i]) < 2e-5


@pytest.mark.parametrize("tag", sorted(NETS))
def test_network_keys_and_param_counts(golden, tag):
    fx = golden("net_" + tag)
    paper, cfg, ncls, agg, rep, aux = NETS[tag]
    rl = (1, 2) if paper == "wacv" else (1, 2, 4, 6)
    Pe, Pd = O.Params(), O.Params()
    x = torch.zeros(tuple(fx["x"].shape))
    with torch.no_grad():
        feats = O.mbv2_encoder(x, Pe, rl)
        if paper == "wacv":
            O.template_decoder(feats, Pd, cfg, O.encoder_out_sizes(rl), ncls, agg, rep)
        else:
            O.micro_decoder(feats, Pd, cfg, O.encoder_out_sizes(rl), ncls, agg, aux, rep)
    mine = {("encoder." + k, tuple(v.shape)) for k, v in Pe.sd.items()}
    mine |= {("decoder." + k, tuple(v.shape)) for k, v in Pd.sd.items()}
    assert mine == set(keys_shapes(fx))
    n_params = sum(int(np.prod(s)) for k, s in keys_shapes(fx)
                   if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert n_params == int(fx["n_params"])
    if tag == "W0":
        assert n_params == 280147  # README.md:79
    if tag == "W1":
        assert n_params == 268235


@pytest.mark.parametrize("tag", ["W0cv", "C0search", "C1search"])
def test_networks_train_step_grads(golden, tag):
    fx = golden("net_" + tag)
    sd = det_state_dict(keys_shapes(fx), seed=7)
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
    feats, out, auxs, Pe, Pd = _run_net(tag, fx, training=True, sd=sd)
    y = t(fx["train_y"])
    loss = O.segm_loss(out, y)
    for a in auxs:
        loss = loss + 0.15 * O.segm_loss(a, y, y.shape[1:])
    assert rel_err(out.detach().numpy(), fx["train_out"]) < 2e-5
    assert abs(float(loss.detach()) - float(fx["train_loss"])) < 1e-5 * max(1.0, abs(float(fx["train_loss"])))
    loss.backward()
    for k in [k for k in fx.files if k.startswith("grad/")]:
        g = sd[k[5:]].grad
        assert rel_err(g.numpy(), fx[k]) < 1e-3, k
    for k in [k for k in fx.files if k.startswith("gnorm/")]:
        ref = float(fx[k])
        g = sd[k[6:]].grad
        if ref < 0:
            assert g is None or float(g.abs().max()) == 0.0, k
        else:
            assert abs(float(g.norm()) - ref) <= 2e-3 * ref + 1e-4, k  # floor: analytically-zero grads are fp noise


def test_metric_fast_cm_and_ius(golden):
    fx = golden("metric")
    for case, C in enumerate((21, 19, 11, 2, 5)):
        p, g = fx["cm%d/p" % case], fx["cm%d/g" % case]
        assert np.array_equal(MO.fast_cm_c(p, g, C), fx["cm%d/cm" % case])
        assert np.array_equal(MO.fast_cm_np(p, g, C), fx["cm%d/cm" % case])
    for case in range(3):
        cm = fx["v%d/cm" % case]
        for fn in (MO.compute_ius_accs_c, MO.compute_ius_accs_np):
            iu, npx, acc = fn(cm)
            assert np.array_equal(iu, fx["v%d/ious" % case])
            assert np.array_equal(npx, fx["v%d/npx" % case])
            assert np.array_equal(acc, fx["v%d/accs" % case])
        assert np.array_equal(MO.compute_ius_accs_c(cm)[0], fx["v%d/iu_only" % case])


def test_metric_uint32_wrap_matches_c_and_numpy():
    cm = np.array([[2 ** 32 + 5, 7], [3, 2 ** 33 + 11]], dtype=np.int64)
    a, b = MO.compute_ius_accs_c(cm), MO.compute_ius_accs_np(cm)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)


def test_validate_reward_path(golden):
    """inference.py:58-91 end to end from logits: upsample -> argmax(first) -> mask -> cm -> reward."""
    fx = golden("metric")
    for case in range(3):
        C = int(fx["v%d/C" % case])
        cm = np.zeros((C, C), np.int64)
        for b in range(2):
            lg, tg = t(fx["v%d/logits%d" % (case, b)]), fx["v%d/target%d" % (case, b)]
            up = torch.nn.functional.interpolate(lg, size=tg.shape[1:], mode="bilinear", align_corners=False)
            cm += MO.cm_from_logits(up.numpy(), tg, C)
        assert np.array_equal(cm, fx["v%d/cm" % case])
        assert abs(MO.reward_from_cm(cm, (0,))[0] - float(fx["v%d/reward" % case])) < 1e-12
