"""Row f2: the multi-tensor clip + SGD/Adam + Polyak kernels against torch's own clip_grad_norm_ / optim.step() /
mul_().add_() loop (reference: src/engine/trainer.py:163-169,258-272, src/utils/solvers.py:35-52), and the one-launch
operand packs against the per-call packs."""
import copy

import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def _params(seed, dev):
    g = torch.Generator().manual_seed(seed)
    shapes = [(64, 32, 1, 1), (32,), (32,), (24, 1, 5, 5), (19, 64, 3, 3), (19,), (7,), (3, 5, 1, 1), (1,), (96, 96, 1, 1),
              (130, 33)]  # 130*33 = 4290 elements: crosses a 4096-element chunk with a ragged tail
    return [nn.Parameter(torch.randn(s, generator=g).to(dev)) for s in shapes]


def _grads(params, seed, scale):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(p.shape, generator=g).to(p.device) * scale for p in params]


@pytest.mark.parametrize("kind", ["sgd_adam", "adam_adam", "sgd_nomom"])
def test_fused_step_matches_torch(kind):
    from nas_segm_b200.optim import FusedStep
    dev = torch.device("cuda")
    pa, pb = _params(1, dev), _params(2, dev)
    qa, qb = [nn.Parameter(p.detach().clone()) for p in pa], [nn.Parameter(p.detach().clone()) for p in pb]

    def make(a, b):
        if kind == "sgd_adam":
            return torch.optim.SGD(a, lr=1e-2, momentum=0.9, weight_decay=1e-4), torch.optim.Adam(b, lr=3e-3, weight_decay=1e-5)
        if kind == "adam_adam":
            return (torch.optim.Adam(a, lr=1e-3, betas=(0.8, 0.95), eps=1e-6),
                    torch.optim.Adam([{"params": b[:4], "lr": 2e-3}, {"params": b[4:], "weight_decay": 1e-3}], lr=5e-3))
        return torch.optim.SGD(a, lr=1e-2), torch.optim.SGD(b, lr=1e-2, momentum=0.5, nesterov=True)

    oa, ob = make(pa, pb)      # fused
    ra, rb = make(qa, qb)      # torch
    avg_f = [p.detach().clone() for p in pa + pb]
    avg_t = copy.deepcopy(avg_f)
    fs = FusedStep([(oa, pa, 1.5), (ob, pb, 0.0 if kind == "sgd_nomom" else 40.0)], pa + pb, avg_f)
    assert FusedStep.supported(fs.entries)
    for it in range(6):
        ga, gb = _grads(pa, 10 + it, 1.0), _grads(pb, 20 + it, 3.0 if it % 2 else 0.01)
        for plist, glist in ((pa, ga), (pb, gb), (qa, ga), (qb, gb)):
            for p, g in zip(plist, glist):
                p.grad = g.clone()
        if it == 2:  # a parameter without gradient is skipped by clip and step alike
            pa[3].grad = qa[3].grad = None
            pb[0].grad = qb[0].grad = None
        fs.step(0.9)
        na = nn.utils.clip_grad_norm_(qa, 1.5)
        if kind != "sgd_nomom":
            nb = nn.utils.clip_grad_norm_(qb, 40.0)
        ra.step(), rb.step()
        for p, a in zip(qa + qb, avg_t):
            a.mul_(0.9).add_(p.data, alpha=0.1)
        norms = fs.grad_norms()
        assert abs(float(norms[0]) - float(na)) <= 1e-5 * float(na)
        if kind != "sgd_nomom":
            assert abs(float(norms[1]) - float(nb)) <= 1e-5 * float(nb)
        for j, (p, q) in enumerate(zip(pa + pb, qa + qb)):
            assert torch.allclose(p, q, rtol=2e-6, atol=2e-7), (it, p.shape, float((p - q).abs().max()))
            # clip_grad_norm_ leaves the scaled gradient behind (torch's foreach SGD additionally overwrites .grad with
            # grad + momentum*buf when nesterov is on -- a side effect of its implementation that is not reproduced)
            if p.grad is not None and not (kind == "sgd_nomom" and j >= len(pa)):
                assert torch.allclose(p.grad, q.grad, rtol=2e-6, atol=1e-7)
        for a, b in zip(avg_f, avg_t):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-7)
    # optimiser state is torch's own: same keys, same values, and a plain torch step continues from it
    for o, r, ps, qs in ((oa, ra, pa, qa), (ob, rb, pb, qb)):
        for p, q in zip(ps, qs):
            assert set(o.state[p].keys()) == set(r.state[q].keys())
            for k in o.state[p]:
                assert torch.allclose(o.state[p][k].float().cpu(), r.state[q][k].float().cpu(), rtol=1e-5, atol=1e-7), k
    sd = oa.state_dict()
    assert len(sd["state"]) == len(ra.state_dict()["state"])


def test_fused_step_declines_what_it_does_not_cover():
    from nas_segm_b200.optim import FusedStep
    p = [nn.Parameter(torch.zeros(4, device="cuda"))]
    assert not FusedStep.supported([(torch.optim.Adam(p, amsgrad=True), p, 1.0)])
    assert not FusedStep.supported([(torch.optim.AdamW(p), p, 1.0)])
    assert not FusedStep.supported([(torch.optim.SGD(p, lr=0.1, momentum=0.9, dampening=0.1), p, 1.0)])
    assert not FusedStep.supported([(torch.optim.RMSprop(p), p, 1.0)])
    assert FusedStep.supported([(torch.optim.SGD(p, lr=0.1, momentum=0.9), p, 1.0)])


def test_pack_scope_equals_per_call_packs():
    import nas_segm_b200
    from nas_segm_b200 import functional as Fn
    from nas_segm_b200 import packs
    from nas_segm_b200.nn.encoders import mbv2
    from nas_segm_b200.nn.micro_decoders import TemplateDecoder
    from golden_util import W0
    nas_segm_b200.set_act_dtype(torch.bfloat16)
    try:
        enc = mbv2(return_layers=[1, 2])
        dec = TemplateDecoder(list(enc.out_sizes), 19, W0, agg_size=64, repeats=2)
        net = nn.Sequential(enc, dec).cuda()
        convs = [m for m in net.modules() if isinstance(m, nn.Conv2d) and m.groups == 1]
        with packs.scope(net):
            n = 0
            for m in convs:
                w = m.weight
                if w.shape[2:] == (1, 1):
                    a, at = packs.get(w, packs.PW), packs.get(w, packs.PW_T)
                    assert a is not None and at is not None
                    packs.end()
                    b, bt = Fn._pack_weight(w, False), Fn._pack_weight(w, True)
                    packs.begin(net)
                    assert torch.equal(a.view(torch.int16), b.view(torch.int16)) and torch.equal(at.view(torch.int16), bt.view(torch.int16))
                    n += 1
                elif w.shape[2:] == (3, 3) and w.shape[1] != 3:
                    a, at = packs.get(w, packs.C3), packs.get(w, packs.C3_T)
                    assert a is not None and at is not None
                    packs.end()
                    b, bt = Fn._pack_conv3(w, 0), Fn._pack_conv3(w, 1)
                    packs.begin(net)
                    assert torch.equal(a.view(torch.int16), b.view(torch.int16)) and torch.equal(at.view(torch.int16), bt.view(torch.int16))
                    n += 1
            assert n > 40
            # a weight update between scopes is picked up by the next begin()
            w = convs[-2].weight
            with torch.no_grad():
                w.add_(1.0)
            packs.begin(net)
            a = packs.get(w, packs.PW)
            packs.end()
            assert torch.equal(a.view(torch.int16), Fn._pack_weight(w, False).view(torch.int16))
        assert packs.get(convs[0].weight, packs.PW) is None  # no scope, no persistent operand
    finally:
        packs.end()
        nas_segm_b200.set_act_dtype(torch.float32)
