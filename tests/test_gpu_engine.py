"""Engine API on the GPU: train_task0 against the real reference's two-epoch Adam trajectory (fixture), the RuntimeError
-> 0 convention, populate_task0 -> train_task0 -> validate end to end."""
import logging
import types

import numpy as np
import pytest
import torch
from torch import nn

import nas_segm_b200
from detweights import det_array
from golden_util import C0, det_state_dict, keys_shapes, t

pytestmark = pytest.mark.gpu


class Seg(nn.Module):
    def __init__(self, enc, dec):
        super().__init__()
        self.encoder, self.decoder = enc, dec

    def forward(self, x):
        return self.decoder(self.encoder(x))


@pytest.mark.parametrize("graphs", [False, True])
def test_train_task0_matches_reference_trajectory(golden, graphs):
    """graphs=True: the same trajectory with every iteration after the first replayed from one captured CUDA graph."""
    from nas_segm_b200.engine import trainer
    nas_segm_b200.config().cuda_graphs = graphs
    nas_segm_b200.config().graph_warmup = 1
    from nas_segm_b200.nn.encoders import mbv2
    from nas_segm_b200.nn.micro_decoders import MicroDecoder
    nas_segm_b200.set_act_dtype(torch.float32)
    fx = golden("task0_step")
    np.random.seed(0)
    enc = mbv2()
    dec = MicroDecoder(list(enc.out_sizes), 21, C0, agg_size=16, aux_cell=True, repeats=1)
    dec.load_state_dict(det_state_dict(keys_shapes(fx), seed=11), strict=True)
    seg = types.SimpleNamespace(module=Seg(enc, dec).cuda())
    N, B = 8, 4
    Xy = {}
    for i, s in enumerate([(24, 16, 16), (32, 8, 8), (96, 4, 4), (320, 2, 2)]):
        Xy[i] = t(det_array("t0/f%d" % i, (N,) + s)).abs().cuda()
    y = t(det_array("t0/y", (N, 16, 16), kind="int", lo=0, hi=21))
    y[:, ::4, ::3] = 255
    Xy["y"] = y.cuda()
    Xy["kd_y"] = t(det_array("t0/kd", (N, 21, 16, 16))).cuda()
    Xy["out_size"] = torch.Size((16, 16))
    optim = torch.optim.Adam(dec.parameters(), lr=3e-3, weight_decay=1e-5)
    avg = [p.data.clone() for p in dec.parameters()]
    losses = []
    orig = trainer.logger.info
    trainer.logger.info = lambda msg, *a: losses.append(float(msg.split("Avg. Loss:")[1].split()[0]))
    try:
        for epoch in range(2):
            r = trainer.train_task0(Xy, seg, optim, epoch, nn.NLLLoss(ignore_index=255), nn.MSELoss(), B, False, True, 0.3,
                                    3.0, True, avg_param=avg, polyak_decay=0.9, aux_weight=0.15)
            assert r is None
    finally:
        trainer.logger.info = orig
        nas_segm_b200.config().cuda_graphs = False
        nas_segm_b200.config().graph_warmup = 1
    if graphs:
        assert dec._nasb_task0_graph.graph is not None and dec._nasb_task0_graph.calls == 4
    assert np.allclose(losses, fx["logged_avg_loss"], atol=2e-3), (losses, fx["logged_avg_loss"])
    sd = dec.state_dict()
    tot = ok = 0
    for k in [k for k in fx.files if k.startswith("post/")]:
        a, b = sd[k[5:]].float().cpu().numpy(), fx[k]
        if a.dtype.kind != "f" or b.dtype.kind != "f":
            assert np.array_equal(a, b), k
            continue
        tot += a.size
        ok += (np.abs(a - b) <= 1e-3 + 1e-3 * np.abs(b)).sum()
    # Adam turns near-zero gradients into +-lr steps, so a handful of coordinates may legitimately differ
    assert ok / tot > 0.995, ok / tot
    for (pn, _), a in zip(dec.named_parameters(), avg):
        b = fx["avg/" + pn]
        assert (np.abs(a.cpu().numpy() - b) <= 1e-3 + 1e-3 * np.abs(b)).mean() > 0.99, pn


def test_engine_runtime_error_returns_zero():
    from nas_segm_b200.engine import trainer
    bad = {0: torch.zeros(4, 24, 8, 8), "y": torch.zeros(4, 8, 8).long(), "out_size": (8, 8)}  # CPU tensors -> RuntimeError

    class Dec(nn.Module):
        def forward(self, x):
            raise RuntimeError("CUDA out of memory (simulated)")

    seg = types.SimpleNamespace(module=types.SimpleNamespace(decoder=Dec(), encoder=None))
    assert trainer.train_task0(bad, seg, None, 0, None, None, 2, False, False, 0.0, 3.0, False) == 0


def test_populate_train_validate_end_to_end():
    """A tiny WACV-style candidate through populate_task0 -> train_task0 -> train_segmenter -> validate."""
    from nas_segm_b200.engine import inference, trainer
    from nas_segm_b200.nn.encoders import mbv2
    from nas_segm_b200.nn.micro_decoders import MicroDecoder
    nas_segm_b200.set_act_dtype(torch.float32)
    torch.manual_seed(0)
    np.random.seed(0)
    enc = mbv2()
    dec = MicroDecoder(list(enc.out_sizes), 5, C0, agg_size=16, aux_cell=True, repeats=1)
    seg = nn.DataParallel(Seg(enc, dec).cuda(), device_ids=[0])

    class DS:
        def set_stage(self, s):
            pass

    class Loader(list):
        dataset = DS()
        batch_sampler = types.SimpleNamespace(batch_size=2)

    g = torch.Generator().manual_seed(1)
    def sample(b):
        img = torch.randn(b, 3, 64, 64, generator=g)
        m = (img[:, 0] > 0).long() + 2 * (img[:, 1] > 0.5).long()  # labels correlated with the image
        return {"image": img.double(), "mask": m.to(torch.uint8)}
    train = Loader(sample(1) for _ in range(12))
    Xy = trainer.populate_task0(seg, train, None, 12, do_kd=False)
    assert Xy != 0 and Xy[0].shape[0] == 12 and tuple(Xy["out_size"]) == (16, 16)
    optim_dec = torch.optim.Adam(seg.module.decoder.parameters(), lr=3e-3)
    optim_enc = torch.optim.SGD(seg.module.encoder.parameters(), lr=1e-3, momentum=0.9)
    crit = nn.NLLLoss(ignore_index=255)
    msgs = []
    h = logging.Handler()
    h.emit = lambda rec: msgs.append(rec.getMessage())
    trainer.logger.addHandler(h)
    trainer.logger.setLevel(logging.INFO)
    for ep in range(6):
        assert trainer.train_task0(Xy, seg, optim_dec, ep, crit, None, 4, False, False, 0.0, 3.0, False, aux_weight=0.15) is None
    losses = [float(m.split("Avg. Loss:")[1].split()[0]) for m in msgs if "Avg. Loss" in m]
    assert losses[-1] < losses[0], losses
    big = Loader(sample(4) for _ in range(3))
    assert trainer.train_segmenter(seg, big, optim_enc, optim_dec, 0, crit, False, 3.0, 3.0, False, aux_weight=0.15) is None
    reward = inference.validate(seg, Loader(sample(4) for _ in range(2)), 0, 0, num_classes=5, omit_classes=[0])
    assert isinstance(reward, float) and 0.0 <= reward <= 1.0


def _check_fixture_tensor(fx, key, a):
    """`a` against a stored tensor, or against the strided sample + sums of a large one (make_golden.gen_segmenter)."""
    a = a.detach().float().cpu().numpy()
    if key in fx.files:
        b = fx[key]
        if b.dtype.kind != "f":
            assert np.array_equal(a, b), key
            return 1.0, a.size
        return float((np.abs(a - b) <= 1e-3 + 1e-3 * np.abs(b)).mean()), a.size
    flat = a.reshape(-1)
    smp = flat[:: max(flat.size // 512, 1)][:512]
    b = fx[key + "@sample"]
    sums = fx[key + "@sums"]
    # Adam turns near-zero gradients into +-lr steps: a few coordinates may legitimately end a whole step apart, so the
    # sums are held to a band of that size (4 steps of lr 3e-3 on 1 % of the coordinates), the sample to the 1e-3 band
    assert abs(float(flat.astype(np.float64).sum()) - sums[0]) <= 1e-3 * (abs(sums[0]) + np.sqrt(sums[1])) + 1.2e-4 * flat.size, key
    assert abs(float((flat.astype(np.float64) ** 2).sum()) - sums[1]) <= 5e-3 * sums[1], key
    return float((np.abs(smp - b) <= 1e-3 + 1e-3 * np.abs(b)).mean()), smp.size


@pytest.mark.parametrize("mode", ["eager", "graph", "eager_torch_optim"])
def test_train_segmenter_matches_reference_trajectory(golden, mode):
    """Row a21: two epochs x two iterations of the UNMODIFIED reference train_segmenter (SGD encoder + Adam decoder, both
    clips, Polyak, aux heads, training-mode BN; fixture tests/golden/segmenter_step.npz) against this engine: eager with the
    fused multi-tensor optimiser step, the same as one replayed CUDA graph, and eager with the caller's torch optimisers."""
    from nas_segm_b200.engine import trainer
    from nas_segm_b200.nn.encoders import mbv2
    from nas_segm_b200.nn.micro_decoders import MicroDecoder
    fx = golden("segmenter_step")
    cfg = nas_segm_b200.config()
    nas_segm_b200.set_act_dtype(torch.float32)
    cfg.cuda_graphs, cfg.graph_warmup, cfg.fused_optim = mode == "graph", 1, mode != "eager_torch_optim"
    try:
        enc = mbv2()
        dec = MicroDecoder(list(enc.out_sizes), 21, C0, agg_size=16, aux_cell=True, repeats=1)
        enc.load_state_dict(det_state_dict(keys_shapes(fx, "enc_"), seed=21), strict=True)
        dec.load_state_dict(det_state_dict(keys_shapes(fx, "dec_"), seed=22), strict=True)

        class Wrap(nn.Module):
            def __init__(self, m):
                super().__init__()
                self.module = m

            def forward(self, x):
                return self.module(x)
        seg = Wrap(Seg(enc, dec)).cuda()

        class Loader(list):
            class _DS:
                def set_stage(self, s):
                    pass
            dataset = _DS()
        loader = Loader()
        for i in range(2):
            msk = det_array("seg/msk%d" % i, (4, 192, 192), kind="int", lo=0, hi=21).astype(np.uint8)
            msk[:, ::5, ::3] = 255
            loader.append({"image": t(det_array("seg/img%d" % i, (4, 3, 192, 192)).astype(np.float64)), "mask": t(msk)})
        optim_enc = torch.optim.SGD(enc.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-5)
        optim_dec = torch.optim.Adam(dec.parameters(), lr=3e-3, weight_decay=1e-5)
        avg = [p.data.clone() for p in seg.parameters()]
        losses = []
        orig = trainer.logger.info
        trainer.logger.info = lambda msg, *a: losses.append(float(msg.split("Avg. Loss:")[1].split()[0]))
        try:
            for epoch in range(2):
                # __wrapped__: the engine function without the RuntimeError -> 0 wrapper, so a failure shows its cause
                r = trainer.train_segmenter.__wrapped__(seg, loader, optim_enc, optim_dec, epoch, nn.NLLLoss(ignore_index=255),
                                                        False, 3.0, 3.0, True, print_every=1, aux_weight=0.15, avg_param=avg,
                                                        polyak_decay=0.99)
                assert r is None
        finally:
            trainer.logger.info = orig
        if mode == "graph":
            assert seg._nasb_step_graph.graph is not None and seg._nasb_step_graph.calls == 4
        assert np.allclose(losses, fx["logged_avg_loss"], atol=3e-3), (losses, fx["logged_avg_loss"])
        sd = seg.module.state_dict()
        ok = tot = 0
        for k in sd:
            frac, n = _check_fixture_tensor(fx, "post/" + k, sd[k])
            ok, tot = ok + frac * n, tot + n
        assert ok / tot > 0.99, ok / tot
        ok = tot = 0
        for (pn, _), a in zip(seg.named_parameters(), avg):
            frac, n = _check_fixture_tensor(fx, "avg/" + pn, a)
            ok, tot = ok + frac * n, tot + n
        assert ok / tot > 0.99, ok / tot
    finally:
        cfg.cuda_graphs, cfg.graph_warmup, cfg.fused_optim = False, 1, True


def test_train_segmenter_graph_is_recaptured_when_what_it_bakes_in_changes():
    """the captured iteration bakes in BN mode, clips, aux weight, Polyak, ignore_index and the optimisers'
    hyper-parameters; a call that changes any of them must not replay the old graph."""
    from nas_segm_b200.engine import trainer
    from nas_segm_b200.nn.encoders import mbv2
    from nas_segm_b200.nn.micro_decoders import TemplateDecoder
    from golden_util import W0
    cfg = nas_segm_b200.config()
    nas_segm_b200.set_act_dtype(torch.float32)
    cfg.cuda_graphs, cfg.graph_warmup = True, 1
    try:
        torch.manual_seed(0)
        enc = mbv2(return_layers=[1, 2])
        dec = TemplateDecoder(list(enc.out_sizes), 5, W0, agg_size=16, repeats=1)
        seg = types.SimpleNamespace(module=Seg(enc, dec).cuda())
        seg_mod = nn.Module()
        seg_mod.module = seg.module
        seg_mod.forward = lambda x: seg.module(x)

        class Loader(list):
            class _DS:
                def set_stage(self, s):
                    pass
            dataset = _DS()
        g = torch.Generator().manual_seed(3)
        loader = Loader({"image": torch.randn(2, 3, 64, 64, generator=g), "mask": torch.randint(0, 5, (2, 64, 64), generator=g).to(torch.uint8)}
                        for _ in range(3))
        oe = torch.optim.SGD(enc.parameters(), lr=1e-3, momentum=0.9)
        od = torch.optim.Adam(dec.parameters(), lr=3e-3)
        crit = nn.NLLLoss(ignore_index=255)

        def run(**kw):
            a = dict(freeze_bn=False, enc_grad_clip=3.0, dec_grad_clip=3.0)
            a.update(kw)
            assert trainer.train_segmenter(seg_mod, loader, oe, od, 0, crit, a["freeze_bn"], a["enc_grad_clip"],
                                           a["dec_grad_clip"], False) is None
            return seg_mod._nasb_step_graph
        g0 = run()
        assert run() is g0                       # identical call: replay
        assert run(dec_grad_clip=1.0) is not g0  # clip norm is a kernel argument of the captured optimiser step
        g1 = run()
        od.param_groups[0]["lr"] = 1e-3
        g2 = run()
        assert g2 is not g1                      # learning rate changed
        w_before = dec.conv_clf.weight.detach().clone()
        g3 = run(freeze_bn=True)
        assert g3 is not g2                      # BatchNorm mode changed
        assert not torch.equal(w_before, dec.conv_clf.weight)
    finally:
        cfg.cuda_graphs, cfg.graph_warmup = False, 1


def test_search_round_on_gpu():
    """Row f1 on hardware: engine.search.search_rounds + evaluate_candidate (task 0 AND task 1, TaskPerformer, per-candidate
    rebuild, CUDA graphs, bf16) for two rounds of sampled candidates on one rank."""
    from nas_segm_b200.engine import search, trainer
    from nas_segm_b200.nn.encoders import mbv2
    from nas_segm_b200.nn.micro_decoders import MicroDecoder
    cfg = nas_segm_b200.config()
    nas_segm_b200.set_act_dtype(torch.bfloat16)
    cfg.cuda_graphs, cfg.graph_warmup = True, 1
    np.random.seed(0)
    try:
        class Wrap(nn.Module):
            def __init__(self, m):
                super().__init__()
                self.module = m

            def forward(self, x):
                return self.module(x)

        class Loader(list):
            class _DS:
                def set_stage(self, s):
                    pass
            dataset = _DS()
            batch_sampler = types.SimpleNamespace(batch_size=1)
        g = torch.Generator().manual_seed(1)

        def sample(b, s):
            img = torch.randn(b, 3, s, s, generator=g)
            return {"image": img, "mask": ((img[:, 0] > 0).long() + 2 * (img[:, 1] > 0.5).long()).to(torch.uint8)}
        seg0 = Wrap(Seg(mbv2(), nn.Identity()).cuda())
        Xy = trainer.populate_task0(seg0, Loader(sample(1, 128) for _ in range(16)), None, 16, do_kd=False)
        assert Xy != 0
        Xy["kd_y"] = torch.randn(16, 32, 32, 5, device="cuda").permute(0, 3, 1, 2)
        args = types.SimpleNamespace(
            num_tasks=2, enc_optim="sgd", dec_optim="adam", enc_lr=[1e-3, 1e-3], dec_lr=[3e-3, 3e-3], enc_mom=[0.9] * 2,
            dec_mom=[0.9] * 2, enc_wd=[1e-5] * 2, dec_wd=[1e-5] * 2, do_polyak=True, num_segm_epochs=[2, 1], val_every=[2, 1],
            segm_crit=nn.NLLLoss(ignore_index=255), kd_crit=nn.MSELoss(), batch_size=[8, 4], freeze_bn=[False, False],
            do_kd=True, kd_coeff=0.3, dec_grad_clip=3.0, enc_grad_clip=3.0, dec_aux_weight=0.15, print_every=20,
            num_classes=[5, 5], val_omit_classes=[0])
        task_ps = search.make_task_performers(args.num_segm_epochs, args.val_every)
        train1, val = Loader(sample(4, 96) for _ in range(3)), Loader(sample(4, 96) for _ in range(2))
        built, seen = [], []

        def build(c):
            torch.manual_seed(0)
            enc = mbv2()
            built.append(c)
            return Wrap(Seg(enc, MicroDecoder(list(enc.out_sizes), 5, c, agg_size=16, aux_cell=True, repeats=1)).cuda())
        hist = search.search_rounds(2, search.uniform_sampler(seed=7), build,
                                    lambda seg, c: search.evaluate_candidate(seg, Xy, train1, val, args, task_ps),
                                    update_fn=seen.append, sync_every=2)
        assert len(hist) == 2 and len(built) == 2 and len(seen) == 2 and built[0] != built[1]
        for tb in hist:
            assert tb.shape == (1, 4) and 0.0 <= float(tb[0, 0]) <= 1.0 and float(tb[0, 3]) > 1e4  # reward, #params
        assert all(len(u) == 1 and u[0][0] in built for u in seen)
    finally:
        cfg.cuda_graphs, cfg.graph_warmup = False, 1
        nas_segm_b200.set_act_dtype(torch.float32)
