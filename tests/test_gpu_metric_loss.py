"""GPU parity of the metric (bit-exact integer confusion matrix, reward within 1e-5) and of the fused losses."""
import numpy as np
import pytest
import torch

import nas_segm_b200
from golden_util import rel_err, t
from oracle import miou_oracle as MO
from oracle import nas_oracle as O

pytestmark = pytest.mark.gpu


def test_fast_cm_dropin_bit_exact(golden):
    from nas_segm_b200.helpers.miou_utils import compute_iu, compute_ius_accs, fast_cm
    fx = golden("metric")
    for case, C in enumerate((21, 19, 11, 2, 5)):
        p, g = fx["cm%d/p" % case], fx["cm%d/g" % case]
        cm = fast_cm(p, g, C)
        assert cm.dtype == np.int64 and cm.shape == (C, C)
        assert np.array_equal(cm, fx["cm%d/cm" % case]), case
    for case in range(3):
        cm = fx["v%d/cm" % case]
        iu, npx, acc = compute_ius_accs(cm)
        assert iu.dtype == np.float64 and npx.dtype == np.int64
        assert np.array_equal(iu, fx["v%d/ious" % case])
        assert np.array_equal(npx, fx["v%d/npx" % case])
        assert np.array_equal(acc, fx["v%d/accs" % case])
        assert np.array_equal(compute_iu(cm), fx["v%d/iu_only" % case])


def test_fast_cm_edge_cases():
    from nas_segm_b200.helpers.miou_utils import compute_ius_accs, fast_cm
    # empty, single element, unaligned tails, all one class, C > shared-memory limit, values >= C are skipped
    assert np.array_equal(fast_cm(np.zeros(0, np.uint8), np.zeros(0, np.uint8), 5), np.zeros((5, 5), np.int64))
    rng = np.random.default_rng(3)
    for n, C in ((1, 2), (15, 3), (17, 21), (4097, 21), (100001, 200), (70000, 64), (70000, 65)):
        p = rng.integers(0, C, n).astype(np.uint8)
        g = rng.integers(0, C, n).astype(np.uint8)
        assert np.array_equal(fast_cm(p, g, C), MO.fast_cm_c(p, g, C)), (n, C)
    p = rng.integers(0, 19, 5000).astype(np.uint8)
    g = rng.integers(0, 25, 5000).astype(np.uint8)
    g[::9] = 255
    keep = g < 19
    assert np.array_equal(fast_cm(p, g, 19), MO.fast_cm_c(p[keep], g[keep], 19))
    # device tensors in -> device tensors out, misaligned views
    pt, gt = torch.from_numpy(p).cuda(), torch.from_numpy(g).cuda()
    cm = fast_cm(pt[3:], gt[3:], 19)
    assert cm.is_cuda and np.array_equal(cm.cpu().numpy(), MO.fast_cm_c(p[3:][keep[3:]], g[3:][keep[3:]], 19))
    # uint32 wrap of the reference's accumulators
    big = np.array([[2 ** 32 + 5, 7], [3, 2 ** 33 + 11]], dtype=np.int64)
    for a, b in zip(compute_ius_accs(big), MO.compute_ius_accs_c(big)):
        assert np.array_equal(a, b)


def test_confmat_full_size_properties():
    """2 Mi pixels (one 2048x1024 label pair): totals, marginals and equality with the oracle."""
    from nas_segm_b200 import functional as Fn
    rng = np.random.default_rng(11)
    n = 2048 * 1024
    for C in (11, 19, 21):
        p = rng.integers(0, C, n).astype(np.uint8)
        g = rng.integers(0, C, n).astype(np.uint8)
        g[rng.random(n) < 0.05] = 255
        cm = Fn.confmat_labels(torch.from_numpy(p).cuda(), torch.from_numpy(g).cuda(), C).cpu().numpy()
        keep = g < C
        assert cm.sum() == keep.sum()
        assert np.array_equal(cm.sum(1), np.bincount(g[keep], minlength=C))
        assert np.array_equal(cm.sum(0), np.bincount(p[keep], minlength=C))
        assert np.array_equal(cm, MO.fast_cm_np(p[keep], g[keep], C))
        # accumulation across calls == one call on the concatenation (checksum of checksums)
        acc = torch.zeros((C, C), dtype=torch.int64, device="cuda")
        for part in np.array_split(np.arange(n), 7):
            Fn.confmat_labels(torch.from_numpy(p[part]).cuda(), torch.from_numpy(g[part]).cuda(), C, acc)
        assert np.array_equal(acc.cpu().numpy(), cm)


def test_validate_fused_vs_reference_fixture(golden):
    """inference.py:58-91 from logits: fused upsample+argmax+mask+histogram vs the real reference's confusion matrix and
    reward (fixtures include exact ties -> first-index arg-max, label values >= C, 255, an absent class)."""
    from nas_segm_b200.engine.inference import validate
    fx = golden("metric")

    class Table(torch.nn.Module):
        def __init__(self, logits):
            super().__init__()
            self.logits, self.i = logits, 0

        def forward(self, x):
            self.i += 1
            return self.logits[self.i - 1]

    class Loader(list):
        class _DS:
            def set_stage(self, s):
                pass
        dataset = _DS()

    for case in range(3):
        C = int(fx["v%d/C" % case])
        from nas_segm_b200 import functional as Fn
        cm = torch.zeros((C, C), dtype=torch.int64, device="cuda")
        logits, loader = [], Loader()
        for b in range(2):
            lg = t(fx["v%d/logits%d" % (case, b)]).cuda()
            tg = t(fx["v%d/target%d" % (case, b)])
            Fn.confmat_logits(lg, tg.cuda(), C, cm)
            logits.append(lg)
            loader.append({"image": torch.zeros(lg.shape[0], 3, 4, 4), "mask": tg})
        assert np.array_equal(cm.cpu().numpy(), fx["v%d/cm" % case]), case
        reward = validate(Table(logits), loader, 0, 0, num_classes=C, omit_classes=[0])
        assert abs(reward - float(fx["v%d/reward" % case])) < 1e-5


def test_confmat_logits_exact_arithmetic_full_size():
    """Full-size fused path, bit-exact by construction: logits are multiples of 1/64 and the x4 up-sampling weights are
    multiples of 1/8, so every implementation of the bilinear formula rounds identically."""
    from nas_segm_b200 import functional as Fn
    rng = np.random.default_rng(2)
    B, C, h, w = 2, 19, 256, 128
    lg = (rng.integers(-512, 512, (B, C, h, w)) / 64.0).astype(np.float32)
    gt = rng.integers(0, C + 2, (B, 4 * h, 4 * w)).astype(np.uint8)
    gt[rng.random(gt.shape) < 0.05] = 255
    up = torch.nn.functional.interpolate(torch.from_numpy(lg), size=(4 * h, 4 * w), mode="bilinear", align_corners=False)
    ref = MO.cm_from_logits(up.numpy(), gt, C)
    for dtype in (torch.float32, torch.bfloat16):
        # bf16 holds multiples of 1/64 below 4 exactly only with <= 8 significant bits: use a coarser grid for it
        src = lg if dtype == torch.float32 else (np.round(lg * 4) / 4).astype(np.float32)
        if dtype == torch.bfloat16:
            up = torch.nn.functional.interpolate(torch.from_numpy(src), size=(4 * h, 4 * w), mode="bilinear",
                                                 align_corners=False)
            ref_d = MO.cm_from_logits(up.numpy(), gt, C)
        else:
            ref_d = ref
        cm = Fn.confmat_logits(torch.from_numpy(src).cuda().to(dtype), torch.from_numpy(gt).cuda(), C)
        assert np.array_equal(cm.cpu().numpy(), ref_d), dtype
    r_mine = MO.reward_from_cm(cm.cpu().numpy())[0]
    assert abs(r_mine - MO.reward_from_cm(ref_d)[0]) < 1e-12


def test_cross_entropy_forward_backward():
    from nas_segm_b200 import functional as Fn
    rng = np.random.default_rng(4)
    for (B, C, H, W) in ((2, 21, 33, 47), (3, 19, 16, 16), (1, 1, 8, 8), (2, 11, 5, 7)):
        lg = rng.normal(0, 2, (B, C, H, W)).astype(np.float32)
        y = rng.integers(0, C, (B, H, W))
        y[rng.random(y.shape) < 0.1] = 255
        xo = torch.from_numpy(lg).requires_grad_(True)
        lo = O.segm_loss(xo, torch.from_numpy(y))
        lo.backward()
        x = torch.from_numpy(lg).cuda().requires_grad_(True)
        l = Fn.cross_entropy2d(x, torch.from_numpy(y).cuda(), 255)
        (3.0 * l).backward()
        assert abs(float(l) - float(lo)) < 1e-5 * max(1.0, abs(float(lo)))
        assert rel_err(x.grad.cpu().numpy(), 3.0 * xo.grad.numpy()) < 1e-4
    # every pixel ignored -> nan like torch
    x = torch.zeros(1, 3, 4, 4, device="cuda")
    assert torch.isnan(Fn.cross_entropy2d(x, torch.full((1, 4, 4), 255, device="cuda"), 255))


def test_task0_loss_composition():
    """trainer.py:137-158: CE(up(out)) + kd*MSE(up(out), kd_y) + aux*sum CE(up(aux))."""
    from nas_segm_b200 import functional as Fn
    rng = np.random.default_rng(6)
    out = rng.normal(0, 1, (2, 21, 9, 9)).astype(np.float32)
    aux = [rng.normal(0, 1, (2, 21, s, s)).astype(np.float32) for s in (3, 5, 9)]
    y = rng.integers(0, 21, (2, 16, 16))
    y[:, ::3, ::4] = 255
    kd = rng.normal(0, 1, (2, 21, 16, 16)).astype(np.float32)
    to = [torch.from_numpy(a).requires_grad_(True) for a in [out] + aux]
    lo = O.task0_loss(to[0], to[1:], torch.from_numpy(y), (16, 16), torch.from_numpy(kd), 0.3, 0.15)
    lo.backward()
    tg = [torch.from_numpy(a).cuda().requires_grad_(True) for a in [out] + aux]
    yy = torch.from_numpy(y).cuda()
    up = Fn.resize(Fn.lib.to_nhwc(tg[0]), (16, 16))
    l = Fn.cross_entropy2d(up, yy) + 0.3 * Fn.mse_loss(up, torch.from_numpy(kd).cuda())
    for a in tg[1:]:
        l = l + 0.15 * Fn.cross_entropy2d(Fn.resize(Fn.lib.to_nhwc(a), (16, 16)), yy)
    l.backward()
    assert abs(float(l) - float(lo)) < 1e-5 * max(1.0, abs(float(lo)))
    for a, b in zip(tg, to):
        assert rel_err(a.grad.cpu().numpy(), b.grad.numpy()) < 1e-4


def test_berhu_forward_backward():
    from nas_segm_b200 import functional as Fn
    rng = np.random.default_rng(8)
    pred = rng.uniform(0.2, 9.0, (2, 1, 30, 40)).astype(np.float32)
    tgt = rng.uniform(0.5, 8.0, (2, 1, 30, 40)).astype(np.float32)
    tgt[rng.random(tgt.shape) < 0.05] = 0.0
    po = torch.from_numpy(pred).requires_grad_(True)
    lo = O.berhu_loss(po, torch.from_numpy(tgt))
    lo.backward()
    p = torch.from_numpy(pred).cuda().requires_grad_(True)
    l = Fn.berhu_loss(p, torch.from_numpy(tgt).cuda())
    l.backward()
    assert abs(float(l) - float(lo)) < 1e-5 * abs(float(lo))
    assert rel_err(p.grad.cpu().numpy(), po.grad.numpy()) < 1e-4
