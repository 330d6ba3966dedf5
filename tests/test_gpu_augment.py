"""Row f3: the batch augmentation kernel (csrc/augment.cu behind nas_segm_b200.data) against the fixture written by the
reference's own transform classes (tests/golden/make_golden_augment.py) and against oracle/augment_oracle.py on larger
random inputs.  Everything is compared for equality: uint8 resize results, masks and the float32 network input."""
import os

import numpy as np
import pytest
import torch

from oracle import augment_oracle as A

FX = np.load(os.path.join(os.path.dirname(__file__), "golden", "augment.npz"))
NORM = (1.0 / 255, np.array([0.485, 0.456, 0.406]).reshape((1, 1, 3)), np.array([0.229, 0.224, 0.225]).reshape((1, 1, 3)))
ONORM = (1.0 / 255, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
N_IMG = 6


def _samples():
    return [{"image": FX["image%d" % i], "mask": FX["mask%d" % i]} for i in range(N_IMG)]


def test_param_draws_follow_the_reference_order():
    """Host logic only (runs without a GPU): same seed, same (scale, mirror, top, left) as the oracle's restatement of the
    reference's draws."""
    from nas_segm_b200.data import GpuTrainTransform
    t = GpuTrainTransform.__new__(GpuTrainTransform)
    t.resize_side, t.low_scale, t.high_scale, t.longer, t.crop_size = 70, 0.5, 2.0, False, 56
    np.random.seed(5)
    mine = [t._params(h, w) for h, w in ((61, 83), (90, 64), (75, 75), (200, 31))]
    np.random.seed(5)
    want = [A.draw_train_params(h, w, 70, 0.5, 2.0, False, 57) for h, w in ((61, 83), (90, 64), (75, 75), (200, 31))]
    for m, o in zip(mine, want):
        assert m[:6] == (o["scale"], o["mirror"], o["top"], o["left"], o["out_h"], o["out_w"])


@pytest.mark.gpu
def test_training_chains_match_the_reference_fixture():
    from nas_segm_b200.data import GpuTrainTransform
    for name in ("trn_a", "trn_b", "trn_c"):
        side, low, high, longer, crop, seed = FX[name + "_cfg"]
        t = GpuTrainTransform(int(side), float(low), float(high), bool(longer), int(crop), NORM)
        np.random.seed(int(seed))
        for i, s in enumerate(_samples()):  # one sample per call: the fixture's crops may be ragged, the draws are sequential
            out = t([s])
            torch.cuda.synchronize()
            want_i, want_m = FX["%s_image%d" % (name, i)], FX["%s_mask%d" % (name, i)]
            assert tuple(out["image"].shape[1:]) == want_i.shape, (name, i)
            assert np.array_equal(out["image"][0].cpu().numpy(), want_i), (name, i)
            assert np.array_equal(out["mask"][0].cpu().numpy(), want_m), (name, i)
    # the same chain as ONE batch (all crops of trn_a are 48 x 48): one launch, same draws
    side, low, high, longer, crop, seed = FX["trn_a_cfg"]
    np.random.seed(int(seed))
    out = GpuTrainTransform(int(side), float(low), float(high), bool(longer), int(crop), NORM)(_samples())
    assert out["image"].dtype == torch.float32 and out["mask"].dtype == torch.uint8
    for i in range(N_IMG):
        assert np.array_equal(out["image"][i].cpu().numpy(), FX["trn_a_image%d" % i])
        assert np.array_equal(out["mask"][i].cpu().numpy(), FX["trn_a_mask%d" % i])


@pytest.mark.gpu
def test_validation_chain_matches_the_reference_fixture():
    from nas_segm_b200.data import GpuValTransform
    side, _, _, longer, crop, seed = FX["val_cfg"]
    out = GpuValTransform(int(side), bool(longer), int(crop), NORM)(_samples())
    for i in range(N_IMG):
        assert np.array_equal(out["image"][i].cpu().numpy(), FX["val_image%d" % i]), i
        assert np.array_equal(out["mask"][i].cpu().numpy(), FX["val_mask%d" % i]), i
    with pytest.raises(ValueError):  # crop larger than the resized image: the reference's negative margins are not reproduced
        GpuValTransform(40, False, 64, NORM)(_samples())


@pytest.mark.gpu
def test_dataset_sized_batch_against_the_oracle():
    """VOC-sized images, the search loop's task-1 geometry (resize_side 400, scales 0.7-1.4, crop 350), device-resident and
    host inputs mixed, more samples than one launch table holds (64)."""
    from nas_segm_b200.data import GpuTrainTransform
    rs = np.random.RandomState(11)
    big = []
    for h, w in ((375, 500), (500, 333), (281, 500), (366, 500)):
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        img[:, : w // 3] = (img[:, : w // 3] // 32) * 32
        big.append({"image": img, "mask": rs.randint(0, 22, (h, w)).astype(np.uint8)})
    big[1] = {"image": torch.from_numpy(big[1]["image"]).cuda(), "mask": torch.from_numpy(big[1]["mask"]).cuda()}
    t = GpuTrainTransform(400, 0.7, 1.4, False, 350, NORM)
    np.random.seed(3)
    out = t(big)
    np.random.seed(3)
    for i, s in enumerate(big):
        img = s["image"].cpu().numpy() if torch.is_tensor(s["image"]) else s["image"]
        msk = s["mask"].cpu().numpy() if torch.is_tensor(s["mask"]) else s["mask"]
        p = A.draw_train_params(img.shape[0], img.shape[1], 400, 0.7, 1.4, False, 350)
        want_i, want_m = A.apply(img, msk, p, *ONORM)
        assert np.array_equal(out["image"][i].cpu().numpy(), want_i), i
        assert np.array_equal(out["mask"][i].cpu().numpy(), want_m), i
    small = [{"image": rs.randint(0, 256, (20 + i % 7, 24 + i % 5, 3)).astype(np.uint8),
              "mask": rs.randint(0, 21, (20 + i % 7, 24 + i % 5)).astype(np.uint8)} for i in range(70)]
    t = GpuTrainTransform(30, 0.9, 1.6, False, 16, NORM)
    np.random.seed(4)
    out = t(small)
    np.random.seed(4)
    for i, s in enumerate(small):
        p = A.draw_train_params(s["image"].shape[0], s["image"].shape[1], 30, 0.9, 1.6, False, 16)
        want_i, want_m = A.apply(s["image"], s["mask"], p, *ONORM)
        assert np.array_equal(out["image"][i].cpu().numpy(), want_i), i
        assert np.array_equal(out["mask"][i].cpu().numpy(), want_m), i


@pytest.mark.gpu
def test_train_segmenter_from_raw_batches_equals_training_on_the_reference_tensors():
    """The augmentation inside the engine loop: `train_segmenter` fed by an AugmentedLoader over raw uint8 samples ends, weight for weight, where the run fed with the tensors the reference's transform chain (oracle) produces ends."""
    import types

    import torch.nn as nn

    from nas_segm_b200.data import AugmentedLoader, GpuTrainTransform
    from nas_segm_b200.engine import trainer
    from nas_segm_b200.nn.encoders import mbv2
    from nas_segm_b200.nn.micro_decoders import MicroDecoder

    class Seg(nn.Module):
        def __init__(self, enc, dec):
            super().__init__()
            self.encoder, self.decoder = enc, dec

        def forward(self, x):
            return self.decoder(self.encoder(x))

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
        batch_sampler = types.SimpleNamespace(batch_size=4)

    rs = np.random.RandomState(21)
    raw = Loader([{"image": rs.randint(0, 256, (70 + 3 * j, 90 - 2 * j, 3)).astype(np.uint8),
                   "mask": rs.randint(0, 5, (70 + 3 * j, 90 - 2 * j)).astype(np.uint8)} for j in range(4)] for _ in range(3))
    cfg = (80, 0.8, 1.3, False, 64)
    np.random.seed(17)
    ref_batches = Loader()
    for batch in raw:
        ims, mks = [], []
        for s in batch:
            p = A.draw_train_params(s["image"].shape[0], s["image"].shape[1], *cfg)
            a, b = A.apply(s["image"], s["mask"], p, *ONORM)
            ims.append(torch.from_numpy(a))
            mks.append(torch.from_numpy(b))
        ref_batches.append({"image": torch.stack(ims), "mask": torch.stack(mks)})

    def run(loader):
        torch.manual_seed(0)
        enc = mbv2()
        c0 = [[8, [0, 0, 5, 2], [0, 2, 8, 8], [0, 5, 1, 4]], [[3, 3], [3, 2], [3, 0]]]
        seg = Wrap(Seg(enc, MicroDecoder(list(enc.out_sizes), 5, c0, agg_size=16, aux_cell=True, repeats=1)).cuda())
        oe = torch.optim.SGD(seg.module.encoder.parameters(), lr=1e-3, momentum=0.9)
        od = torch.optim.SGD(seg.module.decoder.parameters(), lr=3e-3, momentum=0.9)
        r = trainer.train_segmenter(seg, loader, oe, od, 0, nn.NLLLoss(ignore_index=255), True, 3.0, 3.0, False, print_every=1,
                                    aux_weight=0.15)
        assert r is None
        torch.cuda.synchronize()
        return [q.detach().clone() for q in seg.parameters()]

    np.random.seed(17)
    got = run(AugmentedLoader(raw, GpuTrainTransform(*cfg, NORM)))
    want = run(ref_batches)
    for a, b in zip(got, want):  # identical inputs; the atomically-reduced weight gradients may differ in the last bits
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5), float((a - b).abs().max())
