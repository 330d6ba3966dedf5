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
        nas_segm_b200.config().graph_warmup = 3
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
