"""Generate parity fixtures by RUNNING THE REAL REFERENCE (only works where /root/reference exists).

    python tests/golden/make_golden.py

Imports /root/reference/src/{nn,engine} unmodified (plus the alias-patched Cython metric built by
oracle/build_ref.py), fills every module with tests/golden/detweights.py values, and stores
inputs -> outputs (and gradients / post-step parameters) as compressed .npz under tests/golden/.
The fixtures pin oracle/ (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_*.py).
"""
import logging
import os
import sys
import types
import warnings

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("NASB_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from detweights import det_array, det_state_dict, keys_shapes_of  # noqa: E402
from oracle import build_ref  # noqa: E402

# the engine imports helpers.miou_utils (compiled module): provide the built reference extension
_ref_miou = build_ref.load() or build_ref.load() if build_ref.build() else None
assert _ref_miou is not None, "reference Cython metric did not build"
import helpers  # noqa: E402  (reference package)
sys.modules["helpers.miou_utils"] = _ref_miou
helpers.miou_utils = _ref_miou

from nn.encoders import mbv2  # noqa: E402
from nn.layer_factory import AGG_OPS, OPS  # noqa: E402
from nn.micro_decoders import MicroDecoder, TemplateDecoder  # noqa: E402
from engine import trainer as ref_trainer  # noqa: E402
from engine import inference as ref_inference  # noqa: E402

torch.set_num_threads(8)

W0 = [[[3, 0, 1], [4, 1, 1], [3, 1, 1]],
      [[0, 1, 0, 0, 1], [2, 1, 2, 1, 0], [3, 1, 1, 1, 0], [1, 1, 2, 0, 0], [3, 0, 2, 0, 0], [5, 3, 2, 1, 0],
       [0, 5, 0, 1, 0]]]
W1 = [[[1, 1, 0], [1, 3, 0], [3, 4, 0]],
      [[1, 1, 0, 0, 0], [0, 1, 1, 1, 1], [3, 1, 2, 3, 0], [3, 0, 2, 2, 0], [0, 1, 2, 0, 0], [2, 1, 1, 3, 0],
       [4, 0, 2, 2, 0]]]
C0 = [[8, [0, 0, 5, 2], [0, 2, 8, 8], [0, 5, 1, 4]], [[3, 3], [3, 2], [3, 0]]]
C1 = [[2, [1, 0, 3, 6], [0, 1, 2, 8], [2, 0, 6, 1]], [[2, 3], [3, 1], [4, 4]]]
C2 = [[5, [0, 0, 4, 1], [3, 2, 0, 1], [5, 6, 5, 0]], [[1, 3], [4, 3], [2, 2]]]


def fill(module, seed=0):
    ks = keys_shapes_of(module)
    module.load_state_dict(det_state_dict(ks, seed), strict=True)
    return ks


def ks_arrays(ks):
    return {"keys": np.array([k for k, _ in ks]), "shapes": np.array([",".join(map(str, s)) for _, s in ks])}


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("%-28s %8.1f KiB" % (name, os.path.getsize(path) / 1024))


def n(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------ registry ops
def gen_ops():
    out = {}
    cases = []
    for name in OPS:
        for (cin, cout, stride, repeats) in ((8, 8, 1, 1), (8, 16, 2, 2)):
            if name in ("skip_connect", "none") or cin == cout or True:
                cases.append((name, cin, cout, stride, repeats))
    for idx, (name, cin, cout, stride, repeats) in enumerate(cases):
        tag = "op%02d_%s_%d_%d_s%d_r%d" % (idx, name, cin, cout, stride, repeats)
        m = OPS[name](cin, cout, stride, True, repeats)
        ks = fill(m, seed=idx)
        x = torch.from_numpy(det_array(tag + "/x", (2, cin, 13, 17))).requires_grad_(True)
        m.eval()
        y_eval = m(x)
        m.train()
        y_tr = m(x)
        ct = torch.from_numpy(det_array(tag + "/ct", tuple(y_tr.shape)))
        grads = torch.autograd.grad((y_tr * ct).sum(), [x] + [p for p in m.parameters()], allow_unused=True)
        out[tag + "/meta"] = np.array([name, cin, cout, stride, repeats], dtype=object).astype(str)
        out[tag + "/keys"] = ks_arrays(ks)["keys"]
        out[tag + "/shapes"] = ks_arrays(ks)["shapes"]
        out[tag + "/y_eval"] = n(y_eval)
        out[tag + "/y_train"] = n(y_tr)
        out[tag + "/gx"] = n(grads[0]) if grads[0] is not None else np.zeros(tuple(x.shape), np.float32)
        for (pn, _), g in zip(m.named_parameters(), grads[1:]):
            out[tag + "/g/" + pn] = n(g)
        for k, v in m.state_dict().items():  # running stats after one train-mode forward
            if k.endswith("running_mean") or k.endswith("running_var"):
                out[tag + "/post/" + k] = n(v)
    # aggregation ops: both size orders, both `larger` flags, channel adaption on/off
    aggc = []
    for name in AGG_OPS:
        for (c0, c1, cout, larger, s0, s1) in ((8, 8, 8, True, (12, 10), (6, 5)), (8, 16, 16, False, (12, 10), (6, 5)),
                                               (16, 8, 16, True, (5, 7), (10, 14)), (8, 8, 8, False, (9, 9), (9, 9))):
            aggc.append((name, c0, c1, cout, larger, s0, s1))
    for idx, (name, c0, c1, cout, larger, s0, s1) in enumerate(aggc):
        tag = "agg%02d_%s" % (idx, name)
        m = AGG_OPS[name](c0, c1, cout, True, repeats=1, larger=larger)
        ks = fill(m, seed=100 + idx)
        x = torch.from_numpy(det_array(tag + "/x", (2, c0) + s0)).requires_grad_(True)
        y = torch.from_numpy(det_array(tag + "/y", (2, c1) + s1)).requires_grad_(True)
        m.eval()
        z_eval = m(x, y)
        m.train()
        z_tr = m(x, y)
        ct = torch.from_numpy(det_array(tag + "/ct", tuple(z_tr.shape)))
        grads = torch.autograd.grad((z_tr * ct).sum(), [x, y] + [p for p in m.parameters()])
        out[tag + "/meta"] = np.array([name, c0, c1, cout, int(larger), s0[0], s0[1], s1[0], s1[1]]).astype(str)
        out[tag + "/keys"] = ks_arrays(ks)["keys"]
        out[tag + "/shapes"] = ks_arrays(ks)["shapes"]
        out[tag + "/z_eval"] = n(z_eval)
        out[tag + "/z_train"] = n(z_tr)
        out[tag + "/gx"] = n(grads[0])
        out[tag + "/gy"] = n(grads[1])
        for (pn, _), g in zip(m.named_parameters(), grads[2:]):
            out[tag + "/g/" + pn] = n(g)
    save("ops", **out)


# ------------------------------------------------------------------------------ whole networks
class EncDec(nn.Module):
    def __init__(self, enc, dec):
        super().__init__()
        self.encoder, self.decoder = enc, dec

    def forward(self, x):
        return self.decoder(self.encoder(x))


def gen_nets():
    specs = [
        ("W0", "wacv", W0, 19, 64, 2, False, (1, 3, 64, 96)),
        ("W1", "wacv", W1, 19, 64, 2, False, (1, 3, 64, 96)),
        ("W0cv", "wacv", W0, 11, 64, 2, False, (2, 3, 72, 88)),
        ("C0search", "cvpr", C0, 21, 48, 1, True, (2, 3, 65, 65)),
        ("C1search", "cvpr", C1, 21, 48, 1, True, (2, 3, 65, 65)),
        ("C2final", "cvpr", C2, 21, 64, 2, False, (1, 3, 72, 96)),
        ("D0depth", "cvpr", C0, 1, 64, 2, False, (1, 3, 64, 80)),
    ]
    for tag, paper, cfg, ncls, agg, rep, aux, shape in specs:
        if paper == "wacv":
            enc = mbv2(pretrained=False, return_layers=[1, 2])
            dec = TemplateDecoder(list(enc.out_sizes), ncls, cfg, agg_size=agg, repeats=rep)
        else:
            enc = mbv2(pretrained=False)
            dec = MicroDecoder(list(enc.out_sizes), ncls, cfg, agg_size=agg, aux_cell=aux, repeats=rep)
        net = EncDec(enc, dec)
        ks = fill(net, seed=7)
        x = torch.from_numpy(det_array(tag + "/x", shape))
        net.eval()
        with torch.no_grad():
            feats = enc(x)
            o = dec(feats)
        out = {"x": n(x), "n_params": np.array(sum(p.numel() for p in net.parameters())),
               "info": np.array(dec.info), **ks_arrays(ks)}
        for i, f in enumerate(feats):
            out["feat%d" % i] = n(f)
        if isinstance(o, tuple):
            out["out"] = n(o[0])
            for i, a in enumerate(o[1]):
                out["aux%d" % i] = n(a)
        else:
            out["out"] = n(o)
        # one train-mode fwd+bwd through decoder AND encoder with the reference's loss recipe
        if shape[0] > 1 and ncls > 1:
            net.train()
            o = net(x)
            aux_outs = []
            if isinstance(o, tuple):
                o, aux_outs = o
            y = torch.from_numpy(det_array(tag + "/y", (shape[0],) + tuple(o.shape[2:]), kind="int", lo=0, hi=ncls))
            y[:, ::5, ::7] = 255
            loss = nn.NLLLoss2d(ignore_index=255)(nn.LogSoftmax()(o), y)
            for a in aux_outs:
                a = nn.Upsample(size=y.size()[1:], mode="bilinear", align_corners=False)(a)
                loss = loss + 0.15 * nn.NLLLoss2d(ignore_index=255)(nn.LogSoftmax()(a), y)
            net.zero_grad()
            loss.backward()
            out["train_y"] = n(y)
            out["train_loss"] = n(loss)
            out["train_out"] = n(o)
            for pn, p in net.named_parameters():
                # keep fixtures small: store per-parameter gradient norms and a few full tensors
                # parameters the genotype leaves disconnected get no gradient in the reference: -1
                out["gnorm/" + pn] = n(p.grad.norm()) if p.grad is not None else np.array(-1.0, np.float32)
            for pn in ("decoder.conv_clf.weight", "decoder.pre_clf.0.weight", "encoder.layer1.0.weight",
                       "encoder.layer3.0.conv.3.weight"):
                out["grad/" + pn] = n(dict(net.named_parameters())[pn].grad)
            print("  no-grad params:", [pn for pn, p in net.named_parameters() if p.grad is None][:6])
        save("net_" + tag, **out)


# ------------------------------------------------------------------------------ engine: train_task0
def gen_task0():
    torch.manual_seed(0)
    np.random.seed(0)
    enc = mbv2(pretrained=False)
    dec = MicroDecoder(list(enc.out_sizes), 21, C0, agg_size=16, aux_cell=True, repeats=1)
    seg = types.SimpleNamespace(module=EncDec(enc, dec))
    ks = fill(dec, seed=11)
    N, B = 8, 4
    sizes = [(24, 16, 16), (32, 8, 8), (96, 4, 4), (320, 2, 2)]
    Xy = {}
    for i, s in enumerate(sizes):
        Xy[i] = torch.from_numpy(det_array("t0/f%d" % i, (N,) + s)).abs()
    y = torch.from_numpy(det_array("t0/y", (N, 16, 16), kind="int", lo=0, hi=21))
    y[:, ::4, ::3] = 255
    Xy["y"] = y
    Xy["kd_y"] = torch.from_numpy(det_array("t0/kd", (N, 21, 16, 16)))
    Xy["out_size"] = torch.Size((16, 16))
    optim = torch.optim.Adam(dec.parameters(), lr=3e-3, weight_decay=1e-5)
    crit = nn.NLLLoss2d(ignore_index=255)
    avg = [p.data.clone() for p in dec.parameters()]
    losses = []
    orig_info = ref_trainer.logger.info
    ref_trainer.logger.info = lambda msg, *a: losses.append(float(msg.split("Avg. Loss:")[1].split()[0]))
    for epoch in range(2):
        r = ref_trainer.train_task0(Xy, seg, optim, epoch, crit, nn.MSELoss(), B, False, True, 0.3, 3.0, True,
                                    avg_param=avg, polyak_decay=0.9, aux_weight=0.15)
        assert r is None
    ref_trainer.logger.info = orig_info
    out = dict(ks_arrays(ks), logged_avg_loss=np.array(losses))
    for k, v in dec.state_dict().items():
        out["post/" + k] = n(v)
    for (pn, _), a in zip(dec.named_parameters(), avg):
        out["avg/" + pn] = n(a)
    save("task0_step", **out)


# ------------------------------------------------------------------------------ engine: train_segmenter
def gen_segmenter():
    """The UNMODIFIED reference train_segmenter (engine/trainer.py:179-283) on the CPU (Tensor.cuda aliased away): two
    epochs x two iterations, SGD encoder + Adam decoder, both clips, Polyak 0.99, aux heads, BatchNorm in training mode."""
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.manual_seed(0)
    enc = mbv2(pretrained=False)
    dec = MicroDecoder(list(enc.out_sizes), 21, C0, agg_size=16, aux_cell=True, repeats=1)
    net = EncDec(enc, dec)
    ks_e, ks_d = fill(enc, seed=21), fill(dec, seed=22)
    seg = nn.Module()
    seg.module = net
    seg.forward = lambda x: net(x)
    # batch 4 @192x192: the deepest BatchNorm sees 144 samples per channel.  At 2 x 128x128 (32 samples) the REFERENCE's own
    # fourth logged loss moves by 0.06 under a 1e-6 relative perturbation of the weights; here it moves by 1e-3.
    B, H, W, n_it = 4, 192, 192, 2
    loader = ListLoader()
    out = {}
    for i in range(n_it):
        img = det_array("seg/img%d" % i, (B, 3, H, W)).astype(np.float64)
        msk = det_array("seg/msk%d" % i, (B, H, W), kind="int", lo=0, hi=21).astype(np.uint8)
        msk[:, ::5, ::3] = 255
        loader.append({"image": torch.from_numpy(img), "mask": torch.from_numpy(msk)})
    optim_enc = torch.optim.SGD(enc.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-5)
    optim_dec = torch.optim.Adam(dec.parameters(), lr=3e-3, weight_decay=1e-5)
    avg = [p.data.clone() for p in seg.parameters()]
    losses = []
    orig_info = ref_trainer.logger.info
    ref_trainer.logger.info = lambda msg, *a: losses.append(float(msg.split("Avg. Loss:")[1].split()[0]))
    for epoch in range(2):
        r = ref_trainer.train_segmenter(seg, loader, optim_enc, optim_dec, epoch, nn.NLLLoss2d(ignore_index=255), False,
                                        3.0, 3.0, True, print_every=1, aux_weight=0.15, avg_param=avg, polyak_decay=0.99)
        assert r is None
    ref_trainer.logger.info = orig_info
    out.update({"enc_keys": ks_arrays(ks_e)["keys"], "enc_shapes": ks_arrays(ks_e)["shapes"],
                "dec_keys": ks_arrays(ks_d)["keys"], "dec_shapes": ks_arrays(ks_d)["shapes"],
                "logged_avg_loss": np.array(losses)})
    # the 1.8 M encoder parameters would make a 14 MB fixture: tensors above 4096 elements are pinned by a strided sample of
    # 512 elements plus their sum and sum of squares (float64), everything smaller is stored whole
    def put(key, arr):
        arr = np.asarray(arr)
        if arr.size <= 4096 or arr.dtype.kind != "f":
            out[key] = arr
        else:
            flat = arr.reshape(-1)
            out[key + "@sample"] = flat[:: max(flat.size // 512, 1)][:512].copy()
            out[key + "@sums"] = np.array([flat.astype(np.float64).sum(), (flat.astype(np.float64) ** 2).sum()])
    for k, v in net.state_dict().items():
        put("post/" + k, n(v))
    for (pn, _), a in zip(seg.named_parameters(), avg):
        put("avg/" + pn, n(a))
    save("segmenter_step", **out)


# ------------------------------------------------------------------------------ engine: validate / metric
class TableSegmenter(nn.Module):
    """Returns pre-computed logits batch by batch -- pins inference.py:58-91 without a network."""

    def __init__(self, logits):
        super().__init__()
        self.logits, self.i = logits, 0

    def forward(self, x):
        o = self.logits[self.i]
        self.i += 1
        return o


class ListLoader(list):
    class _DS:
        def set_stage(self, s):
            pass
    dataset = _DS()


def gen_validate():
    torch.Tensor.cuda = lambda self, *a, **k: self  # run the unmodified validate() on the CPU
    out = {}
    for case, (C, B, h, w, H, W) in enumerate(((21, 3, 20, 20, 80, 80), (19, 2, 16, 32, 64, 128), (11, 2, 23, 30, 90, 120))):
        logits, batches = [], ListLoader()
        for b in range(2):
            lg = torch.from_numpy(det_array("val%d/l%d" % (case, b), (B, C, h, w)))
            lg[:, :, ::3, ::2] = lg[:, :1, ::3, ::2]  # exact ties across classes -> first-index argmax
            tg = det_array("val%d/t%d" % (case, b), (B, H, W), kind="int", lo=0, hi=C + 3).astype(np.uint8)
            tg[:, ::7, ::5] = 255
            if case == 2:
                tg[tg == 4] = 5  # an absent class
            logits.append(lg)
            batches.append({"image": torch.zeros(B, 3, 4, 4), "mask": torch.from_numpy(tg)})
            out["v%d/logits%d" % (case, b)] = n(lg)
            out["v%d/target%d" % (case, b)] = tg
        cms = []
        real = ref_inference.fast_cm
        ref_inference.fast_cm = lambda p, g, c: cms.append(real(p, g, c)) or cms[-1]
        reward = ref_inference.validate(TableSegmenter(logits), batches, 0, 0, num_classes=C, omit_classes=[0])
        ref_inference.fast_cm = real
        cm = sum(cms)
        ious, npx, accs = _ref_miou.compute_ius_accs(cm)
        out["v%d/C" % case] = np.array(C)
        out["v%d/cm" % case] = cm
        out["v%d/reward" % case] = np.array(reward)
        out["v%d/ious" % case], out["v%d/npx" % case], out["v%d/accs" % case] = ious, npx, accs
        out["v%d/iu_only" % case] = _ref_miou.compute_iu(cm)
    # raw fast_cm known answers incl. adversarial label sets
    rng = np.random.default_rng(5)
    for case, (N, C) in enumerate(((100003, 21), (65536, 19), (4099, 11), (1, 2), (0, 5))):
        p = rng.integers(0, C, N).astype(np.uint8)
        g = rng.integers(0, C, N).astype(np.uint8)
        if case == 1:
            g[:] = 3  # all one class
        out["cm%d/p" % case], out["cm%d/g" % case] = p, g
        out["cm%d/cm" % case] = _ref_miou.fast_cm(p, g, C)
    save("metric", **out)


if __name__ == "__main__":
    logging.basicConfig(level=logging.WARNING)
    which = sys.argv[1:] or ["ops", "nets", "task0", "segmenter", "validate"]
    if "ops" in which:
        gen_ops()
    if "nets" in which:
        gen_nets()
    if "task0" in which:
        gen_task0()
    if "segmenter" in which:
        gen_segmenter()
    if "validate" in which:
        gen_validate()
