"""Functional CPU oracle for the reference's network path (TEST INFRASTRUCTURE ONLY).

A genotype interpreter over a flat ``state_dict``: every function below takes the parameter
store ``P`` (keys are exactly the reference's ``state_dict`` keys) plus a key prefix, and
evaluates the same arithmetic as the reference's ``nn.Module`` graph with stock torch CPU
kernels.  Gradients come from torch autograd over these functions.

Each function cites the reference lines (relative to /root/reference/) it restates.
Pinned against fixtures produced by the real reference: tests/test_oracle_golden.py.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

# rl/genotypes.py:8-35 -- index -> registry name tables
OP_NAMES = ("conv1x1", "conv3x3", "sep_conv_3x3", "sep_conv_5x5", "global_average_pool",
            "conv3x3_dil3", "conv3x3_dil12", "sep_conv_3x3_dil3", "sep_conv_5x5_dil6",
            "skip_connect", "none")
OP_NAMES_WACV = ("sep_conv_3x3", "sep_conv_5x5", "global_average_pool", "max_pool_3x3",
                 "sep_conv_5x5_dil6", "skip_connect")
AGG_OP_NAMES = ("psum", "cat")

# nn/encoders.py:19-27 -- (expansion, channels, repeats, stride)
MBV2_CFG = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2),
            (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1))


class Params:
    """Flat parameter store.  ``Params()`` creates tensors on first access (seeded), recording
    key order; ``Params(sd)`` serves an existing state_dict and fails on a missing key."""

    def __init__(self, sd=None, seed=0, dtype=torch.float32):
        self.create = sd is None
        self.sd = OrderedDict() if sd is None else sd
        self.dtype = dtype
        self.gen = torch.Generator().manual_seed(seed)
        self.bn_momentum = 0.1

    def _new(self, key, shape, kind):
        g = self.gen
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g, dtype=self.dtype) * (1.5 / fan_in) ** 0.5
        elif kind == "bn_w" or kind == "bn_var":
            t = torch.rand(shape, generator=g, dtype=self.dtype) + 0.5
        elif kind == "bn_b" or kind == "bn_mean" or kind == "bias":
            t = torch.randn(shape, generator=g, dtype=self.dtype) * 0.1
        elif kind == "ones":
            t = torch.ones(shape, dtype=self.dtype)
        elif kind == "count":
            t = torch.zeros((), dtype=torch.int64)
        else:
            raise KeyError(kind)
        self.sd[key] = t
        return t

    def get(self, key, shape, kind):
        if key not in self.sd:
            if not self.create:
                raise KeyError("oracle: state_dict lacks %s" % key)
            return self._new(key, tuple(shape), kind)
        t = self.sd[key]
        assert tuple(t.shape) == tuple(shape), (key, tuple(t.shape), tuple(shape))
        return t

    def requires_grad_(self, flag=True):
        for k, v in self.sd.items():
            if v.is_floating_point() and not (k.endswith("running_mean") or k.endswith("running_var")):
                v.requires_grad_(flag)
        return self


# ----------------------------------------------------------------------------- primitives
def _j(pfx, name):
    """Join a module prefix and a leaf key (empty prefix = the op is the root module)."""
    return name if pfx == "" else pfx + "." + name


def _conv(x, P, key, cout, cin_g, k, stride=1, pad=0, dil=1, groups=1, bias_key=None):
    w = P.get(key, (cout, cin_g, k, k), "conv")
    b = P.get(bias_key, (cout,), "bias") if bias_key else None
    return F.conv2d(x, w, b, stride=stride, padding=pad, dilation=dil, groups=groups)


def _bn(x, P, pfx, c, training, affine=True):
    """nn.BatchNorm2d(eps=1e-5, momentum=0.1); keys pfx.{weight,bias,running_mean,running_var,
    num_batches_tracked}."""
    w = P.get(_j(pfx, "weight"), (c,), "bn_w") if affine else None
    b = P.get(_j(pfx, "bias"), (c,), "bn_b") if affine else None
    rm = P.get(_j(pfx, "running_mean"), (c,), "bn_mean")
    rv = P.get(_j(pfx, "running_var"), (c,), "bn_var")
    nb = P.get(_j(pfx, "num_batches_tracked"), (), "count")
    if training:
        nb += 1
    return F.batch_norm(x, rm, rv, w, b, training, P.bn_momentum, 1e-5)


def _up(x, size):
    """nn.Upsample(size, mode='bilinear') with align_corners None == False
    (layer_factory.py:341-349, micro_decoders.py:17-23,48-50); used for down-sizing as well."""
    return F.interpolate(x, size=tuple(size), mode="bilinear", align_corners=False)


def conv_bn_relu(x, P, pfx, cin, cout, k, stride, pad, training, dil=1, affine=True):
    """layer_factory.py:101-106 and the registry lambdas :56-75 (Sequential: 0=conv, 1=BN, 2=ReLU)."""
    x = _conv(x, P, _j(pfx, "0.weight"), cout, cin, k, stride, pad, dil)
    return F.relu(_bn(x, P, _j(pfx, "1"), cout, training, affine))


def sep_conv(x, P, pfx, cin, cout, k, stride, pad, dil, repeats, training, affine=True):
    """layer_factory.py:225-265: repeats x [dw kxk (stride EVERY repeat) -> 1x1 -> BN -> ReLU]."""
    for i in range(repeats):
        q = _j(pfx, "op.sep_%d" % i)
        c = cin if i == 0 else cout
        x = _conv(x, P, q + ".0.weight", c, 1, k, stride, pad, dil, groups=c)
        x = _conv(x, P, q + ".1.weight", cout, c, 1)
        x = F.relu(_bn(x, P, q + ".2", cout, training, affine))
    return x


def dil_conv(x, P, pfx, cin, cout, k, stride, pad, dil, training, affine=True):
    """layer_factory.py:198-222: ReLU -> dw -> 1x1 -> BN."""
    x = F.relu(x)
    x = _conv(x, P, _j(pfx, "op.1.weight"), cin, 1, k, stride, pad, dil, groups=cin)
    x = _conv(x, P, _j(pfx, "op.2.weight"), cout, cin, 1)
    return _bn(x, P, _j(pfx, "op.3"), cout, training, affine)


def pool_op(x, P, pfx, cin, cout, stride, mode, training):
    """layer_factory.py:161-178: 1x1 conv -> BN (no ReLU) -> 3x3 max/avg pool, pad 1."""
    x = _conv(x, P, _j(pfx, "conv1x1.0.weight"), cout, cin, 1)
    x = _bn(x, P, _j(pfx, "conv1x1.1"), cout, training)
    if mode == "max":
        return F.max_pool2d(x, 3, stride=stride, padding=1)
    return F.avg_pool2d(x, 3, stride=stride, padding=1, count_include_pad=False)


def gap_conv(x, P, pfx, cin, cout, training):
    """layer_factory.py:181-195: mean over H then W -> 1x1 conv-BN-ReLU -> bilinear back to HxW."""
    size = x.shape[2:]
    o = x.mean(2, keepdim=True).mean(3, keepdim=True)
    o = conv_bn_relu(o, P, _j(pfx, "conv1x1"), cin, cout, 1, 1, 0, training)
    return _up(o, size)


def skip_op(x, cin, cout):
    """layer_factory.py:268-275: channel tiling, stride ignored."""
    assert cout % cin == 0
    return x.repeat(1, cout // cin, 1, 1)


def zero_op(x, cin, cout, stride):
    """layer_factory.py:286-297."""
    x = skip_op(x, cin, cout)
    if stride != 1:
        x = x[:, :, ::stride, ::stride]
    return x.mul(0.0)


def op_forward(name, x, P, pfx, cin, cout, stride, repeats=1, training=False, affine=True):
    """The OPS registry, layer_factory.py:27-82."""
    if name == "none":
        return zero_op(x, cin, cout, stride)
    if name == "skip_connect":
        return skip_op(x, cin, cout)
    if name == "max_pool_3x3":
        return pool_op(x, P, pfx, cin, cout, stride, "max", training)
    if name == "avg_pool_3x3":
        return pool_op(x, P, pfx, cin, cout, stride, "avg", training)
    if name == "global_average_pool":
        return gap_conv(x, P, pfx, cin, cout, training)
    if name == "conv1x1":
        return conv_bn_relu(x, P, pfx, cin, cout, 1, stride, 0, training, affine=affine)
    if name == "conv3x3":
        return conv_bn_relu(x, P, pfx, cin, cout, 3, stride, 1, training, 1, affine)
    if name == "conv3x3_dil3":
        return conv_bn_relu(x, P, pfx, cin, cout, 3, stride, 3, training, 3, affine)
    if name == "conv3x3_dil12":
        return conv_bn_relu(x, P, pfx, cin, cout, 3, stride, 12, training, 12, affine)
    sep = {"sep_conv_3x3": (3, 1, 1), "sep_conv_5x5": (5, 2, 1), "sep_conv_7x7": (7, 3, 1),
           "sep_conv_3x3_dil3": (3, 3, 3), "sep_conv_5x5_dil6": (5, 12, 6)}
    if name in sep:
        k, pad, dil = sep[name]
        return sep_conv(x, P, pfx, cin, cout, k, stride, pad, dil, repeats, training, affine)
    if name == "dil_conv_3x3":
        return dil_conv(x, P, pfx, cin, cout, 3, stride, 2, 2, training, affine)
    if name == "dil_conv_5x5":
        return dil_conv(x, P, pfx, cin, cout, 5, stride, 4, 2, training, affine)
    raise KeyError(name)


def _resize_pair(x1, x2, largest):
    """layer_factory.py:338-350; torch.Size comparison is a lexicographic tuple comparison."""
    s1, s2 = tuple(x1.shape[2:]), tuple(x2.shape[2:])
    if largest:
        if s1 > s2:
            x2 = _up(x2, s1)
        elif s1 < s2:
            x1 = _up(x1, s2)
    else:
        if s1 < s2:
            x2 = _up(x2, s1)
        elif s1 > s2:
            x1 = _up(x1, s2)
    return x1, x2


def _adapt(x1, x2, P, pfx, c0, c1, cout, larger, training):
    """layer_factory.py:316-335."""
    if c0 != cout:
        x1 = conv_bn_relu(x1, P, _j(pfx, "conv0"), c0, cout, 1, 1, 0, training)
    if c1 != cout:
        x2 = conv_bn_relu(x2, P, _j(pfx, "conv1"), c1, cout, 1, 1, 0, training)
    return _resize_pair(x1, x2, larger)


def agg_forward(name, x, y, P, pfx, c0, c1, cout, larger=True, training=False, affine=True):
    """AGG_OPS, layer_factory.py:84-91,353-382."""
    if name == "psum":
        x, y = _adapt(x, y, P, _j(pfx, "adapt"), c0, c1, cout, larger, training)
        a = P.get(_j(pfx, "a"), (cout,), "ones")
        b = P.get(_j(pfx, "b"), (cout,), "ones")
        return a[None, :, None, None] * x + b[None, :, None, None] * y
    if name == "cat":
        x, y = _adapt(x, y, P, _j(pfx, "adapt"), c0, c1, cout, larger, training)
        z = torch.cat([x, y], 1)
        z = F.relu(_bn(z, P, _j(pfx, "conv1x1.0"), 2 * cout, training, affine))
        return _conv(z, P, _j(pfx, "conv1x1.2.weight"), cout, 2 * cout, 1)
    raise KeyError(name)


def collect_all(feats, inds):
    """micro_decoders.py:11-25 (compares heights only)."""
    out = feats[inds[0]]
    for i in inds[1:]:
        c = feats[i]
        if out.shape[2] > c.shape[2]:
            c = _up(c, out.shape[2:])
        elif c.shape[2] > out.shape[2]:
            out = _up(out, c.shape[2:])
        out = torch.cat([out, c], 1)
    return out


def aggregate_cell(x1, x2, P, pfx, s1, s2, agg, pre_transform, training):
    """micro_decoders.py:28-51."""
    if pre_transform:
        x1 = conv_bn_relu(x1, P, _j(pfx, "branch_1"), s1, agg, 1, 1, 0, training)
        x2 = conv_bn_relu(x2, P, _j(pfx, "branch_2"), s2, agg, 1, 1, 0, training)
    x1, x2 = _resize_pair(x1, x2, True)
    return x1 + x2


def contextual_cell(x, P, pfx, config, c, repeats, training):
    """micro_decoders.py:54-121.  _ops index j runs over ops AND the parameter-free sums."""
    feats = [x]
    loose = []
    j = 0
    for ind, op in enumerate(config):
        if ind == 0:
            feats.append(op_forward(OP_NAMES[op], feats[0], P, _j(pfx, "_ops.%d" % j), c, c, 1,
                                    repeats, training))
            j += 1
            loose.append(1)
        else:
            p1, p2, o1, o2 = op
            for p, o in ((p1, o1), (p2, o2)):
                if p in loose:
                    loose.remove(p)
                feats.append(op_forward(OP_NAMES[o], feats[p], P, _j(pfx, "_ops.%d" % j), c, c, 1,
                                        repeats, training))
                j += 1
            feats.append(aggregate_cell(feats[ind * 3 - 1], feats[ind * 3], P, "", None, None, c,
                                        False, training))
            j += 1
            loose.append(ind * 3 + 1)
    out = feats[loose[0]]
    for i in loose[1:]:
        out = out + feats[i]
    return out


def micro_decoder(feats, P, config, inp_sizes, num_classes, agg_size=64, aux_cell=False, repeats=1,
                  training=False, pfx=""):
    """MicroDecoder, micro_decoders.py:142-254.  Returns (out, aux_outs)."""
    cell_cfg, conns = config
    x = [conv_bn_relu(f, P, "%sadapt%d" % (pfx, i + 1), inp_sizes[i], agg_size, 1, 1, 0, training)
         for i, f in enumerate(feats)]
    n_pools = len(x)
    collect = []
    aux_outs = []
    for k, (i1, i2) in enumerate(conns):
        for i in (i1, i2):
            if i in collect:
                collect.remove(i)
        q = "%scells.%d" % (pfx, k)
        a = contextual_cell(x[i1], P, q + ".op_1", cell_cfg, agg_size, repeats, training)
        b = contextual_cell(x[i2], P, q + ".op_2", cell_cfg, agg_size, repeats, training)
        out = aggregate_cell(a, b, P, q + ".agg", agg_size, agg_size, agg_size, True, training)
        x.append(out)
        aux = out
        if aux_cell:
            aux = contextual_cell(aux, P, "%saux_clfs.%d.aux_cell" % (pfx, k), cell_cfg, agg_size,
                                  repeats, training)
        aux_outs.append(_conv(aux, P, "%saux_clfs.%d.aux_clf.weight" % (pfx, k), num_classes,
                              agg_size, 3, 1, 1, bias_key="%saux_clfs.%d.aux_clf.bias" % (pfx, k)))
        collect.append(k + n_pools)
    out = F.relu(collect_all(x, collect))
    out = conv_bn_relu(out, P, pfx + "pre_clf", agg_size * len(collect), agg_size, 1, 1, 0, training)
    out = _conv(out, P, pfx + "conv_clf.weight", num_classes, agg_size, 3, 1, 1,
                bias_key=pfx + "conv_clf.bias")
    return out, aux_outs


def template_decoder(feats, P, config, inp_sizes, num_classes, agg_size=64, repeats=1,
                     stride_power=1, training=False, pfx=""):
    """TemplateDecoder, micro_decoders.py:257-398.  Returns the logits tensor."""
    cells, structure = config
    n_scales = len(inp_sizes)
    chans = list(inp_sizes) + [0] * len(structure)
    feats = list(feats)
    collect = []
    for b, (pos1, pos2, cell_id, n_rep, s) in enumerate(structure):
        larger = b >= (len(structure) // 2)
        stride = 2 ** s
        o1, o2, oagg = cells[cell_id]
        f1, f2 = feats[pos1], feats[pos2]
        new_c = [0, 0]
        prev_c = [0, 0]
        agg_c = None
        out = None
        for r in range(n_rep + 1):
            outs = []
            for li, (pos, oid, f) in enumerate(((pos1, o1, f1), (pos2, o2, f2))):
                if r == 0:
                    cur = chans[pos]
                    new = cur * int(stride ** stride_power)
                elif li == 0:
                    cur = new = prev_c[-1]
                else:
                    cur = new = agg_c
                new_c[li] = new
                prev_c[li] = cur
                if pos in collect:
                    collect.remove(pos)
                outs.append(op_forward(OP_NAMES_WACV[oid], f, P, "%s_ops.%d.%d" % (pfx, b, r * 3 + li),
                                       cur, new, stride, repeats, training))
            agg_c = max(new_c)
            out = agg_forward(AGG_OP_NAMES[oagg], outs[0], outs[1], P, "%s_ops.%d.%d" % (pfx, b, r * 3 + 2),
                              new_c[0], new_c[1], agg_c, larger, training)
            f1, f2 = f2, out
        chans[n_scales + b] = agg_c
        feats.append(out)
        collect.append(n_scales + b)
    c_pre = sum(chans[i] for i in collect)
    out = F.relu(collect_all(feats, collect))
    out = conv_bn_relu(out, P, pfx + "pre_clf", c_pre, agg_size, 1, 1, 0, training)
    return _conv(out, P, pfx + "conv_clf.weight", num_classes, agg_size, 3, 1, 1,
                 bias_key=pfx + "conv_clf.bias")


def inverted_residual(x, P, pfx, inp, oup, stride, t, training):
    """layer_factory.py:125-158 (Sequential indices 0,1 | 3,4 | 6,7; ReLU6 at 2 and 5)."""
    h = inp * t
    y = _conv(x, P, _j(pfx, "conv.0.weight"), h, inp, 1)
    y = F.relu6(_bn(y, P, _j(pfx, "conv.1"), h, training))
    y = _conv(y, P, _j(pfx, "conv.3.weight"), h, 1, 3, stride, 1, 1, groups=h)
    y = F.relu6(_bn(y, P, _j(pfx, "conv.4"), h, training))
    y = _conv(y, P, _j(pfx, "conv.6.weight"), oup, h, 1)
    y = _bn(y, P, _j(pfx, "conv.7"), oup, training)
    return x + y if (stride == 1 and inp == oup) else y


def mbv2_encoder(x, P, return_layers=(1, 2, 4, 6), training=False, pfx=""):
    """MobileNetV2.forward, encoders.py:15-63."""
    x = _conv(x, P, pfx + "layer1.0.weight", 32, 3, 3, 2, 1)
    x = F.relu6(_bn(x, P, pfx + "layer1.1", 32, training))
    cin = 32
    outs = []
    for li, (t, c, n, s) in enumerate(MBV2_CFG[: max(return_layers) + 1]):
        for i in range(n):
            x = inverted_residual(x, P, "%slayer%d.%d" % (pfx, li + 2, i), cin, c, s if i == 0 else 1, t,
                                  training)
            cin = c
        outs.append(x)
    return [outs[i] for i in return_layers]


def encoder_out_sizes(return_layers):
    return [MBV2_CFG[i][1] for i in return_layers]


# ----------------------------------------------------------------------------- losses
def segm_loss(logits, target, out_size=None, ignore_index=255):
    """trainer.py:141-146: bilinear to out_size -> LogSoftmax(dim=1) -> NLLLoss2d(ignore 255, mean)."""
    if out_size is not None:
        logits = _up(logits, out_size)
    return F.nll_loss(F.log_softmax(logits, dim=1), target, ignore_index=ignore_index)


def task0_loss(out, aux_outs, y, out_size, kd_y=None, kd_coeff=0.0, aux_weight=0.0):
    """trainer.py:137-158: CE + kd_coeff*MSE(upsampled logits, kd) + aux_weight*sum(aux CE)."""
    up = _up(out, out_size)
    loss = F.nll_loss(F.log_softmax(up, dim=1), y, ignore_index=255)
    if kd_y is not None:
        loss = loss + kd_coeff * F.mse_loss(up, kd_y)
    if aux_weight > 0:
        for a in aux_outs:
            loss = loss + aux_weight * segm_loss(a, y, out_size)
    return loss


def berhu_loss(pred, target, valid_min=0.0):
    """Reverse Huber (Laina et al. 2016).  NOT in the reference: parity unpinned,
    defined by this repo: e=|p-t| over valid (t>valid_min) px, c=0.2*max(e),
    L = mean(e if e<=c else (e^2+c^2)/(2c))."""
    mask = target > valid_min
    e = (pred - target).abs()[mask]
    c = 0.2 * e.max().detach()
    return torch.where(e <= c, e, (e * e + c * c) / (2 * c)).mean()


def nearest_labels(mask, size):
    """trainer.py:42-49,236-238: float nearest interpolate of the label map, then long."""
    return F.interpolate(mask[:, None].float(), size=tuple(size), mode="nearest").long()[:, 0]
