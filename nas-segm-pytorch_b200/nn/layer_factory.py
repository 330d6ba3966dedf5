"""Registry of decoder primitives (reference: src/nn/layer_factory.py) on fused sm_100a kernels.

Contract kept from the reference:
  * ``OPS[name](C_in, C_out, stride, affine, repeats=1) -> nn.Module`` (16 names), one tensor in / one out;
    ``AGG_OPS[name](C_in0, C_in1, C_out, affine, repeats=1, larger=True) -> nn.Module`` (2 names), two in / one out.
  * every module exposes the reference's parameter names (state_dict keys) through ordinary nn.Conv2d /
    nn.BatchNorm2d children, so released checkpoints load and ``isinstance(m, nn.BatchNorm2d)`` scans (BN freezing,
    src/engine/trainer.py:124-127) keep working.  Those children are parameter holders only: ``forward`` never
    calls them, it hands their tensors to one fused kernel unit (functional.conv_unit & co).
  * quirks preserved: SepConv re-applies its stride in every repeat; Skip and GAPConv1x1 ignore
    stride; Zero honours it; resolution comparisons are lexicographic tuple comparisons.
Activations are logical NCHW tensors over NHWC storage; inputs in any other layout are converted at entry.
"""
import torch
import torch.nn as nn

from .. import functional as Fn
from ..lib import ACT_NONE, ACT_RELU, ACT_RELU6, POOL_AVG, POOL_MAX


def conv3x3(in_planes, out_planes, stride=1, bias=False, dilation=1):
    """Parameter holder for a padded (dilated) 3x3 convolution (layer_factory.py:7-17)."""
    return nn.Conv2d(in_planes, out_planes, 3, stride=stride, padding=dilation, dilation=dilation, bias=bias)


def conv1x1(in_planes, out_planes, stride=1, bias=False):
    return nn.Conv2d(in_planes, out_planes, 1, stride=stride, padding=0, bias=bias)


def _entry(x):
    """Op-level entry: NHWC storage, dtype untouched (fp32 or bf16); the fp32<->bf16 boundary of the network is at
    the encoder stem / decoder entry."""
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    return Fn.lib.to_nhwc(x)


class FusedConvBN(nn.Sequential):
    """nn.Sequential(conv, BatchNorm2d[, ReLU|ReLU6]) executed as ONE kernel unit.

    Children keep the reference indices (0 = conv, 1 = BN, 2 = activation) so state_dict keys are unchanged."""

    def __init__(self, conv, bn, act_module=None, image=False):
        mods = [conv, bn] + ([act_module] if act_module is not None else [])
        super().__init__(*mods)
        self._act = ACT_NONE if act_module is None else (ACT_RELU6 if isinstance(act_module, nn.ReLU6) else ACT_RELU)
        self._image = image

    def forward(self, x, res=None, out_dtype=None):
        conv, bn = self[0], self[1]
        if self._image:
            Fn.lib.require_cuda(x)
            x = x.float().contiguous()  # planar fp32 NCHW image, read directly by the stem kernel
        else:
            x = _entry(x)
        return Fn.conv_unit(x, conv.weight, bn, ks=conv.kernel_size[0], stride=conv.stride[0], dil=conv.dilation[0],
                            pad=conv.padding[0], act=self._act, res=res, dw=conv.groups > 1, image=self._image,
                            out_dtype=out_dtype)


def conv_bn(C_in, C_out, kernel_size, stride, padding, affine=True):
    return FusedConvBN(nn.Conv2d(C_in, C_out, kernel_size, stride=stride, padding=padding, bias=False),
                       nn.BatchNorm2d(C_out, affine=affine))


def conv_bn_relu(C_in, C_out, kernel_size, stride, padding, affine=True):
    return FusedConvBN(nn.Conv2d(C_in, C_out, kernel_size, stride=stride, padding=padding, bias=False),
                       nn.BatchNorm2d(C_out, affine=affine), nn.ReLU(inplace=False))


def conv_bn_relu6(inp, oup, stride):
    """Encoder stem (encoders.py:38): reads the planar fp32 image directly."""
    return FusedConvBN(nn.Conv2d(inp, oup, 3, stride, 1, bias=False), nn.BatchNorm2d(oup), nn.ReLU6(inplace=True),
                       image=True)


def conv_1x1_bn_relu6(inp, oup):
    return FusedConvBN(nn.Conv2d(inp, oup, 1, 1, 0, bias=False), nn.BatchNorm2d(oup), nn.ReLU6(inplace=True))


def _registry_conv(C_in, C_out, ks, stride, dilation, affine):
    conv = conv1x1(C_in, C_out, stride=stride) if ks == 1 else conv3x3(C_in, C_out, stride=stride, dilation=dilation)
    return FusedConvBN(conv, nn.BatchNorm2d(C_out, affine=affine), nn.ReLU(inplace=False))


class InvertedResidual(nn.Module):
    """MobileNet-v2 block (layer_factory.py:125-158): pw-expand+BN+ReLU6 | dw3x3+BN+ReLU6 | pw-project+BN (+x)."""

    def __init__(self, inp, oup, stride, expand_ratio):
        super().__init__()
        assert stride in (1, 2)
        self.stride = stride
        self.use_res_connect = stride == 1 and inp == oup
        hid = inp * expand_ratio
        self.conv = nn.Sequential(
            nn.Conv2d(inp, hid, 1, 1, 0, bias=False), nn.BatchNorm2d(hid), nn.ReLU6(inplace=True),
            nn.Conv2d(hid, hid, 3, stride, 1, groups=hid, bias=False), nn.BatchNorm2d(hid), nn.ReLU6(inplace=True),
            nn.Conv2d(hid, oup, 1, 1, 0, bias=False), nn.BatchNorm2d(oup))

    def forward(self, x):
        x = _entry(x)
        c = self.conv
        y = Fn.conv_unit(x, c[0].weight, c[1], ks=1, act=ACT_RELU6)
        # depthwise + projection: one fused kernel at inference (Fn.sep_unit), two units in training.
        # `sole`: the expansion feeds the depthwise convolution only, which feeds the projection only
        return Fn.sep_unit(y, c[3].weight, c[4], ACT_RELU6, c[6].weight, c[7], ACT_NONE, ks=3, stride=self.stride, dil=1, pad=1,
                           res=x if self.use_res_connect else None, sole=(True, True))


class Pool(nn.Module):
    """1x1 conv + BN, then 3x3 max/avg pooling (layer_factory.py:161-178)."""

    def __init__(self, C_in, C_out, stride, repeats, ksize, mode):
        super().__init__()
        if ksize != 3:
            raise ValueError("only 3x3 pooling is registered")
        if mode not in ("avg", "max"):
            raise ValueError("Unknown pooling method {}".format(mode))
        self.conv1x1 = conv_bn(C_in, C_out, 1, 1, 0)
        self.pool = (nn.AvgPool2d(3, stride=stride, padding=1, count_include_pad=False) if mode == "avg"
                     else nn.MaxPool2d(3, stride=stride, padding=1))
        self._mode, self._stride = (POOL_AVG if mode == "avg" else POOL_MAX), stride

    def forward(self, x):
        return Fn.pool3x3(self.conv1x1(x), self._mode, self._stride)


class GAPConv1x1(nn.Module):
    """Global average pool -> 1x1 conv-BN-ReLU -> broadcast back (layer_factory.py:181-195)."""

    def __init__(self, C_in, C_out):
        super().__init__()
        self.conv1x1 = conv_bn_relu(C_in, C_out, 1, stride=1, padding=0)

    def forward(self, x):
        x = _entry(x)
        v = Fn.spatial_mean(x)
        v = self.conv1x1(v, out_dtype=torch.float32)
        return Fn.spatial_bcast(v, x.shape[2], x.shape[3], x.dtype)


class DilConv(nn.Module):
    """ReLU -> dilated depthwise -> 1x1 -> BN (layer_factory.py:198-222)."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, dilation, affine=True):
        super().__init__()
        self.op = nn.Sequential(
            nn.ReLU(inplace=False),
            nn.Conv2d(C_in, C_in, kernel_size, stride=stride, padding=padding, dilation=dilation, groups=C_in, bias=False),
            nn.Conv2d(C_in, C_out, 1, padding=0, bias=False),
            nn.BatchNorm2d(C_out, affine=affine))

    def forward(self, x):
        x = _entry(x)
        dwc = self.op[1]
        y = Fn.conv_unit(x, dwc.weight, None, ks=dwc.kernel_size[0], stride=dwc.stride[0], dil=dwc.dilation[0],
                         pad=dwc.padding[0], dw=True, in_relu=1)
        return Fn.conv_unit(y, self.op[2].weight, self.op[3], ks=1, act=ACT_NONE)


class SepConv(nn.Module):
    """repeats x [depthwise kxk -> 1x1 -> BN -> ReLU] (layer_factory.py:225-265); stride applies in every repeat."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, dilation=1, affine=True, repeats=1):
        super().__init__()
        self.op = nn.Sequential()
        for idx in range(repeats):
            c = C_in if idx == 0 else C_out
            self.op.add_module("sep_{}".format(idx), nn.Sequential(
                nn.Conv2d(c, c, kernel_size, stride=stride, padding=padding, dilation=dilation, groups=c, bias=False),
                nn.Conv2d(c, C_out, 1, padding=0, bias=False),
                nn.BatchNorm2d(C_out, affine=affine),
                nn.ReLU(inplace=False)))

    def forward(self, x):
        x = _entry(x)
        for idx, blk in enumerate(self.op):
            dwc = blk[0]
            # from the second repeat on, the depthwise conv is the only consumer of the previous repeat's output (`sole`)
            x = Fn.sep_unit(x, dwc.weight, None, ACT_NONE, blk[1].weight, blk[2], ACT_RELU, ks=dwc.kernel_size[0],
                            stride=dwc.stride[0], dil=dwc.dilation[0], pad=dwc.padding[0], sole=(idx > 0, False))
        return x


class Skip(nn.Module):
    """Channel tiling; the stride argument is ignored (layer_factory.py:268-275)."""

    def __init__(self, C_in, C_out, stride):
        super().__init__()
        assert (C_out % C_in) == 0, "C_out must be divisible by C_in"
        self.repeats = (1, C_out // C_in, 1, 1)

    def forward(self, x):
        x = _entry(x)
        return Fn.channel_tile(x, self.repeats[1], 1, 1.0)


class Identity(nn.Module):
    def forward(self, x):
        return x


class Zero(nn.Module):
    """Tiled, strided, multiplied by zero (layer_factory.py:286-297)."""

    def __init__(self, C_in, C_out, stride):
        super().__init__()
        assert (C_out % C_in) == 0, "C_out must be divisible by C_in"
        self.stride = stride
        self.repeats = (1, C_out // C_in, 1, 1)

    def forward(self, x):
        x = _entry(x)
        return Fn.channel_tile(x, self.repeats[1], self.stride, 0.0)


class FactorizedReduce(nn.Module):
    """ReLU -> two stride-2 1x1 convolutions on the even / odd pixel grids -> cat -> BN (layer_factory.py:300-313).  Not
    reachable from any genotype (it is in no name table); kept for registry completeness on the general kernels."""

    def __init__(self, C_in, C_out, affine=True):
        super().__init__()
        assert C_out % 2 == 0
        self.relu = nn.ReLU(inplace=False)
        self.conv_1 = nn.Conv2d(C_in, C_out // 2, 1, stride=2, padding=0, bias=False)
        self.conv_2 = nn.Conv2d(C_in, C_out // 2, 1, stride=2, padding=0, bias=False)
        self.bn = nn.BatchNorm2d(C_out, affine=affine)

    def forward(self, x):
        x = _entry(x)
        x = Fn.concat_resize([x], x.shape[2:], relu=True)
        a = Fn.conv_unit(x, self.conv_1.weight, None, ks=1, stride=2)
        b = Fn.conv_unit(_entry(x[:, :, 1:, 1:]), self.conv_2.weight, None, ks=1, stride=2)
        if a.shape[2:] != b.shape[2:]:  # odd input sizes: torch.cat fails in the reference as well
            raise RuntimeError("Sizes of tensors must match except in dimension 1")
        return Fn.bn_act(Fn.concat_resize([a, b], a.shape[2:]), self.bn, ACT_NONE)


OPS = {
    "none": lambda C_in, C_out, stride, affine, repeats=1: Zero(C_in, C_out, stride),
    "avg_pool_3x3": lambda C_in, C_out, stride, affine, repeats=1: Pool(C_in, C_out, stride, repeats, ksize=3, mode="avg"),
    "max_pool_3x3": lambda C_in, C_out, stride, affine, repeats=1: Pool(C_in, C_out, stride, repeats, ksize=3, mode="max"),
    "global_average_pool": lambda C_in, C_out, stride, affine, repeats=1: GAPConv1x1(C_in, C_out),
    "skip_connect": lambda C_in, C_out, stride, affine, repeats=1: Skip(C_in, C_out, stride),
    "sep_conv_3x3": lambda C_in, C_out, stride, affine, repeats=1: SepConv(C_in, C_out, 3, stride, 1, affine=affine, repeats=repeats),
    "sep_conv_5x5": lambda C_in, C_out, stride, affine, repeats=1: SepConv(C_in, C_out, 5, stride, 2, affine=affine, repeats=repeats),
    "sep_conv_7x7": lambda C_in, C_out, stride, affine, repeats=1: SepConv(C_in, C_out, 7, stride, 3, affine=affine, repeats=repeats),
    "dil_conv_3x3": lambda C_in, C_out, stride, affine, repeats=1: DilConv(C_in, C_out, 3, stride, 2, 2, affine=affine),
    "dil_conv_5x5": lambda C_in, C_out, stride, affine, repeats=1: DilConv(C_in, C_out, 5, stride, 4, 2, affine=affine),
    "conv1x1": lambda C_in, C_out, stride, affine, repeats=1: _registry_conv(C_in, C_out, 1, stride, 1, affine),
    "conv3x3": lambda C_in, C_out, stride, affine, repeats=1: _registry_conv(C_in, C_out, 3, stride, 1, affine),
    "conv3x3_dil3": lambda C_in, C_out, stride, affine, repeats=1: _registry_conv(C_in, C_out, 3, stride, 3, affine),
    "conv3x3_dil12": lambda C_in, C_out, stride, affine, repeats=1: _registry_conv(C_in, C_out, 3, stride, 12, affine),
    "sep_conv_3x3_dil3": lambda C_in, C_out, stride, affine, repeats=1: SepConv(C_in, C_out, 3, stride, 3, affine=affine, dilation=3, repeats=repeats),
    "sep_conv_5x5_dil6": lambda C_in, C_out, stride, affine, repeats=1: SepConv(C_in, C_out, 5, stride, 12, affine=affine, dilation=6, repeats=repeats),
}


def _pick_size(s1, s2, largest):
    """Target size of resize(): tuple (lexicographic) comparison, as torch.Size compares (layer_factory.py:338-350)."""
    s1, s2 = tuple(s1), tuple(s2)
    if s1 == s2:
        return s1
    if largest:
        return s1 if s1 > s2 else s2
    return s1 if s1 < s2 else s2


def resize(x1, x2, largest=True):
    size = _pick_size(x1.shape[2:], x2.shape[2:], largest)
    return Fn.resize(x1, size), Fn.resize(x2, size)


class Adapt(nn.Module):
    """Bring two inputs to C_out channels (1x1 conv-BN-ReLU when needed) and to a common size (layer_factory.py:316-335).
    ``forward`` returns the two channel-adapted maps plus the common size; the resize itself is fused into the
    aggregation kernel by the callers."""

    def __init__(self, C_in0, C_in1, C_out, larger):
        super().__init__()
        self.C_in0, self.C_in1, self.C_out, self.larger = C_in0, C_in1, C_out, larger
        if C_in0 != C_out:
            self.conv0 = conv_bn_relu(C_in0, C_out, 1, 1, 0)
        if C_in1 != C_out:
            self.conv1 = conv_bn_relu(C_in1, C_out, 1, 1, 0)

    def channels(self, x1, x2):
        x1, x2 = _entry(x1), _entry(x2)
        if self.C_in0 != self.C_out and self.C_in1 != self.C_out:  # two independent 1x1 units: concurrent streams
            x1, x2 = Fn.lib.branch_streams.run2(lambda: self.conv0(x1), lambda: self.conv1(x2), (x2,))
        elif self.C_in0 != self.C_out:
            x1 = self.conv0(x1)
        elif self.C_in1 != self.C_out:
            x2 = self.conv1(x2)
        return x1, x2, _pick_size(x1.shape[2:], x2.shape[2:], self.larger)

    def forward(self, x1, x2):
        x1, x2, size = self.channels(x1, x2)
        return Fn.resize(x1, size), Fn.resize(x2, size)


class ParamSum(nn.Module):
    """a[c]*x + b[c]*y after Adapt (layer_factory.py:353-366); resize + scale + add is one kernel."""

    def __init__(self, C_in0, C_in1, C_out, larger):
        super().__init__()
        self.adapt = Adapt(C_in0, C_in1, C_out, larger)
        self.a = nn.Parameter(torch.ones(C_out))
        self.b = nn.Parameter(torch.ones(C_out))

    def forward(self, x, y):
        x, y, size = self.adapt.channels(x, y)
        if tuple(y.shape[2:]) == size:
            return Fn.resize_add(x, y, self.a, self.b)
        return Fn.resize_add(y, x, self.b, self.a)


class ConcatReduce(nn.Module):
    """Adapt -> cat -> BN -> ReLU -> 1x1 conv (layer_factory.py:369-382)."""

    def __init__(self, C_in0, C_in1, C_out, affine=True, repeats=1, larger=True):
        super().__init__()
        self.adapt = Adapt(C_in0, C_in1, C_out, larger)
        self.conv1x1 = nn.Sequential(nn.BatchNorm2d(2 * C_out, affine=affine), nn.ReLU(inplace=False),
                                     nn.Conv2d(2 * C_out, C_out, 1, stride=1, padding=0, bias=False))

    def forward(self, x, y):
        x, y, size = self.adapt.channels(x, y)
        z = Fn.concat_resize([x, y], size)
        z = Fn.bn_act(z, self.conv1x1[0], ACT_RELU)
        return Fn.conv_unit(z, self.conv1x1[2].weight, None, ks=1, sole=True)  # the only consumer of the BN + ReLU output


AGG_OPS = {
    "psum": lambda C_in0, C_in1, C_out, affine, repeats=1, larger=True: ParamSum(C_in0, C_in1, C_out, larger),
    "cat": lambda C_in0, C_in1, C_out, affine, repeats=1, larger=True: ConcatReduce(C_in0, C_in1, C_out, affine=affine, repeats=repeats, larger=larger),
}
