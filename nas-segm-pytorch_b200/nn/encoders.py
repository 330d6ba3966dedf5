"""MobileNet-v2 feature extractor returning 2 (WACV) or 4 (CVPR) scales (reference: src/nn/encoders.py)."""
import os

import torch
import torch.nn as nn

from .layer_factory import InvertedResidual, conv_bn_relu6

__all__ = ["mbv2"]

model_paths = {"mbv2_voc": "./data/weights/mbv2_voc_rflw.ckpt"}


class MobileNetV2(nn.Module):
    # expansion, output channels, repeats, stride (encoders.py:19-27)
    mobilenet_config = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2],
                        [6, 320, 1, 1]]
    in_planes = 32
    num_layers = len(mobilenet_config)

    def __init__(self, width_mult=1.0, return_layers=[1, 2, 4, 6]):
        super().__init__()
        self.return_layers = return_layers
        self.max_layer = max(return_layers)
        self.out_sizes = [self.mobilenet_config[i][1] for i in return_layers]
        cin = int(self.in_planes * width_mult)
        self.layer1 = conv_bn_relu6(3, cin, 2)
        for layer_idx, (t, c, n, s) in enumerate(self.mobilenet_config[: self.max_layer + 1]):
            cout = int(c * width_mult)
            blocks = []
            for i in range(n):
                blocks.append(InvertedResidual(cin, cout, s if i == 0 else 1, t))
                cin = cout
            setattr(self, "layer{}".format(layer_idx + 2), nn.Sequential(*blocks))

    def forward(self, x):
        outs = []
        x = self.layer1(x)
        for layer_idx in range(self.max_layer + 1):
            x = getattr(self, "layer{}".format(layer_idx + 2))(x)
            outs.append(x)
        return [outs[i] for i in self.return_layers]


def mbv2(pretrained=False, **kwargs):
    model = MobileNetV2(**kwargs)
    if pretrained:
        path = model_paths["mbv2_{}".format(str(pretrained))]
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        model.load_state_dict(torch.load(path, map_location="cpu"), strict=False)
    return model


def create_encoder(pretrained="voc", ctrl_version="cvpr", **kwargs):
    return_layers = [1, 2, 4, 6] if ctrl_version == "cvpr" else [1, 2]
    return mbv2(pretrained=pretrained, return_layers=return_layers, **kwargs)
