"""Diagnostic: bf16-mode error of the smoke network (train / eval BN) with the tensor-core and TMA-tile paths on and off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch
import nas_segm_b200
from golden_util import W0, det_state_dict, rel_err
from nas_segm_b200 import functional as Fn
from nas_segm_b200.nn.encoders import mbv2
from nas_segm_b200.nn.micro_decoders import TemplateDecoder

for shape in ((2, 3, 64, 96), (4, 3, 128, 192)):
    enc = mbv2(return_layers=[1, 2]); dec = TemplateDecoder(list(enc.out_sizes), 19, W0, agg_size=64, repeats=2)
    ks = [("encoder." + k, tuple(v.shape)) for k, v in enc.state_dict().items()] + [("decoder." + k, tuple(v.shape)) for k, v in dec.state_dict().items()]
    sd = det_state_dict(ks, seed=3)
    enc.load_state_dict({k[8:]: v for k, v in sd.items() if k.startswith("encoder.")}); dec.load_state_dict({k[8:]: v for k, v in sd.items() if k.startswith("decoder.")})
    enc, dec = enc.cuda(), dec.cuda()
    g = torch.Generator().manual_seed(9314)
    x = torch.randn(*shape, generator=g).cuda()
    for mode in ("train", "eval"):
        enc.train(mode == "train"); dec.train(mode == "train")
        res = {}
        for tag, dt, tc, tiles in (("fp32", torch.float32, True, True), ("bf16 tc+tiles", torch.bfloat16, True, True), ("bf16 tc only", torch.bfloat16, True, False),
                                   ("bf16 tiles only", torch.bfloat16, False, True), ("bf16 cuda-core", torch.bfloat16, False, False)):
            nas_segm_b200.set_act_dtype(dt); nas_segm_b200.config().use_tcgen05 = tc; nas_segm_b200.config().use_tma_tiles = tiles
            with torch.no_grad():
                feats = enc(x); out = dec(feats)
            res[tag] = (out.float().cpu().numpy(), [f.float().cpu().numpy() for f in feats])
        ref = res["fp32"]
        for tag in list(res)[1:]:
            print(shape, mode, "%-16s out %.3e  feat0 %.3e feat1 %.3e" % (tag, rel_err(res[tag][0], ref[0]), rel_err(res[tag][1][0], ref[1][0]), rel_err(res[tag][1][1], ref[1][1])))
nas_segm_b200.set_act_dtype(torch.float32)
