"""Candidates/hour of the NAS inner loop (BASELINE config 4) on one GPU: a CVPR-style candidate (MobileNet-v2 4-tap encoder +
sampled MicroDecoder, agg 48, aux cells, 21 classes) goes through the reference's task0 recipe with the engine API:
populate_task0 once, then per candidate 5 epochs x (N // 64) iterations of train_task0 (batch 64, 256x256 crops -> 64x64
features, KD-MSE + aux CE, Adam, clip, Polyak) and one validate() -> reward.  Prints one JSON line.

    python tools/search_bench.py [--n-task0 4000] [--val-images 448] [--candidates 3] [--dtype bf16] [--cuda-graph 1]
Under torchrun every rank evaluates its own candidates and the rewards are exchanged with the single all-gather."""
import argparse
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch import nn  # noqa: E402

import nas_segm_b200  # noqa: E402
from nas_segm_b200 import parallel  # noqa: E402
from nas_segm_b200.engine import inference, trainer  # noqa: E402
from nas_segm_b200.helpers.utils import init_polyak  # noqa: E402
from nas_segm_b200.nn.encoders import mbv2  # noqa: E402
from nas_segm_b200.nn.micro_decoders import MicroDecoder  # noqa: E402

GENOTYPES = [  # the three published CVPR genotypes (tests/test_inference.py:22-58 of the reference) + variations
    [[8, [0, 0, 5, 2], [0, 2, 8, 8], [0, 5, 1, 4]], [[3, 3], [3, 2], [3, 0]]],
    [[2, [1, 0, 3, 6], [0, 1, 2, 8], [2, 0, 6, 1]], [[2, 3], [3, 1], [4, 4]]],
    [[5, [0, 0, 4, 1], [3, 2, 0, 1], [5, 6, 5, 0]], [[1, 3], [4, 3], [2, 2]]],
]


class Seg(nn.Module):
    def __init__(self, enc, dec):
        super().__init__()
        self.encoder, self.decoder = enc, dec

    def forward(self, x):
        return self.decoder(self.encoder(x))


class Wrapper(nn.Module):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-task0", type=int, default=4000)
    ap.add_argument("--val-images", type=int, default=448)
    ap.add_argument("--candidates", type=int, default=3)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--cuda-graph", type=int, default=1)
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--profile-out", default=None)
    a = ap.parse_args()
    rank, world, dev = parallel.init()
    nas_segm_b200.set_act_dtype(torch.bfloat16 if a.dtype == "bf16" else torch.float32)
    nas_segm_b200.config().cuda_graphs = bool(a.cuda_graph)
    torch.manual_seed(9314 + rank)
    np.random.seed(9314 + rank)
    g = torch.Generator().manual_seed(rank)
    # ---- task0 cache (once): encoder features of n_task0 256x256 crops, labels, KD targets (synthetic teacher logits)
    enc = mbv2()
    seg0 = Wrapper(Seg(enc, nn.Identity()).to(dev))
    t0 = time.time()
    train = Loader({"image": torch.randn(1, 3, 256, 256, generator=g), "mask": torch.randint(0, 21, (1, 256, 256), generator=g).to(torch.uint8)}
                   for _ in range(min(a.n_task0, 64)))
    Xy = trainer.populate_task0(seg0, train, None, len(train), do_kd=False)
    assert Xy != 0
    reps = (a.n_task0 + len(train) - 1) // len(train)  # tile the cached block to n_task0 samples (content is irrelevant to timing)
    for k in list(Xy.keys()):
        if k != "out_size":
            Xy[k] = Xy[k].repeat(*([reps] + [1] * (Xy[k].dim() - 1)))[: a.n_task0].contiguous() if k == "y" else \
                Xy[k].permute(0, 2, 3, 1).repeat(reps, 1, 1, 1)[: a.n_task0].contiguous().permute(0, 3, 1, 2)
    Xy["kd_y"] = torch.randn(a.n_task0, 64, 64, 21, generator=torch.Generator(device=dev).manual_seed(1), device=dev).permute(0, 3, 1, 2)
    torch.cuda.synchronize()
    t_populate = time.time() - t0
    val = Loader({"image": torch.randn(64, 3, 400, 400, generator=g), "mask": torch.randint(0, 21, (64, 400, 400), generator=g).to(torch.uint8)}
                 for _ in range(max(a.val_images // 64, 1)))
    crit, kd_crit = nn.NLLLoss(ignore_index=255), nn.MSELoss()
    times, recs = [], []
    for ci in range(a.candidates):
        geno = GENOTYPES[(rank + ci * world) % len(GENOTYPES)]
        enc = mbv2()
        dec = MicroDecoder(list(enc.out_sizes), 21, geno, agg_size=48, aux_cell=True, repeats=1)
        seg = Wrapper(Seg(enc, dec).to(dev))
        optim_dec = torch.optim.Adam(seg.module.decoder.parameters(), lr=3e-3, weight_decay=1e-5)
        avg = init_polyak(True, seg.module.decoder)
        torch.cuda.synchronize()
        t0 = time.time()
        for ep in range(a.epochs):
            r = trainer.train_task0(Xy, seg, optim_dec, ep, crit, kd_crit, 64, False, True, 0.3, 3.0, True, avg_param=avg,
                                    polyak_decay=0.9, aux_weight=0.15)
            assert r is None
        torch.cuda.synchronize()
        t_train = time.time() - t0
        t0 = time.time()
        reward = inference.validate(seg, val, ci, a.epochs, num_classes=21, omit_classes=[0])
        torch.cuda.synchronize()
        t_val = time.time() - t0
        times.append((t_train, t_val))
        recs.append([float(reward), 0.0, 0.0, 0.0])
    if a.profile_out and rank == 0:  # per-call CUDA events of ONE eager task0 epoch of the last candidate
        from nas_segm_b200 import lib
        nas_segm_b200.config().cuda_graphs = False
        small = {k: (v[:64] if k != "out_size" else v) for k, v in Xy.items()}
        lib.profile_begin()
        trainer.train_task0(small, seg, optim_dec, 0, crit, kd_crit, 64, False, True, 0.3, 3.0, True, avg_param=avg,
                            polyak_decay=0.9, aux_weight=0.15)
        prof = lib.profile_end()
        by = {}
        for k, (c, t_ms, b) in prof.items():
            e = by.setdefault(k.split("[")[0], [0, 0.0])
            e[0] += c
            e[1] += t_ms
        with open(a.profile_out, "w") as f:
            f.write("# one task0 iteration (batch 64, 64x64 features), eager, CUDA events per C-ABI call: total %.3f ms, %d calls\n" % (
                sum(v[1] for v in by.values()), sum(v[0] for v in by.values())))
            for k, (c, t_ms) in sorted(by.items(), key=lambda kv: -kv[1][1]):
                f.write("%9.3f ms %5d calls %7.1f us/call  %s\n" % (t_ms, c, 1e3 * t_ms / c, k))
            f.write("# top individual calls\n")
            for k, (c, t_ms, b) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:40]:
                f.write("%9.3f ms %4d x %8.4f ms %8.1f GB/s  %s\n" % (t_ms, c, t_ms / c, b / (t_ms / c * 1e-3) / 1e9, k))
    table = parallel.gather_records(recs, dev)  # the single collective
    iters = a.epochs * (a.n_task0 // 64)
    # the first candidate pays the one-off kernel-attribute / graph-capture warm-up; report the steady state
    steady = times[1:] if len(times) > 1 else times
    t_tr = float(np.mean([t[0] for t in steady]))
    t_v = float(np.mean([t[1] for t in steady]))
    if rank == 0:
        print(json.dumps({"metric": "candidates_per_hour_task0", "value": world * 3600.0 / (t_tr + t_v), "unit": "candidates/h",
                          "n_gpus": world, "dtype": a.dtype, "cuda_graph": bool(a.cuda_graph),
                          "per_candidate_s": {"train_task0": t_tr, "validate": t_v, "first_candidate": list(times[0])},
                          "ms_per_task0_iteration": 1e3 * t_tr / iters, "task0_iterations": iters,
                          "populate_task0_s": t_populate, "val_images": len(val) * 64, "n_task0": a.n_task0,
                          "rewards_gathered": table[:, :, 0].tolist()}))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
