"""Latency-bound graphs under one NASB_PDL / NASB_PDL_MAX_CTAS setting (read from the environment by the library):
eval forwards (batch 1) and the depth-head training step.  Prints one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    out = {"NASB_PDL": os.environ.get("NASB_PDL"), "NASB_PDL_MAX_CTAS": os.environ.get("NASB_PDL_MAX_CTAS")}
    for name, g in (("arch0", bench.W0), ("arch1", bench.W1)):
        for h, w in ((1024, 2048), (360, 480)):
            out["%s_%dx%d_bf16_graph_ms" % (name, w, h)] = round(bench.fwd_latency(dev, g, h, w, torch.bfloat16, iters=50, graph=True), 4)
    d = bench.depth_head_step(dev)
    out["depth_head"] = d
    print(json.dumps(out))


if __name__ == "__main__":
    main()
