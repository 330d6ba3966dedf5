"""Where one candidate of the search loop (bench.py --workload search) spends its time: every StepGraph call is bracketed by
device synchronisations and booked as eager warm-up / capture / replay per training function.  The synchronisations remove
the host/device overlap, so the sum is an upper bound of the un-instrumented candidate time; the split is what matters."""
import collections
import json
import os
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    from nas_segm_b200 import graphs
    stats = collections.defaultdict(lambda: [0, 0.0])
    orig = graphs.StepGraph.__call__

    def timed(self, *a):
        torch.cuda.synchronize()
        t = time.time()
        had = self.graph is not None
        r = orig(self, *a)
        torch.cuda.synchronize()
        dt = time.time() - t
        kind = "replay" if had else ("capture" if self.graph is not None else "eager_warmup")
        name = getattr(self.fn, "__qualname__", "?").split(".")[0]
        s = stats[name + ":" + kind]
        s[0] += 1
        s[1] += dt
        return r

    graphs.StepGraph.__call__ = timed
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    a = types.SimpleNamespace()
    rounds = int(os.environ.get("ROUNDS", "2"))
    snap = {}

    res = bench.search_numbers(a, dev, 0, 1, rounds, 1, 4000, 297, 1024)
    out = {"per_candidate_s": res["per_candidate_s"], "phase_seconds_per_candidate": res["phase_seconds_per_candidate_rank0"],
           "rounds_incl_warmup": rounds + 1,
           "calls": {k: {"n": v[0], "total_s": round(v[1], 3), "ms_per_call": round(1e3 * v[1] / max(v[0], 1), 3)}
                     for k, v in sorted(stats.items())}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
