"""Per-call CUDA-event table of the search loop's two training functions at their own geometry (task 0: decoder only on
cached 64x64 features, batch 64; task 1: end to end, batch 32 @350x350), eager, one stream, a few iterations of one
candidate.  Usage: python tools/search_profile.py OUT.txt   (or --graphs: the same candidate with
CUDA graphs and no instrumentation, to be run under ncu; tools/summarize_search_launches.py reads the launch list)"""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import nas_segm_b200
    from nas_segm_b200 import lib
    from nas_segm_b200.engine import trainer
    out_path = sys.argv[1]
    cfg = nas_segm_b200.config()
    tables, iters = {}, {"train_task0": 0, "train_segmenter": 0}

    def wrap(name, fn, n_it):
        def w(*a, **k):
            cfg.cuda_graphs = False
            cfg.async_wgrad = cfg.branch_streams = False
            lib.profile_begin()
            try:
                return fn(*a, **k)
            finally:
                prof = lib.profile_end()
                t = tables.setdefault(name, {})
                for key, (c, ms, b) in prof.items():
                    c0, ms0, _ = t.get(key, (0, 0.0, b))
                    t[key] = (c0 + c, ms0 + ms, b)
                iters[name] += n_it
        return w

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    if out_path == "--graphs":  # plain run with CUDA graphs for an `ncu --metrics gpu__time_duration.sum` launch list
        from nas_segm_b200.engine import search

        class _Go:  # a few iterations cannot earn the reward the TaskPerformer asks for: let the candidate reach task 1
            def step(self, reward):
                return True
        search.make_task_performers = lambda n_epochs, val_every: [[_Go() for _ in range(n // v)] for n, v in zip(n_epochs, val_every)]
        res = bench.search_numbers(types.SimpleNamespace(), dev, 0, 1, 1, 0, 256, 5, 64)
        print(res["per_candidate_s"], res["errors"])
        return
    n_task0, task1_iters = 128, 3
    trainer.train_task0 = wrap("train_task0", trainer.train_task0, n_task0 // 64)
    trainer.train_segmenter = wrap("train_segmenter", trainer.train_segmenter, task1_iters)
    res = bench.search_numbers(types.SimpleNamespace(), dev, 0, 1, 1, 0, n_task0, task1_iters, 64)
    with open(out_path, "w") as f:
        f.write("# errors: %s\n" % res["errors"])
        for name, t in tables.items():
            n = max(iters[name], 1)
            tot = sum(v[1] for v in t.values())
            f.write("\n## %s: %d iterations, %.3f ms of kernel time per iteration, %d calls per iteration\n"
                    % (name, n, tot / n, sum(v[0] for v in t.values()) // n))
            by = {}
            for k, (c, ms, b) in t.items():
                e = by.setdefault(k.split("[")[0], [0, 0.0])
                e[0] += c
                e[1] += ms
            f.write("# by entry point: ms_per_iteration  calls_per_iteration  avg_us\n")
            for k, (c, ms) in sorted(by.items(), key=lambda kv: -kv[1][1]):
                f.write("%9.3f %6.1f %8.2f  %s\n" % (ms / n, c / n, 1e3 * ms / c, k))
            f.write("# by call: ms_per_iteration  calls_per_iteration  avg_us  GB/s(algorithmic)  key\n")
            for k, (c, ms, b) in sorted(t.items(), key=lambda kv: -kv[1][1])[:60]:
                f.write("%9.3f %6.1f %8.2f %8.1f  %s\n" % (ms / n, c / n, 1e3 * ms / c, b / (ms / c * 1e-3) / 1e9, k))


if __name__ == "__main__":
    main()
