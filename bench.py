#!/usr/bin/env python
"""bench.py -- the hot path's headline measurement (driver contract: one JSON line on rank 0).

Workload (BASELINE.json configs[1]): WACV arch0 (MobileNet-v2 2-tap encoder + TemplateDecoder, 19 classes) end-to-end
training iteration -- forward, per-pixel CE, backward, grad clips, SGD(encoder)+Adam(decoder) -- bf16 activations,
batch 8 at 2048x1024, synthetic data, BatchNorm in training mode (the reference's default, FREEZE_BN = False).

  python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's CUDA path
  python bench.py --impl reference ...                         the reference algorithm on the host cores (oracle port)

value = images/s with inputs resident in HBM (CUDA events, max over ranks); e2e = the same iteration through the
engine API (train_segmenter) with pinned HOST batches: host->device copies and the device->host loss read are inside
the timed region.  One process per GPU; candidates are independent (weak scaling), the only collective is the all-gather
of one 16-byte reward record per rank per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

W0 = [[[3, 0, 1], [4, 1, 1], [3, 1, 1]],
      [[0, 1, 0, 0, 1], [2, 1, 2, 1, 0], [3, 1, 1, 1, 0], [1, 1, 2, 0, 0], [3, 0, 2, 0, 0], [5, 3, 2, 1, 0], [0, 5, 0, 1, 0]]]
W1 = [[[1, 1, 0], [1, 3, 0], [3, 4, 0]],
      [[1, 1, 0, 0, 0], [0, 1, 1, 1, 1], [3, 1, 2, 3, 0], [3, 0, 2, 2, 0], [0, 1, 2, 0, 0], [2, 1, 1, 3, 0], [4, 0, 2, 2, 0]]]
NUM_CLASSES = 19
METRIC = "arch0_train_images_per_sec_2048x1024"
# algorithmic bytes of W0 @2048x1024, bf16, forward, per image; fwd+bwd convention = 3x
ALGO_BYTES_FWD_PER_IMG = 0.416e9


def args_():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--cuda-graph", type=int, default=1, help="replay the training iteration as one CUDA graph (default on)")
    ap.add_argument("--profile-out", default=None, help="write the per-call CUDA-event table of one instrumented step")
    ap.add_argument("--depth-head-only", action="store_true", help="(internal) print the BASELINE config-5 timings as JSON and exit")
    return ap.parse_args()


def synth(batch, h, w, seed=9314):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(batch, 3, h, w, generator=g)
    lab = torch.randint(0, NUM_CLASSES, (batch, h, w), generator=g).to(torch.uint8)
    lab[torch.rand(batch, h, w, generator=g) < 0.05] = 255
    return img, lab


class Seg(nn.Module):
    def __init__(self, enc, dec):
        super().__init__()
        self.encoder, self.decoder = enc, dec

    def forward(self, x):
        return self.decoder(self.encoder(x))


class Wrapper(nn.Module):
    """DataParallel-shaped wrapper (`.module`) without the scatter/gather: one process owns one GPU."""

    def __init__(self, m):
        super().__init__()
        self.module = m

    def forward(self, x):
        return self.module(x)


class HostLoader(list):
    class _DS:
        def set_stage(self, s):
            pass
    dataset = _DS()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms.  Started before the warm-up (nvidia-smi needs a moment to
    come up, longer on an 8-GPU box); `mark()` brackets the timed region and the summary uses the samples that arrived
    inside it (falling back to every sample taken under load if the region was shorter than one period)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.t0, self.t1 = [], None, index, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 - 0.05 <= t <= (self.t1 or 1e18) + 0.15]
        src = "timed region"
        if not inside:  # region shorter than a sampling period: use the samples taken while the warm-up kept the GPU busy
            inside = [r for t, r in self.rows if self.t1 is None or t <= self.t1 + 0.15][-5:]
            src = "warm-up + timed region"
        sm, mx, reasons = [], 0, set()
        for r in inside:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "window": src}


# ------------------------------------------------------------------------------------------------------------ reference arm
def oracle_step_fn(h, w, batch):
    """One training iteration of the same network in the CPU oracle (fp32, all host threads)."""
    from oracle import nas_oracle as O
    torch.manual_seed(0)
    Pe, Pd = O.Params(seed=1), O.Params(seed=2)
    img, lab = synth(batch, h, w)
    with torch.no_grad():  # create parameters
        O.template_decoder(O.mbv2_encoder(torch.zeros(2, 3, 32, 32), Pe, (1, 2)), Pd, W0, [24, 32], NUM_CLASSES, 64, 2)
    for P in (Pe, Pd):
        P.requires_grad_()
    p_enc = [v for v in Pe.sd.values() if v.requires_grad]
    p_dec = [v for v in Pd.sd.values() if v.requires_grad]
    # the same iteration as the CUDA arm: forward, CE, backward, the two grad-norm clips, SGD (encoder) + Adam (decoder)
    optim_enc = torch.optim.SGD(p_enc, lr=1e-3, momentum=0.9, weight_decay=1e-5)
    optim_dec = torch.optim.Adam(p_dec, lr=3e-3, weight_decay=1e-5)

    def step():
        out = O.template_decoder(O.mbv2_encoder(img, Pe, (1, 2), True), Pd, W0, [24, 32], NUM_CLASSES, 64, 2, training=True)
        y = O.nearest_labels(lab, out.shape[2:])
        loss = O.segm_loss(out, y)
        optim_enc.zero_grad()
        optim_dec.zero_grad()
        loss.backward()
        nn.utils.clip_grad_norm_(p_enc, 3.0)
        nn.utils.clip_grad_norm_(p_dec, 3.0)
        optim_enc.step()
        optim_dec.step()
        return float(loss.detach())
    return step


def pick_threads():
    """The thread count at which the oracle's torch CPU kernels run fastest on this host (a 128-core box is ~50x SLOWER
    with 128 intra-op threads than with 16 on these small / depthwise convolutions).  Calibrated on a 512x256 image."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (4, 8, 16, 32, 64, cores) if c <= cores})
    step = oracle_step_fn(256, 512, 1)
    best, best_t = cands[0], None
    for c in cands:
        torch.set_num_threads(c)
        step()
        t0 = time.time()
        step()
        dt = time.time() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
        if dt > 4 * best_t or dt > 10.0:
            break
    torch.set_num_threads(best)
    return best, cores


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads, cores = pick_threads()
    step = oracle_step_fn(a.height, a.width, 1)
    t0 = time.time()
    step()
    first = time.time() - t0
    # bounded: the whole run stays within a few minutes whatever K and W are
    n_warm = max(0, min(a.warmup, 3) - 1) if first < 15 else 0
    n_steps = max(1, min(a.steps, int(120.0 / max(first, 1e-3))))
    for _ in range(n_warm):
        step()
    t0 = time.time()
    for _ in range(n_steps):
        step()
    dt = (time.time() - t0) / n_steps
    v = 1.0 / dt
    sample = ("1 image per step (of the batch-%d workload), %d timed steps, fp32, torch CPU kernels, %d intra-op threads "
              "(fastest setting on this %d-core host)" % (a.batch, n_steps, threads, cores))
    cores = threads
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(a),
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(a):
    return {"workload": "WACV arch0 (mbv2 2-tap encoder + TemplateDecoder agg64 rep2, %d classes) training iteration "
                        "fwd+CE+bwd+clip+SGD/Adam, BN train mode, batch %d @%dx%d" % (NUM_CLASSES, a.batch, a.width, a.height),
            "global_batch_per_gpu": a.batch, "resolution": [a.width, a.height], "l2": "inputs_exceed_L2",
            "parallelism": "one candidate per GPU (replicas), 1 all-gather of 16 B per step",
            "cuda_graph": bool(getattr(a, "cuda_graph", 0))}


# ------------------------------------------------------------------------------------------------------------ CUDA arm
def build_model(dev, genotype=W0):
    from nas_segm_b200.nn.encoders import mbv2
    from nas_segm_b200.nn.micro_decoders import TemplateDecoder
    torch.manual_seed(0)
    enc = mbv2(return_layers=[1, 2])
    dec = TemplateDecoder(list(enc.out_sizes), NUM_CLASSES, genotype, agg_size=64, repeats=2)
    return Wrapper(Seg(enc, dec)).to(dev)


def fwd_latency(dev, genotype, h, w, dtype, iters=20, graph=False):
    """Eval-mode forward latency, batch 1 (the reference README's table); graph=True replays a captured CUDA graph."""
    import nas_segm_b200
    from nas_segm_b200.graphs import GraphedForward
    nas_segm_b200.set_act_dtype(dtype)
    m = build_model(dev, genotype).eval()
    x = torch.randn(1, 3, h, w, device=dev)
    run = GraphedForward(m, x) if graph else m
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for _ in range(3):
            run(x)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            run(x)
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def depth_head_step(dev, batch=8, h=480, w=640, iters=10):
    """BASELINE config 5: the CVPR arch0 genotype in its final settings as a depth head (MobileNet-v2 4-tap encoder +
    MicroDecoder agg 64, repeats 2, ONE output channel), 640x480, reverse-Huber loss, forward + backward, bf16 activations,
    synthetic depth targets in (0.5, 8) m with 5 % invalid (0) pixels.  Eager and as one CUDA graph."""
    import nas_segm_b200
    from nas_segm_b200 import functional as Fn
    from nas_segm_b200.graphs import StepGraph
    from nas_segm_b200.nn.encoders import mbv2
    from nas_segm_b200.nn.micro_decoders import MicroDecoder
    c0 = [[8, [0, 0, 5, 2], [0, 2, 8, 8], [0, 5, 1, 4]], [[3, 3], [3, 2], [3, 0]]]
    nas_segm_b200.set_act_dtype(torch.bfloat16)
    torch.manual_seed(0)
    enc = mbv2()
    dec = MicroDecoder(list(enc.out_sizes), 1, c0, agg_size=64, aux_cell=False, repeats=2)
    net = Seg(enc, dec).to(dev).train()
    params = [q for q in net.parameters() if q.requires_grad]
    g = torch.Generator().manual_seed(9314)
    x = torch.randn(batch, 3, h, w, generator=g).to(dev)

    def step(im, tg):
        out = net(im)
        out = out[0] if isinstance(out, tuple) else out
        loss = Fn.berhu_loss(out, tg)
        for q in params:
            q.grad = None
        loss.backward()
        return loss

    with torch.no_grad():
        o = net(x)
        o = o[0] if isinstance(o, tuple) else o
    tg = torch.empty(o.shape, dtype=torch.float32).uniform_(0.5, 8.0, generator=g)
    tg[torch.rand(o.shape, generator=g) < 0.05] = 0.0
    tg = tg.to(dev)
    res = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, run in (("eager", step), ("cudagraph", StepGraph(step, [x.clone(), tg.clone()]))):
        for _ in range(5):
            loss = run(x, tg)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            loss = run(x, tg)
        e1.record()
        torch.cuda.synchronize()
        res["d0_depth_berhu_fwdbwd_ms_b%d_%dx%d_bf16_%s" % (batch, w, h, name)] = e0.elapsed_time(e1) / iters
    res["d0_depth_berhu_loss"] = float(loss.detach())
    return res


def main():
    a = args_()
    if a.impl == "reference":
        return run_reference(a)
    if a.depth_head_only:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        print(json.dumps(depth_head_step(torch.device("cuda", torch.cuda.current_device()))))
        return
    import nas_segm_b200
    from nas_segm_b200 import lib
    from nas_segm_b200.engine import trainer
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = world > 1
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl b200) needs a GPU; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if dist:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    nas_segm_b200.set_act_dtype(dtype)
    lib.load()
    seg = build_model(dev)
    seg.train()
    optim_enc = torch.optim.SGD(seg.module.encoder.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-5)
    optim_dec = torch.optim.Adam(seg.module.decoder.parameters(), lr=3e-3, weight_decay=1e-5)
    crit = nn.NLLLoss(ignore_index=255)
    img_h, lab_h = synth(a.batch, a.height, a.width, seed=9314 + rank)
    img_d, lab_d = img_h.to(dev), lab_h.to(dev)
    rec = torch.zeros(4, dtype=torch.float32, device=dev)
    gathered = torch.zeros(4 * world, dtype=torch.float32, device=dev) if dist else None

    nas_segm_b200.config().cuda_graphs = bool(a.cuda_graph)
    if a.cuda_graph:
        from nas_segm_b200.graphs import StepGraph, make_capturable
        make_capturable(optim_enc)
        make_capturable(optim_dec)
        graphed = StepGraph(lambda im, tg: trainer.segmenter_step(seg, im, tg, optim_enc, optim_dec, crit, 3.0, 3.0, False),
                            [img_d, lab_d])

    def step():
        if a.cuda_graph:
            loss = graphed(img_d, lab_d)
        else:
            loss = trainer.segmenter_step(seg, img_d, lab_d, optim_enc, optim_dec, crit, 3.0, 3.0, False)
        if dist:
            rec[0] = loss.detach()
            td.all_gather_into_tensor(gathered, rec)
        return loss

    def barrier():
        if dist:
            td.barrier()
        torch.cuda.synchronize()

    clk = ClockSampler(local).start()
    for _ in range(max(a.warmup, 3) + (2 if a.cuda_graph else 0)):  # graph mode: 3 eager iterations, then capture + replays
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.launches
    barrier()
    clk.mark_begin()
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    barrier()
    clk.mark_end()
    clk.stop()
    launches = lib.launches - l0
    ms = e0.elapsed_time(e1)
    tmax = torch.tensor([ms], device=dev)
    if dist:
        td.all_reduce(tmax, op=td.ReduceOp.MAX)
    ms = float(tmax.item())
    ms_step = ms / a.steps
    value = world * a.batch * a.steps / (ms / 1e3)
    final_loss = float(loss.detach())

    # ---- end to end through the engine API: pinned host batches in, loss read back every iteration
    pinned = [{"image": img_h.clone().pin_memory(), "mask": lab_h.clone().pin_memory()} for _ in range(2)]
    loader = HostLoader(pinned[i % 2] for i in range(a.steps))
    warm = HostLoader(pinned[i % 2] for i in range(2 + (4 if a.cuda_graph else 0)))
    trainer.train_segmenter(seg, warm, optim_enc, optim_dec, 0, crit, False, 3.0, 3.0, False, print_every=1)
    barrier()
    e0.record()
    r = trainer.train_segmenter(seg, loader, optim_enc, optim_dec, 0, crit, False, 3.0, 3.0, False, print_every=1)
    e1.record()
    barrier()
    assert r is None, "train_segmenter swallowed a RuntimeError"
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dist:
        td.all_reduce(ms_e2e, op=td.ReduceOp.MAX)
    e2e_value = world * a.batch * a.steps / (float(ms_e2e.item()) / 1e3)
    h2d = img_h.numel() * 4 + lab_h.numel()
    e2e = {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
           "ms_per_step": float(ms_e2e.item()) / a.steps}

    # ---- roofline: CUDA-event duration of every C-ABI call over instrumented iterations (same stream, same data)
    roof, top_list = None, []
    if rank == 0:
        n_prof = min(a.steps, 3)
        lib.profile_begin()
        for _ in range(n_prof):
            trainer.segmenter_step(seg, img_d, lab_d, optim_enc, optim_dec, crit, 3.0, 3.0, False)
        prof = lib.profile_end()
        tot = sum(v[1] for v in prof.values())
        if a.profile_out:
            with open(a.profile_out, "w") as f:
                f.write("# ms_per_step  calls_per_step  avg_ms  GB/s(algorithmic)  key\n")
                for k, (c, t_ms, b) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
                    f.write("%9.3f %4d %8.4f %8.1f  %s\n" % (t_ms / n_prof, c // n_prof, t_ms / c, b / (t_ms / c * 1e-3) / 1e9, k))
        by_entry = {}
        for k, (c, t_ms, b) in prof.items():
            e = by_entry.setdefault(k.split("[")[0], [0, 0.0, 0.0])
            e[0] += c
            e[1] += t_ms
            e[2] += b * c
        peak, peak_src = peaks()
        # dominant kernel = the entry point with the largest total CUDA-event time; achieved = its algorithmic bytes over
        # its time, summed over all of its launches in the instrumented iterations (launch-weighted average)
        top_entry = max(by_entry, key=lambda k: by_entry[k][1])
        ecalls, ems, ebytes = by_entry[top_entry]
        achieved = ebytes / (ems * 1e-3) / 1e9
        inst = max((k for k in prof if k.split("[")[0] == top_entry), key=lambda k: prof[k][1])
        ic, ims, ib = prof[inst]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "top_kernel_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(top_entry)
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": top_entry, "launches_per_step": ecalls // n_prof, "avg_launch_ms": ems / ecalls,
                "share_of_step": ems / tot, "algorithmic_bytes_per_launch": ebytes / ecalls, "peak_source": peak_src,
                "largest_instance": {"key": inst, "avg_launch_ms": ims / ic, "achieved_gbs": ib / (ims / ic * 1e-3) / 1e9},
                "step_level": {"algorithmic_bytes_per_step": 3 * ALGO_BYTES_FWD_PER_IMG * a.batch,
                               "achieved_gbs": 3 * ALGO_BYTES_FWD_PER_IMG * a.batch / (ms_step * 1e-3) / 1e9,
                               "frac": 3 * ALGO_BYTES_FWD_PER_IMG * a.batch / (ms_step * 1e-3) / 1e9 / peak}}
        top_list = sorted(((v[1] / n_prof, v[0] // n_prof, k) for k, v in by_entry.items()), reverse=True)[:8]

    extras = {}
    if rank == 0 and not a.no_extras and not dist:
        del img_d, lab_d
        torch.cuda.empty_cache()
        try:
            for name, g in (("arch0", W0), ("arch1", W1)):
                for (hh, ww) in ((1024, 2048), (360, 480)):
                    for dn, dt in (("bf16", torch.bfloat16), ("f32", torch.float32)):
                        extras["%s_fwd_ms_b1_%dx%d_%s" % (name, ww, hh, dn)] = fwd_latency(dev, g, hh, ww, dt)
                        extras["%s_fwd_ms_b1_%dx%d_%s_cudagraph" % (name, ww, hh, dn)] = fwd_latency(dev, g, hh, ww, dt, graph=True)
            try:  # BASELINE config 5 (depth head) in a child process with a deadline: it can never cost the headline line
                child = subprocess.run([sys.executable, os.path.abspath(__file__), "--depth-head-only"], capture_output=True,
                                       text=True, timeout=240, env=dict(os.environ, LOCAL_RANK=str(local)))
                line = [ln for ln in child.stdout.splitlines() if ln.startswith("{")]
                if child.returncode == 0 and line:
                    extras.update(json.loads(line[-1]))
                else:
                    extras["d0_depth_error"] = ("rc=%d " % child.returncode) + child.stderr[-300:]
            except Exception as e:  # noqa: BLE001
                extras["d0_depth_error"] = repr(e)[:300]
            try:  # BASELINE config 4 (NAS inner loop, candidates/hour on this GPU), same arrangement
                child = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "search_bench.py"), "--candidates", "2"],
                                       capture_output=True, text=True, timeout=300,
                                       env={k: v for k, v in dict(os.environ, LOCAL_RANK=str(local)).items()
                                            if k not in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")})
                line = [ln for ln in child.stdout.splitlines() if ln.startswith("{")]
                if child.returncode == 0 and line:
                    sb = json.loads(line[-1])
                    extras["search_loop_task0_candidates_per_hour_n1"] = sb.get("value")
                    extras["search_loop_ms_per_task0_iteration"] = sb.get("ms_per_task0_iteration")
                    extras["search_loop_per_candidate_s"] = sb.get("per_candidate_s")
                else:
                    extras["search_loop_error"] = ("rc=%d " % child.returncode) + child.stderr[-300:]
            except Exception as e:  # noqa: BLE001
                extras["search_loop_error"] = repr(e)[:300]
        finally:
            nas_segm_b200.set_act_dtype(dtype)

    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        threads, cores = pick_threads()
        ostep = oracle_step_fn(a.height, a.width, 1)
        t0 = time.time()
        ostep()
        first = time.time() - t0
        n = 2 if first < 10 else 0
        t0 = time.time()
        for _ in range(n):
            ostep()
        dt = (time.time() - t0) / n if n else first
        cpu = {"value": 1.0 / dt, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": "oracle (torch CPU fp32) training iteration on 1 image @%dx%d, %d timed + 1 warm-up, %d intra-op "
                         "threads (fastest setting on this %d-core host)" % (a.width, a.height, n, threads, cores)}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": a.steps,
               "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": a.dtype, "data": "synthetic", "config": workload_config(a),
               "clocks": clk.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
               "final_loss": final_loss, "top_entry_points_ms_per_step": [[round(t, 3), c, k] for t, c, k in top_list],
               "extras": extras, "lib": lib.version()}
        print(json.dumps(out))
    if dist:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
