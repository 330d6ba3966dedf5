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
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch_b200"],
                    help="b200: this repo's CUDA path; reference: the reference algorithm on the host cores; torch_b200: the "
                         "same network through stock PyTorch (cuDNN / ATen) on this GPU, the re-baselined comparator")
    ap.add_argument("--workload", default="train", choices=["train", "search"],
                    help="train: BASELINE config 2 (arch0 training iteration); search: BASELINE config 4 (NAS candidates/hour)")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--cuda-graph", type=int, default=1, help="replay the training iteration as one CUDA graph (default on)")
    ap.add_argument("--profile-out", default=None, help="write the per-call CUDA-event table of one instrumented step")
    ap.add_argument("--depth-head-only", action="store_true", help="(internal) print the BASELINE config-5 timings as JSON and exit")
    ap.add_argument("--metric-only", action="store_true", help="(internal) print the confusion-matrix timings as JSON and exit")
    ap.add_argument("--augment-only", action="store_true", help="(internal) print the batch-augmentation timings as JSON and exit")
    ap.add_argument("--n-task0", type=int, default=4000, help="search workload: cached task0 crops per candidate")
    ap.add_argument("--task1-iters", type=int, default=297, help="search workload: train_segmenter iterations of task 1 "
                    "(297 = one epoch of batch 32 over the reference's 90 %% meta-train split of VOC train+)")
    ap.add_argument("--val-images", type=int, default=1024, help="search workload: validation images per validate() call "
                    "(the reference's 10 %% meta-val split is 1058)")
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
        "dtype": "f32", "data": "synthetic", "config": workload_config(a, sample_batch=1),
        "note": "CPU arm: ONE host process whatever --gpus is (rank 0 only); images/s normalises the 1-image sample",
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(a, sample_batch=None):
    """`sample_batch`: the batch a bounded CPU sample actually ran (the reference arm times 1 image per step of the batch-8
    workload; images/s is what makes the two arms comparable)."""
    b = a.batch if sample_batch is None else sample_batch
    return {"workload": "WACV arch0 (mbv2 2-tap encoder + TemplateDecoder agg64 rep2, %d classes) training iteration "
                        "fwd+CE+bwd+clip+SGD/Adam, BN train mode, batch %d @%dx%d%s"
                        % (NUM_CLASSES, b, a.width, a.height,
                           "" if sample_batch is None else " (bounded sample of the batch-%d workload)" % a.batch),
            "global_batch_per_gpu": b, "resolution": [a.width, a.height], "l2": "inputs_exceed_L2",
            "parallelism": "one candidate per GPU (replicas), 1 all-gather of 16 B per step",
            "cuda_graph": bool(getattr(a, "cuda_graph", 0))}



# ------------------------------------------------------------------------------------------------------------ stock PyTorch on the B200
def torch_b200_numbers(dev, a, budget_s=150.0):
    """The same network and iteration through stock PyTorch (cuDNN / ATen kernels) on THIS GPU: the reference's own module
    graph as restated by oracle/nas_oracle.py (the reference itself cannot travel to the GPU box), fp32 and bf16 autocast +
    channels_last, eager and as a CUDA graph.  This is the "52.25 ms on a 1080Ti, re-baselined on 1xB200" comparator of the
    north star (reference README.md:80-81).  Comparator only: nothing of the product path runs here."""
    from oracle import nas_oracle as O
    t_start = time.time()
    res = {}
    torch.backends.cudnn.benchmark = True  # let cuDNN pick its fastest algorithm per shape (the warm-up absorbs the search)

    def params(seed_e=1, seed_d=2):
        Pe, Pd = O.Params(seed=seed_e), O.Params(seed=seed_d)
        with torch.no_grad():
            O.template_decoder(O.mbv2_encoder(torch.zeros(2, 3, 32, 32), Pe, (1, 2)), Pd, W0, [24, 32], NUM_CLASSES, 64, 2)
        for P in (Pe, Pd):
            for k in list(P.sd):
                v = P.sd[k].to(dev)
                P.sd[k] = v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v
        return Pe, Pd

    def timed(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    # ---- eval forward, batch 1 (README table shapes)
    for genotype, gname in ((W0, "arch0"), (W1, "arch1")):
        Pe, Pd = O.Params(seed=1), O.Params(seed=2)
        with torch.no_grad():
            O.template_decoder(O.mbv2_encoder(torch.zeros(2, 3, 32, 32), Pe, (1, 2)), Pd, genotype, [24, 32], NUM_CLASSES, 64, 2)
        for P in (Pe, Pd):
            for k in list(P.sd):
                v = P.sd[k].to(dev)
                P.sd[k] = v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v
        for (hh, ww) in ((1024, 2048), (360, 480)):
            x = torch.randn(1, 3, hh, ww, device=dev).contiguous(memory_format=torch.channels_last)
            for dn in ("f32", "bf16"):
                if time.time() - t_start > budget_s:
                    break

                def fwd():
                    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=dn == "bf16"):
                        return O.template_decoder(O.mbv2_encoder(x, Pe, (1, 2)), Pd, genotype, [24, 32], NUM_CLASSES, 64, 2)
                key = "%s_fwd_ms_b1_%dx%d_%s" % (gname, ww, hh, dn)
                try:
                    res[key] = timed(fwd, 10)
                    s = torch.cuda.Stream()
                    s.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(s):
                        fwd()
                    torch.cuda.current_stream().wait_stream(s)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        fwd()
                    res[key + "_cudagraph"] = timed(g.replay, 10)
                    del g
                except Exception as e:  # noqa: BLE001
                    res[key + "_error"] = repr(e)[:200]
        del Pe, Pd
        torch.cuda.empty_cache()

    # ---- training iteration, batch 8 @2048x1024 (the headline workload)
    img, lab = synth(a.batch, a.height, a.width)
    for dn in ("bf16", "f32"):
        if time.time() - t_start > budget_s:
            res["train_%s_skipped" % dn] = "time budget"
            continue
        try:
            Pe, Pd = params()
            for P in (Pe, Pd):
                P.requires_grad_()
            p_enc = [v for v in Pe.sd.values() if v.requires_grad]
            p_dec = [v for v in Pd.sd.values() if v.requires_grad]
            optim_enc = torch.optim.SGD(p_enc, lr=1e-3, momentum=0.9, weight_decay=1e-5)
            optim_dec = torch.optim.Adam(p_dec, lr=3e-3, weight_decay=1e-5, capturable=True)
            x = img.to(dev).contiguous(memory_format=torch.channels_last)
            lab_d = lab.to(dev)

            def step():
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dn == "bf16"):
                    out = O.template_decoder(O.mbv2_encoder(x, Pe, (1, 2), True), Pd, W0, [24, 32], NUM_CLASSES, 64, 2,
                                             training=True)
                y = O.nearest_labels(lab_d, out.shape[2:])
                loss = O.segm_loss(out.float(), y)
                optim_enc.zero_grad()
                optim_dec.zero_grad()
                loss.backward()
                nn.utils.clip_grad_norm_(p_enc, 3.0)
                nn.utils.clip_grad_norm_(p_dec, 3.0)
                optim_enc.step()
                optim_dec.step()
                return loss
            ms = timed(step, max(3, min(a.steps, 8)))
            res["train_ms_per_step_b%d_%s_eager" % (a.batch, dn)] = ms
            res["train_images_per_sec_%s_eager" % dn] = a.batch / ms * 1e3
            res["train_peak_mem_gb_%s" % dn] = torch.cuda.max_memory_allocated() / 1e9
            if dn == "bf16" and time.time() - t_start < budget_s:
                try:
                    s = torch.cuda.Stream()
                    s.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(s):
                        step()
                    torch.cuda.current_stream().wait_stream(s)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        step()
                    ms = timed(g.replay, max(3, min(a.steps, 8)))
                    res["train_ms_per_step_b%d_bf16_cudagraph" % a.batch] = ms
                    res["train_images_per_sec_bf16_cudagraph"] = a.batch / ms * 1e3
                    del g
                except Exception as e:  # noqa: BLE001
                    res["train_bf16_cudagraph_error"] = repr(e)[:200]
            del Pe, Pd, p_enc, p_dec, optim_enc, optim_dec, x
        except Exception as e:  # noqa: BLE001
            res["train_%s_error" % dn] = repr(e)[:300]
        torch.cuda.empty_cache()
    res["what"] = ("stock PyTorch %s (cuDNN/ATen, cudnn.benchmark on; channels_last; bf16 = torch.autocast) running the reference's module graph "
                   "(oracle/nas_oracle.py restatement) on this GPU" % torch.__version__)
    return res


def run_torch_b200(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    res = torch_b200_numbers(dev, a, budget_s=400.0)
    best = max((v for k, v in res.items() if k.startswith("train_images_per_sec")), default=None)
    ms = min((v for k, v in res.items() if k.startswith("train_ms_per_step")), default=None)
    print(json.dumps({"impl": "torch_b200", "metric": METRIC, "value": best, "unit": "images/s", "n_gpus": 1, "steps": a.steps,
                      "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "bf16", "data": "synthetic", "config": workload_config(a), "details": res}))


# ------------------------------------------------------------------------------------------------- augmentation (row f3)
def augment_numbers(dev):
    """The training transform chain of src/data/loaders.py:43-56 for a batch of 32 VOC-sized images (375x500 -> ResizeScale(400,
    0.7-1.4) -> mirror -> 350x350 crop -> normalise) through nas_segm_b200.data (one kernel launch): from pinned-able host
    uint8 arrays (raw bytes uploaded inside the timed region) and from device-resident raw images; beside it the same chain
    with cv2 / numpy calls on ONE host core (what one loader worker of the reference does per image)."""
    from nas_segm_b200.data import GpuTrainTransform
    rs = np.random.RandomState(0)
    norm = (1.0 / 255, np.array([0.485, 0.456, 0.406]).reshape((1, 1, 3)), np.array([0.229, 0.224, 0.225]).reshape((1, 1, 3)))
    host = [{"image": rs.randint(0, 256, (375, 500, 3)).astype(np.uint8), "mask": rs.randint(0, 21, (375, 500)).astype(np.uint8)}
            for _ in range(32)]
    resident = [{"image": torch.from_numpy(h["image"]).to(dev), "mask": torch.from_numpy(h["mask"]).to(dev)} for h in host]
    t = GpuTrainTransform(400, 0.7, 1.4, False, 350, norm, device=dev)
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, batch in (("host_uint8", host), ("device_uint8", resident)):
        np.random.seed(1)
        for _ in range(3):
            t(batch)
        torch.cuda.synchronize()
        w0 = time.time()
        e0.record()
        for _ in range(10):
            o = t(batch)
        e1.record()
        torch.cuda.synchronize()
        out["augment_b32_375x500_to_350_%s_ms" % name] = e0.elapsed_time(e1) / 10
        out["augment_b32_%s_images_per_s_wall" % name] = 320.0 / (time.time() - w0)
    out["h2d_bytes_per_image_raw_uint8"] = 375 * 500 * 4
    out["h2d_bytes_per_image_reference_float32_crop"] = 350 * 350 * 3 * 4 + 350 * 350
    try:
        import cv2
        cv2.setNumThreads(1)
        np.random.seed(1)
        w0, n = time.time(), 0
        while time.time() - w0 < 3.0:
            s = host[n % 32]
            sc = np.random.uniform(0.7, 1.4)
            sc = max(sc, 400.0 / 375)
            im = cv2.resize(s["image"], None, fx=sc, fy=sc, interpolation=cv2.INTER_CUBIC)
            mk = cv2.resize(s["mask"], None, fx=sc, fy=sc, interpolation=cv2.INTER_NEAREST)
            if np.random.randint(2):
                im, mk = cv2.flip(im, 1), cv2.flip(mk, 1)
            top, left = np.random.randint(0, im.shape[0] - 350 + 1), np.random.randint(0, im.shape[1] - 350 + 1)
            im, mk = im[top:top + 350, left:left + 350], mk[top:top + 350, left:left + 350]
            im = ((norm[0] * im - norm[1]) / norm[2]).transpose(2, 0, 1)
            torch.from_numpy(np.ascontiguousarray(im)).float()
            n += 1
        out["cv2_chain_images_per_s_1core"] = n / (time.time() - w0)
        out["cv2_version"] = cv2.__version__
    except Exception as e:  # noqa: BLE001
        out["cv2_chain_error"] = repr(e)[:200]
    out["check"] = "bit-exact vs the reference's transform classes: tests/test_gpu_augment.py (fixture tests/golden/augment.npz)"
    return out


# ------------------------------------------------------------------------------------------------------------ metric (fast_cm)
def metric_numbers(dev):
    """the confusion-matrix kernels against the HBM roofline (label path = 2 bytes per pixel; fused
    upsample+argmax path = the bf16/fp32 logits + 1 label byte per full-resolution pixel) next to the reference's own Cython
    fast_cm (oracle/_ref, built from src/helpers/miou_utils.pyx:7-30; single-threaded by construction) and the C
    restatement on the same arrays."""
    import nas_segm_b200  # noqa: F401
    from nas_segm_b200 import functional as Fn
    peak, peak_src = peaks()
    res = {"peak_gbs": peak, "peak_source": peak_src}
    g = torch.Generator().manual_seed(9314)
    B, H, W, C = 8, 1024, 2048, NUM_CLASSES
    n = B * H * W
    # several label sets, rotated, so that no launch finds its input in the 126 MB L2
    sets = []
    for _ in range(6):
        gt = torch.randint(0, C, (n,), generator=g, dtype=torch.int64).to(torch.uint8)
        gt[torch.rand(n, generator=g) < 0.05] = 255
        pr = torch.randint(0, C, (n,), generator=g, dtype=torch.int64).to(torch.uint8)
        sets.append((pr.to(dev), gt.to(dev)))
    cm = torch.zeros((C, C), dtype=torch.int64, device=dev)

    def timed(fns, iters=30):
        """Average device time per call: the calls are captured into ONE CUDA graph (one pass over the rotating buffer sets)
        and the graph is replayed -- eagerly a 25 us kernel would be timed at the ~20 us the host needs to issue it."""
        for f in fns:
            f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for f in fns:
                f()
        reps = max(iters // len(fns), 2)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (reps * len(fns))

    ms = timed([lambda p=p, q=q: Fn.confmat_labels(p, q, C, cm) for p, q in sets])
    res["confmat_labels_ms_%dMpx_C%d" % (n >> 20, C)] = ms
    res["confmat_labels_gbs"] = 2.0 * n / (ms * 1e-3) / 1e9
    res["confmat_labels_frac_of_hbm_peak"] = res["confmat_labels_gbs"] / peak
    res["confmat_labels_gpx_per_s"] = n / (ms * 1e-3) / 1e9
    # the same on label maps with spatial structure (64x64-pixel regions of one class, 5 % ignore pixels): uniform random
    # labels are the worst case of any histogram -- every pixel is a different (gt, pred) pair -- real maps are not
    bsets = []
    for _ in range(6):
        small = torch.randint(0, C, (B, H // 64, W // 64), generator=g)
        gtb = small.repeat_interleave(64, 1).repeat_interleave(64, 2).to(torch.uint8)
        prb = gtb.clone()
        flip = torch.rand(B, H, W, generator=g)
        prb[flip < 0.1] = torch.randint(0, C, (int((flip < 0.1).sum()),), generator=g).to(torch.uint8)  # 10 % errors
        gtb[flip > 0.95] = 255
        bsets.append((prb.view(-1).to(dev), gtb.view(-1).to(dev)))
    ms = timed([lambda p=p, q=q: Fn.confmat_labels(p, q, C, cm) for p, q in bsets])
    res["confmat_labels_structured_ms"] = ms
    res["confmat_labels_structured_gbs"] = 2.0 * n / (ms * 1e-3) / 1e9
    res["confmat_labels_structured_frac_of_hbm_peak"] = res["confmat_labels_structured_gbs"] / peak
    del bsets
    for dt, dn in ((torch.float32, "f32"),):
        lsets = []
        for i in range(4):
            lg = torch.randn(B, H // 4, W // 4, C, generator=g).to(dev).to(dt).permute(0, 3, 1, 2)
            lsets.append((lg, sets[i][1].view(B, H, W)))
        ms = timed([lambda lg=lg, q=q: Fn.confmat_logits(lg, q, C, cm) for lg, q in lsets], iters=20)
        nbytes = B * (H // 4) * (W // 4) * C * (4 if dt == torch.float32 else 2) + n
        res["confmat_logits_ms_b8_19x256x512_to_2048x1024_%s" % dn] = ms
        res["confmat_logits_gbs_%s" % dn] = nbytes / (ms * 1e-3) / 1e9
        res["confmat_logits_frac_of_hbm_peak_%s" % dn] = res["confmat_logits_gbs_%s" % dn] / peak
        res["confmat_logits_gpx_per_s_%s" % dn] = n / (ms * 1e-3) / 1e9
    # CPU: the reference's Cython and the C restatement on ONE 2048x1024 image (2.1 Mpx), 1 core each
    pr_h, gt_h = sets[0][0][:H * W].cpu().numpy(), sets[0][1][:H * W].cpu().numpy()
    valid = gt_h < C  # inference.py:65: the reference caller pre-masks
    pr_v, gt_v = np.ascontiguousarray(pr_h[valid]), np.ascontiguousarray(gt_h[valid])
    try:
        from oracle import build_ref, miou_oracle
        ref = build_ref.load()
        if ref is not None:
            ref.fast_cm(pr_v, gt_v, C)
            t0 = time.time()
            for _ in range(5):
                cm_ref = ref.fast_cm(pr_v, gt_v, C)
            dt_ref = (time.time() - t0) / 5
            res["cython_fast_cm_ms_2Mpx_1core"] = dt_ref * 1e3
            res["cython_fast_cm_gpx_per_s_1core"] = pr_v.size / dt_ref / 1e9
            cm_gpu = Fn.confmat_labels(sets[0][0][:H * W].contiguous(), sets[0][1][:H * W].contiguous(), C).cpu().numpy()
            res["bit_exact_vs_cython"] = bool(np.array_equal(cm_gpu, np.asarray(cm_ref)))
        else:
            res["cython_fast_cm"] = "oracle/_ref not built on this box"
        miou_oracle.fast_cm_c(pr_v, gt_v, C)
        t0 = time.time()
        for _ in range(5):
            miou_oracle.fast_cm_c(pr_v, gt_v, C)
        dt_c = (time.time() - t0) / 5
        res["c_port_fast_cm_gpx_per_s_1core"] = pr_v.size / dt_c / 1e9
    except Exception as e:  # noqa: BLE001
        res["cpu_metric_error"] = repr(e)[:200]
    return res



# ------------------------------------------------------------------------------------------------------------ search loop (config 4)
def search_numbers(a, dev, rank, world, rounds, warm_rounds, n_task0, task1_iters, val_images, deadline_s=None):
    """BASELINE config 4 through the product's own outer loop: engine.search.search_rounds + evaluate_candidate, i.e. the
    per-candidate recipe of the reference's main_search.py:548-680 -- a FRESH segmenter per candidate (MobileNet-v2 4-tap
    encoder + sampled MicroDecoder, agg 48, aux cells, 21 classes; model build and CUDA-graph capture are inside the timed
    region), task 0 = 5 epochs of train_task0 on n_task0 cached 256x256 crops (batch 64, KD-MSE + aux CE, Adam, clip,
    Polyak) + validate, TaskPerformer decision, task 1 = one epoch of train_segmenter (batch 32 @350x350, host batches)
    + validate, reward.  One candidate per rank per round; records are exchanged with the single all-gather
    (sync_every = rounds: a rank stopped early moves on instead of idling; the synchronous figure is derived from the
    per-candidate times of the same run).  Candidates come from search.uniform_sampler (the controller is outside the hot
    path).  populate_task0 runs once, outside the timed region, and is reported separately."""
    import types
    import nas_segm_b200
    from nas_segm_b200 import parallel
    from nas_segm_b200.engine import search, trainer
    from nas_segm_b200.nn.encoders import mbv2
    from nas_segm_b200.nn.micro_decoders import MicroDecoder
    nas_segm_b200.set_act_dtype(torch.bfloat16)
    nas_segm_b200.config().cuda_graphs = True
    np.random.seed(9314 + rank)
    g = torch.Generator().manual_seed(9314 + rank)

    class Loader(list):
        class _DS:
            def set_stage(self, s):
                pass
        dataset = _DS()
        batch_sampler = types.SimpleNamespace(batch_size=1)

    t0 = time.time()
    enc0 = mbv2()
    seg0 = Wrapper(Seg(enc0, nn.Identity()).to(dev))
    n_real = min(n_task0, 64)
    train0 = Loader({"image": torch.randn(1, 3, 256, 256, generator=g),
                     "mask": torch.randint(0, 21, (1, 256, 256), generator=g).to(torch.uint8)} for _ in range(n_real))
    Xy = trainer.populate_task0(seg0, train0, None, n_real, do_kd=False)
    assert Xy != 0, "populate_task0 swallowed a RuntimeError"
    torch.cuda.synchronize()
    t_populate_per_image = (time.time() - t0) / n_real
    reps = (n_task0 + n_real - 1) // n_real  # the cached block is tiled to n_task0 samples: content is irrelevant to timing
    for k in list(Xy.keys()):
        if k == "out_size":
            continue
        if k == "y":
            Xy[k] = Xy[k].repeat(reps, 1, 1)[:n_task0].contiguous()
        else:
            Xy[k] = Xy[k].permute(0, 2, 3, 1).repeat(reps, 1, 1, 1)[:n_task0].contiguous().permute(0, 3, 1, 2)
    Xy["kd_y"] = torch.randn(n_task0, 64, 64, 21, device=dev,
                             generator=torch.Generator(device=dev).manual_seed(1)).permute(0, 3, 1, 2)
    del seg0, enc0
    pinned1 = [{"image": torch.randn(32, 3, 350, 350, generator=g).pin_memory(),
                "mask": torch.randint(0, 21, (32, 350, 350), generator=g).to(torch.uint8).pin_memory()} for _ in range(2)]
    train1 = Loader(pinned1[i % 2] for i in range(task1_iters))
    pinned_v = [{"image": torch.randn(64, 3, 400, 400, generator=g).pin_memory(),
                 "mask": torch.randint(0, 21, (64, 400, 400), generator=g).to(torch.uint8).pin_memory()} for _ in range(2)]
    val = Loader(pinned_v[i % 2] for i in range(max(val_images // 64, 1)))
    args = types.SimpleNamespace(
        num_tasks=2, enc_optim="sgd", dec_optim="adam", enc_lr=[1e-3, 1e-3], dec_lr=[3e-3, 3e-3], enc_mom=[0.9] * 3,
        dec_mom=[0.9] * 3, enc_wd=[1e-5] * 3, dec_wd=[1e-5] * 3, do_polyak=True, num_segm_epochs=[5, 1], val_every=[5, 1],
        segm_crit=nn.NLLLoss(ignore_index=255), kd_crit=nn.MSELoss(), batch_size=[64, 32], freeze_bn=[False, False],
        do_kd=True, kd_coeff=0.3, dec_grad_clip=3.0, enc_grad_clip=3.0, dec_aux_weight=0.15, print_every=20,
        num_classes=[21, 21], val_omit_classes=[0])  # utils/default_args.py of the reference
    task_ps = search.make_task_performers(args.num_segm_epochs, args.val_every)
    sampler = search.uniform_sampler(seed=9314)
    per_candidate, errors = [], []

    def build(cfg):
        torch.manual_seed(0)
        enc = mbv2()
        dec = MicroDecoder(list(enc.out_sizes), 21, cfg, agg_size=48, aux_cell=True, repeats=1)
        return Wrapper(Seg(enc, dec).to(dev))

    from nas_segm_b200.engine import inference as _inference
    phases = {"train_task0": 0.0, "train_segmenter": 0.0, "validate": 0.0}

    def _timed(name, fn):  # per-phase host time (each engine call ends with a device sync of its own: the logged loss / the reward)
        def wrapped(*a, **k):
            t = time.time()
            try:
                return fn(*a, **k)
            finally:
                torch.cuda.synchronize()
                phases[name] += time.time() - t
        return wrapped

    class engine_ns:  # noqa: N801
        train_task0 = staticmethod(_timed("train_task0", trainer.train_task0))
        train_segmenter = staticmethod(_timed("train_segmenter", trainer.train_segmenter))
        validate = staticmethod(_timed("validate", _inference.validate))

    def evaluate(seg, cfg):
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        try:
            out = search.evaluate_candidate(seg, Xy, train1, val, args, task_ps, engine=engine_ns)
        except Exception as e:  # noqa: BLE001 -- keep the ranks' collectives matched whatever one candidate does
            errors.append(repr(e)[:200])
            out = (0.0, 1)
        ev1.record()
        torch.cuda.synchronize()
        per_candidate.append((ev0.elapsed_time(ev1) / 1e3, int(out[1]), float(out[0])))
        return out

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    if warm_rounds:
        search.search_rounds(warm_rounds, lambda r, s: sampler(10000 + r, s), build, evaluate, sync_every=warm_rounds)
    n_warm = len(per_candidate)
    for k in phases:
        phases[k] = 0.0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.time()
    e0.record()
    hist = search.search_rounds(rounds, sampler, build, evaluate, sync_every=rounds)
    e1.record()
    barrier()
    t_wall = time.time() - t_wall
    t_mine = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev)
    times = torch.tensor([[c[0], float(c[1])] for c in per_candidate[n_warm:]], dtype=torch.float32, device=dev)  # [rounds, 2]
    if world > 1:
        import torch.distributed as td
        td.all_reduce(t_mine, op=td.ReduceOp.MAX)
        allt = torch.empty((world,) + tuple(times.shape), dtype=torch.float32, device=dev)
        td.all_gather_into_tensor(allt, times)
    else:
        allt = times[None]
    allt = allt.cpu().numpy()  # [world, rounds, 2]
    t_block = float(t_mine.item())                          # max over ranks, device timeline, one exchange at the end
    t_sync = float(allt[:, :, 0].max(axis=0).sum())         # what per-round all-gathers would have cost
    n = world * rounds
    return {"metric": "search_loop_candidates_per_hour", "value": n * 3600.0 / t_block, "unit": "candidates/h",
            "n_gpus": world, "rounds": rounds, "warmup_rounds": warm_rounds, "candidates": n,
            "seconds_per_round": t_block / rounds, "wall_s": t_wall,
            "candidates_per_hour_synchronous_rounds": n * 3600.0 / t_sync,
            "straggler_loss_of_synchronous_rounds": 1.0 - t_block / t_sync if t_sync > 0 else None,
            "per_candidate_s": [[round(float(v), 3) for v in allt[r, :, 0]] for r in range(world)],
            "epochs_run": [[int(v) for v in allt[r, :, 1]] for r in range(world)],
            "rewards": [[round(float(t[s, 0]), 5) for s in range(world)] for t in hist],
            "phase_seconds_per_candidate_rank0": {k: round(v / max(rounds, 1), 3) for k, v in phases.items()},
            "populate_task0_s_per_image": t_populate_per_image, "n_task0": n_task0, "task1_iterations": task1_iters,
            "max_memory_reserved_gb": round(torch.cuda.max_memory_reserved() / 2 ** 30, 2),
            "val_images": len(val) * 64, "errors": errors[:3],
            "recipe": "task0: 5 epochs x %d it (batch 64, 64x64 feats, KD+aux, Adam, clip, Polyak) + validate; TaskPerformer; "
                      "task1: %d it of train_segmenter (batch 32 @350x350, host batches) + validate; fresh model + CUDA-graph "
                      "capture per candidate inside the timed region; bf16" % (n_task0 // 64, task1_iters)}


def run_search(a):
    from nas_segm_b200 import parallel
    rank, world, dev = parallel.init()
    res = search_numbers(a, dev, rank, world, rounds=max(a.steps, 1), warm_rounds=min(max(a.warmup, 0), 1),
                         n_task0=a.n_task0, task1_iters=a.task1_iters, val_images=a.val_images)
    if rank == 0:
        res.update({"steps": a.steps, "warmup": a.warmup, "ms_per_step": res["seconds_per_round"] * 1e3, "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                    "config": {"workload": "CVPR search loop (BASELINE config 4): " + res.pop("recipe"),
                               "parallelism": "one candidate per GPU per round, 1 all-gather of 16 B per candidate"}})
        print(json.dumps(res))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


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
            run(x)  # eager: plain module call, BN fold + operand pack per unit and call (weights may change between calls)
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
    if a.impl == "torch_b200":
        return run_torch_b200(a)
    if a.metric_only:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        print(json.dumps(metric_numbers(torch.device("cuda", torch.cuda.current_device()))))
        return
    if a.augment_only:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        print(json.dumps(augment_numbers(torch.device("cuda", torch.cuda.current_device()))))
        return
    if a.workload == "search":
        return run_search(a)
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
        # per-kernel durations need the kernels one after another: the side-stream weight gradients and the concurrent
        # decoder branches of the timed runs above would overlap (and stretch) the launches being measured
        cfg_ = nas_segm_b200.config()
        saved = (cfg_.async_wgrad, cfg_.branch_streams)
        cfg_.async_wgrad = cfg_.branch_streams = False
        try:
            trainer.segmenter_step(seg, img_d, lab_d, optim_enc, optim_dec, crit, 3.0, 3.0, False)
            lib.profile_begin()
            for _ in range(n_prof):
                trainer.segmenter_step(seg, img_d, lab_d, optim_enc, optim_dec, crit, 3.0, 3.0, False)
            prof = lib.profile_end()
        finally:
            cfg_.async_wgrad, cfg_.branch_streams = saved
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
            def child(argv, timeout):
                """A measurement in a child process with a deadline: it can never cost the headline line."""
                env = {k: v for k, v in dict(os.environ, LOCAL_RANK=str(local)).items()
                       if k not in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
                try:
                    c = subprocess.run([sys.executable, os.path.abspath(__file__)] + argv, capture_output=True, text=True,
                                       timeout=timeout, env=env)
                    line = [ln for ln in c.stdout.splitlines() if ln.startswith("{")]
                    if c.returncode == 0 and line:
                        return json.loads(line[-1])
                    return {"error": ("rc=%d " % c.returncode) + c.stderr[-300:]}
                except Exception as e:  # noqa: BLE001
                    return {"error": repr(e)[:300]}
            # the same network through stock PyTorch on this GPU (north star: "re-baselined on 1xB200")
            tb = child(["--impl", "torch_b200", "--steps", str(min(a.steps, 8)), "--warmup", "3"], 420)
            extras["torch_b200"] = tb.get("details", tb)
            # confusion-matrix kernels vs HBM roofline, Cython fast_cm beside them
            extras["metric"] = child(["--metric-only"], 180)
            # row f3: the loaders' transform chain as one kernel over raw uint8 batches, one cv2 worker beside it
            extras["augment"] = child(["--augment-only"], 120)
            # BASELINE config 4 through engine.search (task0 + task1, TaskPerformer, per-candidate rebuild + capture)
            sb = child(["--workload", "search", "--steps", "3", "--warmup", "1"], 480)
            sb.pop("config", None)
            extras["search_loop"] = sb
        finally:
            nas_segm_b200.set_act_dtype(dtype)

    if dist and not a.no_extras:
        # candidates/hour at this N through the same outer loop (one candidate per rank per round); every rank takes part.
        # Errors inside a candidate are recorded, not raised, so the ranks' collectives stay matched.
        del img_d, lab_d
        seg = optim_enc = optim_dec = graphed = None
        torch.cuda.empty_cache()
        try:
            sb = search_numbers(a, dev, rank, world, rounds=3, warm_rounds=1, n_task0=a.n_task0, task1_iters=a.task1_iters,
                                val_images=a.val_images)
            sb.pop("recipe", None)
            extras["search_loop"] = sb
        except Exception as e:  # noqa: BLE001
            extras["search_loop"] = {"error": repr(e)[:300]}
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
