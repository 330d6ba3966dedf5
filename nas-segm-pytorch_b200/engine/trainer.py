"""Training functions with the reference's signatures (src/engine/trainer.py) on the fused kernel path.

  populate_task0(segmenter, train_loader, kd_net, n_train, do_kd=False) -> dict            (trainer.py:17-74)
  train_task0(Xy_train, segmenter, optim_dec, epoch, segm_crit, kd_crit, batch_size, freeze_bn, do_kd, kd_coeff,
              dec_grad_clip, do_polyak, avg_param=None, polyak_decay=0.9, aux_weight=0) -> None   (trainer.py:78-175)
  train_segmenter(segmenter, train_loader, optim_enc, optim_dec, epoch, segm_crit, freeze_bn, enc_grad_clip,
              dec_grad_clip, do_polyak, print_every=10, aux_weight=-1, avg_param=None, polyak_decay=0.99) -> None
                                                                                                 (trainer.py:179-283)
``segmenter`` is the DataParallel-style wrapper: ``segmenter.module.{encoder,decoder}``.  All four engine functions
are wrapped by ``try_except``: RuntimeError -> return 0.

Differences that do not change results: the loss (bilinear resize + LogSoftmax + NLL(ignore) [+ KD-MSE] [+ aux]) runs in
the fused loss kernels instead of four torch passes, and the per-iteration ``loss.item()`` device sync of the reference
(trainer.py:165) is replaced by one sync per epoch for the logged average.
"""
import logging
import time
from collections import defaultdict

import numpy as np
import torch
from torch import nn

from .. import config as _config
from .. import functional as Fn
from .. import lib, packs
from ..graphs import StepGraph, make_capturable
from ..optim import FusedStep
from ..helpers.utils import try_except

logger = logging.getLogger(__name__)


def _ignore_index(crit):
    return int(getattr(crit, "ignore_index", 255))


def _segm_loss(crit, logits, target, size=None):
    """segm_crit(LogSoftmax(resize(logits)), target) -- fused when crit is the reference's NLL criterion."""
    if size is not None:
        logits = Fn.resize(Fn.lib.to_nhwc(logits), size)
    if crit is None or isinstance(crit, (nn.NLLLoss,)) or hasattr(crit, "ignore_index"):
        return Fn.cross_entropy2d(logits, target, _ignore_index(crit) if crit is not None else 255)
    return crit(nn.functional.log_softmax(logits.float(), dim=1), target)


def _finish_step(owner, entries, polyak_params, avg_param, polyak_decay):
    """clip_grad_norm_ + optimiser.step() for every (optim, clip_params, max_norm) entry, then the Polyak average
    (trainer.py:163-169,258-272).  Plain SGD / Adam run as two multi-tensor launches (optim.FusedStep, cached on `owner`
    with strong references to everything its key names); anything else keeps the caller's own torch calls."""
    do_polyak = avg_param is not None and polyak_params is not None
    lib.wgrad_stream.join()  # the weight gradients of this iteration were computed on the side stream
    if _config().fused_optim and FusedStep.supported(entries):
        key = tuple(id(o) for o, _, _ in entries) + tuple(float(mn) for _, _, mn in entries) + (id(avg_param) if do_polyak else 0,)
        cached = getattr(owner, "_nasb_fused_step", None)
        if cached is None or cached[0] != key:
            plist = list(polyak_params) if do_polyak else None
            cached = (key, FusedStep([(o, list(cp), mn) for o, cp, mn in entries], plist, avg_param if do_polyak else None))
            owner._nasb_fused_step = cached
        cached[1].step(polyak_decay if do_polyak else 0.0)
        return
    for o, clip_params, max_norm in entries:
        if max_norm > 0:
            nn.utils.clip_grad_norm_(clip_params, max_norm)
        o.step()
    if do_polyak:
        for p, avg_p in zip(polyak_params, avg_param):
            avg_p.mul_(polyak_decay).add_(p.data, alpha=1.0 - polyak_decay)


def _hyper_key(*optims):
    """Hyper-parameters a captured iteration bakes in (kernel arguments of the fused step / torch's own scalars)."""
    key = []
    for o in optims:
        for g in o.param_groups:
            key.append(tuple((k, v) for k, v in sorted(g.items()) if k != "params" and isinstance(v, (int, float, bool, tuple, type(None)))))
    return tuple(key)


def _set_stage(loader, stage):
    try:
        loader.dataset.set_stage(stage)
    except AttributeError:
        try:
            loader.dataset.dataset.set_stage(stage)
        except AttributeError:
            pass


@try_except
def populate_task0(segmenter, train_loader, kd_net, n_train, do_kd=False):
    """Cache the encoder's outputs (NHWC, activation dtype), nearest-resized int64 labels and optional KD targets."""
    Xy_train = defaultdict(list)
    segmenter.eval()
    _set_stage(train_loader, "train")
    try:
        train_loader.batch_sampler.batch_size = 1  # reference: batch 1 "to not run out of memory"
    except AttributeError:
        pass
    # weights are constant over the whole loop: one multi-tensor pack of the encoder's tensor-core operands
    with torch.no_grad(), packs.scope(segmenter.module.encoder):
        n_curr = 0
        for sample in train_loader:
            image = sample["image"].float().cuda()
            target = sample["mask"].float()
            enc_outputs = segmenter.module.encoder(image)
            size = enc_outputs[0].size()[2:]
            for i, enc_output in enumerate(enc_outputs):
                Xy_train[i].append(enc_output.permute(0, 2, 3, 1))
            Xy_train["y"].append(nn.functional.interpolate(target[:, None], size=size, mode="nearest").long()
                                 .squeeze(dim=1).cuda())
            if do_kd:
                kd_y = kd_net(image)
                Xy_train["kd_y"].append(Fn.resize(Fn.lib.to_nhwc(kd_y.float()), size).permute(0, 2, 3, 1))
            n_curr += image.size(0)
            if n_curr >= n_train:
                Xy_train["out_size"] = size
                logger.info(" Populated Xy_train, N = {}".format(n_curr))
                break
        for k, v in list(Xy_train.items()):
            if k == "out_size":
                continue
            cat = torch.cat(v, 0)
            Xy_train[k] = cat if k == "y" else cat.permute(0, 3, 1, 2)  # logical NCHW over NHWC storage
    return Xy_train


def _gather(t, idx):
    """Rows `idx` of a cached tensor, NHWC storage (GPU-side gather; trainer.py:131-136)."""
    if t.dim() == 4:
        return torch.index_select(t.permute(0, 2, 3, 1), 0, idx).permute(0, 3, 1, 2)
    return torch.index_select(t, 0, idx)


@try_except
def train_task0(Xy_train, segmenter, optim_dec, epoch, segm_crit, kd_crit, batch_size, freeze_bn, do_kd, kd_coeff,
                dec_grad_clip, do_polyak, avg_param=None, polyak_decay=0.9, aux_weight=0):
    """Decoder-only training on cached encoder features."""
    decoder = segmenter.module.decoder
    n_examples = Xy_train[0].size(0)
    batch_size = min(batch_size, n_examples)
    n_passes = n_examples // batch_size
    indices = np.arange(n_examples)
    decoder.train()
    if freeze_bn:
        for m in decoder.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
    np.random.shuffle(indices)
    dev = Xy_train[0].device
    feat_keys = [k for k in Xy_train.keys() if k not in ("y", "kd_y", "out_size")]
    out_size = tuple(Xy_train["out_size"])
    loss_sum = torch.zeros((), dtype=torch.float32, device=dev)
    start = time.time()

    def iteration(idx):
        lib.zero_arena.begin(dev)
        packs.begin(decoder)
        lib.wgrad_stream.begin(dev)
        try:
            return _iteration(idx)
        finally:
            lib.wgrad_stream.end()
            packs.end()
            lib.zero_arena.end()

    def _iteration(idx):
        encoder_outputs = [_gather(Xy_train[k], idx) for k in feat_keys]
        y = _gather(Xy_train["y"], idx)
        output = decoder(encoder_outputs)
        aux_outs = []
        if isinstance(output, tuple):
            output, aux_outs = output
        output = Fn.resize(output, out_size)  # NOTE (reference): output size can change with the connectivity
        loss = _segm_loss(segm_crit, output, y)
        if do_kd:
            loss = loss + kd_coeff * Fn.mse_loss(output, _gather(Xy_train["kd_y"], idx))
        if aux_weight > 0:
            for aux_out in aux_outs:
                loss = loss + _segm_loss(segm_crit, aux_out, y, out_size) * aux_weight
        optim_dec.zero_grad(set_to_none=True)
        loss.backward()
        if dec_grad_clip > 0:
            _finish_step(decoder, [(optim_dec, dec_params, dec_grad_clip)], dec_params if do_polyak else None,
                         avg_param if do_polyak else None, polyak_decay)
        else:  # the reference clips unconditionally (trainer.py:163); a non-positive norm keeps torch's own arithmetic
            nn.utils.clip_grad_norm_(dec_params, dec_grad_clip)
            _finish_step(decoder, [(optim_dec, (), 0.0)], dec_params if do_polyak else None,
                         avg_param if do_polyak else None, polyak_decay)
        return loss

    dec_params = list(decoder.parameters())

    step = iteration
    if _config().cuda_graphs:
        make_capturable(optim_dec)
        sg = getattr(decoder, "_nasb_task0_graph", None)
        key = (id(optim_dec), id(Xy_train), id(avg_param), batch_size, bool(do_kd), float(kd_coeff), float(aux_weight),
               bool(do_polyak), float(polyak_decay), bool(freeze_bn), float(dec_grad_clip), _ignore_index(segm_crit),
               _hyper_key(optim_dec), _config().act_dtype)
        if sg is None or sg.key != key:
            sg = StepGraph(iteration, [torch.zeros(batch_size, dtype=torch.int64, device=dev)])
            sg.key = key
            sg.keep = (optim_dec, Xy_train, avg_param)  # the ids in the key stay unique while the graph is cached
            decoder._nasb_task0_graph = sg
        step = sg
    for i in range(n_passes):
        idx = torch.from_numpy(indices[i * batch_size:(i + 1) * batch_size]).to(dev, non_blocking=True)
        loss = step(idx)
        loss_sum += loss.detach()
    avg_loss = float(loss_sum.item()) / max(n_passes, 1)
    logger.info(" Train epoch: {}\tAvg. Loss: {:.3f}\tAvg. Time: {:.3f}".format(
        epoch, avg_loss, (time.time() - start) / max(n_passes, 1)))


def segmenter_step(segmenter, image, target, optim_enc, optim_dec, segm_crit, enc_grad_clip, dec_grad_clip, do_polyak,
                   aux_weight=-1, avg_param=None, polyak_decay=0.99):
    """One end-to-end iteration on device-resident tensors (the body of trainer.py:226-272): forward, nearest-resized
    target, CE (+ aux), backward, the two grad-norm clips, the two optimiser steps, Polyak.  Returns the loss tensor."""
    lib.zero_arena.begin(image.device)
    packs.begin(segmenter)
    lib.wgrad_stream.begin(image.device)
    try:
        return _segmenter_step(segmenter, image, target, optim_enc, optim_dec, segm_crit, enc_grad_clip, dec_grad_clip,
                               do_polyak, aux_weight, avg_param, polyak_decay)
    finally:
        lib.wgrad_stream.end()
        packs.end()
        lib.zero_arena.end()


def _segmenter_step(segmenter, image, target, optim_enc, optim_dec, segm_crit, enc_grad_clip, dec_grad_clip, do_polyak,
                    aux_weight, avg_param, polyak_decay):
    output = segmenter(image)
    aux_outs = []
    if isinstance(output, tuple):
        output, aux_outs = output
    target_var = nn.functional.interpolate(target[:, None].float(), size=output.size()[2:], mode="nearest").long()[:, 0]
    loss = _segm_loss(segm_crit, output, target_var)
    if aux_weight > 0:
        for aux_out in aux_outs:
            loss = loss + _segm_loss(segm_crit, aux_out, target_var, tuple(target_var.size()[1:])) * aux_weight
    optim_enc.zero_grad(set_to_none=True)
    optim_dec.zero_grad(set_to_none=True)
    loss.backward()
    _finish_step(segmenter, [(optim_enc, segmenter.module.encoder.parameters(), enc_grad_clip),
                             (optim_dec, segmenter.module.decoder.parameters(), dec_grad_clip)],
                 segmenter.parameters() if do_polyak else None, avg_param if do_polyak else None, polyak_decay)
    return loss


@try_except
def train_segmenter(segmenter, train_loader, optim_enc, optim_dec, epoch, segm_crit, freeze_bn, enc_grad_clip,
                    dec_grad_clip, do_polyak, print_every=10, aux_weight=-1, avg_param=None, polyak_decay=0.99):
    """End-to-end training (encoder + decoder)."""
    _set_stage(train_loader, "train")
    segmenter.train()
    if freeze_bn:
        for m in segmenter.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
    loss_sum, n_it, start = None, 0, time.time()
    use_graph = _config().cuda_graphs
    if use_graph:
        make_capturable(optim_enc)
        make_capturable(optim_dec)
    # Host batches are uploaded on a side stream one iteration ahead: the (pinned) host -> device copy of batch i+1 runs
    # under the kernels of batch i instead of in front of them (201 MB of fp32 image per 8 x 2048x1024 batch).
    copy_stream = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    batches = iter(train_loader)

    def upload():
        try:
            sample = next(batches)
        except StopIteration:
            return None
        with torch.cuda.stream(copy_stream):
            im = sample["image"].float().cuda(non_blocking=True)
            tg = sample["mask"].cuda(non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return im, tg, ev

    # The running loss is read back with ONE iteration of lag: iteration i's value is copied into pinned memory behind its
    # kernels and logged after iteration i+1 has been launched, so the host never drains the GPU to print a line (the
    # reference's `.item()` right after the step leaves the device idle for the host's launch latency every print_every
    # iterations; with print_every = 1 that was 1.2 ms of a 23 ms iteration).  Same values, same lines, one launch later.
    host_loss = torch.empty(1, dtype=torch.float32).pin_memory()
    pending = []

    def flush_log():
        while pending:
            pi, pn, buf, ev = pending.pop(0)
            ev.synchronize()
            logger.info(" Train epoch: {} [{}/{}]\tAvg. Loss: {:.3f}\tAvg. Time: {:.3f}".format(
                epoch, pi, len(train_loader), float(buf[0]) / pn, (time.time() - start) / pn))

    ahead, i = upload(), -1
    while ahead is not None:
        image, target, ready = ahead
        i += 1
        main_stream.wait_event(ready)
        image.record_stream(main_stream)
        target.record_stream(main_stream)
        ahead = upload()
        if use_graph:
            sg = getattr(segmenter, "_nasb_step_graph", None)
            # everything the captured iteration bakes in (the same discipline as train_task0's key); `keep` holds strong
            # references so that no id() in the key can be recycled by a new object while the graph is cached
            key = (id(optim_enc), id(optim_dec), id(avg_param), bool(freeze_bn), float(enc_grad_clip), float(dec_grad_clip),
                   float(aux_weight), bool(do_polyak), float(polyak_decay), _ignore_index(segm_crit),
                   _hyper_key(optim_enc, optim_dec), _config().act_dtype)
            if sg is None or sg.key != key or not sg.matches((image, target)):
                sg = StepGraph(lambda im, tg: segmenter_step(segmenter, im, tg, optim_enc, optim_dec, segm_crit, enc_grad_clip,
                                                             dec_grad_clip, do_polyak, aux_weight, avg_param, polyak_decay),
                               [torch.empty_like(image), torch.empty_like(target)])
                sg.key = key
                sg.keep = (optim_enc, optim_dec, avg_param)
                segmenter._nasb_step_graph = sg
            loss = sg(image, target)
        else:
            loss = segmenter_step(segmenter, image, target, optim_enc, optim_dec, segm_crit, enc_grad_clip, dec_grad_clip,
                                  do_polyak, aux_weight, avg_param, polyak_decay)
        # (clone: a replayed graph returns the SAME static loss tensor every iteration -- an alias would be overwritten)
        loss_sum = loss.detach().clone() if loss_sum is None else loss_sum + loss.detach()
        n_it += 1
        flush_log()  # the line of the previous print iteration: its loss arrived while this iteration was being launched
        if i % print_every == 0:
            buf = host_loss  # free again: flush_log() above has consumed the previous line
            buf.copy_(loss_sum.reshape(1), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pending.append((i, n_it, buf, ev))
    flush_log()
