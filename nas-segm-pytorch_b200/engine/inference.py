"""Validation -> reward with the reference's signature (src/engine/inference.py:18-97).

The reference up-samples the logits on the GPU, copies the full-resolution B x C x H x W fp32 tensor to the host,
takes numpy arg-max and runs the single-threaded Cython histogram (inference.py:58-66).  Here one kernel does the
bilinear up-sampling, arg-max (first maximal index), ``gt < num_classes`` mask and confusion-matrix accumulation on
the device; only the C x C int64 matrix crosses PCIe, once, after the last batch."""
import logging

import numpy as np
import torch

from .. import functional as Fn
from .. import packs
from ..helpers.utils import try_except

logger = logging.getLogger(__name__)


def reward_from_cm(cm, omit_classes=(0,)):
    """inference.py:78-91 on a host int64 confusion matrix -> (reward, miou, macc, mfwiou, ious, accs)."""
    d = torch.from_numpy(np.ascontiguousarray(cm, dtype=np.int64)).cuda()
    iu, npx, acc = Fn.ius_accs(d)
    ious, n_pixels, accs = iu.cpu().numpy(), npx.cpu().numpy(), acc.cpu().numpy()
    present_ind = np.array([idx for idx, v in enumerate(ious) if v <= 1.0])
    present_ind = np.setdiff1d(present_ind, list(omit_classes))
    present_ious, present_pixels, present_accs = ious[present_ind], n_pixels[present_ind], accs[present_ind]
    miou = np.mean(present_ious)
    macc = np.mean(present_accs)
    mfwiou = np.sum(present_ious * present_pixels) / np.sum(present_pixels)
    reward = np.prod([miou, macc, mfwiou]) ** (1.0 / 3)
    return reward, miou, macc, mfwiou, ious, accs


@try_except
def validate(segmenter, val_loader, epoch, epoch2, num_classes=-1, print_every=10, omit_classes=[0]):
    """Returns the reward: geometric mean of mean-IoU, mean-accuracy and frequency-weighted IoU."""
    try:
        val_loader.dataset.set_stage("val")
    except AttributeError:
        try:
            val_loader.dataset.dataset.set_stage("val")
        except AttributeError:
            pass
    segmenter.eval()
    cm = None
    with torch.no_grad(), packs.scope(segmenter):  # weights are constant here: one multi-tensor operand pack
        for i, sample in enumerate(val_loader):
            image, target = sample["image"], sample["mask"]
            output = segmenter(image.float().cuda(non_blocking=True))
            if isinstance(output, tuple):
                output, _ = output
            gt = target.to(torch.uint8).cuda(non_blocking=True)  # inference.py:63: labels wrap to uint8
            if cm is None:
                cm = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=output.device)
            Fn.confmat_logits(output, gt, num_classes, cm)
            if i % print_every == 0:
                ious = Fn.ius_accs(cm)[0].cpu().numpy()
                logger.info(" Val epoch: {} [{}/{}]\tMean IoU: {:.3f}".format(
                    epoch, i, len(val_loader), np.mean([iu for iu in ious if iu <= 1.0])))
    reward, miou, macc, mfwiou, ious, accs = reward_from_cm(cm.cpu().numpy(), omit_classes)
    logger.info(" IoUs: {}, accs: {}".format(ious, accs))
    logger.info(" Val epoch: {}/{}\tMean IoU: {:.3f}\tMean FW-IoU: {:.3f}\tMean Acc: {:.3f}\tReward: {:.3f}".format(
        epoch, epoch2, miou, mfwiou, macc, reward))
    return reward
