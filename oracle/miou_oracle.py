"""Metric oracle (TEST INFRASTRUCTURE ONLY): numpy + plain-C restatement of
src/helpers/miou_utils.pyx and of the reward tail of src/engine/inference.py:62-91."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmiou_oracle.so")
_lib = None


def build(force=False):
    """gcc the C restatement into oracle/_build/ (git-ignored, travels with gpurun)."""
    src = os.path.join(_HERE, "miou_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, src])
    return _SO


def _c():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_fast_cm.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                        ctypes.c_void_p]
        _lib.oracle_compute_ius_accs.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 3
    return _lib


def fast_cm_c(preds, gt, n_classes):
    """miou_utils.pyx:7-30 through the C restatement."""
    preds = np.ascontiguousarray(preds, dtype=np.uint8)
    gt = np.ascontiguousarray(gt, dtype=np.uint8)
    assert preds.ndim == 1 and preds.shape == gt.shape
    cm = np.zeros((n_classes, n_classes), dtype=np.int64)
    _c().oracle_fast_cm(preds.ctypes.data, gt.ctypes.data, preds.shape[0], n_classes, cm.ctypes.data)
    return cm


def fast_cm_np(preds, gt, n_classes):
    """Same, as one numpy bincount (rows = gt, cols = prediction)."""
    idx = gt.astype(np.int64) * n_classes + preds.astype(np.int64)
    return np.bincount(idx, minlength=n_classes * n_classes).reshape(n_classes, n_classes).astype(np.int64)


def compute_ius_accs_c(cm):
    cm = np.ascontiguousarray(cm, dtype=np.int64)
    c = cm.shape[0]
    iu, acc = np.empty(c, np.float64), np.empty(c, np.float64)
    npx = np.empty(c, np.int64)
    _c().oracle_compute_ius_accs(cm.ctypes.data, c, iu.ctypes.data, npx.ctypes.data, acc.ctypes.data)
    return iu, npx, acc


def compute_ius_accs_np(cm):
    """miou_utils.pyx:59-90 (uint32 wrap included)."""
    m = cm.astype(np.uint64)
    pi = (m.sum(0) & 0xFFFFFFFF)
    gi = (m.sum(1) & 0xFFFFFFFF)
    ii = (np.diag(m) & 0xFFFFFFFF)
    denom = (pi + gi - ii) & 0xFFFFFFFF
    iu = np.where(denom > 0, ii / np.maximum(denom, 1), 2.0)
    acc = np.where(gi > 0, ii / np.maximum(gi, 1), 2.0)
    return iu.astype(np.float64), gi.astype(np.int64), acc.astype(np.float64)


def labels_from_logits(logits_up):
    """inference.py:62 -- numpy argmax over the class axis (first maximal index), cast to uint8."""
    return np.asarray(logits_up).argmax(axis=1).astype(np.uint8)


def cm_from_logits(logits_up, target, num_classes):
    """inference.py:62-66: argmax -> uint8, gt uint8, keep gt < num_classes, fast_cm."""
    pred = labels_from_logits(logits_up)
    gt = np.asarray(target).astype(np.uint8)
    keep = gt < num_classes
    return fast_cm_c(pred[keep], gt[keep], num_classes)


def reward_from_cm(cm, omit_classes=(0,)):
    """inference.py:78-91.  Returns (reward, miou, macc, mfwiou)."""
    ious, n_pixels, accs = compute_ius_accs_c(cm)
    present = np.array([i for i, iu in enumerate(ious) if iu <= 1.0])
    present = np.setdiff1d(present, list(omit_classes))
    p_iou, p_px, p_acc = ious[present], n_pixels[present], accs[present]
    miou, macc = np.mean(p_iou), np.mean(p_acc)
    mfwiou = np.sum(p_iou * p_px) / np.sum(p_px)
    reward = np.prod([miou, macc, mfwiou]) ** (1.0 / 3)
    return float(reward), float(miou), float(macc), float(mfwiou)
