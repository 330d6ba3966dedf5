"""CUDA replacement of the reference's Cython module src/helpers/miou_utils.pyx (same three names, same returns).

``fast_cm(preds, gt, n_classes)``: uint8 label arrays -> int64 confusion matrix, rows = ground truth
(miou_utils.pyx:7-30).  ``compute_iu`` / ``compute_ius_accs``: :32-90 (32-bit unsigned accumulators, default 2).
Inputs may be numpy arrays (host, as the reference passes them: they are copied to the device) or CUDA uint8
tensors (no copy); numpy in -> numpy out, tensor in -> tensor out.  There is no CPU implementation here."""
import numpy as np
import torch

from .. import functional as Fn


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("helpers.miou_utils runs on a CUDA device only (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev_u8(a):
    if isinstance(a, torch.Tensor):
        Fn.lib.require_cuda(a)
        return a.contiguous().view(-1), True
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint8:
        raise ValueError("Buffer dtype mismatch, expected 'unsigned char' but got '{}'".format(a.dtype))
    if a.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got {})".format(a.ndim))
    return torch.from_numpy(a).to(_dev(), non_blocking=False), False


def fast_cm(preds, gt, n_classes):
    p, on_dev = _to_dev_u8(preds)
    g, _ = _to_dev_u8(gt)
    if p.numel() != g.numel():
        raise IndexError("preds and gt differ in length")
    cm = Fn.confmat_labels(p, g, int(n_classes))
    return cm if on_dev else cm.cpu().numpy()


def _cm_dev(cm):
    if isinstance(cm, torch.Tensor):
        Fn.lib.require_cuda(cm)
        return cm.to(torch.int64).contiguous(), True
    cm = np.ascontiguousarray(cm)
    if cm.dtype != np.int64:
        raise ValueError("Buffer dtype mismatch, expected 'int64_t' but got '{}'".format(cm.dtype))
    return torch.from_numpy(cm).to(_dev()), False


def compute_ius_accs(cm):
    d, on_dev = _cm_dev(cm)
    iu, npx, acc = Fn.ius_accs(d)
    if on_dev:
        return iu, npx, acc
    return iu.cpu().numpy(), npx.cpu().numpy(), acc.cpu().numpy()


def compute_iu(cm):
    return compute_ius_accs(cm)[0]
