"""The reference's per-sample transform chains (src/data/datasets.py, composed in src/data/loaders.py:43-64) for a whole batch
in one kernel launch (csrc/augment.cu, `nasb_augment_batch`).

    train : ResizeScale(resize_side, low_scale, high_scale, longer) -> RandomMirror -> RandomCrop(crop_size) -> Normalise -> ToTensor
    val   : ResizeScale(val_resize_side, 1, 1, longer)             -> CentralCrop(val_crop_size)           -> Normalise -> ToTensor

The host side here only does what the reference's `__call__`s do on the host anyway: it draws the random numbers -- from
`np.random`, in the reference's order (uniform scale, mirror bit, crop top, crop left, per sample), so a seeded run sees the
same crops -- and ships the RAW uint8 image and mask (a quarter of the bytes of the float32 tensor the reference's workers
produce).  The arithmetic -- cv2's 8-bit cubic / nearest resize, flip, crop, normalisation -- runs on the device and is
bit-exact with the reference run on OpenCV's own resize code (tests/test_gpu_augment.py against tests/golden/augment.npz)."""
import ctypes as C

import numpy as np
import torch

from .. import lib


def make_even(x):
    """datasets.py:17-21."""
    return x - 1 if x % 2 else x


def resized_size(h, w, scale):
    """cv2.resize(src, None, fx=scale, fy=scale): (round(h*scale), round(w*scale)), halves to even (saturate_cast<int>)."""
    return int(np.rint(h * scale)), int(np.rint(w * scale))


def _resize_scale(h, w, resize_side, scale, longer):
    """ResizeScale.__call__ after its random draw (datasets.py:147-155)."""
    if longer:
        mside = max(h, w)
        if mside * scale > resize_side:
            scale = resize_side * 1.0 / mside
    else:
        mside = min(h, w)
        if mside * scale < resize_side:
            scale = resize_side * 1.0 / mside
    return scale


def _u8(t, device, stream_tensors):
    """uint8 numpy / torch, host or device -> contiguous device tensor (host data goes through pinned memory)."""
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    if t.dtype != torch.uint8:
        raise TypeError("raw images and masks are uint8 (got %s)" % t.dtype)
    if not t.is_cuda:
        t = t.contiguous().pin_memory().to(device, non_blocking=True)
    t = t.contiguous()
    stream_tensors.append(t)
    return t


class _GpuTransform(object):
    def __init__(self, normalise_params, device=None):
        scale, mean, std = normalise_params
        self.scale = float(scale)
        self.mean = (C.c_double * 3)(*[float(v) for v in np.asarray(mean, dtype=np.float64).reshape(-1)])
        self.std = (C.c_double * 3)(*[float(v) for v in np.asarray(std, dtype=np.float64).reshape(-1)])
        self.device = torch.device("cuda") if device is None else torch.device(device)

    def _params(self, h, w):
        raise NotImplementedError

    def __call__(self, samples):
        """samples: sequence of {"image": uint8 [h, w, 3], "mask": uint8 [h, w]} (numpy or torch, any sizes).
        -> {"image": float32 [B, 3, out_h, out_w], "mask": uint8 [B, out_h, out_w]} on the device."""
        if not torch.cuda.is_available():
            raise RuntimeError("nas_segm_b200.data needs the CUDA library (no CPU fallback)")
        n = len(samples)
        table = (lib.NasbAugSample * max(n, 1))()
        keep, out_hw = [], None
        for i, s in enumerate(samples):
            img, msk = _u8(s["image"], self.device, keep), _u8(s["mask"], self.device, keep)
            h, w = int(img.shape[0]), int(img.shape[1])
            if img.dim() != 3 or img.shape[2] != 3 or tuple(msk.shape) != (h, w):
                raise ValueError("sample %d: image must be [h, w, 3] and mask [h, w]" % i)
            scale, mirror, top, left, oh, ow, rh, rw = self._params(h, w)
            if out_hw is None:
                out_hw = (oh, ow)
            elif out_hw != (oh, ow):  # the reference's default collate would fail on ragged crops as well
                raise ValueError("sample %d yields a %dx%d crop, the batch started with %dx%d" % (i, oh, ow, out_hw[0], out_hw[1]))
            e = table[i]
            e.image, e.mask, e.h, e.w, e.rh, e.rw = img.data_ptr(), msk.data_ptr(), h, w, rh, rw
            e.scale, e.top, e.left, e.mirror = scale, top, left, mirror
        oh, ow = out_hw if out_hw is not None else (0, 0)
        image = torch.empty((n, 3, oh, ow), dtype=torch.float32, device=self.device)
        mask = torch.empty((n, oh, ow), dtype=torch.uint8, device=self.device)
        if n and oh and ow:
            with torch.cuda.device(self.device):
                lib.call("nasb_augment_batch", table, n, oh, ow, self.scale, self.mean, self.std, lib.ptr(image), lib.ptr(mask))
                for t in keep:  # raw inputs may be released as soon as the kernel (on the current stream) has read them
                    t.record_stream(torch.cuda.current_stream())
        return {"image": image, "mask": mask}


class GpuTrainTransform(_GpuTransform):
    """loaders.py:43-56 for a batch.  Random draws come from `np.random` (seed it as the reference does), in the reference's
    per-sample order: np.random.uniform(low, high), np.random.randint(2), np.random.randint(0, rh - new_h + 1),
    np.random.randint(0, rw - new_w + 1)."""

    def __init__(self, resize_side, low_scale, high_scale, longer, crop_size, normalise_params, device=None):
        super().__init__(normalise_params, device)
        assert isinstance(resize_side, int) and isinstance(crop_size, int)
        self.resize_side, self.low_scale, self.high_scale, self.longer = resize_side, low_scale, high_scale, bool(longer)
        self.crop_size = make_even(crop_size)

    def _params(self, h, w):
        scale = _resize_scale(h, w, self.resize_side, np.random.uniform(self.low_scale, self.high_scale), self.longer)
        mirror = int(np.random.randint(2))
        rh, rw = resized_size(h, w, scale)
        new_h, new_w = min(rh, self.crop_size), min(rw, self.crop_size)
        top = int(np.random.randint(0, rh - new_h + 1))
        left = int(np.random.randint(0, rw - new_w + 1))
        return float(scale), mirror, top, left, new_h, new_w, rh, rw


class GpuValTransform(_GpuTransform):
    """loaders.py:57-64 for a batch.  `consume_rng` draws the np.random.uniform(1, 1) ResizeScale makes per sample, for runs
    that must leave the global generator where the reference leaves it."""

    def __init__(self, val_resize_side, longer, val_crop_size, normalise_params, device=None, consume_rng=True):
        super().__init__(normalise_params, device)
        self.resize_side, self.longer, self.crop_size = int(val_resize_side), bool(longer), make_even(int(val_crop_size))
        self.consume_rng = consume_rng

    def _params(self, h, w):
        u = np.random.uniform(1, 1) if self.consume_rng else 1.0
        scale = _resize_scale(h, w, self.resize_side, u, self.longer)
        rh, rw = resized_size(h, w, scale)
        top, left = (rh - self.crop_size) // 2, (rw - self.crop_size) // 2
        if top < 0 or left < 0:
            raise ValueError("CentralCrop(%d) of a %dx%d image: the reference's negative margins (python slicing from the end) "
                             "are not reproduced" % (self.crop_size, rh, rw))
        return float(scale), 0, top, left, self.crop_size, self.crop_size, rh, rw


class AugmentedLoader(object):
    """`AugmentedLoader(loader, transform)`: iterates a loader that yields lists of RAW samples (dataset without transform,
    `collate_fn=lambda batch: batch`) and hands the engine device-side batches -- `engine.trainer.train_segmenter`,
    `populate_task0` and `engine.inference.validate` consume it like the reference's own loaders (their
    `sample["image"].float().cuda()` is then a no-op).  Attributes the engine looks up on a loader (`dataset.set_stage`,
    `batch_sampler.batch_size`) are forwarded."""

    def __init__(self, loader, transform):
        self.loader, self.transform = loader, transform

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for raw in self.loader:
            yield self.transform(raw)

    def __getattr__(self, name):
        return getattr(self.loader, name)
