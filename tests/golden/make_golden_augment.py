"""Generate the augmentation fixture by RUNNING THE REFERENCE'S OWN TRANSFORM CLASSES (src/data/datasets.py, composed as in
src/data/loaders.py:43-64) on synthetic images -- only works where /root/reference and cv2 exist.

    python tests/golden/make_golden_augment.py        -> tests/golden/augment.npz

cv2's IPP dispatch is switched off (cv2.ipp.setUseIPP(False)): the fixture pins OpenCV's own 8-bit resize, which
oracle/augment_oracle.py restates; how far the IPP-accelerated wheel is from it is stored alongside
(`ipp_max_abs_diff`, `ipp_diff_fraction`)."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("NASB_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "src"))

from data import datasets as D  # noqa: E402  (the reference's module, unmodified)


class Compose(object):  # torchvision.transforms.Compose without the dependency
    def __init__(self, ts):
        self.ts = ts

    def __call__(self, s):
        for t in self.ts:
            s = t(s)
        return s


NORM = [1.0 / 255, np.array([0.485, 0.456, 0.406]).reshape((1, 1, 3)), np.array([0.229, 0.224, 0.225]).reshape((1, 1, 3))]


def images(rs, sizes):
    out = []
    for h, w in sizes:
        yy, xx = np.mgrid[0:h, 0:w]
        base = np.stack([(xx * 5 + yy * 3) % 256, (xx * yy) % 256, (xx * 2 + 200 - yy) % 256], 2)
        img = ((base + rs.randint(0, 96, (h, w, 3))) % 256).astype(np.uint8)   # structure + noise: hits both clamps
        img[: h // 5, : w // 4] = 255
        img[-h // 6:, -w // 3:] = 0
        mask = (rs.randint(0, 21, (h // 7 + 1, w // 7 + 1)).repeat(7, 0).repeat(7, 1)[:h, :w]).astype(np.uint8)
        mask[rs.rand(h, w) < 0.03] = 255
        out.append((img, mask))
    return out


def run(compose, samples, seed):
    np.random.seed(seed)
    ims, mks = [], []
    for img, mask in samples:
        o = compose({"image": img.copy(), "mask": mask.copy()})
        ims.append(o["image"].float().numpy())
        mks.append(o["mask"].numpy())
    return ims, mks


def main():
    cv2.ipp.setUseIPP(False)
    rs = np.random.RandomState(9314)
    fx = {}
    sizes = [(61, 83), (90, 64), (75, 75), (120, 57), (58, 131), (97, 101)]
    samples = images(rs, sizes)
    for i, (img, mask) in enumerate(samples):
        fx["image%d" % i], fx["mask%d" % i] = img, mask
    # training chains: shorter-side and longer-side rule, two crop sizes (odd crop size -> make_even)
    for name, (side, low, high, longer, crop, seed) in {"trn_a": (64, 0.7, 1.4, False, 48, 1), "trn_b": (70, 0.5, 2.0, False, 57, 2),
                                                         "trn_c": (150, 0.7, 1.4, True, 40, 3)}.items():
        comp = Compose([D.ResizeScale(side, low, high, longer), D.RandomMirror(), D.RandomCrop(crop), D.Normalise(*NORM), D.ToTensor()])
        ims, mks = run(comp, samples, seed)
        fx[name + "_cfg"] = np.array([side, low, high, float(longer), crop, seed], dtype=np.float64)
        for i, (a, b) in enumerate(zip(ims, mks)):
            fx["%s_image%d" % (name, i)], fx["%s_mask%d" % (name, i)] = a, b
    # validation chain
    comp = Compose([D.ResizeScale(80, 1, 1, False), D.CentralCrop(64), D.Normalise(*NORM), D.ToTensor()])
    ims, mks = run(comp, samples, 4)
    fx["val_cfg"] = np.array([80, 1, 1, 0.0, 64, 4], dtype=np.float64)
    for i, (a, b) in enumerate(zip(ims, mks)):
        fx["val_image%d" % i], fx["val_mask%d" % i] = a, b
    # plain resizes (no crop) at a few factors, image and mask: the resize restatement on its own
    for j, s in enumerate([0.613, 1.0, 1.377, 2.05]):
        img, mask = samples[j]
        fx["rs%d_scale" % j] = np.float64(s)
        fx["rs%d_image" % j] = cv2.resize(img, None, fx=s, fy=s, interpolation=cv2.INTER_CUBIC)
        fx["rs%d_mask" % j] = cv2.resize(mask, None, fx=s, fy=s, interpolation=cv2.INTER_NEAREST)
    # distance of the IPP-dispatched wheel from OpenCV's own code on the same inputs
    cv2.ipp.setUseIPP(True)
    mx, nd, nt = 0, 0, 0
    for j, s in enumerate([0.613, 1.0, 1.377, 2.05]):
        d = np.abs(cv2.resize(samples[j][0], None, fx=s, fy=s, interpolation=cv2.INTER_CUBIC).astype(int) - fx["rs%d_image" % j].astype(int))
        mx, nd, nt = max(mx, int(d.max())), nd + int((d > 0).sum()), nt + d.size
    fx["ipp_max_abs_diff"], fx["ipp_diff_fraction"] = np.int64(mx), np.float64(nd / nt)
    fx["cv2_version"] = np.array(cv2.__version__)
    out = os.path.join(HERE, "augment.npz")
    np.savez_compressed(out, **fx)
    print("wrote", out, os.path.getsize(out), "bytes; IPP differs by <=", mx, "in", round(100.0 * nd / nt, 2), "% of the pixels")


if __name__ == "__main__":
    main()
