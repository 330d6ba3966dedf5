"""Row f3 (augmentation): oracle/augment_oracle.py against the fixture written by the reference's own transform classes
(tests/golden/make_golden_augment.py; src/data/datasets.py, src/data/loaders.py:43-64).  Bit-exact: uint8 resize, uint8
masks, and the float32 the trainer sees after Normalise / ToTensor / .float()."""
import os

import numpy as np

from oracle import augment_oracle as A

FX = np.load(os.path.join(os.path.dirname(__file__), "golden", "augment.npz"))
NORM = (1.0 / 255, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
N_IMG = 6


def _samples():
    return [(FX["image%d" % i], FX["mask%d" % i]) for i in range(N_IMG)]


def test_cv2_resize_restatement_is_bit_exact():
    for j in range(4):
        s = float(FX["rs%d_scale" % j])
        img, mask = FX["image%d" % j], FX["mask%d" % j]
        assert np.array_equal(A.resize_cubic_u8(img, s), FX["rs%d_image" % j]), s
        assert np.array_equal(A.resize_nearest(mask, s), FX["rs%d_mask" % j]), s
    # what the stock (IPP-dispatched) wheel does differently is bounded and recorded, not hidden
    assert int(FX["ipp_max_abs_diff"]) <= 1 and float(FX["ipp_diff_fraction"]) < 0.08


def test_training_chains_follow_the_reference_draw_for_draw():
    for name in ("trn_a", "trn_b", "trn_c"):
        side, low, high, longer, crop, seed = FX[name + "_cfg"]
        np.random.seed(int(seed))
        for i, (img, mask) in enumerate(_samples()):
            p = A.draw_train_params(img.shape[0], img.shape[1], int(side), float(low), float(high), bool(longer), int(crop))
            out_i, out_m = A.apply(img, mask, p, *NORM)
            want_i, want_m = FX["%s_image%d" % (name, i)], FX["%s_mask%d" % (name, i)]
            assert out_i.shape == want_i.shape and out_i.dtype == np.float32, (name, i, out_i.shape, want_i.shape)
            assert np.array_equal(out_i, want_i), (name, i, float(np.abs(out_i - want_i).max()))
            assert np.array_equal(out_m, want_m), (name, i)


def test_validation_chain():
    side, _, _, longer, crop, seed = FX["val_cfg"]
    for i, (img, mask) in enumerate(_samples()):
        p = A.val_params(img.shape[0], img.shape[1], int(side), bool(longer), int(crop))
        out_i, out_m = A.apply(img, mask, p, *NORM)
        assert np.array_equal(out_i, FX["val_image%d" % i]) and np.array_equal(out_m, FX["val_mask%d" % i]), i


def test_edge_cases():
    rs = np.random.RandomState(0)
    img = rs.randint(0, 256, (9, 5, 3)).astype(np.uint8)
    assert np.array_equal(A.resize_cubic_u8(img, 1.0), img)              # identity scale: taps (0, 2048, 0, 0)
    one = np.full((1, 1, 3), 77, np.uint8)
    assert np.array_equal(A.resize_cubic_u8(one, 4.0), np.full((4, 4, 3), 77, np.uint8))  # all taps clamp onto one pixel
    assert A.make_even(57) == 56 and A.make_even(48) == 48
    m = np.arange(12, dtype=np.uint8).reshape(3, 4)
    assert np.array_equal(A.resize_nearest(m, 2.0), m.repeat(2, 0).repeat(2, 1))
