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


_SHIM = r'''
#include <cmath>
#include <cstddef>
#include <cstdint>
#define __device__
#define __forceinline__ inline
// the CUDA round-to-nearest intrinsics are plain IEEE operations; -ffp-contract=off keeps g++ from fusing them
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline int __float2int_rn(float a) { return (int)nearbyintf(a); }
#include "nasb200.h"
struct AugTable { NasbAugSample s[1]; };
struct AugParams { int out_h, out_w; double scale, mean[3], stdv[3]; float *out_image; uint8_t *out_mask; };
%(helpers)s
static void pixel(const AugTable &tab, const AugParams &p, int x, int y, int si) {
%(pixel)s
}
extern "C" void run(const NasbAugSample *smp, int out_h, int out_w, double scale, const double *mean, const double *stdv,
                    float *oi, uint8_t *om) {
    AugTable tab;
    tab.s[0] = *smp;
    AugParams p;
    p.out_h = out_h; p.out_w = out_w; p.scale = scale;
    for (int c = 0; c < 3; ++c) { p.mean[c] = mean[c]; p.stdv[c] = stdv[c]; }
    p.out_image = oi; p.out_mask = om;
    for (int y = 0; y < out_h; ++y)
        for (int x = 0; x < out_w; ++x) pixel(tab, p, x, y, 0);
}
'''


def test_kernel_source_arithmetic_on_the_host(tmp_path):
    """The per-pixel arithmetic of csrc/augment.cu -- the very source text between its [host-replica] markers -- compiled for
    the host and run over the reference fixture: what the GPU test checks on the device, checked here without one (launch
    geometry, the by-value table and the device intrinsics themselves remain the GPU test's business)."""
    import ctypes as C
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "nas-segm-pytorch_b200", "csrc", "augment.cu")).read()

    def between(a, b):
        i, j = src.index(a), src.index(b)
        return src[src.index("\n", i) + 1:j]
    helpers = between("// [host-replica: helpers begin]", "// [host-replica: helpers end]")
    helpers = helpers[helpers.index("// OpenCV interpolateCubic"):]  # the marker comment spans two lines
    code = _SHIM % {"helpers": helpers, "pixel": between("// [host-replica: pixel begin]", "// [host-replica: pixel end]")}
    cpp, so = tmp_path / "replica.cpp", tmp_path / "libreplica.so"
    cpp.write_text(code)
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(root, "include"),
                           "-o", str(so), str(cpp)])

    class S(C.Structure):
        _fields_ = [("image", C.c_void_p), ("mask", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("rh", C.c_int32),
                    ("rw", C.c_int32), ("scale", C.c_double), ("top", C.c_int32), ("left", C.c_int32), ("mirror", C.c_int32),
                    ("reserved", C.c_int32)]
    lib = C.CDLL(str(so))
    lib.run.argtypes = [C.POINTER(S), C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p,
                        C.c_void_p]
    mean, std = (C.c_double * 3)(*NORM[1]), (C.c_double * 3)(*NORM[2])

    def run(img, msk, p):
        img, msk = np.ascontiguousarray(img), np.ascontiguousarray(msk)
        rh, rw = A.dst_size(img.shape[0], img.shape[1], p["scale"])
        s = S(img.ctypes.data, msk.ctypes.data, img.shape[0], img.shape[1], rh, rw, p["scale"], p["top"], p["left"], p["mirror"], 0)
        oi = np.empty((3, p["out_h"], p["out_w"]), np.float32)
        om = np.empty((p["out_h"], p["out_w"]), np.uint8)
        lib.run(C.byref(s), p["out_h"], p["out_w"], NORM[0], mean, std, oi.ctypes.data, om.ctypes.data)
        return oi, om
    for name in ("trn_a", "trn_b", "trn_c"):
        side, low, high, longer, crop, seed = FX[name + "_cfg"]
        np.random.seed(int(seed))
        for i, (img, mask) in enumerate(_samples()):
            p = A.draw_train_params(img.shape[0], img.shape[1], int(side), float(low), float(high), bool(longer), int(crop))
            oi, om = run(img, mask, p)
            assert np.array_equal(oi, FX["%s_image%d" % (name, i)]) and np.array_equal(om, FX["%s_mask%d" % (name, i)]), (name, i)
    side, _, _, longer, crop, _ = FX["val_cfg"]
    for i, (img, mask) in enumerate(_samples()):
        oi, om = run(img, mask, A.val_params(img.shape[0], img.shape[1], int(side), bool(longer), int(crop)))
        assert np.array_equal(oi, FX["val_image%d" % i]) and np.array_equal(om, FX["val_mask%d" % i]), i
