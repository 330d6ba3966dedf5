"""Augmentation oracle (TEST INFRASTRUCTURE ONLY): numpy restatement of the reference's per-sample transform chains
(src/data/datasets.py; composed in src/data/loaders.py:43-64)

    train : ResizeScale(resize_side, low, high, longer) -> RandomMirror -> RandomCrop(crop) -> Normalise -> ToTensor
    val   : ResizeScale(val_resize_side, 1, 1, longer)  -> CentralCrop(val_crop)            -> Normalise -> ToTensor

The resize is cv2.resize(image, None, fx=s, fy=s, INTER_CUBIC) for the uint8 image and INTER_NEAREST for the mask
(datasets.py:156-165).  cv2 is a third-party dependency of the reference (opencv-python, unpinned in requirements.txt; 4.13.0
in this image); its published algorithm for 8-bit images (modules/imgproc/src/resize.cpp) is restated here:

  * destination size  = (round(w*s), round(h*s))  [saturate_cast<int>], source step = 1/s (NOT w/dst_w) in double;
  * cubic: source coordinate f = float((d + 0.5)/s - 0.5), tap base floor(f) - 1, four taps with Keys coefficients A = -0.75
    evaluated in float32 exactly as interpolateCubic() does, converted to 11-bit fixed point by saturate_cast<short>(c * 2048)
    (no renormalisation: the four integers need not sum to 2048), tap indices clamped to the image (replicate);
    horizontal pass in int32; vertical pass: float32 vector form for the first floor(n/8)*8 elements of a row, int32 scalar
    form ((v + 2^21) >> 22) for the tail -- see resize_cubic_u8;
  * nearest: source index = min(floor(d / s), size - 1);
  * a destination of the source's size is a plain copy, whatever the factor (cv::resize's early exit).

PINNED: bit-exact against cv2.resize with cv2.ipp.setUseIPP(False) (tests/golden/make_golden_augment.py runs the reference's
own transform classes that way and stores inputs, drawn parameters and outputs in tests/golden/augment.npz).  The stock
opencv-python wheel dispatches 8-bit cubic resizing to Intel IPP, whose result differs from OpenCV's own code by one grey level
in ~4 % of the pixels (never more) -- recorded in the fixture as `ipp_max_abs_diff` / `ipp_diff_fraction`.

Random draws follow the reference's order per sample: np.random.uniform(low, high) [ResizeScale], np.random.randint(2)
[RandomMirror], np.random.randint(0, h - new_h + 1), np.random.randint(0, w - new_w + 1) [RandomCrop]."""
import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS
SIMD_LANES = 8  # v_uint16 lanes of the universal intrinsics at the wheel's baseline (SSE3, 128-bit)


def _round_half_even(x):
    return int(np.rint(x))


def dst_size(h, w, s):
    """cv2.resize(..., None, fx=s, fy=s): Size(saturate_cast<int>(w*s), saturate_cast<int>(h*s))."""
    return _round_half_even(h * s), _round_half_even(w * s)


def cubic_table(ssize, dsize, s):
    """-> (tap base index [dsize] (first of the four taps, unclamped), int16 coefficients [dsize, 4])."""
    f32 = np.float32
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * (1.0 / s) - 0.5).astype(f32)
    base = np.floor(f).astype(np.int64)
    x = (f - base.astype(f32)).astype(f32)
    A, one = f32(-0.75), f32(1)
    xp1 = (x + one).astype(f32)
    omx = (one - x).astype(f32)
    c0 = (((A * xp1 - f32(5) * A) * xp1 + f32(8) * A) * xp1 - f32(4) * A).astype(f32)
    c1 = (((A + f32(2)) * x - (A + f32(3))) * x * x + one).astype(f32)
    c2 = (((A + f32(2)) * omx - (A + f32(3))) * omx * omx + one).astype(f32)
    c3 = (one - c0 - c1 - c2).astype(f32)
    c = np.stack([c0, c1, c2, c3], 1)
    ic = np.clip(np.rint((c * f32(COEF_SCALE)).astype(f32)), -32768, 32767).astype(np.int64)
    return base - 1, ic


def resize_cubic_u8(img, s):
    """uint8 [h, w, c] -> uint8 [round(h*s), round(w*s), c]."""
    h, w = img.shape[:2]
    dh, dw = dst_size(h, w, s)
    if (dh, dw) == (h, w):  # cv::resize: "Source and destination are of same size. Use simple copy." -- whatever fx says
        return img.copy()
    xb, xa = cubic_table(w, dw, s)
    yb, ya = cubic_table(h, dh, s)
    src = img.astype(np.int64)
    xi = np.clip(xb[:, None] + np.arange(4)[None, :], 0, w - 1)          # [dw, 4]
    hor = (src[:, xi, :] * xa[None, :, :, None]).sum(2)                    # [h, dw, c]
    yi = np.clip(yb[:, None] + np.arange(4)[None, :], 0, h - 1)          # [dh, 4]
    rows = hor[yi, :, :]                                                   # [dh, 4, dw, c]
    # vertical pass, scalar form (VResizeCubic<uchar, int, short, FixedPtCast<int, uchar, 22>>)
    ver = (rows * ya[:, :, None, None]).sum(1)
    out_int = np.clip((ver + (1 << (2 * COEF_BITS - 1))) >> (2 * COEF_BITS), 0, 255)
    # vertical pass, vector form (VResizeCubicVec_32s8u, compiled for the wheel's SSE3 baseline: 8 elements per step, mulps and
    # addps -- NOT fused): float32 b_k = beta_k * (1.f / (2048 * 2048)), r = S0*b0 + (S1*b1 + (S2*b2 + S3*b3)), each product and
    # sum rounded to float32, then cvtps2dq (round half to even) and saturating packs.  It covers the first floor(n / 8) * 8
    # elements of a row of n = dw * channels; the scalar form does the tail.  The two forms differ when the exact value lies
    # within float32 rounding of a half (about one pixel in 10^4).
    f32 = np.float32
    b = (ya.astype(f32) * (f32(1.0) / f32(COEF_SCALE * COEF_SCALE))).astype(f32)      # [dh, 4]
    rf = rows.astype(f32)
    acc = (rf[:, 3] * b[:, 3, None, None]).astype(f32)
    for k in (2, 1, 0):
        acc = ((rf[:, k] * b[:, k, None, None]).astype(f32) + acc).astype(f32)
    out_vec = np.clip(np.rint(acc), 0, 255).astype(np.int64)
    n = dw * img.shape[2]
    nvec = (n // SIMD_LANES) * SIMD_LANES
    out = out_vec.reshape(dh, n).copy()
    out[:, nvec:] = out_int.reshape(dh, n)[:, nvec:]
    return out.reshape(dh, dw, img.shape[2]).astype(np.uint8)


def nearest_index(ssize, dsize, s):
    d = np.arange(dsize, dtype=np.float64)
    return np.minimum(np.floor(d * (1.0 / s)).astype(np.int64), ssize - 1)


def resize_nearest(mask, s):
    h, w = mask.shape[:2]
    dh, dw = dst_size(h, w, s)
    if (dh, dw) == (h, w):
        return mask.copy()
    return mask[nearest_index(h, dh, s)[:, None], nearest_index(w, dw, s)[None, :]]


def resize_scale_factor(h, w, resize_side, scale, longer):
    """ResizeScale.__call__ after its np.random.uniform draw (datasets.py:147-155)."""
    if longer:
        mside = max(h, w)
        if mside * scale > resize_side:
            scale = resize_side * 1.0 / mside
    else:
        mside = min(h, w)
        if mside * scale < resize_side:
            scale = resize_side * 1.0 / mside
    return scale


def make_even(x):
    return x - 1 if x % 2 else x


def normalise(image_u8, scale, mean, std):
    """Normalise + ToTensor + the trainer's .float(): float64 arithmetic (numpy promotes uint8 * python float), CHW, then
    float32 (datasets.py:205-208,216-220; engine/trainer.py:210)."""
    out = (scale * image_u8 - np.asarray(mean, dtype=np.float64)) / np.asarray(std, dtype=np.float64)
    return out.transpose(2, 0, 1).astype(np.float32)


def draw_train_params(h, w, resize_side, low, high, longer, crop_size, rng=np.random):
    """The reference's random draws for one sample, in its order.  -> dict(scale, mirror, top, left, out_h, out_w)."""
    scale = resize_scale_factor(h, w, resize_side, rng.uniform(low, high), longer)
    mirror = int(rng.randint(2))
    rh, rw = dst_size(h, w, scale)
    crop = make_even(crop_size)
    new_h, new_w = min(rh, crop), min(rw, crop)
    top = int(rng.randint(0, rh - new_h + 1))
    left = int(rng.randint(0, rw - new_w + 1))
    return {"scale": float(scale), "mirror": mirror, "top": top, "left": left, "out_h": new_h, "out_w": new_w}


def val_params(h, w, val_resize_side, longer, val_crop):
    """ResizeScale(side, 1, 1, longer) (its uniform(1, 1) draw included by the caller if the RNG stream matters) + CentralCrop."""
    scale = resize_scale_factor(h, w, val_resize_side, 1.0, longer)
    rh, rw = dst_size(h, w, scale)
    crop = make_even(val_crop)
    top, left = (rh - crop) // 2, (rw - crop) // 2
    return {"scale": float(scale), "mirror": 0, "top": top, "left": left, "out_h": crop, "out_w": crop}


def apply(image_u8, mask_u8, p, scale, mean, std):
    """Resize -> mirror -> crop -> normalise for one sample with drawn parameters p.  -> (float32 [3, oh, ow], uint8 [oh, ow])."""
    img = resize_cubic_u8(image_u8, p["scale"])
    msk = resize_nearest(mask_u8, p["scale"])
    if p["mirror"]:
        img, msk = img[:, ::-1], msk[:, ::-1]
    t, l, oh, ow = p["top"], p["left"], p["out_h"], p["out_w"]
    img, msk = img[t:t + oh, l:l + ow], msk[t:t + oh, l:l + ow]
    return normalise(img, scale, mean, std), np.ascontiguousarray(msk)
