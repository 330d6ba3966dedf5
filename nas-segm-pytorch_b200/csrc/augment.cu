// augment.cu -- the reference's per-sample training / validation transform chain as ONE kernel over a batch of raw images
// (src/data/datasets.py: ResizeScale -> RandomMirror -> RandomCrop | CentralCrop -> Normalise -> ToTensor; composed in
// src/data/loaders.py:43-64).  The host draws the random parameters exactly as the reference does (nn side:
// nas_segm_b200/data/augment.py); the device produces, for every output pixel of every sample, the value the reference's
// cv2.resize / slicing / numpy arithmetic would have produced -- bit for bit:
//
//   * cv2.resize(image, None, fx=s, fy=s, INTER_CUBIC) on uint8: Keys cubic (A = -0.75) coefficients evaluated in float32 in
//     OpenCV's operation order, 11-bit fixed point, int32 horizontal pass, vertical pass in OpenCV's float32 vector form for
//     the first floor(n/8)*8 elements of a destination row (n = width * 3) and in its int32 scalar form for the tail
//     (oracle/augment_oracle.py documents and pins both), replicate border, destination == source size is a plain copy;
//   * cv2.resize(mask, ..., INTER_NEAREST): index min(floor(d / s), size - 1);
//   * mirror and crop are index arithmetic; Normalise is (scale * v - mean) / std in float64, rounded once to float32.
//
// Every floating-point step uses the round-to-nearest intrinsics (__fmul_rn, __dadd_rn, ...): the compiler must not contract
// a multiply and an add into an FMA, OpenCV's SSE code and numpy do not.  Byte work on 12 M output values per batch of
// 32 x 350 x 350: the kernel is latency-, not bandwidth-bound (48 L1-resident byte loads per pixel); the point of having it
// is that the host uploads raw uint8 images (a quarter of the float32 bytes) and no cv2 worker pool has to keep up with a
// GPU that trains 4000 such crops per second.
#include "common.cuh"

namespace nasb {

constexpr int AUG_MAX = 64;  // samples per launch (the table travels in the kernel parameters)

struct AugTable {
    NasbAugSample s[AUG_MAX];
};

struct AugParams {
    int out_h, out_w;
    double scale, mean[3], stdv[3];
    float *out_image;      // [n][3][out_h][out_w]
    uint8_t *out_mask;     // [n][out_h][out_w]
};

// [host-replica: helpers begin]  (tests/test_augment_oracle.py compiles the text between these markers with g++, the
// round-to-nearest intrinsics mapped to plain IEEE operations under -ffp-contract=off, and checks it against the fixture)
// OpenCV interpolateCubic() + saturate_cast<short>(c * 2048), float32, no contraction
__device__ __forceinline__ void cubic_coeffs(float x, int (&ic)[4]) {
    const float A = -0.75f;
    const float xp1 = __fadd_rn(x, 1.f), omx = __fsub_rn(1.f, x);
    float c[4];
    // ((A*(x+1) - 5A)*(x+1) + 8A)*(x+1) - 4A
    c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, xp1), 5.f * A), xp1), 8.f * A), xp1), 4.f * A);
    // ((A+2)*x - (A+3))*x*x + 1
    c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, x), A + 3.f), x), x), 1.f);
    c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, omx), A + 3.f), omx), omx), 1.f);
    c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int v = __float2int_rn(__fmul_rn(c[k], 2048.f));
        ic[k] = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
    }
}

// destination index d -> (first tap index, four fixed-point coefficients)
__device__ __forceinline__ int cubic_taps(int d, double inv, int (&ic)[4]) {
    const float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, inv), -0.5);
    const float fl = floorf(f);
    cubic_coeffs(__fsub_rn(f, fl), ic);
    return (int)fl - 1;
}

// [host-replica: helpers end]

__global__ void __launch_bounds__(256) augment_kernel(const __grid_constant__ AugTable tab, const AugParams p) {
    pdl_sync();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, si = blockIdx.z;
    if (x >= p.out_w) return;
    // [host-replica: pixel begin]
    const NasbAugSample &s = tab.s[si];
    const int ry = s.top + y;
    int rx = s.left + x;
    if (s.mirror) rx = s.rw - 1 - rx;
    int pix[3], m;
    if (s.rh == s.h && s.rw == s.w) {  // cv::resize copies when the destination has the source's size
        const uint8_t *q = s.image + ((size_t)ry * s.w + rx) * 3;
        pix[0] = q[0];
        pix[1] = q[1];
        pix[2] = q[2];
        m = s.mask[(size_t)ry * s.w + rx];
    } else {
        const double inv = __ddiv_rn(1.0, s.scale);
        int ax[4], ay[4];
        const int bx = cubic_taps(rx, inv, ax), by = cubic_taps(ry, inv, ay);
        int ox[4], oy[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ix = bx + k, iy = by + k;
            ox[k] = (ix < 0 ? 0 : (ix > s.w - 1 ? s.w - 1 : ix)) * 3;
            oy[k] = iy < 0 ? 0 : (iy > s.h - 1 ? s.h - 1 : iy);
        }
        int hor[4][3];
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            const uint8_t *row = s.image + (size_t)oy[ky] * s.w * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                int acc = 0;
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) acc += (int)row[ox[kx] + c] * ax[kx];
                hor[ky][c] = acc;
            }
        }
        const int nvec = (s.rw * 3) & ~7;  // elements of a destination row done by OpenCV's 8-lane vector loop
        const float vs = 1.f / (2048.f * 2048.f);
        float bf[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) bf[k] = __fmul_rn((float)ay[k], vs);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int v;
            if (rx * 3 + c < nvec) {  // VResizeCubicVec_32s8u: S0*b0 + (S1*b1 + (S2*b2 + S3*b3)), mulps / addps, cvtps2dq
                float acc = __fmul_rn((float)hor[3][c], bf[3]);
                acc = __fadd_rn(__fmul_rn((float)hor[2][c], bf[2]), acc);
                acc = __fadd_rn(__fmul_rn((float)hor[1][c], bf[1]), acc);
                acc = __fadd_rn(__fmul_rn((float)hor[0][c], bf[0]), acc);
                v = __float2int_rn(acc);
            } else {                   // VResizeCubic scalar: FixedPtCast<int, uchar, 22>
                const int acc = hor[0][c] * ay[0] + hor[1][c] * ay[1] + hor[2][c] * ay[2] + hor[3][c] * ay[3];
                v = (acc + (1 << 21)) >> 22;
            }
            pix[c] = v < 0 ? 0 : (v > 255 ? 255 : v);
        }
        int nx = (int)floor(__dmul_rn((double)rx, inv)), ny = (int)floor(__dmul_rn((double)ry, inv));
        nx = nx < s.w - 1 ? nx : s.w - 1;
        ny = ny < s.h - 1 ? ny : s.h - 1;
        m = s.mask[(size_t)ny * s.w + nx];
    }
    const size_t plane = (size_t)p.out_h * p.out_w, o = (size_t)y * p.out_w + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double v = __ddiv_rn(__dsub_rn(__dmul_rn(p.scale, (double)pix[c]), p.mean[c]), p.stdv[c]);
        p.out_image[((size_t)si * 3 + c) * plane + o] = (float)v;
    }
    p.out_mask[(size_t)si * plane + o] = (uint8_t)m;
    // [host-replica: pixel end]
}

}  // namespace nasb

using namespace nasb;

extern "C" int nasb_augment_batch(const NasbAugSample *samples, int n, int out_h, int out_w, double scale, const double *mean,
                                  const double *stdv, float *out_image, uint8_t *out_mask, void *stream) {
    if (!samples || !mean || !stdv || !out_image || !out_mask || n < 0 || out_h <= 0 || out_w <= 0) return NASB_ERR_BAD_ARG;
    for (int i = 0; i < n; ++i) {
        const NasbAugSample &s = samples[i];
        if (!s.image || !s.mask || s.h <= 0 || s.w <= 0 || s.rh <= 0 || s.rw <= 0 || !(s.scale > 0.0)) return NASB_ERR_BAD_ARG;
        // the crop window must lie inside the resized image (the reference's negative CentralCrop margins are not reproduced)
        if (s.top < 0 || s.left < 0 || s.top + out_h > s.rh || s.left + out_w > s.rw) return NASB_ERR_UNSUPPORTED;
    }
    AugParams p{};
    p.out_h = out_h;
    p.out_w = out_w;
    p.scale = scale;
    for (int c = 0; c < 3; ++c) {
        p.mean[c] = mean[c];
        p.stdv[c] = stdv[c];
    }
    const size_t plane = (size_t)out_h * out_w;
    for (int i0 = 0; i0 < n; i0 += AUG_MAX) {
        const int cnt = n - i0 < AUG_MAX ? n - i0 : AUG_MAX;
        AugTable tab;
        for (int i = 0; i < cnt; ++i) tab.s[i] = samples[i0 + i];
        for (int i = cnt; i < AUG_MAX; ++i) tab.s[i] = samples[i0];
        p.out_image = out_image + (size_t)i0 * 3 * plane;
        p.out_mask = out_mask + (size_t)i0 * plane;
        nasb::launch_pdl((augment_kernel), dim3(cdiv(out_w, 128), out_h, cnt), dim3(128), 0, (cudaStream_t)stream, tab, p);
        NASB_CHECK_LAUNCH();
    }
    return 0;
}
