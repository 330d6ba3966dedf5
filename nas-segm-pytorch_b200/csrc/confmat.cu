// confmat.cu -- the CUDA replacement of the reference's only native module, src/helpers/miou_utils.pyx (Cython, CPU,
// single thread), plus the fusion of the validation tail of src/engine/inference.py:58-66.
//
//  * nasb_confmat_labels: fast_cm.  Pure streaming over two uint8 arrays (2 B / pixel): 16-byte vector loads, a
//    per-CTA shared-memory histogram (C*C int32 counters, C <= 64) flushed with one 64-bit atomic per non-zero bin;
//    grid = a multiple of the SM count, grid-stride so every CTA flushes exactly once.
//  * nasb_confmat_logits: bilinear x-up (align_corners=False) + arg-max + (gt < C) mask + histogram in one pass, so
//    the full-resolution logits are never materialised and never cross PCIe (the reference copies B*C*H*W fp32 to the
//    host per batch, inference.py:62).
//  * nasb_ius_accs: compute_iu / compute_ius_accs with the reference's 32-bit unsigned accumulators.
#include "common.cuh"

namespace nasb {

constexpr int CM_SMEM_MAXC = 64;  // 64*64*4 = 16 KiB of counters

__device__ __forceinline__ void cm_count(int *hist, long long *cm, int C, bool use_smem, unsigned g, unsigned p) {
    if (g < (unsigned)C && p < (unsigned)C) {
        if (use_smem)
            atomicAdd(&hist[g * C + p], 1);
        else
            atomicAdd(reinterpret_cast<unsigned long long *>(&cm[(long long)g * C + p]), 1ULL);
    }
}

__global__ void __launch_bounds__(256) confmat_labels_kernel(const uint8_t *pred, const uint8_t *gt, long long n, int C,
                                                             long long *cm) {
    pdl_sync();
    extern __shared__ int hist[];
    const bool use_smem = C <= CM_SMEM_MAXC;
    if (use_smem) {
        for (int i = threadIdx.x; i < C * C; i += blockDim.x) hist[i] = 0;
        __syncthreads();
    }
    // 16-byte body (both arrays must be 16 B aligned for it), scalar tail
    const bool aligned = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(gt)) & 15) == 0;
    const long long nvec = aligned ? n / 16 : 0;
    const uint4 *pv = reinterpret_cast<const uint4 *>(pred), *gv = reinterpret_cast<const uint4 *>(gt);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        uint4 a = pv[i], b = gv[i];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int w = 0; w < 4; ++w)
#pragma unroll
            for (int k = 0; k < 4; ++k) cm_count(hist, cm, C, use_smem, (bw[w] >> (8 * k)) & 0xff, (aw[w] >> (8 * k)) & 0xff);
    }
    for (long long i = nvec * 16 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        cm_count(hist, cm, C, use_smem, gt[i], pred[i]);
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
            int v = hist[i];
            if (v) atomicAdd(reinterpret_cast<unsigned long long *>(&cm[i]), (unsigned long long)v);
        }
    }
}

// fast_cm for C <= 32: one PRIVATE histogram per warp (8 x C*C counters, no contention between warps; the single per-CTA
// histogram spent its time on shared-memory atomic conflicts: 23 % of the HBM roofline on 16 Mpx of uniform random labels),
// two 16-byte vectors of each array in flight per thread, and consecutive equal (gt, pred) pairs -- the common case in real
// label maps -- merged in registers into one atomic.
constexpr int CM_WARP_MAXC = 32;

__global__ void __launch_bounds__(256) confmat_labels_warp_kernel(const uint8_t *pred, const uint8_t *gt, long long nvec, int C,
                                                                  long long *cm) {
    pdl_sync();
    extern __shared__ int hist[];  // [8 warps][C*C]
    const int CC = C * C;
    for (int i = threadIdx.x; i < 8 * CC; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    int *mine = hist + (threadIdx.x >> 5) * CC;
    const uint4 *pv = reinterpret_cast<const uint4 *>(pred), *gv = reinterpret_cast<const uint4 *>(gt);
    const long long stride = (long long)gridDim.x * blockDim.x;
    int run_bin = -1, run_len = 0;
    auto count = [&](unsigned g, unsigned p) {
        const int bin = (g < (unsigned)C && p < (unsigned)C) ? (int)(g * C + p) : -1;
        if (bin == run_bin) {
            ++run_len;
        } else {
            if (run_bin >= 0) atomicAdd(&mine[run_bin], run_len);
            run_bin = bin;
            run_len = 1;
        }
    };
    auto vec = [&](const uint4 &a, const uint4 &b) {
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int w = 0; w < 4; ++w)
#pragma unroll
            for (int k = 0; k < 4; ++k) count((bw[w] >> (8 * k)) & 0xff, (aw[w] >> (8 * k)) & 0xff);
    };
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {  // eight 16-byte loads in flight per thread
        const uint4 a0 = __ldg(pv + i), b0 = __ldg(gv + i), a1 = __ldg(pv + i + stride), b1 = __ldg(gv + i + stride);
        const uint4 a2 = __ldg(pv + i + 2 * stride), b2 = __ldg(gv + i + 2 * stride);
        const uint4 a3 = __ldg(pv + i + 3 * stride), b3 = __ldg(gv + i + 3 * stride);
        vec(a0, b0);
        vec(a1, b1);
        vec(a2, b2);
        vec(a3, b3);
    }
    for (; i < nvec; i += stride) vec(__ldg(pv + i), __ldg(gv + i));
    if (run_bin >= 0) atomicAdd(&mine[run_bin], run_len);
    __syncthreads();
    for (int b = threadIdx.x; b < CC; b += blockDim.x) {
        int v = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += hist[w * CC + b];
        if (v) atomicAdd(reinterpret_cast<unsigned long long *>(&cm[b]), (unsigned long long)v);
    }
}

// one thread = one full-resolution pixel
template <typename T>
__global__ void __launch_bounds__(256) confmat_logits_kernel(const T *x, int cs, int N, int h, int w, int C, const uint8_t *gt,
                                                             int H, int W, int n_classes, long long *cm, float rh, float rw) {
    pdl_sync();
    extern __shared__ int hist[];
    const bool use_smem = n_classes <= CM_SMEM_MAXC;
    if (use_smem) {
        for (int i = threadIdx.x; i < n_classes * n_classes; i += blockDim.x) hist[i] = 0;
        __syncthreads();
    }
    const long long total = (long long)N * H * W;
    const bool identity = (h == H && w == W);
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        unsigned g = gt[p];
        if (g >= (unsigned)n_classes) continue;
        int X = (int)(p % W);
        long long t = p / W;
        int Y = (int)(t % H);
        int n = (int)(t / H);
        Lerp ly, lx;
        if (identity) {
            ly.i0 = ly.i1 = Y; ly.l0 = 1.f; ly.l1 = 0.f;
            lx.i0 = lx.i1 = X; lx.l0 = 1.f; lx.l1 = 0.f;
        } else {
            ly = lerp_coord(Y, rh, h);
            lx = lerp_coord(X, rw, w);
        }
        const T *base = x + (long long)n * h * w * cs;
        const T *p00 = base + ((long long)ly.i0 * w + lx.i0) * cs, *p01 = base + ((long long)ly.i0 * w + lx.i1) * cs;
        const T *p10 = base + ((long long)ly.i1 * w + lx.i0) * cs, *p11 = base + ((long long)ly.i1 * w + lx.i1) * cs;
        float best = 0.f;
        int arg = 0;
        for (int c = 0; c < C; ++c) {
            // same expression as ATen's upsample_bilinear2d
            float v = ly.l0 * (lx.l0 * to_f(p00[c]) + lx.l1 * to_f(p01[c])) + ly.l1 * (lx.l0 * to_f(p10[c]) + lx.l1 * to_f(p11[c]));
            if (c == 0 || v > best) {  // numpy argmax: first maximal index
                best = v;
                arg = c;
            }
        }
        // inference.py:62 casts the arg-max to uint8 before counting
        cm_count(hist, cm, n_classes, use_smem, g, (unsigned)(arg & 0xff));
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_classes * n_classes; i += blockDim.x) {
            int v = hist[i];
            if (v) atomicAdd(reinterpret_cast<unsigned long long *>(&cm[i]), (unsigned long long)v);
        }
    }
}

__global__ void ius_accs_kernel(const long long *cm, int C, double *iu, long long *npx, double *accs) {
    pdl_sync();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C) return;
    unsigned int pi = 0, gi = 0;
    for (int j = 0; j < C; ++j) {
        pi += (unsigned int)cm[(long long)j * C + i];
        gi += (unsigned int)cm[(long long)i * C + j];
    }
    unsigned int ii = (unsigned int)cm[(long long)i * C + i];
    unsigned int denom = pi + gi - ii;
    double u = 2.0, a = 2.0;
    if (denom > 0) u = (double)ii / (double)denom;
    if (gi > 0) a = (double)ii / (double)gi;
    if (iu) iu[i] = u;
    if (accs) accs[i] = a;
    if (npx) npx[i] = (long long)gi;
}

}  // namespace nasb

using namespace nasb;
#define ST ((cudaStream_t)stream)

extern "C" int nasb_confmat_labels(const uint8_t *pred, const uint8_t *gt, long long n, int n_classes, long long *cm,
                                   void *stream) {
    if (!cm || n_classes <= 0 || n_classes > 256 || n < 0 || (n > 0 && (!pred || !gt))) return NASB_ERR_BAD_ARG;
    if (n == 0) return 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(gt)) & 15) == 0;
    if (aligned && n_classes <= CM_WARP_MAXC && n >= 16) {  // 16-byte body on per-warp histograms, scalar tail below
        const long long nvec = n / 16;
        // 3 CTAs per SM: every CTA ends with up to C*C 64-bit global atomics onto the SAME C*C addresses, and with 8 CTAs per
        // SM those 427 k serialised L2 atomics (C = 19), not the streaming, set the time (B200: 0.022 ms for 33 MB)
        long long blocks = (nvec + 1023) / 1024, cap = (long long)NASB_SM_COUNT * 3;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        nasb::launch_pdl((confmat_labels_warp_kernel), dim3((int)blocks), dim3(256), (size_t)8 * n_classes * n_classes * sizeof(int), (cudaStream_t)(ST), pred, gt, nvec, n_classes, cm);
        NASB_CHECK_LAUNCH();
        pred += nvec * 16, gt += nvec * 16, n -= nvec * 16;
        if (n == 0) return 0;
    }
    long long vecs = (n + 15) / 16;
    long long blocks = (vecs + 255) / 256;
    long long cap = (long long)NASB_SM_COUNT * 8;
    if (blocks > cap) blocks = cap;
    size_t smem = n_classes <= CM_SMEM_MAXC ? (size_t)n_classes * n_classes * sizeof(int) : 0;
    nasb::launch_pdl((confmat_labels_kernel), dim3((int)blocks), dim3(256), smem, (cudaStream_t)(ST), pred, gt, n, n_classes, cm);
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_confmat_logits(const NasbTensor *logits, const uint8_t *gt, int H, int W, int n_classes, long long *cm,
                                   void *stream) {
    if (!logits || !gt || !cm || n_classes <= 0 || n_classes > 256 || H <= 0 || W <= 0) return NASB_ERR_BAD_ARG;
    if (logits->dtype != NASB_F32 && logits->dtype != NASB_BF16) return NASB_ERR_BAD_ARG;
    long long total = (long long)logits->n * H * W;
    if (total == 0) return 0;
    long long blocks = (total + 255) / 256, cap = (long long)NASB_SM_COUNT * 8;
    if (blocks > cap) blocks = cap;
    size_t smem = n_classes <= CM_SMEM_MAXC ? (size_t)n_classes * n_classes * sizeof(int) : 0;
    float rh = (float)logits->h / (float)H, rw = (float)logits->w / (float)W;
    if (logits->dtype == NASB_BF16)
        nasb::launch_pdl((confmat_logits_kernel<bf16>), dim3((int)blocks), dim3(256), smem, (cudaStream_t)(ST), (const bf16 *)logits->ptr, logits->cstride, logits->n, logits->h,
                                                                    logits->w, logits->c, gt, H, W, n_classes, cm, rh, rw);
    else
        nasb::launch_pdl((confmat_logits_kernel<float>), dim3((int)blocks), dim3(256), smem, (cudaStream_t)(ST), (const float *)logits->ptr, logits->cstride, logits->n,
                                                                     logits->h, logits->w, logits->c, gt, H, W, n_classes, cm, rh,
                                                                     rw);
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_ius_accs(const long long *cm, int n_classes, double *iu, long long *n_pixels, double *accs, void *stream) {
    if (!cm || n_classes <= 0) return NASB_ERR_BAD_ARG;
    nasb::launch_pdl((ius_accs_kernel), dim3(cdiv(n_classes, 64)), dim3(64), 0, (cudaStream_t)(ST), cm, n_classes, iu, n_pixels, accs);
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" const char *nasb_version(void) { return "nasb200 0.1 sm_100a"; }
