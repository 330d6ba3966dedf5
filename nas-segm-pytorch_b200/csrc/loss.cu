// loss.cu -- per-pixel losses of the training engines (reference: src/engine/trainer.py:144-158,
// src/main_search.py:435,458): LogSoftmax+NLLLoss2d(ignore_index) fused into one pass over the logits, MSE (knowledge
// distillation term) and the reverse-Huber depth loss.  One thread = one pixel for the class-axis reductions; the
// scalar results are reduced in fp64 (block reduce + one fp64 atomic per CTA) and finalised by a one-thread kernel.
#include "common.cuh"

namespace nasb {

template <typename T>
__global__ void __launch_bounds__(256) ce_fwd_kernel(const T *x, int cs, int C, const int64_t *target, int ignore, long long P,
                                                     double *acc2) {
    pdl_sync();
    __shared__ double sm[256];
    double loss = 0.0, cnt = 0.0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        long long t = target[p];
        if (t == ignore) continue;
        if (t < 0 || t >= C) continue;  // out-of-range labels are a device assert in torch; skipped here
        const T *row = x + p * cs;
        float m = -INFINITY;
        for (int c = 0; c < C; ++c) m = fmaxf(m, to_f(row[c]));
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += expf(to_f(row[c]) - m);
        loss += (double)(m + logf(s) - to_f(row[t]));
        cnt += 1.0;
    }
    double l = block_sum<256>(loss, sm);
    double n = block_sum<256>(cnt, sm);
    if (threadIdx.x == 0) {
        atomicAdd(&acc2[0], l);
        atomicAdd(&acc2[1], n);
    }
}

__global__ void ce_finalize_kernel(const double *acc2, float *out2) {
    pdl_sync();
    out2[0] = (float)(acc2[0] / acc2[1]);  // 0/0 -> nan like torch when every pixel is ignored
    out2[1] = (float)acc2[1];
}

template <typename T, typename TG>
__global__ void __launch_bounds__(256) ce_bwd_kernel(const T *x, int cs, int C, const int64_t *target, int ignore, long long P,
                                                     const float *out2, const float *gscale, TG *dx, int dcs) {
    pdl_sync();
    const float k = (gscale ? gscale[0] : 1.f) / out2[1];
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        long long t = target[p];
        TG *drow = dx + p * dcs;
        if (t == ignore || t < 0 || t >= C) {
            for (int c = 0; c < C; ++c) drow[c] = from_f<TG>(0.f);
            continue;
        }
        const T *row = x + p * cs;
        float m = -INFINITY;
        for (int c = 0; c < C; ++c) m = fmaxf(m, to_f(row[c]));
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += expf(to_f(row[c]) - m);
        float inv = 1.f / s;
        for (int c = 0; c < C; ++c) {
            float sm = expf(to_f(row[c]) - m) * inv;
            drow[c] = from_f<TG>(k * (sm - (c == t ? 1.f : 0.f)));
        }
    }
}

// generic strided element walk over an NHWC tensor: element e -> (pixel, channel)
template <typename TX, typename TY>
__global__ void __launch_bounds__(256) mse_fwd_kernel(const TX *x, int xcs, const TY *y, int ycs, int C, long long P, double *acc) {
    pdl_sync();
    __shared__ double sm[256];
    double a = 0.0;
    const long long total = P * C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long p = e / C;
        int c = (int)(e - p * C);
        float d = to_f(x[p * xcs + c]) - to_f(y[p * ycs + c]);
        a += (double)(d * d);
    }
    double s = block_sum<256>(a, sm);
    if (threadIdx.x == 0) atomicAdd(acc, s);
}

__global__ void mse_finalize_kernel(const double *acc, double total, float *out1) {
    pdl_sync(); out1[0] = (float)(acc[0] / total); }

template <typename TX, typename TY>
__global__ void __launch_bounds__(256) mse_bwd_kernel(const TX *x, int xcs, const TY *y, int ycs, int C, long long P,
                                                      const float *gscale, TX *dx, int dcs) {
    pdl_sync();
    const long long total = P * C;
    const float k = 2.f * (gscale ? gscale[0] : 1.f) / (float)total;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long p = e / C;
        int c = (int)(e - p * C);
        dx[p * dcs + c] = from_f<TX>(k * (to_f(x[p * xcs + c]) - to_f(y[p * ycs + c])));
    }
}

// berHu pass 1: max |e| over valid ; non-negative floats order like their bit patterns
template <typename TX, typename TY>
__global__ void __launch_bounds__(256) berhu_max_kernel(const TX *x, int xcs, const TY *y, int ycs, int C, long long P,
                                                        float vmin, unsigned int *maxbits) {
    pdl_sync();
    float m = 0.f;
    const long long total = P * C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long p = e / C;
        int c = (int)(e - p * C);
        float t = to_f(y[p * ycs + c]);
        if (t > vmin) m = fmaxf(m, fabsf(to_f(x[p * xcs + c]) - t));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(maxbits, __float_as_uint(m));
}

template <typename TX, typename TY>
__global__ void __launch_bounds__(256) berhu_sum_kernel(const TX *x, int xcs, const TY *y, int ycs, int C, long long P, float vmin,
                                                        const unsigned int *maxbits, double *acc2) {
    pdl_sync();
    __shared__ double sm[256];
    const float cth = 0.2f * __uint_as_float(maxbits[0]);
    double a = 0.0, n = 0.0;
    const long long total = P * C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long p = e / C;
        int c = (int)(e - p * C);
        float t = to_f(y[p * ycs + c]);
        if (t > vmin) {
            float d = fabsf(to_f(x[p * xcs + c]) - t);
            a += (double)(d <= cth ? d : (d * d + cth * cth) / (2.f * cth));
            n += 1.0;
        }
    }
    double s = block_sum<256>(a, sm);
    double k = block_sum<256>(n, sm);
    if (threadIdx.x == 0) {
        atomicAdd(&acc2[0], s);
        atomicAdd(&acc2[1], k);
    }
}

__global__ void berhu_finalize_kernel(const double *acc2, const unsigned int *maxbits, float *out3) {
    pdl_sync();
    out3[0] = (float)(acc2[0] / acc2[1]);
    out3[1] = (float)acc2[1];
    out3[2] = 0.2f * __uint_as_float(maxbits[0]);
}

template <typename TX, typename TY>
__global__ void __launch_bounds__(256) berhu_bwd_kernel(const TX *x, int xcs, const TY *y, int ycs, int C, long long P, float vmin,
                                                        const float *out3, const float *gscale, TX *dx, int dcs) {
    pdl_sync();
    const float k = (gscale ? gscale[0] : 1.f) / out3[1], cth = out3[2];
    const long long total = P * C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long p = e / C;
        int c = (int)(e - p * C);
        float t = to_f(y[p * ycs + c]);
        float g = 0.f;
        if (t > vmin) {
            float d = to_f(x[p * xcs + c]) - t;
            float ad = fabsf(d);
            float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
            g = ad <= cth ? sg : d / cth;
        }
        dx[p * dcs + c] = from_f<TX>(k * g);
    }
}

static inline int lgrid(long long total) {
    long long b = (total + 255) / 256, cap = (long long)NASB_SM_COUNT * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}
static inline bool fdt(const NasbTensor *t) { return t && (t->dtype == NASB_F32 || t->dtype == NASB_BF16); }

}  // namespace nasb

using namespace nasb;
#define ST ((cudaStream_t)stream)

extern "C" long long nasb_loss_workspace(void) { return 64; }

extern "C" int nasb_ce_fwd(const NasbTensor *logits, const int64_t *target, int ignore_index, float *out2, void *workspace,
                           void *stream) {
    if (!fdt(logits) || !target || !out2 || !workspace) return NASB_ERR_BAD_ARG;
    long long P = npix(*logits);
    double *acc = (double *)workspace;
    cudaError_t e = cudaMemsetAsync(acc, 0, 16, ST);
    if (e != cudaSuccess) return (int)e;
    if (P > 0) {
        if (logits->dtype == NASB_BF16)
            nasb::launch_pdl((ce_fwd_kernel<bf16>), dim3(lgrid(P)), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)logits->ptr, logits->cstride, logits->c, target, ignore_index,
                                                          P, acc);
        else
            nasb::launch_pdl((ce_fwd_kernel<float>), dim3(lgrid(P)), dim3(256), 0, (cudaStream_t)(ST), (const float *)logits->ptr, logits->cstride, logits->c, target,
                                                           ignore_index, P, acc);
        NASB_CHECK_LAUNCH();
    }
    nasb::launch_pdl((ce_finalize_kernel), dim3(1), dim3(1), 0, (cudaStream_t)(ST), acc, out2);
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_ce_bwd(const NasbTensor *logits, const int64_t *target, int ignore_index, const float *out2,
                           const float *gscale_dev, const NasbTensor *dlogits, void *stream) {
    if (!fdt(logits) || !fdt(dlogits) || !target || !out2 || logits->c != dlogits->c || npix(*logits) != npix(*dlogits))
        return NASB_ERR_BAD_ARG;
    long long P = npix(*logits);
    if (P == 0) return 0;
#define CEB(T, TG)                                                                                                          \
    nasb::launch_pdl((ce_bwd_kernel<T, TG>), dim3(lgrid(P)), dim3(256), 0, (cudaStream_t)(ST), (const T *)logits->ptr, logits->cstride, logits->c, target, ignore_index, P, \
                                                   out2, gscale_dev, (TG *)dlogits->ptr, dlogits->cstride)
    if (logits->dtype == NASB_BF16) {
        if (dlogits->dtype == NASB_BF16) CEB(bf16, bf16); else CEB(bf16, float);
    } else {
        if (dlogits->dtype == NASB_BF16) CEB(float, bf16); else CEB(float, float);
    }
#undef CEB
    NASB_CHECK_LAUNCH();
    return 0;
}

#define NASB_XY(KERNEL, ...)                                                                       \
    do {                                                                                           \
        if (x->dtype == NASB_BF16) {                                                               \
            if (y->dtype == NASB_BF16) nasb::launch_pdl((KERNEL<bf16, bf16>), dim3(lgrid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)x->ptr, x->cstride, (const bf16 *)y->ptr, y->cstride, C, P, __VA_ARGS__); \
            else nasb::launch_pdl((KERNEL<bf16, float>), dim3(lgrid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)x->ptr, x->cstride, (const float *)y->ptr, y->cstride, C, P, __VA_ARGS__); \
        } else {                                                                                   \
            if (y->dtype == NASB_BF16) nasb::launch_pdl((KERNEL<float, bf16>), dim3(lgrid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const float *)x->ptr, x->cstride, (const bf16 *)y->ptr, y->cstride, C, P, __VA_ARGS__); \
            else nasb::launch_pdl((KERNEL<float, float>), dim3(lgrid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const float *)x->ptr, x->cstride, (const float *)y->ptr, y->cstride, C, P, __VA_ARGS__); \
        }                                                                                          \
    } while (0)

extern "C" int nasb_mse_fwd(const NasbTensor *x, const NasbTensor *y, float *out1, void *workspace, void *stream) {
    if (!fdt(x) || !fdt(y) || !out1 || !workspace || x->c != y->c || npix(*x) != npix(*y)) return NASB_ERR_BAD_ARG;
    long long P = npix(*x);
    int C = x->c;
    double *acc = (double *)workspace;
    cudaError_t e = cudaMemsetAsync(acc, 0, 8, ST);
    if (e != cudaSuccess) return (int)e;
    if (P > 0) {
        NASB_XY(mse_fwd_kernel, acc);
        NASB_CHECK_LAUNCH();
    }
    nasb::launch_pdl((mse_finalize_kernel), dim3(1), dim3(1), 0, (cudaStream_t)(ST), acc, (double)P * (double)C, out1);
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_mse_bwd(const NasbTensor *x, const NasbTensor *y, const float *gscale_dev, const NasbTensor *dx,
                            void *stream) {
    if (!fdt(x) || !fdt(y) || !fdt(dx) || x->c != y->c || dx->c != x->c || dx->dtype != x->dtype || npix(*x) != npix(*y) ||
        npix(*x) != npix(*dx))
        return NASB_ERR_BAD_ARG;
    long long P = npix(*x);
    int C = x->c;
    if (P == 0) return 0;
    if (x->dtype == NASB_BF16) {
        bf16 *d = (bf16 *)dx->ptr;
        if (y->dtype == NASB_BF16) nasb::launch_pdl((mse_bwd_kernel<bf16, bf16>), dim3(lgrid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)x->ptr, x->cstride, (const bf16 *)y->ptr, y->cstride, C, P, gscale_dev, d, dx->cstride);
        else nasb::launch_pdl((mse_bwd_kernel<bf16, float>), dim3(lgrid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)x->ptr, x->cstride, (const float *)y->ptr, y->cstride, C, P, gscale_dev, d, dx->cstride);
    } else {
        float *d = (float *)dx->ptr;
        if (y->dtype == NASB_BF16) nasb::launch_pdl((mse_bwd_kernel<float, bf16>), dim3(lgrid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const float *)x->ptr, x->cstride, (const bf16 *)y->ptr, y->cstride, C, P, gscale_dev, d, dx->cstride);
        else nasb::launch_pdl((mse_bwd_kernel<float, float>), dim3(lgrid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const float *)x->ptr, x->cstride, (const float *)y->ptr, y->cstride, C, P, gscale_dev, d, dx->cstride);
    }
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_berhu_fwd(const NasbTensor *x, const NasbTensor *y, float valid_min, float *out3, void *workspace,
                              void *stream) {
    if (!fdt(x) || !fdt(y) || !out3 || !workspace || x->c != y->c || npix(*x) != npix(*y)) return NASB_ERR_BAD_ARG;
    long long P = npix(*x);
    int C = x->c;
    double *acc = (double *)workspace;
    unsigned int *mb = (unsigned int *)(acc + 2);
    cudaError_t e = cudaMemsetAsync(acc, 0, 24, ST);
    if (e != cudaSuccess) return (int)e;
    if (P > 0) {
        NASB_XY(berhu_max_kernel, valid_min, mb);
        NASB_CHECK_LAUNCH();
        NASB_XY(berhu_sum_kernel, valid_min, mb, acc);
        NASB_CHECK_LAUNCH();
    }
    nasb::launch_pdl((berhu_finalize_kernel), dim3(1), dim3(1), 0, (cudaStream_t)(ST), acc, mb, out3);
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_berhu_bwd(const NasbTensor *x, const NasbTensor *y, float valid_min, const float *out3,
                              const float *gscale_dev, const NasbTensor *dx, void *stream) {
    if (!fdt(x) || !fdt(y) || !fdt(dx) || !out3 || x->c != y->c || dx->c != x->c || dx->dtype != x->dtype ||
        npix(*x) != npix(*y) || npix(*x) != npix(*dx))
        return NASB_ERR_BAD_ARG;
    long long P = npix(*x);
    int C = x->c;
    if (P == 0) return 0;
    if (x->dtype == NASB_BF16) {
        bf16 *d = (bf16 *)dx->ptr;
        if (y->dtype == NASB_BF16) nasb::launch_pdl((berhu_bwd_kernel<bf16, bf16>), dim3(lgrid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)x->ptr, x->cstride, (const bf16 *)y->ptr, y->cstride, C, P, valid_min, out3, gscale_dev, d, dx->cstride);
        else nasb::launch_pdl((berhu_bwd_kernel<bf16, float>), dim3(lgrid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)x->ptr, x->cstride, (const float *)y->ptr, y->cstride, C, P, valid_min, out3, gscale_dev, d, dx->cstride);
    } else {
        float *d = (float *)dx->ptr;
        if (y->dtype == NASB_BF16) nasb::launch_pdl((berhu_bwd_kernel<float, bf16>), dim3(lgrid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const float *)x->ptr, x->cstride, (const bf16 *)y->ptr, y->cstride, C, P, valid_min, out3, gscale_dev, d, dx->cstride);
        else nasb::launch_pdl((berhu_bwd_kernel<float, float>), dim3(lgrid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const float *)x->ptr, x->cstride, (const float *)y->ptr, y->cstride, C, P, valid_min, out3, gscale_dev, d, dx->cstride);
    }
    NASB_CHECK_LAUNCH();
    return 0;
}
