// stem.cu -- the encoder stem: dense 3x3 convolution of the 3-channel planar fp32 image (the layout the reference's
// DataLoader produces, src/data/datasets.py:222-232) into 32 NHWC channels (src/nn/encoders.py:38, layer_factory.py:109).
// With K = 27 this layer is far too thin for an implicit GEMM: one thread owns one output pixel and all 32 output channels
// (27 image loads, 864 FMAs against a shared-memory copy of the weights, one 64-byte coalesced store).  HBM-bound:
// 12 B/pixel of image in, 64 B/pixel out.  The weight gradient stages 64 pixels of dz and of the 27-tap input patch in shared
// memory and lets thread (co, k-group) accumulate its 3-4 products per pixel.
#include "common.cuh"

namespace nasb {

constexpr int STEM_CO = 32, STEM_K = 27;

struct StemP {
    const float *img;  // [N][3][IH][IW]
    int N, IH, IW, OH, OW;
    int stride, pad, dil;
    const float *w;  // [32][3][3][3]
    const float *scale, *shift;
    int act;
    void *out;
    int out_cs, out_dtype;
};

__global__ void __launch_bounds__(128) stem_fwd_kernel(const StemP p) {
    pdl_sync();
    __shared__ __align__(16) float ws[STEM_K][STEM_CO];
    __shared__ float s_sc[STEM_CO], s_sh[STEM_CO];
    for (int i = threadIdx.x; i < STEM_K * STEM_CO; i += blockDim.x) {
        int k = i / STEM_CO, co = i - k * STEM_CO;
        ws[k][co] = p.w[co * STEM_K + k];
    }
    if (threadIdx.x < STEM_CO) {
        s_sc[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
        s_sh[threadIdx.x] = p.shift ? p.shift[threadIdx.x] : 0.f;
    }
    __syncthreads();
    const long long total = (long long)p.N * p.OH * p.OW;
    const long long plane = (long long)p.IH * p.IW;
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
        int ox = (int)(pix % p.OW);
        long long t = pix / p.OW;
        int oy = (int)(t % p.OH);
        int n = (int)(t / p.OH);
        float acc[STEM_CO];
#pragma unroll
        for (int c = 0; c < STEM_CO; ++c) acc[c] = 0.f;
        const float *ib = p.img + (long long)n * 3 * plane;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                int iy = oy * p.stride - p.pad + ky * p.dil;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    int ix = ox * p.stride - p.pad + kx * p.dil;
                    float x = (iy >= 0 && iy < p.IH && ix >= 0 && ix < p.IW) ? ib[ci * plane + (long long)iy * p.IW + ix] : 0.f;
                    const float4 *wr = reinterpret_cast<const float4 *>(ws[ci * 9 + ky * 3 + kx]);
#pragma unroll
                    for (int q = 0; q < STEM_CO / 4; ++q) {
                        float4 w4 = wr[q];
                        acc[q * 4 + 0] = fmaf(x, w4.x, acc[q * 4 + 0]);
                        acc[q * 4 + 1] = fmaf(x, w4.y, acc[q * 4 + 1]);
                        acc[q * 4 + 2] = fmaf(x, w4.z, acc[q * 4 + 2]);
                        acc[q * 4 + 3] = fmaf(x, w4.w, acc[q * 4 + 3]);
                    }
                }
            }
#pragma unroll
        for (int c = 0; c < STEM_CO; ++c) acc[c] = apply_act(acc[c] * s_sc[c] + s_sh[c], p.act);
        if (p.out_dtype == NASB_BF16) {
            bf16 *o = reinterpret_cast<bf16 *>(p.out) + pix * p.out_cs;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = acc[q * 8 + j];
                store_vec<bf16, 8>(o + q * 8, v);
            }
        } else {
            float *o = reinterpret_cast<float *>(p.out) + pix * p.out_cs;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = acc[q * 4 + j];
                store_vec<float, 4>(o + q * 4, v);
            }
        }
    }
}


// ---- speed mode: the stem as a tensor-core GEMM.  One pass turns the planar fp32 image into the bf16 patch matrix
// [N*OH*OW][32] (k = ci*9 + ky*3 + kx for k < 27, zero for k = 27..31 -- 64 bytes per output pixel); the 3x3x3 -> 32
// convolution is then the pointwise tcgen05 kernel with K = 32 (fused BN statistics included) and its weight gradient is
// nasb_pw_tc_wgrad on the same matrix.  Thread = output pixel: 27 loads (adjacent threads read adjacent columns),
// four 16-byte stores.
__global__ void __launch_bounds__(256) stem_im2col_kernel(const StemP p) {
    pdl_sync();
    const long long total = (long long)p.N * p.OH * p.OW;
    const long long plane = (long long)p.IH * p.IW;
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(pix % p.OW);
        const long long t = pix / p.OW;
        const int oy = (int)(t % p.OH);
        const int n = (int)(t / p.OH);
        const float *ib = p.img + (long long)n * 3 * plane;
        float v[32];
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int iy = oy * p.stride - p.pad + ky * p.dil;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int ix = ox * p.stride - p.pad + kx * p.dil;
                    v[ci * 9 + ky * 3 + kx] =
                        (iy >= 0 && iy < p.IH && ix >= 0 && ix < p.IW) ? ib[ci * plane + (long long)iy * p.IW + ix] : 0.f;
                }
            }
#pragma unroll
        for (int k = STEM_K; k < 32; ++k) v[k] = 0.f;
        uint4 *o = reinterpret_cast<uint4 *>(reinterpret_cast<bf16 *>(p.out) + pix * p.out_cs);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            o[q] = make_uint4(pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]), pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]),
                              pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]));
    }
}

constexpr int STEM_WP = 64;  // pixels staged per step of the weight gradient

template <typename T>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const StemP p, const T *dz, int dz_cs, float *dw, long long rows_per_cta) {
    pdl_sync();
    __shared__ float Zs[STEM_WP][STEM_CO];
    __shared__ float Xs[STEM_WP][STEM_K + 1];
    const int tid = threadIdx.x;
    const int co = tid & 31, kg = tid >> 5;  // warp = one k-group, lanes = output channels
    const long long M = (long long)p.N * p.OH * p.OW;
    const long long plane = (long long)p.IH * p.IW;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long m0 = r0; m0 < r1; m0 += STEM_WP) {
        __syncthreads();
        for (int i = tid; i < STEM_WP * STEM_CO; i += blockDim.x) {
            int pp = i >> 5, c = i & 31;
            long long m = m0 + pp;
            Zs[pp][c] = m < r1 ? to_f(dz[m * dz_cs + c]) : 0.f;
        }
        for (int i = tid; i < STEM_WP * STEM_K; i += blockDim.x) {
            int pp = i / STEM_K, k = i - pp * STEM_K;
            long long m = m0 + pp;
            float x = 0.f;
            if (m < r1) {
                int ox = (int)(m % p.OW);
                long long t = m / p.OW;
                int oy = (int)(t % p.OH);
                int n = (int)(t / p.OH);
                int ci = k / 9, r = k - ci * 9, ky = r / 3, kx = r - ky * 3;
                int iy = oy * p.stride - p.pad + ky * p.dil, ix = ox * p.stride - p.pad + kx * p.dil;
                if (iy >= 0 && iy < p.IH && ix >= 0 && ix < p.IW) x = p.img[((long long)n * 3 + ci) * plane + (long long)iy * p.IW + ix];
            }
            Xs[pp][k] = x;
        }
        __syncthreads();
#pragma unroll 4
        for (int pp = 0; pp < STEM_WP; ++pp) {
            float z = Zs[pp][co];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int k = kg + 8 * j;
                if (k < STEM_K) acc[j] = fmaf(z, Xs[pp][k], acc[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int k = kg + 8 * j;
        if (k < STEM_K) atomicAdd(&dw[co * STEM_K + k], acc[j]);
    }
}

}  // namespace nasb

using namespace nasb;

static bool stem_shape_ok(const NasbTensor *img, const NasbTensor *out, int ks) {
    return img && out && img->dtype == NASB_F32_NCHW && img->c == 3 && out->c == STEM_CO && ks == 3 &&
           (out->dtype == NASB_BF16 || out->dtype == NASB_F32) && out->n == img->n;
}

extern "C" int nasb_stem_fwd(const NasbTensor *img, const float *weight, int ks, int stride, int dil, int pad,
                             const float *out_scale, const float *out_shift, int act, const NasbTensor *out, void *stream) {
    if (!stem_shape_ok(img, out, ks) || !weight) return NASB_ERR_UNSUPPORTED;
    if (!vec_ok(*out, 8)) return NASB_ERR_UNSUPPORTED;
    int eh = (img->h + 2 * pad - dil * 2 - 1) / stride + 1, ew = (img->w + 2 * pad - dil * 2 - 1) / stride + 1;
    if (eh != out->h || ew != out->w) return NASB_ERR_BAD_ARG;
    StemP p{};
    p.img = (const float *)img->ptr;
    p.N = img->n;
    p.IH = img->h;
    p.IW = img->w;
    p.OH = out->h;
    p.OW = out->w;
    p.stride = stride;
    p.pad = pad;
    p.dil = dil;
    p.w = weight;
    p.scale = out_scale;
    p.shift = out_shift;
    p.act = act;
    p.out = out->ptr;
    p.out_cs = out->cstride;
    p.out_dtype = out->dtype;
    long long total = npix(*out);
    if (total == 0) return 0;
    long long blocks = (total + 127) / 128, cap = (long long)NASB_SM_COUNT * 32;
    if (blocks > cap) blocks = cap;
    nasb::launch_pdl((stem_fwd_kernel), dim3((int)blocks), dim3(128), 0, (cudaStream_t)((cudaStream_t)stream), p);
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_stem_wgrad(const NasbTensor *img, const NasbTensor *dz, int ks, int stride, int dil, int pad,
                               float *dweight, void *stream) {
    if (!stem_shape_ok(img, dz, ks) || !dweight) return NASB_ERR_UNSUPPORTED;
    StemP p{};
    p.img = (const float *)img->ptr;
    p.N = img->n;
    p.IH = img->h;
    p.IW = img->w;
    p.OH = dz->h;
    p.OW = dz->w;
    p.stride = stride;
    p.pad = pad;
    p.dil = dil;
    long long M = npix(*dz);
    if (M == 0) return 0;
    long long want = (long long)NASB_SM_COUNT * 8;
    long long rows = (M + want - 1) / want;
    rows = (rows + STEM_WP - 1) / STEM_WP * STEM_WP;
    if (rows < STEM_WP * 4) rows = STEM_WP * 4;
    int blocks = cdiv(M, rows);
    if (dz->dtype == NASB_BF16)
        nasb::launch_pdl((stem_wgrad_kernel<bf16>), dim3(blocks), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), p, (const bf16 *)dz->ptr, dz->cstride, dweight, rows);
    else
        nasb::launch_pdl((stem_wgrad_kernel<float>), dim3(blocks), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), p, (const float *)dz->ptr, dz->cstride, dweight, rows);
    NASB_CHECK_LAUNCH();
    return 0;
}

// image [N][3][IH][IW] fp32 planar -> bf16 patch matrix `out` [N, OH, OW, 32] (see stem_im2col_kernel)
extern "C" int nasb_stem_im2col(const NasbTensor *img, int ks, int stride, int dil, int pad, const NasbTensor *out, void *stream) {
    if (!img || !out || img->dtype != NASB_F32_NCHW || img->c != 3 || ks != 3 || out->dtype != NASB_BF16 || out->c != 32 ||
        out->n != img->n || !vec_ok(*out, 8))
        return NASB_ERR_UNSUPPORTED;
    int eh = (img->h + 2 * pad - dil * 2 - 1) / stride + 1, ew = (img->w + 2 * pad - dil * 2 - 1) / stride + 1;
    if (eh != out->h || ew != out->w) return NASB_ERR_BAD_ARG;
    StemP p{};
    p.img = (const float *)img->ptr;
    p.N = img->n;
    p.IH = img->h;
    p.IW = img->w;
    p.OH = out->h;
    p.OW = out->w;
    p.stride = stride;
    p.pad = pad;
    p.dil = dil;
    p.out = out->ptr;
    p.out_cs = out->cstride;
    long long total = npix(*out);
    if (total == 0) return 0;
    long long blocks = (total + 255) / 256, cap = (long long)NASB_SM_COUNT * 16;
    if (blocks > cap) blocks = cap;
    nasb::launch_pdl((stem_im2col_kernel), dim3((int)blocks), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), p);
    NASB_CHECK_LAUNCH();
    return 0;
}
