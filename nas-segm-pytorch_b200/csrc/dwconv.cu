// dwconv.cu -- depthwise k x k convolution (groups == channels) in NHWC: forward, data gradient, weight gradient.
// HBM-bound: each thread owns one pixel x one V-channel vector (16 B for bf16, 16 B for fp32), neighbouring
// threads own neighbouring channel vectors so every warp access is a contiguous 512 B run; taps re-read the
// input through L1/L2.  Weights ([C][1][k][k] fp32) are staged once per CTA into shared memory as [tap][C].
#include "common.cuh"

namespace nasb {

struct DwP {
    const void *x;
    int x_cs;
    void *out;
    int out_cs;
    int N, IH, IW, OH, OW, C;
    int ks, stride, dil, pad;
    int in_relu;
    const float *w;
    const float *scale, *shift;
    int act;
    int mode;  // 0 forward (rows = output pixels), 1 data gradient (rows = input pixels, x = dz)
};

template <typename T, int V>
__global__ void __launch_bounds__(256) dwconv_kernel(const DwP p) {
    pdl_sync();
    extern __shared__ float wsm[];  // [ks*ks][C]
    const int KK = p.ks * p.ks;
    for (int i = threadIdx.x; i < KK * p.C; i += blockDim.x) {
        int tap = i / p.C, c = i - tap * p.C;
        wsm[i] = p.w[(long long)c * KK + tap];
    }
    __syncthreads();
    const int CV = p.C / V;
    const int RH = p.mode == 0 ? p.OH : p.IH, RW = p.mode == 0 ? p.OW : p.IW;  // row space
    const int SH = p.mode == 0 ? p.IH : p.OH, SW = p.mode == 0 ? p.IW : p.OW;  // space read
    const long long total = (long long)p.N * RH * RW * CV;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        int rx = (int)(pix % RW);
        long long t = pix / RW;
        int ry = (int)(t % RH);
        int n = (int)(t / RH);
        const int c0 = cv * V;
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = 0.f;
        for (int ky = 0; ky < p.ks; ++ky) {
            int sy;
            if (p.mode == 0) {
                sy = ry * p.stride - p.pad + ky * p.dil;
            } else {
                int ty = ry + p.pad - ky * p.dil;
                if (ty < 0 || ty % p.stride) continue;
                sy = ty / p.stride;
            }
            if (sy < 0 || sy >= SH) continue;
            for (int kx = 0; kx < p.ks; ++kx) {
                int sx;
                if (p.mode == 0) {
                    sx = rx * p.stride - p.pad + kx * p.dil;
                } else {
                    int tx = rx + p.pad - kx * p.dil;
                    if (tx < 0 || tx % p.stride) continue;
                    sx = tx / p.stride;
                }
                if (sx < 0 || sx >= SW) continue;
                float v[V];
                load_vec<T, V>(reinterpret_cast<const T *>(p.x) + (((long long)n * SH + sy) * SW + sx) * p.x_cs + c0, v);
                const float *wt = wsm + (ky * p.ks + kx) * p.C + c0;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    float xv = p.in_relu ? fmaxf(v[j], 0.f) : v[j];
                    acc[j] = fmaf(xv, wt[j], acc[j]);
                }
            }
        }
        if (p.scale || p.shift || p.act) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
                float s = p.scale ? p.scale[c0 + j] : 1.f, b = p.shift ? p.shift[c0 + j] : 0.f;
                acc[j] = apply_act(acc[j] * s + b, p.act);
            }
        }
        store_vec<T, V>(reinterpret_cast<T *>(p.out) + pix * p.out_cs + c0, acc);
    }
}

// weight gradient: dw[c][tap] += sum_{n,oy,ox} dz[n,oy,ox,c] * pro(x)[n, oy*s-pad+ky*d, ox*s-pad+kx*d, c]
// block = (CL channel lanes) x (PL pixel lanes); every thread keeps KS*KS partial sums for its channel.
template <typename T, int KS>
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const DwP p, const void *dz, int dz_cs, float *dw,
                                                           long long rows_per_cta) {
    pdl_sync();
    constexpr int KK = KS * KS;
    __shared__ float red[256];
    const int CL = blockDim.x, PL = blockDim.y;
    const int c = blockIdx.x * CL + threadIdx.x;
    const long long M = (long long)p.N * p.OH * p.OW;
    const long long r0 = (long long)blockIdx.y * rows_per_cta;
    const long long r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
    float acc[KK];
#pragma unroll
    for (int i = 0; i < KK; ++i) acc[i] = 0.f;
    if (c < p.C) {
        for (long long m = r0 + threadIdx.y; m < r1; m += PL) {
            int ox = (int)(m % p.OW);
            long long t = m / p.OW;
            int oy = (int)(t % p.OH);
            int n = (int)(t / p.OH);
            float g = to_f(reinterpret_cast<const T *>(dz)[m * dz_cs + c]);
#pragma unroll
            for (int ky = 0; ky < KS; ++ky) {
                int sy = oy * p.stride - p.pad + ky * p.dil;
                if (sy < 0 || sy >= p.IH) continue;
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    int sx = ox * p.stride - p.pad + kx * p.dil;
                    if (sx < 0 || sx >= p.IW) continue;
                    float xv = to_f(reinterpret_cast<const T *>(p.x)[(((long long)n * p.IH + sy) * p.IW + sx) * p.x_cs + c]);
                    if (p.in_relu) xv = fmaxf(xv, 0.f);
                    acc[ky * KS + kx] = fmaf(g, xv, acc[ky * KS + kx]);
                }
            }
        }
    }
    // reduce over the pixel lanes (threadIdx.y) through shared memory, one tap at a time
    const int lin = threadIdx.y * CL + threadIdx.x;
#pragma unroll
    for (int tap = 0; tap < KK; ++tap) {
        red[lin] = acc[tap];
        __syncthreads();
        if (threadIdx.y == 0 && c < p.C) {
            float s = 0.f;
            for (int y = 0; y < PL; ++y) s += red[y * CL + threadIdx.x];
            atomicAdd(&dw[(long long)c * KK + tap], s);
        }
        __syncthreads();
    }
}


// vectorised weight gradient: thread = (channel vector of V channels) x (pixel lane); a warp reads whole pixels
// contiguously (V*sizeof(T) = 16 B for k=3, 8 B for k=5, 4 B for k=7 keeps k*k*V accumulators in registers).
template <typename T, int KS, int V>
__global__ void __launch_bounds__(256) dwconv_wgrad_vec_kernel(const DwP p, const void *dz_, int dz_cs, float *dw,
                                                               long long rows_per_cta) {
    pdl_sync();
    constexpr int KK = KS * KS;
    extern __shared__ float red_sm[];  // [PL][C]
    const T *dz = reinterpret_cast<const T *>(dz_);
    const T *x = reinterpret_cast<const T *>(p.x);
    const int CV = p.C / V, PL = blockDim.x / CV;
    const int t = threadIdx.x, cv = t % CV, pl = t / CV, c0 = cv * V;
    const long long M = (long long)p.N * p.OH * p.OW;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
    float acc[KK][V];
#pragma unroll
    for (int i = 0; i < KK; ++i)
#pragma unroll
        for (int j = 0; j < V; ++j) acc[i][j] = 0.f;
    if (pl < PL) {
        for (long long m = r0 + pl; m < r1; m += PL) {
            int ox = (int)(m % p.OW);
            long long tt = m / p.OW;
            int oy = (int)(tt % p.OH);
            int n = (int)(tt / p.OH);
            float g[V];
            load_vec<T, V>(dz + m * dz_cs + c0, g);
#pragma unroll
            for (int ky = 0; ky < KS; ++ky) {
                int sy = oy * p.stride - p.pad + ky * p.dil;
                if (sy < 0 || sy >= p.IH) continue;
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    int sx = ox * p.stride - p.pad + kx * p.dil;
                    if (sx < 0 || sx >= p.IW) continue;
                    float xv[V];
                    load_vec<T, V>(x + (((long long)n * p.IH + sy) * p.IW + sx) * p.x_cs + c0, xv);
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        float v = p.in_relu ? fmaxf(xv[j], 0.f) : xv[j];
                        acc[ky * KS + kx][j] = fmaf(g[j], v, acc[ky * KS + kx][j]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int tap = 0; tap < KK; ++tap) {
        if (pl < PL) {
#pragma unroll
            for (int j = 0; j < V; ++j) red_sm[(size_t)pl * p.C + c0 + j] = acc[tap][j];
        }
        __syncthreads();
        for (int c = t; c < p.C; c += blockDim.x) {
            float sum = 0.f;
            for (int i = 0; i < PL; ++i) sum += red_sm[(size_t)i * p.C + c];
            atomicAdd(&dw[(long long)c * KK + tap], sum);
        }
        __syncthreads();
    }
}

template <typename T, int KS, int V>
static bool launch_dw_wgrad_vec(const DwP &p, const NasbTensor *x, const NasbTensor *dz, float *dweight, cudaStream_t st) {
    if (!vec_ok(*x, V) || !vec_ok(*dz, V)) return false;
    int CV = p.C / V;
    if (CV < 1 || CV > 256) return false;
    int PL = 256 / CV;
    size_t smem = (size_t)PL * p.C * sizeof(float);
    if (smem > 48 * 1024) return false;
    long long M = (long long)p.N * p.OH * p.OW;
    long long want = (long long)NASB_SM_COUNT * 8;
    long long rows = (M + want - 1) / want;
    if (rows < (long long)PL * 48) rows = (long long)PL * 48;  // >= 48 pixels per thread before the per-tap reduction
    nasb::launch_pdl((dwconv_wgrad_vec_kernel<T, KS, V>), dim3(cdiv(M, rows)), dim3(256), smem, (cudaStream_t)(st), p, dz->ptr, dz->cstride, dweight, rows);
    return true;
}

template <typename T>
static int launch_dw(const DwP &p, bool vec, long long rows, cudaStream_t st) {
    constexpr int V = 16 / sizeof(T);
    size_t smem = (size_t)p.ks * p.ks * p.C * sizeof(float);
    int cv = vec ? p.C / V : p.C;
    long long total = rows * cv;
    int blocks = (int)((total + 255) / 256);
    int cap = NASB_SM_COUNT * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (vec) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(dwconv_kernel<T, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        nasb::launch_pdl((dwconv_kernel<T, V>), dim3(blocks), dim3(256), smem, (cudaStream_t)(st), p);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(dwconv_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        nasb::launch_pdl((dwconv_kernel<T, 1>), dim3(blocks), dim3(256), smem, (cudaStream_t)(st), p);
    }
    return 0;
}

}  // namespace nasb

using namespace nasb;

static int dw_common(const NasbTensor *x, const NasbTensor *out, const float *weight, int ks, int stride, int dil, int pad,
                     int in_relu, const float *scale, const float *shift, int act, int mode, void *stream) {
    if (!x || !out || !weight || (ks != 3 && ks != 5 && ks != 7)) return NASB_ERR_BAD_ARG;
    if (x->dtype != out->dtype || x->dtype == NASB_F32_NCHW || x->c != out->c || x->n != out->n) return NASB_ERR_BAD_ARG;
    if ((size_t)ks * ks * x->c * 4 > 200 * 1024) return NASB_ERR_UNSUPPORTED;
    DwP p{};
    p.x = x->ptr;
    p.x_cs = x->cstride;
    p.out = out->ptr;
    p.out_cs = out->cstride;
    p.N = x->n;
    p.C = x->c;
    // geometry is always stated for the FORWARD conv: (IH,IW) input, (OH,OW) output
    const NasbTensor *fin = mode == 0 ? x : out, *fout = mode == 0 ? out : x;
    p.IH = fin->h;
    p.IW = fin->w;
    p.OH = fout->h;
    p.OW = fout->w;
    int eh = (p.IH + 2 * pad - dil * (ks - 1) - 1) / stride + 1, ew = (p.IW + 2 * pad - dil * (ks - 1) - 1) / stride + 1;
    if (eh != p.OH || ew != p.OW) return NASB_ERR_BAD_ARG;
    p.ks = ks;
    p.stride = stride;
    p.dil = dil;
    p.pad = pad;
    p.in_relu = in_relu;
    p.w = weight;
    p.scale = scale;
    p.shift = shift;
    p.act = act;
    p.mode = mode;
    long long rows = npix(*out);
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (x->dtype == NASB_BF16)
        launch_dw<bf16>(p, vec_ok(*x, 8) && vec_ok(*out, 8), rows, st);
    else
        launch_dw<float>(p, vec_ok(*x, 4) && vec_ok(*out, 4), rows, st);
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_dwconv_fwd(const NasbTensor *x, const float *weight, int ks, int stride, int dil, int pad,
                               int in_relu, const float *out_scale, const float *out_shift, int act,
                               const NasbTensor *out, void *stream) {
    return dw_common(x, out, weight, ks, stride, dil, pad, in_relu, out_scale, out_shift, act, 0, stream);
}

extern "C" int nasb_dwconv_dgrad(const NasbTensor *dz, const float *weight, int ks, int stride, int dil, int pad,
                                 const NasbTensor *dx, void *stream) {
    return dw_common(dz, dx, weight, ks, stride, dil, pad, 0, nullptr, nullptr, NASB_ACT_NONE, 1, stream);
}

extern "C" int nasb_dwconv_wgrad(const NasbTensor *x, int in_relu, const NasbTensor *dz, int ks, int stride, int dil,
                                 int pad, float *dweight, void *stream) {
    if (!x || !dz || !dweight || (ks != 3 && ks != 5 && ks != 7)) return NASB_ERR_BAD_ARG;
    if (x->dtype != dz->dtype || x->dtype == NASB_F32_NCHW || x->c != dz->c) return NASB_ERR_BAD_ARG;
    DwP p{};
    p.x = x->ptr;
    p.x_cs = x->cstride;
    p.N = x->n;
    p.C = x->c;
    p.IH = x->h;
    p.IW = x->w;
    p.OH = dz->h;
    p.OW = dz->w;
    p.ks = ks;
    p.stride = stride;
    p.dil = dil;
    p.pad = pad;
    p.in_relu = in_relu;
    long long M = npix(*dz);
    if (M == 0) return 0;
    {
        cudaStream_t st = (cudaStream_t)stream;
        bool done = false;
        if (x->dtype == NASB_BF16) {
            if (ks == 3) done = launch_dw_wgrad_vec<bf16, 3, 8>(p, x, dz, dweight, st);
            else if (ks == 5) done = launch_dw_wgrad_vec<bf16, 5, 4>(p, x, dz, dweight, st);
            else done = launch_dw_wgrad_vec<bf16, 7, 2>(p, x, dz, dweight, st);
        } else {
            if (ks == 3) done = launch_dw_wgrad_vec<float, 3, 4>(p, x, dz, dweight, st);
            else if (ks == 5) done = launch_dw_wgrad_vec<float, 5, 4>(p, x, dz, dweight, st);
            else done = launch_dw_wgrad_vec<float, 7, 2>(p, x, dz, dweight, st);
        }
        if (done) {
            NASB_CHECK_LAUNCH();
            return 0;
        }
    }
    const int CL = 32, PL = 8;
    int cblocks = cdiv(p.C, CL);
    long long want = (long long)NASB_SM_COUNT * 8 / cblocks;
    if (want < 1) want = 1;
    long long rows = (M + want - 1) / want;
    if (rows < 64) rows = 64;
    dim3 grid(cblocks, cdiv(M, rows)), block(CL, PL);
    cudaStream_t st = (cudaStream_t)stream;
#define NASB_DW_WG(T, KS) nasb::launch_pdl((dwconv_wgrad_kernel<T, KS>), dim3(grid), dim3(block), 0, (cudaStream_t)(st), p, dz->ptr, dz->cstride, dweight, rows)
    if (x->dtype == NASB_BF16) {
        if (ks == 3) NASB_DW_WG(bf16, 3);
        else if (ks == 5) NASB_DW_WG(bf16, 5);
        else NASB_DW_WG(bf16, 7);
    } else {
        if (ks == 3) NASB_DW_WG(float, 3);
        else if (ks == 5) NASB_DW_WG(float, 5);
        else NASB_DW_WG(float, 7);
    }
#undef NASB_DW_WG
    NASB_CHECK_LAUNCH();
    return 0;
}
