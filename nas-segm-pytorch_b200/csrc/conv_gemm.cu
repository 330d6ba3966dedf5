// conv_gemm.cu -- dense k x k convolution (k in {1,3}) as an implicit GEMM on the CUDA cores with fp32
// accumulation: forward, data gradient and weight gradient.  This is the precision-exact path (fp32 "parity
// mode") and the shape-generic path; the bf16 pointwise convolutions that dominate the network are routed to the
// tcgen05 kernel in pw_tcgen05.cu when their shapes allow it.
//
// GEMM view:  rows = output pixels (n,oy,ox), cols = output channels, reduction k = (tap, ci), tap-major so
// that consecutive k are consecutive NHWC channels.
#include "common.cuh"

namespace nasb {

struct ConvP {
    const void *src[2];
    int sc[2], scs[2];  // channels / pixel stride of each concatenated source
    int src_dtype;
    int N, IH, IW;  // spatial dims of the tensor the reduction reads (fwd: x ; dgrad: dz)
    int OH, OW;     // spatial dims of the GEMM rows          (fwd: out ; dgrad: dx)
    int Kc;         // channels on the reduction side          (fwd: C_in ; dgrad: C_out)
    int Nc;         // GEMM columns                            (fwd: C_out ; dgrad: C_in)
    int ks, stride, dil, pad;
    int mode;  // 0 forward, 1 data gradient
    const float *w;
    int w_cin;  // conv C_in (weight is [C_out][C_in][ks][ks])
    const float *in_scale, *in_shift;
    int in_relu;
    const float *out_scale, *out_shift;
    int act;
    const void *res;
    int res_cs;
    void *out[2];
    int oc[2], ocs[2];
    int out_dtype;
    int vec8;  // sources addressable as 8-channel vectors and taps/source boundaries 8-aligned
    int vec4o;  // outputs addressable as 4-channel vectors
};

// Source pixel of GEMM row (n, ry, rx) for tap (ky,kx); returns linear pixel index or -1 (zero padding).
__device__ __forceinline__ long long src_pixel(const ConvP &p, int n, int ry, int rx, int ky, int kx) {
    int sy, sx;
    if (p.mode == 0) {
        sy = ry * p.stride - p.pad + ky * p.dil;
        sx = rx * p.stride - p.pad + kx * p.dil;
    } else {
        int ty = ry + p.pad - ky * p.dil, tx = rx + p.pad - kx * p.dil;
        if (ty < 0 || tx < 0) return -1;
        if (p.stride > 1) {
            if ((ty % p.stride) | (tx % p.stride)) return -1;
            ty /= p.stride;
            tx /= p.stride;
        }
        sy = ty;
        sx = tx;
    }
    if (sy < 0 || sy >= p.IH || sx < 0 || sx >= p.IW) return -1;
    return ((long long)n * p.IH + sy) * p.IW + sx;
}

__device__ __forceinline__ float load_src_elem(const ConvP &p, long long pix, int c) {
    if (p.src_dtype == NASB_F32_NCHW) {
        long long hw = (long long)p.IH * p.IW;
        long long n = pix / hw, r = pix - n * hw;
        return reinterpret_cast<const float *>(p.src[0])[(n * p.sc[0] + c) * hw + r];
    }
    int s = 0, cc = c;
    if (c >= p.sc[0]) {
        s = 1;
        cc = c - p.sc[0];
    }
    long long off = pix * p.scs[s] + cc;
    if (p.src_dtype == NASB_BF16) return __bfloat162float(reinterpret_cast<const bf16 *>(p.src[s])[off]);
    return reinterpret_cast<const float *>(p.src[s])[off];
}

__device__ __forceinline__ float prologue(const ConvP &p, float v, int c) {
    if (p.in_scale) v = v * p.in_scale[c] + p.in_shift[c];
    if (p.in_relu) v = fmaxf(v, 0.f);
    return v;
}

// NV consecutive reduction elements k..k+NV-1 of GEMM row (n,ry,rx) -> fp32 registers.
template <int NV>
__device__ __forceinline__ void load_a(const ConvP &p, bool row_ok, int n, int ry, int rx, int k, int K,
                                       float (&a)[NV]) {
#pragma unroll
    for (int j = 0; j < NV; ++j) a[j] = 0.f;
    if (!row_ok || k >= K) return;
    if (p.vec8) {
        int tap = k / p.Kc, c = k - tap * p.Kc;
        int ky = tap / p.ks, kx = tap - ky * p.ks;
        long long pix = src_pixel(p, n, ry, rx, ky, kx);
        if (pix < 0) return;
        int s = 0, cc = c;
        if (c >= p.sc[0]) {
            s = 1;
            cc = c - p.sc[0];
        }
        long long off = pix * p.scs[s] + cc;
        if (p.src_dtype == NASB_BF16) {
            const bf16 *q = reinterpret_cast<const bf16 *>(p.src[s]) + off;
            if constexpr (NV == 8) {
                load_vec<bf16, 8>(q, a);
            } else {
                load_vec<bf16, NV>(q, a);
            }
        } else {
            const float *q = reinterpret_cast<const float *>(p.src[s]) + off;
#pragma unroll
            for (int j = 0; j < NV; j += 4) {
                float4 v = *reinterpret_cast<const float4 *>(q + j);
                a[j] = v.x;
                a[j + 1] = v.y;
                a[j + 2] = v.z;
                a[j + 3] = v.w;
            }
        }
        if (p.in_scale || p.in_relu) {
#pragma unroll
            for (int j = 0; j < NV; ++j) a[j] = prologue(p, a[j], c + j);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            int kk = k + j;
            if (kk < K) {
                int tap = kk / p.Kc, c = kk - tap * p.Kc;
                int ky = tap / p.ks, kx = tap - ky * p.ks;
                long long pix = src_pixel(p, n, ry, rx, ky, kx);
                if (pix >= 0) a[j] = prologue(p, load_src_elem(p, pix, c), c);
            }
        }
    }
}

// B[n][k]: forward n = co, k = (tap,ci) ; dgrad n = ci, k = (tap,co).  Weight [C_out][C_in][ks*ks].
__device__ __forceinline__ float load_w(const ConvP &p, int n, int k, int K) {
    if (n >= p.Nc || k >= K) return 0.f;
    int tap = k / p.Kc, c = k - tap * p.Kc;
    int KK = p.ks * p.ks;
    long long idx = (p.mode == 0) ? ((long long)n * p.w_cin + c) * KK + tap : ((long long)c * p.w_cin + n) * KK + tap;
    return p.w[idx];
}

constexpr int BM = 128, BK = 16, NT = 256;

template <int BN>
__global__ void __launch_bounds__(NT) conv_gemm_kernel(const ConvP p) {
    pdl_sync();
    constexpr int TX = BN / 4;      // thread columns
    constexpr int TY = NT / TX;     // thread rows
    constexpr int RM = BM / TY;     // output rows per thread (8 for BN=64, 4 for BN=32)
    constexpr int BNP = BN + 4;     // padded B row (keeps 16 B alignment, breaks the store conflicts)
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Bs[BK][BNP];

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const long long M = (long long)p.N * p.OH * p.OW;
    const int K = p.ks * p.ks * p.Kc;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // A loader: thread -> (row = tid/2, 8 consecutive k starting at (tid&1)*8)
    const int a_row = tid >> 1, a_k = (tid & 1) * 8;
    const long long am = m0 + a_row;
    const bool a_ok = am < M;
    int an = 0, ay = 0, ax = 0;
    if (a_ok) {
        long long hw = (long long)p.OH * p.OW;
        an = (int)(am / hw);
        int r = (int)(am - (long long)an * hw);
        ay = r / p.OW;
        ax = r - ay * p.OW;
    }
    // B loader: BK*BN elements, k fastest
    constexpr int B_PER = (BK * BN) / NT;

    float acc[RM][4];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float a_reg[8], b_reg[B_PER];
    load_a<8>(p, a_ok, an, ay, ax, a_k, K, a_reg);
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
        int idx = tid + i * NT;
        b_reg[i] = load_w(p, n0 + idx / BK, idx % BK, K);
    }

    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int j = 0; j < 8; ++j) As[a_k + j][a_row] = a_reg[j];
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int idx = tid + i * NT;
            Bs[idx % BK][idx / BK] = b_reg[i];
        }
        __syncthreads();
        if (k0 + BK < K) {  // prefetch the next tile while this one is consumed
            load_a<8>(p, a_ok, an, ay, ax, k0 + BK + a_k, K, a_reg);
#pragma unroll
            for (int i = 0; i < B_PER; ++i) {
                int idx = tid + i * NT;
                b_reg[i] = load_w(p, n0 + idx / BK, k0 + BK + idx % BK, K);
            }
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            float a[RM];
#pragma unroll
            for (int g = 0; g < RM / 4; ++g) {
                float4 v = *reinterpret_cast<const float4 *>(&As[kk][g * (BM / (RM / 4)) + ty * 4]);
                a[g * 4] = v.x;
                a[g * 4 + 1] = v.y;
                a[g * 4 + 2] = v.z;
                a[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < RM; ++i) {
                acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
                acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
                acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
                acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
            }
        }
        __syncthreads();
    }

    // epilogue
    const int nb = n0 + tx * 4;
    if (nb >= p.Nc) return;
    float sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int n = nb + j;
        sc[j] = (p.out_scale && n < p.Nc) ? p.out_scale[n] : 1.f;
        sh[j] = (p.out_shift && n < p.Nc) ? p.out_shift[n] : 0.f;
    }
    // destination slice of this 4-column group (split outputs are used by the data gradient of a concat)
    int d = 0, nd = nb;
    if (nb >= p.oc[0]) {
        d = 1;
        nd = nb - p.oc[0];
    }
    const bool full4 = p.vec4o && (nb + 3 < p.Nc);
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int g = i / 4;
        long long m = m0 + g * (BM / (RM / 4)) + ty * 4 + (i % 4);
        if (m >= M) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = apply_act(acc[i][j] * sc[j] + sh[j], p.act);
        if (p.res) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (nb + j < p.Nc) {
                    long long ro = m * p.res_cs + nb + j;
                    v[j] += (p.out_dtype == NASB_BF16) ? __bfloat162float(reinterpret_cast<const bf16 *>(p.res)[ro])
                                                       : reinterpret_cast<const float *>(p.res)[ro];
                }
            }
        }
        if (full4) {
            long long off = m * p.ocs[d] + nd;
            if (p.out_dtype == NASB_BF16)
                store_vec<bf16, 4>(reinterpret_cast<bf16 *>(p.out[d]) + off, v);
            else
                store_vec<float, 4>(reinterpret_cast<float *>(p.out[d]) + off, v);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int n = nb + j;
                if (n >= p.Nc) continue;
                int dd = 0, nn = n;
                if (n >= p.oc[0]) {
                    dd = 1;
                    nn = n - p.oc[0];
                }
                long long off = m * p.ocs[dd] + nn;
                if (p.out_dtype == NASB_BF16)
                    reinterpret_cast<bf16 *>(p.out[dd])[off] = __float2bfloat16_rn(v[j]);
                else
                    reinterpret_cast<float *>(p.out[dd])[off] = v[j];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[co][ci][tap] += sum_p dz[p][co] * pro(x)[src(p,tap)][ci].  CTA tile: 64 (co) x 64 (k) outputs, reduction
// over a slab of pixels in steps of 16; fp32 atomics merge the slabs.
constexpr int WB = 64, WP = 16;

struct WgradP {
    ConvP c;  // forward-mode geometry + sources (x) ; c.Nc = C_out
    const void *dz;
    int dz_cs, dz_dtype;
    float *dw;
    long long rows_per_cta;
    int dz_vec4;
};

__global__ void __launch_bounds__(NT) conv_wgrad_kernel(const WgradP q) {
    pdl_sync();
    const ConvP &p = q.c;
    __shared__ __align__(16) float Zs[WP][WB];
    __shared__ __align__(16) float Xs[WP][WB];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int K = p.ks * p.ks * p.Kc;
    const int tiles_k = cdiv(K, WB);
    const int co0 = (blockIdx.x / tiles_k) * WB, k0 = (blockIdx.x % tiles_k) * WB;
    const long long M = (long long)p.N * p.OH * p.OW;
    const long long r_begin = (long long)blockIdx.y * q.rows_per_cta;
    const long long r_end = r_begin + q.rows_per_cta < M ? r_begin + q.rows_per_cta : M;

    const int l_p = tid >> 4, l_c = (tid & 15) * 4;  // loader: pixel-in-step, 4 consecutive columns
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (long long r0 = r_begin; r0 < r_end; r0 += WP) {
        long long m = r0 + l_p;
        bool ok = m < r_end;
        float z[4] = {0.f, 0.f, 0.f, 0.f}, x[4];
        if (ok) {
            int co = co0 + l_c;
            long long off = m * q.dz_cs + co;
            if (q.dz_vec4 && co + 3 < p.Nc) {
                if (q.dz_dtype == NASB_BF16)
                    load_vec<bf16, 4>(reinterpret_cast<const bf16 *>(q.dz) + off, z);
                else
                    load_vec<float, 4>(reinterpret_cast<const float *>(q.dz) + off, z);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (co + j < p.Nc)
                        z[j] = (q.dz_dtype == NASB_BF16)
                                   ? __bfloat162float(reinterpret_cast<const bf16 *>(q.dz)[off + j])
                                   : reinterpret_cast<const float *>(q.dz)[off + j];
            }
        }
        int n = 0, oy = 0, ox = 0;
        if (ok) {
            long long hw = (long long)p.OH * p.OW;
            n = (int)(m / hw);
            int r = (int)(m - (long long)n * hw);
            oy = r / p.OW;
            ox = r - oy * p.OW;
        }
        load_a<4>(p, ok, n, oy, ox, k0 + l_c, K, x);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            Zs[l_p][l_c + j] = z[j];
            Xs[l_p][l_c + j] = x[j];
        }
        __syncthreads();
#pragma unroll
        for (int pp = 0; pp < WP; ++pp) {
            float4 zz = *reinterpret_cast<const float4 *>(&Zs[pp][ty * 4]);
            float4 xx = *reinterpret_cast<const float4 *>(&Xs[pp][tx * 4]);
            float zv[4] = {zz.x, zz.y, zz.z, zz.w}, xv[4] = {xx.x, xx.y, xx.z, xx.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(zv[i], xv[j], acc[i][j]);
        }
    }
    const int KK = p.ks * p.ks;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int co = co0 + ty * 4 + i;
        if (co >= p.Nc) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + tx * 4 + j;
            if (k >= K) continue;
            int tap = k / p.Kc, ci = k - tap * p.Kc;
            atomicAdd(&q.dw[((long long)co * p.w_cin + ci) * KK + tap], acc[i][j]);
        }
    }
}

static bool fill_sources(ConvP &p, const NasbTensor *x0, const NasbTensor *x1) {
    p.src[0] = x0->ptr;
    p.sc[0] = x0->c;
    p.scs[0] = x0->cstride;
    p.src[1] = nullptr;
    p.sc[1] = 0;
    p.scs[1] = 0;
    p.src_dtype = x0->dtype;
    if (x1) {
        if (x1->n != x0->n || x1->h != x0->h || x1->w != x0->w || x1->dtype != x0->dtype) return false;
        if (x0->dtype == NASB_F32_NCHW) return false;
        p.src[1] = x1->ptr;
        p.sc[1] = x1->c;
        p.scs[1] = x1->cstride;
    }
    p.N = x0->n;
    p.IH = x0->h;
    p.IW = x0->w;
    p.Kc = x0->c + (x1 ? x1->c : 0);
    bool v = x0->dtype != NASB_F32_NCHW && vec_ok(*x0, 8) && (!x1 || vec_ok(*x1, 8));
    p.vec8 = v ? 1 : 0;
    return true;
}

}  // namespace nasb

using namespace nasb;

extern "C" int nasb_conv_fwd(const NasbTensor *x0, const NasbTensor *x1, const float *weight, int ks, int stride,
                             int dil, int pad, const float *in_scale, const float *in_shift, int in_relu,
                             const float *out_scale, const float *out_shift, int act, const NasbTensor *res,
                             const NasbTensor *out, void *stream) {
    if (!x0 || !out || !weight || (ks != 1 && ks != 3) || stride < 1) return NASB_ERR_BAD_ARG;
    ConvP p{};
    if (!fill_sources(p, x0, x1)) return NASB_ERR_BAD_ARG;
    int eh = (x0->h + 2 * pad - dil * (ks - 1) - 1) / stride + 1, ew = (x0->w + 2 * pad - dil * (ks - 1) - 1) / stride + 1;
    if (out->n != x0->n || out->h != eh || out->w != ew || out->dtype == NASB_F32_NCHW) return NASB_ERR_BAD_ARG;
    p.OH = out->h;
    p.OW = out->w;
    p.Nc = out->c;
    p.ks = ks;
    p.stride = stride;
    p.dil = dil;
    p.pad = pad;
    p.mode = 0;
    p.w = weight;
    p.w_cin = p.Kc;
    p.in_scale = in_scale;
    p.in_shift = in_shift;
    p.in_relu = in_relu;
    p.out_scale = out_scale;
    p.out_shift = out_shift;
    p.act = act;
    p.res = res ? res->ptr : nullptr;
    p.res_cs = res ? res->cstride : 0;
    if (res && (res->dtype != out->dtype || res->c != out->c)) return NASB_ERR_BAD_ARG;
    p.out[0] = out->ptr;
    p.out[1] = nullptr;
    p.oc[0] = out->c;
    p.oc[1] = 0;
    p.ocs[0] = out->cstride;
    p.ocs[1] = 0;
    p.out_dtype = out->dtype;
    p.vec4o = vec_ok(*out, 4) ? 1 : 0;
    long long M = npix(*out);
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (p.Nc <= 32) {
        dim3 grid(cdiv(M, BM), cdiv(p.Nc, 32));
        nasb::launch_pdl((conv_gemm_kernel<32>), dim3(grid), dim3(NT), 0, (cudaStream_t)(st), p);
    } else {
        dim3 grid(cdiv(M, BM), cdiv(p.Nc, 64));
        nasb::launch_pdl((conv_gemm_kernel<64>), dim3(grid), dim3(NT), 0, (cudaStream_t)(st), p);
    }
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_conv_dgrad(const NasbTensor *dz, const float *weight, int ks, int stride, int dil, int pad,
                               const NasbTensor *dx0, const NasbTensor *dx1, void *stream) {
    if (!dz || !dx0 || !weight || (ks != 1 && ks != 3)) return NASB_ERR_BAD_ARG;
    if (dz->dtype == NASB_F32_NCHW) return NASB_ERR_BAD_ARG;
    ConvP p{};
    if (!fill_sources(p, dz, nullptr)) return NASB_ERR_BAD_ARG;
    p.OH = dx0->h;
    p.OW = dx0->w;
    p.Nc = dx0->c + (dx1 ? dx1->c : 0);
    p.ks = ks;
    p.stride = stride;
    p.dil = dil;
    p.pad = pad;
    p.mode = 1;
    p.w = weight;
    p.w_cin = p.Nc;
    p.act = NASB_ACT_NONE;
    p.out[0] = dx0->ptr;
    p.oc[0] = dx0->c;
    p.ocs[0] = dx0->cstride;
    p.out[1] = dx1 ? dx1->ptr : nullptr;
    p.oc[1] = dx1 ? dx1->c : 0;
    p.ocs[1] = dx1 ? dx1->cstride : 0;
    p.out_dtype = dx0->dtype;
    if (dx1 && (dx1->dtype != dx0->dtype || dx1->n != dx0->n || dx1->h != dx0->h || dx1->w != dx0->w))
        return NASB_ERR_BAD_ARG;
    p.vec4o = (vec_ok(*dx0, 4) && (!dx1 || vec_ok(*dx1, 4))) ? 1 : 0;
    long long M = npix(*dx0);
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (p.Nc <= 32) {
        dim3 grid(cdiv(M, BM), cdiv(p.Nc, 32));
        nasb::launch_pdl((conv_gemm_kernel<32>), dim3(grid), dim3(NT), 0, (cudaStream_t)(st), p);
    } else {
        dim3 grid(cdiv(M, BM), cdiv(p.Nc, 64));
        nasb::launch_pdl((conv_gemm_kernel<64>), dim3(grid), dim3(NT), 0, (cudaStream_t)(st), p);
    }
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_conv_wgrad(const NasbTensor *x0, const NasbTensor *x1, const float *in_scale,
                               const float *in_shift, int in_relu, const NasbTensor *dz, int ks, int stride, int dil,
                               int pad, float *dweight, void *stream) {
    if (!x0 || !dz || !dweight || (ks != 1 && ks != 3)) return NASB_ERR_BAD_ARG;
    if (dz->dtype == NASB_F32_NCHW) return NASB_ERR_BAD_ARG;
    WgradP q{};
    ConvP &p = q.c;
    if (!fill_sources(p, x0, x1)) return NASB_ERR_BAD_ARG;
    // 4-wide loads need 4-aligned sources; fall back to the scalar loader otherwise
    p.vec8 = (x0->dtype != NASB_F32_NCHW && vec_ok(*x0, 4) && (!x1 || vec_ok(*x1, 4))) ? 1 : 0;
    p.OH = dz->h;
    p.OW = dz->w;
    p.Nc = dz->c;
    p.ks = ks;
    p.stride = stride;
    p.dil = dil;
    p.pad = pad;
    p.mode = 0;
    p.w_cin = p.Kc;
    p.in_scale = in_scale;
    p.in_shift = in_shift;
    p.in_relu = in_relu;
    q.dz = dz->ptr;
    q.dz_cs = dz->cstride;
    q.dz_dtype = dz->dtype;
    q.dz_vec4 = vec_ok(*dz, 4) ? 1 : 0;
    q.dw = dweight;
    long long M = npix(*dz);
    if (M == 0) return 0;
    int K = ks * ks * p.Kc;
    int tiles = cdiv(p.Nc, WB) * cdiv(K, WB);
    // enough slabs to fill the GPU a few times over, but at least 256 rows per slab
    long long want = (long long)NASB_SM_COUNT * 8 / tiles;
    if (want < 1) want = 1;
    long long rows = (M + want - 1) / want;
    if (rows < 256) rows = 256;
    rows = (rows + WP - 1) / WP * WP;
    q.rows_per_cta = rows;
    dim3 grid(tiles, cdiv(M, rows));
    nasb::launch_pdl((conv_wgrad_kernel), dim3(grid), dim3(NT), 0, (cudaStream_t)((cudaStream_t)stream), q);
    NASB_CHECK_LAUNCH();
    return 0;
}
