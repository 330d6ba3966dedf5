// spatial.cu -- HBM-bound NHWC data-movement kernels: 3x3 pooling, bilinear resize fused with the aggregation that
// follows it, channel tiling (Skip/Zero), global-average-pool pieces, and the small per-channel reductions.
// One thread = one pixel x one V-channel vector; neighbouring threads own neighbouring channel vectors.
#include <stdlib.h>

#include "common.cuh"

namespace nasb {

static inline int grid_for(long long total, int threads = 256) {
    long long b = (total + threads - 1) / threads;
    long long cap = (long long)NASB_SM_COUNT * 32;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

#define NASB_GRID_STRIDE(idx, total) \
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (total); idx += (long long)gridDim.x * blockDim.x)

// ------------------------------------------------------------------------------------------------ pooling
// (A variant moving the eight arg-max bytes of a channel vector as one 8-byte access was measured on a B200 in round 2:
// forward unchanged, backward 4-9 % SLOWER -- the compiler already merges the byte accesses -- and was removed.)
template <typename T, int V>
__global__ void __launch_bounds__(256) pool_fwd_kernel(const T *x, int x_cs, T *out, int out_cs, uint8_t *argmax, int N,
                                                       int IH, int IW, int OH, int OW, int C, int stride, int mode) {
    pdl_sync();
    const int CV = C / V;
    const long long total = (long long)N * OH * OW * CV;
    NASB_GRID_STRIDE(idx, total) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        int ox = (int)(pix % OW);
        long long t = pix / OW;
        int oy = (int)(t % OH);
        int n = (int)(t / OH);
        float acc[V];
        int arg[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            acc[j] = mode == NASB_POOL_MAX ? -INFINITY : 0.f;
            arg[j] = 0;
        }
        int cnt = 0;
        bool first = true;
        for (int ky = 0; ky < 3; ++ky) {
            int sy = oy * stride - 1 + ky;
            if (sy < 0 || sy >= IH) continue;
            for (int kx = 0; kx < 3; ++kx) {
                int sx = ox * stride - 1 + kx;
                if (sx < 0 || sx >= IW) continue;
                float v[V];
                load_vec<T, V>(x + (((long long)n * IH + sy) * IW + sx) * x_cs + cv * V, v);
                ++cnt;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    if (mode == NASB_POOL_MAX) {
                        // first maximal element wins (strict >), NaN propagates like ATen's max_pool2d
                        if (first || v[j] > acc[j] || v[j] != v[j]) {
                            acc[j] = v[j];
                            arg[j] = ky * 3 + kx;
                        }
                    } else {
                        acc[j] += v[j];
                    }
                }
                first = false;
            }
        }
        if (mode == NASB_POOL_AVG) {
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = acc[j] / (float)cnt;
        }
        store_vec<T, V>(out + pix * out_cs + cv * V, acc);
        if (argmax) {
#pragma unroll
            for (int j = 0; j < V; ++j) argmax[pix * C + cv * V + j] = (uint8_t)arg[j];
        }
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) pool_bwd_kernel(const T *dy, int dy_cs, T *dx, int dx_cs, const uint8_t *argmax,
                                                       int N, int IH, int IW, int OH, int OW, int C, int stride, int mode) {
    pdl_sync();
    const int CV = C / V;
    const long long total = (long long)N * IH * IW * CV;
    NASB_GRID_STRIDE(idx, total) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        int ix = (int)(pix % IW);
        long long t = pix / IW;
        int iy = (int)(t % IH);
        int n = (int)(t / IH);
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = 0.f;
        for (int ky = 0; ky < 3; ++ky) {
            int ty = iy + 1 - ky;
            if (ty < 0 || ty % stride) continue;
            int oy = ty / stride;
            if (oy >= OH) continue;
            for (int kx = 0; kx < 3; ++kx) {
                int tx = ix + 1 - kx;
                if (tx < 0 || tx % stride) continue;
                int ox = tx / stride;
                if (ox >= OW) continue;
                long long opix = ((long long)n * OH + oy) * OW + ox;
                float g[V];
                load_vec<T, V>(dy + opix * dy_cs + cv * V, g);
                if (mode == NASB_POOL_MAX) {
#pragma unroll
                    for (int j = 0; j < V; ++j)
                        if (argmax[opix * C + cv * V + j] == ky * 3 + kx) acc[j] += g[j];
                } else {
                    int y0 = oy * stride - 1, x0 = ox * stride - 1;
                    int cy = min(y0 + 3, IH) - max(y0, 0), cx = min(x0 + 3, IW) - max(x0, 0);
                    float inv = 1.f / (float)(cy * cx);
#pragma unroll
                    for (int j = 0; j < V; ++j) acc[j] += g[j] * inv;
                }
            }
        }
        store_vec<T, V>(dx + pix * dx_cs + cv * V, acc);
    }
}

// ------------------------------------------------------------------------------------------------ bilinear (+ axpby)
template <typename T, typename TY, int V>
__global__ void __launch_bounds__(256) resize_axpby_kernel(const T *x, int x_cs, int IH, int IW, const float *sa,
                                                           const TY *y, int y_cs, const float *sb, T *out, int out_cs,
                                                           int N, int OH, int OW, int C, float rh, float rw, int identity,
                                                           int relu) {
    pdl_sync();
    const int CV = C / V;
    const long long total = (long long)N * OH * OW * CV;
    NASB_GRID_STRIDE(idx, total) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        int ox = (int)(pix % OW);
        long long t = pix / OW;
        int oy = (int)(t % OH);
        int n = (int)(t / OH);
        const int c0 = cv * V;
        float r[V];
        if (identity) {
            load_vec<T, V>(x + pix * x_cs + c0, r);
        } else {
            Lerp ly = lerp_coord(oy, rh, IH), lx = lerp_coord(ox, rw, IW);
            const T *base = x + (long long)n * IH * IW * x_cs + c0;
            float v00[V], v01[V], v10[V], v11[V];
            load_vec<T, V>(base + ((long long)ly.i0 * IW + lx.i0) * x_cs, v00);
            load_vec<T, V>(base + ((long long)ly.i0 * IW + lx.i1) * x_cs, v01);
            load_vec<T, V>(base + ((long long)ly.i1 * IW + lx.i0) * x_cs, v10);
            load_vec<T, V>(base + ((long long)ly.i1 * IW + lx.i1) * x_cs, v11);
#pragma unroll
            for (int j = 0; j < V; ++j)
                r[j] = ly.l0 * (lx.l0 * v00[j] + lx.l1 * v01[j]) + ly.l1 * (lx.l0 * v10[j] + lx.l1 * v11[j]);
        }
        if (sa) {
#pragma unroll
            for (int j = 0; j < V; ++j) r[j] *= sa[c0 + j];
        }
        if (y) {
            float yv[V];
            load_vec<TY, V>(y + pix * y_cs + c0, yv);
#pragma unroll
            for (int j = 0; j < V; ++j) r[j] += (sb ? sb[c0 + j] : 1.f) * yv[j];
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < V; ++j) r[j] = fmaxf(r[j], 0.f);
        }
        store_vec<T, V>(out + pix * out_cs + c0, r);
    }
}

// contributions of input index i along one axis: candidate outputs [lo,hi]
__device__ __forceinline__ void adj_range(int i, float scale, int out_size, int &lo, int &hi) {
    float inv = 1.f / scale;
    lo = (int)floorf(((float)i - 1.f + 0.5f) * inv - 0.5f) - 1;
    hi = (int)ceilf(((float)i + 1.f + 0.5f) * inv - 0.5f) + 1;
    if (lo < 0) lo = 0;
    if (hi > out_size - 1) hi = out_size - 1;
}
__device__ __forceinline__ float adj_weight(int o, int i, float scale, int in_size) {
    Lerp l = lerp_coord(o, scale, in_size);
    float w = 0.f;
    if (l.i0 == i) w += l.l0;
    if (l.i1 == i) w += l.l1;
    return w;
}

// dx[n,iy,ix,c] = sum_{oy,ox} wy(oy,iy) wx(ox,ix) * sa[c] * dz[n,oy,ox,c]     (deterministic gather)
template <typename T, int V>
__global__ void __launch_bounds__(256) resize_bwd_kernel(const T *dz, int dz_cs, int OH, int OW, const float *sa, T *dx,
                                                         int dx_cs, int N, int IH, int IW, int C, float rh, float rw,
                                                         int identity) {
    pdl_sync();
    const int CV = C / V;
    const long long total = (long long)N * IH * IW * CV;
    NASB_GRID_STRIDE(idx, total) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        int ix = (int)(pix % IW);
        long long t = pix / IW;
        int iy = (int)(t % IH);
        int n = (int)(t / IH);
        const int c0 = cv * V;
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = 0.f;
        if (identity) {
            load_vec<T, V>(dz + pix * dz_cs + c0, acc);
        } else {
            int ylo, yhi, xlo, xhi;
            adj_range(iy, rh, OH, ylo, yhi);
            adj_range(ix, rw, OW, xlo, xhi);
            // The tent weights are separable: the column weights of this thread's window are computed once (<= RB_MAXT
            // candidates for up-sampling factors up to 8) instead of once per (row, column) pair -- the x8 gradient was
            // bound by that arithmetic (0.8 TB/s), not by its loads.
            constexpr int RB_MAXT = 24;
            const int nx = xhi - xlo + 1;
            if (nx <= RB_MAXT) {
                float wxs[RB_MAXT];
#pragma unroll
                for (int k = 0; k < RB_MAXT; ++k) wxs[k] = k < nx ? adj_weight(xlo + k, ix, rw, IW) : 0.f;
                int k0 = 0, k1 = nx - 1;  // trim the zero-weight fringe of the candidate range
                while (k0 < k1 && wxs[k0] == 0.f) ++k0;
                while (k1 > k0 && wxs[k1] == 0.f) --k1;
                // Four taps per step with unconditional (clamped) loads: a thread keeps four 16-byte loads in flight instead
                // of one -- with one load per iteration the x8 gradient was bound by 256 serialised L2 round trips per thread.
                for (int oy = ylo; oy <= yhi; ++oy) {
                    const float wy = adj_weight(oy, iy, rh, IH);
                    if (wy == 0.f) continue;
                    const T *rowp = dz + (((long long)n * OH + oy) * OW + xlo) * dz_cs + c0;
                    for (int k = k0; k <= k1; k += 4) {
                        float g[4][V], w[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int kk = k + u <= k1 ? k + u : k1;
                            w[u] = k + u <= k1 ? wy * wxs[kk] : 0.f;
                            load_vec<T, V>(rowp + (long long)kk * dz_cs, g[u]);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int j = 0; j < V; ++j) acc[j] = fmaf(w[u], g[u][j], acc[j]);
                    }
                }
            } else {
                for (int oy = ylo; oy <= yhi; ++oy) {
                    float wy = adj_weight(oy, iy, rh, IH);
                    if (wy == 0.f) continue;
                    for (int ox = xlo; ox <= xhi; ++ox) {
                        float wx = adj_weight(ox, ix, rw, IW);
                        if (wx == 0.f) continue;
                        float g[V];
                        load_vec<T, V>(dz + (((long long)n * OH + oy) * OW + ox) * dz_cs + c0, g);
                        float w = wy * wx;
#pragma unroll
                        for (int j = 0; j < V; ++j) acc[j] = fmaf(w, g[j], acc[j]);
                    }
                }
            }
        }
        if (sa) {
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] *= sa[c0 + j];
        }
        store_vec<T, V>(dx + pix * dx_cs + c0, acc);
    }
}

// dsa[c] += sum dz*resize(x) ; dsb[c] += sum dz*y.   block (32 channel lanes x 8 pixel lanes), slab per CTA.
template <typename T>
__global__ void __launch_bounds__(256) axpby_bwd_params_kernel(const T *dz, int dz_cs, const T *x, int x_cs, int IH, int IW,
                                                               const T *y, int y_cs, float *dsa, float *dsb, int N, int OH,
                                                               int OW, int C, float rh, float rw, int identity,
                                                               long long rows_per_cta) {
    pdl_sync();
    __shared__ float ra[256], rb[256];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const long long M = (long long)N * OH * OW;
    const long long r0 = (long long)blockIdx.y * rows_per_cta, r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
    float a = 0.f, b = 0.f;
    if (c < C) {
        for (long long m = r0 + threadIdx.y; m < r1; m += 8) {
            float g = to_f(dz[m * dz_cs + c]);
            float xv;
            if (identity) {
                xv = to_f(x[m * x_cs + c]);
            } else {
                int ox = (int)(m % OW);
                long long t = m / OW;
                int oy = (int)(t % OH);
                int n = (int)(t / OH);
                Lerp ly = lerp_coord(oy, rh, IH), lx = lerp_coord(ox, rw, IW);
                const T *base = x + (long long)n * IH * IW * x_cs + c;
                xv = ly.l0 * (lx.l0 * to_f(base[((long long)ly.i0 * IW + lx.i0) * x_cs]) +
                              lx.l1 * to_f(base[((long long)ly.i0 * IW + lx.i1) * x_cs])) +
                     ly.l1 * (lx.l0 * to_f(base[((long long)ly.i1 * IW + lx.i0) * x_cs]) +
                              lx.l1 * to_f(base[((long long)ly.i1 * IW + lx.i1) * x_cs]));
            }
            a = fmaf(g, xv, a);
            if (y) b = fmaf(g, to_f(y[m * y_cs + c]), b);
        }
    }
    const int lin = threadIdx.y * 32 + threadIdx.x;
    ra[lin] = a;
    rb[lin] = b;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float sa_ = 0.f, sb_ = 0.f;
        for (int i = 0; i < 8; ++i) {
            sa_ += ra[i * 32 + threadIdx.x];
            sb_ += rb[i * 32 + threadIdx.x];
        }
        if (dsa) atomicAdd(&dsa[c], sa_);
        if (dsb && y) atomicAdd(&dsb[c], sb_);
    }
}

// ------------------------------------------------------------------------------------------------ copies / tiling
template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(256) scale_copy_kernel(const TI *x, int x_cs, const float *s, int relu, TO *out,
                                                         int out_cs, long long P, int C) {
    pdl_sync();
    const int CV = C / V;
    const long long total = P * CV;
    NASB_GRID_STRIDE(idx, total) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        float v[V];
        load_vec<TI, V>(x + pix * x_cs + cv * V, v);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            if (s) v[j] *= s[cv * V + j];
            if (relu) v[j] = fmaxf(v[j], 0.f);
        }
        store_vec<TO, V>(out + pix * out_cs + cv * V, v);
    }
}

// out[n,oy,ox,r*C+c] = scale * x[n,oy*stride,ox*stride,c]
template <typename T, int V>
__global__ void __launch_bounds__(256) channel_tile_kernel(const T *x, int x_cs, int IH, int IW, T *out, int out_cs, int N,
                                                           int OH, int OW, int C, int R, int stride, float scale) {
    pdl_sync();
    const int CV = C / V;
    const long long total = (long long)N * OH * OW * CV;
    NASB_GRID_STRIDE(idx, total) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        int ox = (int)(pix % OW);
        long long t = pix / OW;
        int oy = (int)(t % OH);
        int n = (int)(t / OH);
        float v[V];
        load_vec<T, V>(x + (((long long)n * IH + oy * stride) * IW + ox * stride) * x_cs + cv * V, v);
        // scale == 0 must give exact zeros even for inf/nan inputs?  The reference computes x.mul(0.0) (nan stays nan);
        // keep the multiplication.
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] *= scale;
        for (int r = 0; r < R; ++r) store_vec<T, V>(out + pix * out_cs + r * C + cv * V, v);
    }
}

// dx[n,iy,ix,c] = (iy%stride==0 && ix%stride==0) ? scale * sum_r dz[n,iy/stride,ix/stride,r*C+c] : 0
template <typename T, int V>
__global__ void __launch_bounds__(256) channel_tile_bwd_kernel(const T *dz, int dz_cs, int OH, int OW, T *dx, int dx_cs,
                                                               int N, int IH, int IW, int C, int R, int stride, float scale) {
    pdl_sync();
    const int CV = C / V;
    const long long total = (long long)N * IH * IW * CV;
    NASB_GRID_STRIDE(idx, total) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        int ix = (int)(pix % IW);
        long long t = pix / IW;
        int iy = (int)(t % IH);
        int n = (int)(t / IH);
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = 0.f;
        if (iy % stride == 0 && ix % stride == 0 && iy / stride < OH && ix / stride < OW) {
            long long opix = ((long long)n * OH + iy / stride) * OW + ix / stride;
            for (int r = 0; r < R; ++r) {
                float g[V];
                load_vec<T, V>(dz + opix * dz_cs + r * C + cv * V, g);
#pragma unroll
                for (int j = 0; j < V; ++j) acc[j] += g[j];
            }
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] *= scale;
        }
        store_vec<T, V>(dx + pix * dx_cs + cv * V, acc);
    }
}

// out[n,c] (+)= scale * sum_{pixels of image n} x     block (32 channel lanes x 8 pixel lanes); grid (cblocks, slabs, N)
template <typename T>
__global__ void __launch_bounds__(256) spatial_sum_kernel(const T *x, int x_cs, float *out, int HW, int C, float scale,
                                                          int rows_per_cta) {
    pdl_sync();
    __shared__ float red[256];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int n = blockIdx.z;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(r0 + rows_per_cta, HW);
    float a = 0.f;
    if (c < C)
        for (int m = r0 + threadIdx.y; m < r1; m += 8) a += to_f(x[((long long)n * HW + m) * x_cs + c]);
    red[threadIdx.y * 32 + threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i * 32 + threadIdx.x];
        atomicAdd(&out[(long long)n * C + c], s * scale);
    }
}

template <typename TV, typename T, int V>
__global__ void __launch_bounds__(256) spatial_bcast_kernel(const TV *v, int v_cs, float s, T *out, int out_cs, int N, int HW,
                                                            int C) {
    pdl_sync();
    const int CV = C / V;
    const long long total = (long long)N * HW * CV;
    NASB_GRID_STRIDE(idx, total) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        int n = (int)(pix / HW);
        float r[V];
#pragma unroll
        for (int j = 0; j < V; ++j) r[j] = s * to_f(v[(long long)n * v_cs + cv * V + j]);
        store_vec<T, V>(out + pix * out_cs + cv * V, r);
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) relu_bwd_kernel(const T *dy, int dy_cs, const T *y, int y_cs, T *dx, int dx_cs,
                                                       long long P, int C) {
    pdl_sync();
    const int CV = C / V;
    const long long total = P * CV;
    NASB_GRID_STRIDE(idx, total) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        float g[V], yy[V];
        load_vec<T, V>(dy + pix * dy_cs + cv * V, g);
        load_vec<T, V>(y + pix * y_cs + cv * V, yy);
#pragma unroll
        for (int j = 0; j < V; ++j) g[j] = yy[j] > 0.f ? g[j] : 0.f;
        store_vec<T, V>(dx + pix * dx_cs + cv * V, g);
    }
}

__global__ void sumsq_kernel(const float *x, long long n, float *out) {
    pdl_sync();
    __shared__ double sm[256];
    double a = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        a += (double)x[i] * (double)x[i];
    double s = block_sum<256>(a, sm);
    if (threadIdx.x == 0) atomicAdd(out, (float)s);
}

template <typename T>
constexpr int vw() {
    return 16 / sizeof(T);
}

}  // namespace nasb

using namespace nasb;

#define ST ((cudaStream_t)stream)

// dispatch on dtype (f32 / bf16) and vectorisability of every tensor involved
#define NASB_DISPATCH(dtype, ok, CALL)                 \
    do {                                               \
        if ((dtype) == NASB_BF16) {                    \
            typedef bf16 T;                            \
            if (ok) { constexpr int V = 8; CALL; }     \
            else { constexpr int V = 1; CALL; }        \
        } else {                                       \
            typedef float T;                           \
            if (ok) { constexpr int V = 4; CALL; }     \
            else { constexpr int V = 1; CALL; }        \
        }                                              \
    } while (0)

static inline int vfor(int dtype) { return dtype == NASB_BF16 ? 8 : 4; }
static inline bool same_nhw(const NasbTensor *a, const NasbTensor *b) { return a->n == b->n && a->h == b->h && a->w == b->w; }
static inline bool act_dtype(const NasbTensor *a) { return a->dtype == NASB_F32 || a->dtype == NASB_BF16; }

extern "C" int nasb_pool3x3_fwd(const NasbTensor *x, int mode, int stride, const NasbTensor *out, uint8_t *argmax,
                                void *stream) {
    if (!x || !out || !act_dtype(x) || x->dtype != out->dtype || x->c != out->c || x->n != out->n || stride < 1)
        return NASB_ERR_BAD_ARG;
    if (out->h != (x->h + 2 - 3) / stride + 1 || out->w != (x->w + 2 - 3) / stride + 1) return NASB_ERR_BAD_ARG;
    long long rows = npix(*out);
    if (rows == 0) return 0;
    bool ok = vec_ok(*x, vfor(x->dtype)) && vec_ok(*out, vfor(x->dtype));
    NASB_DISPATCH(x->dtype, ok, (nasb::launch_pdl((pool_fwd_kernel<T, V>), dim3(grid_for(rows * (x->c / V))), dim3(256), 0, (cudaStream_t)(ST), 
                                    (const T *)x->ptr, x->cstride, (T *)out->ptr, out->cstride, argmax, x->n, x->h, x->w,
                                    out->h, out->w, x->c, stride, mode)));
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_pool3x3_bwd(const NasbTensor *dy, int mode, int stride, const uint8_t *argmax, const NasbTensor *dx,
                                void *stream) {
    if (!dy || !dx || !act_dtype(dy) || dy->dtype != dx->dtype || dy->c != dx->c || dy->n != dx->n) return NASB_ERR_BAD_ARG;
    if (mode == NASB_POOL_MAX && !argmax) return NASB_ERR_BAD_ARG;
    long long rows = npix(*dx);
    if (rows == 0) return 0;
    bool ok = vec_ok(*dy, vfor(dy->dtype)) && vec_ok(*dx, vfor(dy->dtype));
    NASB_DISPATCH(dy->dtype, ok, (nasb::launch_pdl((pool_bwd_kernel<T, V>), dim3(grid_for(rows * (dx->c / V))), dim3(256), 0, (cudaStream_t)(ST), 
                                     (const T *)dy->ptr, dy->cstride, (T *)dx->ptr, dx->cstride, argmax, dx->n, dx->h, dx->w,
                                     dy->h, dy->w, dx->c, stride, mode)));
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_resize_axpby(const NasbTensor *x, const float *sa, const NasbTensor *y, const float *sb, int relu,
                                 const NasbTensor *out, void *stream) {
    if (!x || !out || !act_dtype(x) || x->dtype != out->dtype || x->c != out->c || x->n != out->n) return NASB_ERR_BAD_ARG;
    if (y && (!same_nhw(y, out) || y->c != out->c || !act_dtype(y))) return NASB_ERR_BAD_ARG;
    long long rows = npix(*out);
    if (rows == 0) return 0;
    int identity = (x->h == out->h && x->w == out->w) ? 1 : 0;
    float rh = (float)x->h / (float)out->h, rw = (float)x->w / (float)out->w;
    bool ok = vec_ok(*x, vfor(x->dtype)) && vec_ok(*out, vfor(x->dtype)) && (!y || vec_ok(*y, vfor(x->dtype)));
    if (y && y->dtype != x->dtype) {
        // mixed operand dtype (fp32 second operand with bf16 activations): scalar path only
        if (x->dtype == NASB_BF16)
            nasb::launch_pdl((resize_axpby_kernel<bf16, float, 1>), dim3(grid_for(rows * x->c)), dim3(256), 0, (cudaStream_t)(ST), 
                (const bf16 *)x->ptr, x->cstride, x->h, x->w, sa, (const float *)y->ptr, y->cstride, sb, (bf16 *)out->ptr,
                out->cstride, out->n, out->h, out->w, out->c, rh, rw, identity, relu);
        else
            nasb::launch_pdl((resize_axpby_kernel<float, bf16, 1>), dim3(grid_for(rows * x->c)), dim3(256), 0, (cudaStream_t)(ST), 
                (const float *)x->ptr, x->cstride, x->h, x->w, sa, (const bf16 *)y->ptr, y->cstride, sb, (float *)out->ptr,
                out->cstride, out->n, out->h, out->w, out->c, rh, rw, identity, relu);
    } else {
        NASB_DISPATCH(x->dtype, ok, (nasb::launch_pdl((resize_axpby_kernel<T, T, V>), dim3(grid_for(rows * (x->c / V))), dim3(256), 0, (cudaStream_t)(ST), 
                                        (const T *)x->ptr, x->cstride, x->h, x->w, sa, y ? (const T *)y->ptr : nullptr,
                                        y ? y->cstride : 0, sb, (T *)out->ptr, out->cstride, out->n, out->h, out->w, out->c,
                                        rh, rw, identity, relu)));
    }
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_resize_bwd(const NasbTensor *dz, const float *sa, const NasbTensor *dx, void *stream) {
    if (!dz || !dx || !act_dtype(dz) || dz->dtype != dx->dtype || dz->c != dx->c || dz->n != dx->n) return NASB_ERR_BAD_ARG;
    long long rows = npix(*dx);
    if (rows == 0) return 0;
    int identity = (dz->h == dx->h && dz->w == dx->w) ? 1 : 0;
    float rh = (float)dx->h / (float)dz->h, rw = (float)dx->w / (float)dz->w;
    bool ok = vec_ok(*dz, vfor(dz->dtype)) && vec_ok(*dx, vfor(dz->dtype));
    NASB_DISPATCH(dz->dtype, ok, (nasb::launch_pdl((resize_bwd_kernel<T, V>), dim3(grid_for(rows * (dx->c / V))), dim3(256), 0, (cudaStream_t)(ST), 
                                     (const T *)dz->ptr, dz->cstride, dz->h, dz->w, sa, (T *)dx->ptr, dx->cstride, dx->n,
                                     dx->h, dx->w, dx->c, rh, rw, identity)));
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_axpby_bwd_params(const NasbTensor *dz, const NasbTensor *x, const NasbTensor *y, float *dsa, float *dsb,
                                     void *workspace, void *stream) {
    (void)workspace;
    if (!dz || !x || !act_dtype(dz) || x->dtype != dz->dtype || x->c != dz->c || x->n != dz->n) return NASB_ERR_BAD_ARG;
    if (y && (y->dtype != dz->dtype || !same_nhw(y, dz) || y->c != dz->c)) return NASB_ERR_BAD_ARG;
    long long M = npix(*dz);
    if (M == 0) return 0;
    int identity = (x->h == dz->h && x->w == dz->w) ? 1 : 0;
    float rh = (float)x->h / (float)dz->h, rw = (float)x->w / (float)dz->w;
    int cblocks = cdiv(dz->c, 32);
    long long want = (long long)NASB_SM_COUNT * 8 / cblocks;
    if (want < 1) want = 1;
    long long rows = (M + want - 1) / want;
    if (rows < 64) rows = 64;
    dim3 grid(cblocks, cdiv(M, rows)), block(32, 8);
    if (dz->dtype == NASB_BF16)
        nasb::launch_pdl((axpby_bwd_params_kernel<bf16>), dim3(grid), dim3(block), 0, (cudaStream_t)(ST), (const bf16 *)dz->ptr, dz->cstride, (const bf16 *)x->ptr, x->cstride,
                                                              x->h, x->w, y ? (const bf16 *)y->ptr : nullptr, y ? y->cstride : 0,
                                                              dsa, dsb, dz->n, dz->h, dz->w, dz->c, rh, rw, identity, rows);
    else
        nasb::launch_pdl((axpby_bwd_params_kernel<float>), dim3(grid), dim3(block), 0, (cudaStream_t)(ST), (const float *)dz->ptr, dz->cstride, (const float *)x->ptr,
                                                               x->cstride, x->h, x->w, y ? (const float *)y->ptr : nullptr,
                                                               y ? y->cstride : 0, dsa, dsb, dz->n, dz->h, dz->w, dz->c, rh, rw,
                                                               identity, rows);
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_scale_copy(const NasbTensor *x, const float *s, int relu, const NasbTensor *out, void *stream) {
    if (!x || !out || !act_dtype(x) || !act_dtype(out) || !same_nhw(x, out) || x->c != out->c) return NASB_ERR_BAD_ARG;
    long long P = npix(*x);
    if (P == 0) return 0;
    int C = x->c;
    bool ok4 = vec_ok(*x, 4) && vec_ok(*out, 4), ok8 = vec_ok(*x, 8) && vec_ok(*out, 8);
#define SC(TI, TO, V)                                                                                                   \
    nasb::launch_pdl((scale_copy_kernel<TI, TO, V>), dim3(grid_for(P * (C / V))), dim3(256), 0, (cudaStream_t)(ST), (const TI *)x->ptr, x->cstride, s, relu, (TO *)out->ptr, \
                                                                          out->cstride, P, C)
    if (x->dtype == NASB_BF16 && out->dtype == NASB_BF16) {
        if (ok8) SC(bf16, bf16, 8); else SC(bf16, bf16, 1);
    } else if (x->dtype == NASB_F32 && out->dtype == NASB_F32) {
        if (ok4) SC(float, float, 4); else SC(float, float, 1);
    } else if (x->dtype == NASB_BF16) {
        if (ok4) SC(bf16, float, 4); else SC(bf16, float, 1);
    } else {
        if (ok4) SC(float, bf16, 4); else SC(float, bf16, 1);
    }
#undef SC
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_channel_tile(const NasbTensor *x, int stride, float scale, const NasbTensor *out, void *stream) {
    if (!x || !out || !act_dtype(x) || x->dtype != out->dtype || x->n != out->n || out->c % x->c || stride < 1)
        return NASB_ERR_BAD_ARG;
    if (out->h != (x->h + stride - 1) / stride || out->w != (x->w + stride - 1) / stride) return NASB_ERR_BAD_ARG;
    long long rows = npix(*out);
    if (rows == 0) return 0;
    int R = out->c / x->c;
    bool ok = vec_ok(*x, vfor(x->dtype)) && (out->cstride % vfor(x->dtype) == 0) &&
              ((uintptr_t)out->ptr % 16 == 0);
    NASB_DISPATCH(x->dtype, ok, (nasb::launch_pdl((channel_tile_kernel<T, V>), dim3(grid_for(rows * (x->c / V))), dim3(256), 0, (cudaStream_t)(ST), 
                                    (const T *)x->ptr, x->cstride, x->h, x->w, (T *)out->ptr, out->cstride, out->n, out->h,
                                    out->w, x->c, R, stride, scale)));
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_channel_tile_bwd(const NasbTensor *dz, int stride, float scale, const NasbTensor *dx, void *stream) {
    if (!dz || !dx || !act_dtype(dz) || dz->dtype != dx->dtype || dz->n != dx->n || dz->c % dx->c || stride < 1)
        return NASB_ERR_BAD_ARG;
    long long rows = npix(*dx);
    if (rows == 0) return 0;
    int R = dz->c / dx->c;
    bool ok = vec_ok(*dx, vfor(dx->dtype)) && (dz->cstride % vfor(dx->dtype) == 0) && ((uintptr_t)dz->ptr % 16 == 0);
    NASB_DISPATCH(dx->dtype, ok, (nasb::launch_pdl((channel_tile_bwd_kernel<T, V>), dim3(grid_for(rows * (dx->c / V))), dim3(256), 0, (cudaStream_t)(ST), 
                                     (const T *)dz->ptr, dz->cstride, dz->h, dz->w, (T *)dx->ptr, dx->cstride, dx->n, dx->h,
                                     dx->w, dx->c, R, stride, scale)));
    NASB_CHECK_LAUNCH();
    return 0;
}

static int spatial_reduce(const NasbTensor *x, float *out_nc, float scale, void *stream) {
    if (!x || !out_nc || !act_dtype(x)) return NASB_ERR_BAD_ARG;
    int HW = x->h * x->w;
    if (npix(*x) == 0) return 0;
    cudaError_t e = cudaMemsetAsync(out_nc, 0, sizeof(float) * (size_t)x->n * x->c, ST);
    if (e != cudaSuccess) return (int)e;
    int cblocks = cdiv(x->c, 32);
    int want = NASB_SM_COUNT * 4 / (cblocks * x->n);
    if (want < 1) want = 1;
    int rows = (HW + want - 1) / want;
    if (rows < 64) rows = 64;
    dim3 grid(cblocks, cdiv(HW, rows), x->n), block(32, 8);
    if (x->dtype == NASB_BF16)
        nasb::launch_pdl((spatial_sum_kernel<bf16>), dim3(grid), dim3(block), 0, (cudaStream_t)(ST), (const bf16 *)x->ptr, x->cstride, out_nc, HW, x->c, scale, rows);
    else
        nasb::launch_pdl((spatial_sum_kernel<float>), dim3(grid), dim3(block), 0, (cudaStream_t)(ST), (const float *)x->ptr, x->cstride, out_nc, HW, x->c, scale, rows);
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_spatial_mean(const NasbTensor *x, float *out_nc, void *stream) {
    if (!x) return NASB_ERR_BAD_ARG;
    return spatial_reduce(x, out_nc, 1.f / (float)((long long)x->h * x->w), stream);
}
extern "C" int nasb_spatial_sum(const NasbTensor *x, float *out_nc, void *stream) {
    return spatial_reduce(x, out_nc, 1.f, stream);
}

extern "C" int nasb_spatial_bcast(const NasbTensor *v, float s, const NasbTensor *out, void *stream) {
    if (!v || !out || !act_dtype(v) || !act_dtype(out) || v->n != out->n || v->c != out->c || v->h != 1 || v->w != 1)
        return NASB_ERR_BAD_ARG;
    long long rows = npix(*out);
    if (rows == 0) return 0;
    int HW = out->h * out->w;
#define BC(TV, T, V)                                                                                                          \
    nasb::launch_pdl((spatial_bcast_kernel<TV, T, V>), dim3(grid_for(rows * (out->c / V))), dim3(256), 0, (cudaStream_t)(ST), (const TV *)v->ptr, v->cstride, s, (T *)out->ptr, \
                                                                                    out->cstride, out->n, HW, out->c)
    if (out->dtype == NASB_BF16) {
        bool ok = vec_ok(*out, 8);
        if (v->dtype == NASB_BF16) { if (ok) BC(bf16, bf16, 8); else BC(bf16, bf16, 1); }
        else { if (ok) BC(float, bf16, 8); else BC(float, bf16, 1); }
    } else {
        bool ok = vec_ok(*out, 4);
        if (v->dtype == NASB_BF16) { if (ok) BC(bf16, float, 4); else BC(bf16, float, 1); }
        else { if (ok) BC(float, float, 4); else BC(float, float, 1); }
    }
#undef BC
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_channel_sum(const NasbTensor *x, float *out_c, void *workspace, void *stream) {
    (void)workspace;
    if (!x || !out_c || !act_dtype(x)) return NASB_ERR_BAD_ARG;
    // sum over all pixels of all images == spatial sum with the batch folded into the pixel axis
    long long P = npix(*x);
    if (P == 0) return 0;
    if (P > 0x7fffffffLL) return NASB_ERR_UNSUPPORTED;
    int cblocks = cdiv(x->c, 32);
    int want = NASB_SM_COUNT * 4 / cblocks;
    int rows = (int)((P + want - 1) / want);
    if (rows < 64) rows = 64;
    dim3 grid(cblocks, cdiv(P, rows), 1), block(32, 8);
    if (x->dtype == NASB_BF16)
        nasb::launch_pdl((spatial_sum_kernel<bf16>), dim3(grid), dim3(block), 0, (cudaStream_t)(ST), (const bf16 *)x->ptr, x->cstride, out_c, (int)P, x->c, 1.f, rows);
    else
        nasb::launch_pdl((spatial_sum_kernel<float>), dim3(grid), dim3(block), 0, (cudaStream_t)(ST), (const float *)x->ptr, x->cstride, out_c, (int)P, x->c, 1.f, rows);
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_relu_bwd(const NasbTensor *dy, const NasbTensor *y, const NasbTensor *dx, void *stream) {
    if (!dy || !y || !dx || !act_dtype(dy) || dy->dtype != y->dtype || dy->dtype != dx->dtype || !same_nhw(dy, y) ||
        !same_nhw(dy, dx) || dy->c != y->c || dy->c != dx->c)
        return NASB_ERR_BAD_ARG;
    long long P = npix(*dy);
    if (P == 0) return 0;
    bool ok = vec_ok(*dy, vfor(dy->dtype)) && vec_ok(*y, vfor(dy->dtype)) && vec_ok(*dx, vfor(dy->dtype));
    NASB_DISPATCH(dy->dtype, ok, (nasb::launch_pdl((relu_bwd_kernel<T, V>), dim3(grid_for(P * (dy->c / V))), dim3(256), 0, (cudaStream_t)(ST), 
                                     (const T *)dy->ptr, dy->cstride, (const T *)y->ptr, y->cstride, (T *)dx->ptr, dx->cstride, P,
                                     dy->c)));
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_sumsq(const float *x, long long n, float *out1, void *stream) {
    if (!x || !out1) return NASB_ERR_BAD_ARG;
    if (n == 0) return 0;
    nasb::launch_pdl((sumsq_kernel), dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)(ST), x, n, out1);
    NASB_CHECK_LAUNCH();
    return 0;
}
