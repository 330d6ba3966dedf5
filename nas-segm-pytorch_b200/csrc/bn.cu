// bn.cu -- BatchNorm2d pieces: eval-mode folding, training-mode batch statistics (+ running-stat update), the
// affine+activation pass, and the backward of y = act(gamma*xhat + beta).  All reductions accumulate in fp64
// (per-thread and across CTAs through fp64 atomics), so the fp32 results do not depend on summation order.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace nasb {

__global__ void bn_fold_kernel(const float *gamma, const float *beta, const float *mean, const float *var, float eps, int C,
                               float *scale, float *shift) {
    pdl_sync();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    float s = g / sqrtf(var[c] + eps);
    scale[c] = s;
    shift[c] = b - mean[c] * s;
}

__device__ __forceinline__ float xhat_from_y(float y, float gamma, float beta) {
    return gamma != 0.f ? (y - beta) / gamma : 0.f;
}

// ---------------------------------------------------------------------------------------------------------------
// Column (per-channel) reduction skeleton for NHWC tensors.  A CTA owns a slab of pixels; thread t owns the channel
// vector cv = t % CV (V channels = one 16-byte load) and the pixel lane pl = t / CV, so a warp reads whole pixels
// contiguously.  Per-thread partials are fp32 (a few hundred to a few thousand terms), the cross-lane reduction and the
// cross-CTA merge (fp64 atomics into ws[0..C) and ws[C..2C)) are fp64.
// F: void f(long long pixel, int c0, float (&a)[V], float (&b)[V])  accumulates into a / b.
template <int V, typename F>
__device__ __forceinline__ void colreduce2(long long r0, long long r1, int C, double *ws, F f) {
    extern __shared__ float red_sm[];  // [2][PL][C]
    const int CV = C / V;
    const int PL = blockDim.x / CV;
    const int t = threadIdx.x;
    const int cv = t % CV, pl = t / CV;
    float a[V], b[V];
#pragma unroll
    for (int j = 0; j < V; ++j) a[j] = b[j] = 0.f;
    if (pl < PL)
        for (long long m = r0 + pl; m < r1; m += PL) f(m, cv * V, a, b);
    if (pl < PL) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            red_sm[(size_t)pl * C + cv * V + j] = a[j];
            red_sm[(size_t)(PL + pl) * C + cv * V + j] = b[j];
        }
    }
    __syncthreads();
    for (int c = t; c < C; c += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int i = 0; i < PL; ++i) {
            s1 += (double)red_sm[(size_t)i * C + c];
            s2 += (double)red_sm[(size_t)(PL + i) * C + c];
        }
        atomicAdd(&ws[c], s1);
        atomicAdd(&ws[C + c], s2);
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) bn_sums_vec_kernel(const T *z, int cs, long long P, int C, double *ws,
                                                          long long rows_per_cta) {
    pdl_sync();
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = r0 + rows_per_cta < P ? r0 + rows_per_cta : P;
    colreduce2<V>(r0, r1, C, ws, [&](long long m, int c0, float (&a)[V], float (&b)[V]) {
        float v[V];
        load_vec<T, V>(z + m * cs + c0, v);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            a[j] += v[j];
            b[j] = fmaf(v[j], v[j], b[j]);
        }
    });
}

// training: g = dy*act'(z*scale+shift), xhat = (z-mean)*rstd ; eval: g = dy*act'(y), xhat = (y-beta)/gamma
// Two pixels per iteration are in flight per thread (4 x 16-byte loads) and three CTAs (768 threads) are resident per SM: the pass is a pure HBM stream.
__device__ __forceinline__ float pre_mask(float y_pre, int act) {
    if (act == NASB_ACT_RELU) return y_pre > 0.f ? 1.f : 0.f;
    if (act == NASB_ACT_RELU6) return (y_pre > 0.f && y_pre < 6.f) ? 1.f : 0.f;
    return 1.f;
}

template <typename T, int V>
__global__ void __launch_bounds__(256, 3) bn_bwd_sums_vec_kernel(const T *dy, int dy_cs, const T *yz, int yz_cs, int training,
                                                                 const float *scale, const float *shift, const float *mean,
                                                                 const float *rstd, const float *gamma, const float *beta,
                                                                 int act, long long P, int C, double *ws,
                                                                 long long rows_per_cta) {
    pdl_sync();
    extern __shared__ float red_sm[];  // [2][PL][C]
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = r0 + rows_per_cta < P ? r0 + rows_per_cta : P;
    const int CV = C / V, PL = blockDim.x / CV;
    const int t = threadIdx.x, cv = t % CV, pl = t / CV, c0 = cv * V;
    // per-channel constants of this thread's fixed channel vector: mask from y_pre = v*k_s + k_b, xhat = (v - k_mu)*k_rs
    float k_s[V], k_b[V], k_mu[V], k_rs[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int c = c0 + j;
        if (training) {
            k_s[j] = scale[c];
            k_b[j] = shift[c];
            k_mu[j] = mean[c];
            k_rs[j] = rstd[c];
        } else {  // eval: the mask is taken on y itself (identity pre-map), xhat = (y - beta) / gamma
            float ga = gamma ? gamma[c] : 1.f;
            k_s[j] = 1.f;
            k_b[j] = 0.f;
            k_mu[j] = beta ? beta[c] : 0.f;
            k_rs[j] = ga != 0.f ? 1.f / ga : 0.f;
        }
    }
    float a[V], b[V];
#pragma unroll
    for (int j = 0; j < V; ++j) a[j] = b[j] = 0.f;
    if (pl < PL) {
        for (long long m = r0 + pl; m < r1; m += 2 * PL) {
            const long long m1 = m + PL;
            const bool two = m1 < r1;
            float g0[V], v0[V], g1[V], v1[V];
            load_vec<T, V>(dy + m * dy_cs + c0, g0);
            load_vec<T, V>(yz + m * yz_cs + c0, v0);
            if (two) {
                load_vec<T, V>(dy + m1 * dy_cs + c0, g1);
                load_vec<T, V>(yz + m1 * yz_cs + c0, v1);
            }
#pragma unroll
            for (int j = 0; j < V; ++j) {
                float gm = g0[j] * pre_mask(v0[j] * k_s[j] + k_b[j], act);
                a[j] += gm;
                b[j] = fmaf(gm, (v0[j] - k_mu[j]) * k_rs[j], b[j]);
            }
            if (two) {
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    float gm = g1[j] * pre_mask(v1[j] * k_s[j] + k_b[j], act);
                    a[j] += gm;
                    b[j] = fmaf(gm, (v1[j] - k_mu[j]) * k_rs[j], b[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
            red_sm[(size_t)pl * C + c0 + j] = a[j];
            red_sm[(size_t)(PL + pl) * C + c0 + j] = b[j];
        }
    }
    __syncthreads();
    for (int c = t; c < C; c += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int i = 0; i < PL; ++i) {
            s1 += (double)red_sm[(size_t)i * C + c];
            s2 += (double)red_sm[(size_t)(PL + i) * C + c];
        }
        atomicAdd(&ws[c], s1);
        atomicAdd(&ws[C + c], s2);
    }
}

// scalar fallbacks (channel counts / pitches that are not 16-byte addressable)
// ws[0..C) += sum z ; ws[C..2C) += sum z^2        block (32 channel lanes x 8 pixel lanes), one pixel slab per CTA
template <typename T>
__global__ void __launch_bounds__(256) bn_sums_kernel(const T *z, int cs, long long P, int C, double *ws,
                                                      long long rows_per_cta) {
    pdl_sync();
    __shared__ double r1[256], r2[256];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * rows_per_cta, rend = r0 + rows_per_cta < P ? r0 + rows_per_cta : P;
    double a = 0.0, b = 0.0;
    if (c < C)
        for (long long m = r0 + threadIdx.y; m < rend; m += 8) {
            double v = (double)to_f(z[m * cs + c]);
            a += v;
            b += v * v;
        }
    const int lin = threadIdx.y * 32 + threadIdx.x;
    r1[lin] = a;
    r2[lin] = b;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double s1 = 0.0, s2 = 0.0;
        for (int i = 0; i < 8; ++i) {
            s1 += r1[i * 32 + threadIdx.x];
            s2 += r2[i * 32 + threadIdx.x];
        }
        atomicAdd(&ws[c], s1);
        atomicAdd(&ws[C + c], s2);
    }
}

__global__ void bn_stats_finalize_kernel(const double *ws, long long P, int C, const float *gamma, const float *beta,
                                         float eps, float momentum, float *running_mean, float *running_var,
                                         float *save_mean, float *save_rstd, float *scale, float *shift,
                                         long long *num_batches_tracked) {
    pdl_sync();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;  // nn.BatchNorm2d's step counter, no extra launch
    double mean = ws[c] / (double)P;
    double var = ws[C + c] / (double)P - mean * mean;  // biased
    if (var < 0.0) var = 0.0;
    float rstd = (float)(1.0 / sqrt(var + (double)eps));
    if (save_mean) save_mean[c] = (float)mean;
    if (save_rstd) save_rstd[c] = rstd;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) {
        double unb = P > 1 ? var * (double)P / (double)(P - 1) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
    float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    float s = g * rstd;
    scale[c] = s;
    shift[c] = b - (float)mean * s;
}

template <typename T, int V>
__global__ void __launch_bounds__(256) affine_act_kernel(const T *z, int z_cs, const float *scale, const float *shift, int act,
                                                         T *y, int y_cs, long long P, int C) {
    pdl_sync();
    const int CV = C / V;
    const long long total = P * CV;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        float v[V];
        load_vec<T, V>(z + pix * z_cs + cv * V, v);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float s = scale ? scale[cv * V + j] : 1.f, b = shift ? shift[cv * V + j] : 0.f;
            v[j] = apply_act(v[j] * s + b, act);
        }
        store_vec<T, V>(y + pix * y_cs + cv * V, v);
    }
}

// ws[0..C) += sum g ; ws[C..2C) += sum g*xhat   with g = dy * act'(y)
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_sums_kernel(const T *dy, int dy_cs, const T *y, int y_cs, const T *z, int z_cs,
                                                          const float *scale, const float *shift,
                                                          const float *mean, const float *rstd, int act,
                                                          const float *gamma, const float *beta, long long P, int C,
                                                          double *ws, long long rows_per_cta) {
    pdl_sync();
    __shared__ double r1[256], r2[256];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * rows_per_cta, rend = r0 + rows_per_cta < P ? r0 + rows_per_cta : P;
    double a = 0.0, b = 0.0;
    if (c < C) {
        float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
        float mu = z ? mean[c] : 0.f, rs = z ? rstd[c] : 0.f;
        for (long long m = r0 + threadIdx.y; m < rend; m += 8) {
            float zz = z ? to_f(z[m * z_cs + c]) : 0.f;
            float yy = z ? apply_act(zz * scale[c] + shift[c], act) : to_f(y[m * y_cs + c]);
            float g = to_f(dy[m * dy_cs + c]) * act_mask(yy, act);
            float xh = z ? (zz - mu) * rs : xhat_from_y(yy, ga, be);
            a += (double)g;
            b += (double)(g * xh);
        }
    }
    const int lin = threadIdx.y * 32 + threadIdx.x;
    r1[lin] = a;
    r2[lin] = b;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double s1 = 0.0, s2 = 0.0;
        for (int i = 0; i < 8; ++i) {
            s1 += r1[i * 32 + threadIdx.x];
            s2 += r2[i * 32 + threadIdx.x];
        }
        atomicAdd(&ws[c], s1);
        atomicAdd(&ws[C + c], s2);
    }
}

// dgamma += sum g*xhat ; dbeta += sum g ; coef[0..C) = mean g ; coef[C..2C) = mean g*xhat  (fp32, aliased after ws)
__global__ void bn_bwd_finalize_kernel(const double *ws, long long P, int C, float *dgamma, float *dbeta, float *coef) {
    pdl_sync();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s1 = ws[c], s2 = ws[C + c];
    if (dbeta) dbeta[c] += (float)s1;
    if (dgamma) dgamma[c] += (float)s2;
    coef[c] = (float)(s1 / (double)P);
    coef[C + c] = (float)(s2 / (double)P);
}

template <typename T, int V>
__global__ void __launch_bounds__(256) bn_bwd_dz_kernel(const T *dy, int dy_cs, const T *y, int y_cs, const T *z, int z_cs,
                                                        const float *mean, const float *rstd, int act,
                                                        const float *scale, const float *shift,
                                                        const float *coef, int training, T *dz, int dz_cs, long long P,
                                                        int C) {
    pdl_sync();
    const int CV = C / V;
    const long long total = P * CV;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        float g[V], yy[V], zz[V];
        load_vec<T, V>(dy + pix * dy_cs + cv * V, g);
        if (training) {
            load_vec<T, V>(z + pix * z_cs + cv * V, zz);
#pragma unroll
            for (int j = 0; j < V; ++j) yy[j] = apply_act(zz[j] * scale[cv * V + j] + shift[cv * V + j], act);
        } else {
            load_vec<T, V>(y + pix * y_cs + cv * V, yy);
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
            int c = cv * V + j;
            float gm = g[j] * act_mask(yy[j], act);
            float s = scale ? scale[c] : 1.f;
            if (training) {
                float xh = (zz[j] - mean[c]) * rstd[c];
                g[j] = s * (gm - coef[c] - xh * coef[C + c]);
            } else {
                g[j] = s * gm;
            }
        }
        store_vec<T, V>(dz + pix * dz_cs + cv * V, g);
    }
}


// Element-wise passes with a FIXED channel vector per thread: thread t owns channel vector cv = t % CV for its whole
// life, so the per-channel constants live in registers and the inner loop is 16-byte loads / stores only.
template <typename T, int V>
__global__ void __launch_bounds__(256, 4) affine_act_fixed_kernel(const T *z, int z_cs, const float *scale, const float *shift,
                                                               int act, T *y, int y_cs, long long P, int C) {
    pdl_sync();
    const int CV = C / V, PL = blockDim.x / CV;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV, c0 = cv * V;
    if (pl >= PL) return;
    float s[V], b[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        s[j] = scale ? scale[c0 + j] : 1.f;
        b[j] = shift ? shift[c0 + j] : 0.f;
    }
    const long long G = (long long)gridDim.x * PL;
    for (long long m = (long long)blockIdx.x * PL + pl; m < P; m += 2 * G) {
        const long long m1 = m + G;
        float v0[V], v1[V];
        load_vec<T, V>(z + m * z_cs + c0, v0);
        if (m1 < P) load_vec<T, V>(z + m1 * z_cs + c0, v1);
#pragma unroll
        for (int j = 0; j < V; ++j) v0[j] = apply_act(v0[j] * s[j] + b[j], act);
        store_vec<T, V>(y + m * y_cs + c0, v0);
        if (m1 < P) {
#pragma unroll
            for (int j = 0; j < V; ++j) v1[j] = apply_act(v1[j] * s[j] + b[j], act);
            store_vec<T, V>(y + m1 * y_cs + c0, v1);
        }
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(256, 3) bn_bwd_dz_fixed_kernel(const T *dy, int dy_cs, const T *yz, int yz_cs, int training,
                                                                 const float *scale, const float *shift, const float *mean,
                                                                 const float *rstd, const float *coef, int act, T *dz, int dz_cs,
                                                                 long long P, int C) {
    pdl_sync();
    const int CV = C / V, PL = blockDim.x / CV;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV, c0 = cv * V;
    if (pl >= PL) return;
    // training: dz = s*(g*mask - k1 - xhat*k2) = (s*mask)*g + A*(z - mu) + B  with A = -s*k2*rstd, B = -s*k1;
    //           mask from y_pre = z*s + b.   eval: dz = s*g*mask(y)  (A = B = 0, y_pre = y)
    // The mean is subtracted BEFORE the multiply: folding it into B (A*z + (B - A*mu)) cancels catastrophically in fp32
    // whenever |mean| >> std, which cost the fp32 parity mode a factor 5 in gradient accuracy.
    float s[V], b[V], A[V], B[V], mu[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int c = c0 + j;
        s[j] = scale ? scale[c] : 1.f;
        if (training) {
            const float k1 = coef[c], k2 = coef[C + c], rs = rstd[c];
            b[j] = shift[c];
            mu[j] = mean[c];
            A[j] = -s[j] * k2 * rs;
            B[j] = -s[j] * k1;
        } else {
            b[j] = 0.f;
            A[j] = B[j] = mu[j] = 0.f;
        }
    }
    const long long G = (long long)gridDim.x * PL;
    for (long long m = (long long)blockIdx.x * PL + pl; m < P; m += 2 * G) {
        const long long m1 = m + G;
        const bool two = m1 < P;
        float g0[V], v0[V], g1[V], v1[V];
        load_vec<T, V>(dy + m * dy_cs + c0, g0);
        load_vec<T, V>(yz + m * yz_cs + c0, v0);
        if (two) {
            load_vec<T, V>(dy + m1 * dy_cs + c0, g1);
            load_vec<T, V>(yz + m1 * yz_cs + c0, v1);
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float ypre = training ? v0[j] * s[j] + b[j] : v0[j];
            g0[j] = s[j] * pre_mask(ypre, act) * g0[j] + (A[j] * (v0[j] - mu[j]) + B[j]);
        }
        store_vec<T, V>(dz + m * dz_cs + c0, g0);
        if (two) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float ypre = training ? v1[j] * s[j] + b[j] : v1[j];
                g1[j] = s[j] * pre_mask(ypre, act) * g1[j] + (A[j] * (v1[j] - mu[j]) + B[j]);
            }
            store_vec<T, V>(dz + m1 * dz_cs + c0, g1);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 streaming passes (speed mode).  Same thread layout as the *_fixed kernels (thread = fixed 8-channel vector, pixel
// lane), but FOUR pixels are in flight per thread (raw 16-byte registers, converted after all loads were issued): with
// ~1.6 us of loaded HBM latency the two-pixel versions ran at 58-70 % of the DRAM peak on long_scoreboard stalls (ncu,
// profiles/r1_iter_sections_summary.txt).  Arithmetic is packed fp32 (FFMA2).  DESC walks the tensor from its last pixel
// to its first: the pass that follows a producer which wrote the tensor front-to-back then starts on the part still in L2.
constexpr int BN_PX = 4;

__device__ __forceinline__ uint4 ldg16(const bf16 *p) { return *reinterpret_cast<const uint4 *>(p); }

__device__ __forceinline__ float2 act2(float2 v, int act) { return make_float2(apply_act(v.x, act), apply_act(v.y, act)); }
// g where lo < ypre < hi, else 0
__device__ __forceinline__ float2 gate2(float2 g, float2 ypre, float lo, float hi) {
    return make_float2((ypre.x > lo && ypre.x < hi) ? g.x : 0.f, (ypre.y > lo && ypre.y < hi) ? g.y : 0.f);
}
__device__ __forceinline__ void ldc8(const float *p, int c0, float def, float2 (&o)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = p ? make_float2(p[c0 + 2 * j], p[c0 + 2 * j + 1]) : make_float2(def, def);
}

// Optional statistics finalisation folded into the pass (training mode with statistics accumulated by the conv kernel):
// every thread derives scale / shift of ITS 8 channels from the fp64 sums with exactly the arithmetic of
// bn_stats_finalize_kernel; block 0 additionally publishes save_mean / save_rstd / scale / shift, updates the running
// statistics and the step counter.  Saves one launch (and one dependency bubble) per BatchNorm.
struct BnFin {
    const double *sums;  // [2][C] or null
    long long P;
    const float *gamma, *beta;
    float eps, momentum;
    float *running_mean, *running_var, *save_mean, *save_rstd, *scale, *shift;
    long long *nbt;
};

// COOP (validated on a B200 in round 2; chosen by bn_coop() for tensors up to ~100 MB): the CTA
// derives the constants cooperatively -- thread c handles channel c (coalesced loads, one L2 round trip, one fp64 chain)
// and publishes them through shared memory -- instead of every thread walking its 8 channels one dependent load after the
// other (SASS: 8 serial LDG.64 + fp64 division chains, ~3 us before the first streaming load of every CTA; measured as
// 0.104 -> 0.128 ms on the 8x256x512x144 tensor against the unmerged apply pass).
constexpr int BN_MAXC = 2048;  // fixed_cfg: C / 8 <= 256

template <bool DESC, bool COOP>
__global__ void __launch_bounds__(256, 4) affine_act_bf16_kernel(const bf16 *z, int z_cs, const float *scale, const float *shift,
                                                                 int act, bf16 *y, int y_cs, long long P, int C, const BnFin fin,
                                                                 const bf16 *res, int res_cs) {
    pdl_sync();
    __shared__ float cst[COOP ? 2 : 1][COOP ? BN_MAXC : 1];
    const int CV = C / 8, PL = blockDim.x / CV;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV, c0 = cv * 8;
    float2 s[4], b[4];
    if (COOP && fin.sums) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const double mean = fin.sums[c] / (double)fin.P;
            double var = fin.sums[C + c] / (double)fin.P - mean * mean;  // biased
            if (var < 0.0) var = 0.0;
            const float rstd = (float)(1.0 / sqrt(var + (double)fin.eps));
            const float g = fin.gamma ? fin.gamma[c] : 1.f, be = fin.beta ? fin.beta[c] : 0.f;
            const float sc = g * rstd, sh = be - (float)mean * sc;
            cst[0][c] = sc;
            cst[COOP ? 1 : 0][c] = sh;
            if (blockIdx.x == 0) {
                if (fin.save_mean) fin.save_mean[c] = (float)mean;
                if (fin.save_rstd) fin.save_rstd[c] = rstd;
                if (fin.running_mean) fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * (float)mean;
                if (fin.running_var) {
                    const double unb = fin.P > 1 ? var * (double)fin.P / (double)(fin.P - 1) : var;
                    fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * (float)unb;
                }
                fin.scale[c] = sc;
                fin.shift[c] = sh;
            }
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && fin.nbt) *fin.nbt += 1;
        __syncthreads();
        if (pl >= PL) return;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s[j] = make_float2(cst[0][c0 + 2 * j], cst[0][c0 + 2 * j + 1]);
            b[j] = make_float2(cst[COOP ? 1 : 0][c0 + 2 * j], cst[COOP ? 1 : 0][c0 + 2 * j + 1]);
        }
    } else if (pl >= PL) {
        return;
    } else if (fin.sums) {
        float sc[8], sh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            const double mean = fin.sums[c] / (double)fin.P;
            double var = fin.sums[C + c] / (double)fin.P - mean * mean;  // biased
            if (var < 0.0) var = 0.0;
            const float rstd = (float)(1.0 / sqrt(var + (double)fin.eps));
            const float g = fin.gamma ? fin.gamma[c] : 1.f, be = fin.beta ? fin.beta[c] : 0.f;
            sc[j] = g * rstd;
            sh[j] = be - (float)mean * sc[j];
            if (blockIdx.x == 0 && pl == 0) {
                if (fin.save_mean) fin.save_mean[c] = (float)mean;
                if (fin.save_rstd) fin.save_rstd[c] = rstd;
                if (fin.running_mean) fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * (float)mean;
                if (fin.running_var) {
                    const double unb = fin.P > 1 ? var * (double)fin.P / (double)(fin.P - 1) : var;
                    fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * (float)unb;
                }
                fin.scale[c] = sc[j];
                fin.shift[c] = sh[j];
            }
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && fin.nbt) *fin.nbt += 1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s[j] = make_float2(sc[2 * j], sc[2 * j + 1]);
            b[j] = make_float2(sh[2 * j], sh[2 * j + 1]);
        }
    } else {
        ldc8(scale, c0, 1.f, s);
        ldc8(shift, c0, 0.f, b);
    }
    const long long G = (long long)gridDim.x * PL;
    for (long long m0 = (long long)blockIdx.x * PL + pl; m0 < P; m0 += BN_PX * G) {
        uint4 r[BN_PX], rr[BN_PX];
        long long m[BN_PX];
#pragma unroll
        for (int k = 0; k < BN_PX; ++k) {
            const long long mm = m0 + k * G;
            m[k] = mm < P ? (DESC ? P - 1 - mm : mm) : -1;
            if (m[k] >= 0) {
                r[k] = ldg16(z + m[k] * z_cs + c0);
                if (res) rr[k] = ldg16(res + m[k] * res_cs + c0);
            }
        }
#pragma unroll
        for (int k = 0; k < BN_PX; ++k) {
            if (m[k] < 0) continue;
            float2 v[4];
            cvt8(r[k], v);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = act2(ffma2(v[j], s[j], b[j]), act);
            if (res) {  // residual added AFTER the activation (InvertedResidual: x + conv(x)); the backward pass works from z
                float2 q[4];
                cvt8(rr[k], q);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j].x += q[j].x;
                    v[j].y += q[j].y;
                }
            }
            uint4 o;
            o.x = pack_bf16x2(v[0].x, v[0].y);
            o.y = pack_bf16x2(v[1].x, v[1].y);
            o.z = pack_bf16x2(v[2].x, v[2].y);
            o.w = pack_bf16x2(v[3].x, v[3].y);
            *reinterpret_cast<uint4 *>(y + m[k] * y_cs + c0) = o;
        }
    }
}

// ws[0..C) += sum g ; ws[C..2C) += sum g*xhat, g = dy gated by lo < v*k_s + k_b < hi, xhat = (v - mu) * rs.
// Per thread the raw moments sum g and sum g*v are accumulated (two constant vectors instead of four in registers); the
// CTA converts its partial to the centred form in fp64 before the atomics.  Slabs are walked from the END of the tensor.
__global__ void __launch_bounds__(256, 3) bn_bwd_sums_bf16_kernel(const bf16 *dy, int dy_cs, const bf16 *yz, int yz_cs,
                                                                  const float *k_s, const float *k_b, const float *mu,
                                                                  const float *rs, float lo, float hi, long long P, int C,
                                                                  double *ws, long long rows_per_cta) {
    pdl_sync();
    extern __shared__ float red_sm[];  // [2][PL][C]
    const long long e1 = P - (long long)blockIdx.x * rows_per_cta;  // this CTA's slab is [e0, e1), counted from the end
    const long long e0 = e1 - rows_per_cta > 0 ? e1 - rows_per_cta : 0;
    const int CV = C / 8, PL = blockDim.x / CV;
    const int t = threadIdx.x, cv = t % CV, pl = t / CV, c0 = cv * 8;
    float2 a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = b[j] = make_float2(0.f, 0.f);
    if (pl < PL) {
        float2 ks[4], kb[4];
        ldc8(k_s, c0, 1.f, ks);
        ldc8(k_b, c0, 0.f, kb);
        for (long long m0 = e1 - 1 - pl; m0 >= e0; m0 -= (long long)BN_PX * PL) {
            uint4 rg[BN_PX], rv[BN_PX];
            bool ok[BN_PX];
#pragma unroll
            for (int k = 0; k < BN_PX; ++k) {
                const long long m = m0 - (long long)k * PL;
                ok[k] = m >= e0;
                if (ok[k]) {
                    rg[k] = ldg16(dy + m * dy_cs + c0);
                    rv[k] = ldg16(yz + m * yz_cs + c0);
                }
            }
#pragma unroll
            for (int k = 0; k < BN_PX; ++k) {
                if (!ok[k]) continue;
                float2 g[4], v[4];
                cvt8(rg[k], g);
                cvt8(rv[k], v);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 gm = gate2(g[j], ffma2(v[j], ks[j], kb[j]), lo, hi);
                    a[j].x += gm.x;
                    a[j].y += gm.y;
                    b[j] = ffma2(gm, v[j], b[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            red_sm[(size_t)pl * C + c0 + 2 * j] = a[j].x;
            red_sm[(size_t)pl * C + c0 + 2 * j + 1] = a[j].y;
            red_sm[(size_t)(PL + pl) * C + c0 + 2 * j] = b[j].x;
            red_sm[(size_t)(PL + pl) * C + c0 + 2 * j + 1] = b[j].y;
        }
    }
    __syncthreads();
    for (int c = t; c < C; c += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int i = 0; i < PL; ++i) {
            s1 += (double)red_sm[(size_t)i * C + c];
            s2 += (double)red_sm[(size_t)(PL + i) * C + c];
        }
        atomicAdd(&ws[c], s1);
        atomicAdd(&ws[C + c], (double)rs[c] * (s2 - (double)mu[c] * s1));
    }
}

// dz = sg * gate(g) + A*v + B   (training mode only: sg = scale, A = -scale*k2*rstd, B = scale*(k2*rstd*mean - k1)).
template <bool COOP>
__global__ void __launch_bounds__(256, 3) bn_bwd_dz_bf16_kernel(const bf16 *dy, int dy_cs, const bf16 *yz, int yz_cs,
                                                                const double *ws, const float *scale, const float *shift,
                                                                const float *mean, const float *rstd, float *dgamma,
                                                                float *dbeta, float lo, float hi, bf16 *dz, int dz_cs,
                                                                long long P, int C) {
    pdl_sync();
    __shared__ float cst[COOP ? 4 : 1][COOP ? BN_MAXC : 1];
    const int CV = C / 8, PL = blockDim.x / CV;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV, c0 = cv * 8;
    if (COOP) {  // constants per channel by thread c (see affine_act_bf16_kernel), block 0 accumulates dgamma / dbeta
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const double s1 = ws[c], s2 = ws[C + c];
            const float k1 = (float)(s1 / (double)P), k2 = (float)(s2 / (double)P);
            const float sc = scale ? scale[c] : 1.f, r = rstd[c];
            cst[0][c] = sc;
            cst[COOP ? 1 : 0][c] = shift[c];
            cst[COOP ? 2 : 0][c] = -sc * k2 * r;
            cst[COOP ? 3 : 0][c] = sc * (k2 * r * mean[c] - k1);
            if (blockIdx.x == 0) {
                if (dbeta) dbeta[c] += (float)s1;
                if (dgamma) dgamma[c] += (float)s2;
            }
        }
        __syncthreads();
    }
    if (pl >= PL) return;
    // The reduction's finalisation is folded in: every thread turns the fp64 sums of ITS 8 channels into the constants of
    // dz = sg * gate(g) + A*v + B; block 0 also accumulates dgamma / dbeta (what bn_bwd_finalize_kernel did in a launch of
    // its own).  Training mode: the mask's pre-activation is z*scale + shift, so its scale is sg.
    float2 sg[4], mb[4], A[4], B[4];
    if (COOP) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sg[j] = make_float2(cst[0][c0 + 2 * j], cst[0][c0 + 2 * j + 1]);
            mb[j] = make_float2(cst[COOP ? 1 : 0][c0 + 2 * j], cst[COOP ? 1 : 0][c0 + 2 * j + 1]);
            A[j] = make_float2(cst[COOP ? 2 : 0][c0 + 2 * j], cst[COOP ? 2 : 0][c0 + 2 * j + 1]);
            B[j] = make_float2(cst[COOP ? 3 : 0][c0 + 2 * j], cst[COOP ? 3 : 0][c0 + 2 * j + 1]);
        }
    } else {
        float t[4][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            const double s1 = ws[c], s2 = ws[C + c];
            const float k1 = (float)(s1 / (double)P), k2 = (float)(s2 / (double)P);
            const float sc = scale ? scale[c] : 1.f, r = rstd[c];
            t[0][j] = sc;
            t[1][j] = shift[c];
            t[2][j] = -sc * k2 * r;
            t[3][j] = sc * (k2 * r * mean[c] - k1);
            if (blockIdx.x == 0 && pl == 0) {
                if (dbeta) dbeta[c] += (float)s1;
                if (dgamma) dgamma[c] += (float)s2;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sg[j] = make_float2(t[0][2 * j], t[0][2 * j + 1]);
            mb[j] = make_float2(t[1][2 * j], t[1][2 * j + 1]);
            A[j] = make_float2(t[2][2 * j], t[2][2 * j + 1]);
            B[j] = make_float2(t[3][2 * j], t[3][2 * j + 1]);
        }
    }
    const long long G = (long long)gridDim.x * PL;
    for (long long m0 = (long long)blockIdx.x * PL + pl; m0 < P; m0 += BN_PX * G) {
        uint4 rg[BN_PX], rv[BN_PX];
        bool ok[BN_PX];
#pragma unroll
        for (int kk = 0; kk < BN_PX; ++kk) {
            const long long m = m0 + kk * G;
            ok[kk] = m < P;
            if (ok[kk]) {
                rg[kk] = ldg16(dy + m * dy_cs + c0);
                rv[kk] = ldg16(yz + m * yz_cs + c0);
            }
        }
#pragma unroll
        for (int kk = 0; kk < BN_PX; ++kk) {
            if (!ok[kk]) continue;
            float2 g[4], v[4];
            cvt8(rg[kk], g);
            cvt8(rv[kk], v);
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 gm = gate2(g[j], ffma2(v[j], sg[j], mb[j]), lo, hi);
                const float2 r = ffma2(gm, sg[j], ffma2(A[j], v[j], B[j]));
                o[j] = pack_bf16x2(r.x, r.y);
            }
            *reinterpret_cast<uint4 *>(dz + (m0 + kk * G) * dz_cs + c0) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

static inline bool bn_safe() {
    static int v = -1;
    if (v < 0) v = getenv("NASB_BN_SAFE") ? atoi(getenv("NASB_BN_SAFE")) : 0;
    return v;
}
// Cooperative finalise prologue (thread c -> channel c, constants through shared memory) or every thread deriving its own
// eight channels: measured on a B200 (profiles/r2_switches_kbench.txt) the cooperative form wins 8-22 % up to ~100 MB
// tensors and loses 6-10 % on the 270-800 MB ones (its __syncthreads delays the first streaming loads of every CTA).
// NASB_BN_COOP=0 / 1 forces one form.
static inline bool bn_coop(long long P, int C) {
    static int v = -1;
    if (v < 0) v = getenv("NASB_BN_COOP") ? atoi(getenv("NASB_BN_COOP")) : 2;
    return v == 1 || (v == 2 && P * C <= (56LL << 20));
}
static inline bool fixed_cfg(int C, int V, long long P, int &blocks) {
    if (bn_safe() & 1) return false;
    if (C % V || C / V > 256 || C / V < 1) return false;
    int PL = 256 / (C / V);
    long long b = (P + (long long)PL * 4 - 1) / ((long long)PL * 4);  // >= 4 pixels per thread
    long long cap = (long long)NASB_SM_COUNT * 12;  // whole waves for 3 or 4 resident CTAs per SM
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    blocks = (int)b;
    return true;
}

static inline void slab_grid(long long P, int C, dim3 &grid, long long &rows) {
    int cblocks = cdiv(C, 32);
    long long want = (long long)NASB_SM_COUNT * 8 / cblocks;
    if (want < 1) want = 1;
    rows = (P + want - 1) / want;
    if (rows < 64) rows = 64;
    grid = dim3(cblocks, cdiv(P, rows));
}

// vectorised reductions: one slab per CTA, ~8 CTAs per SM; returns false if the tensor is not vector-addressable
template <int V>
static inline bool vec_reduce_cfg(int C, long long P, int &blocks, long long &rows, size_t &smem) {
    if (bn_safe() & 2) return false;
    if (C % V || C / V > 256 || C / V < 1) return false;
    int PL = 256 / (C / V);
    long long want = (long long)NASB_SM_COUNT * 6;  // two full waves of the 3 resident CTAs per SM: no partial tail wave
    rows = (P + want - 1) / want;
    long long minrows = (long long)PL * 8;
    if (rows < minrows) rows = minrows;
    blocks = cdiv(P, rows);
    smem = (size_t)2 * PL * C * sizeof(float);
    return smem <= 48 * 1024;
}

static inline int ew_grid(long long total) {
    long long b = (total + 255) / 256, cap = (long long)NASB_SM_COUNT * 32;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace nasb

using namespace nasb;
#define ST ((cudaStream_t)stream)

extern "C" int nasb_bn_fold(const float *gamma, const float *beta, const float *mean, const float *var, float eps, int C,
                            float *scale, float *shift, void *stream) {
    if (!mean || !var || !scale || !shift || C <= 0) return NASB_ERR_BAD_ARG;
    nasb::launch_pdl((bn_fold_kernel), dim3(cdiv(C, 128)), dim3(128), 0, (cudaStream_t)(ST), gamma, beta, mean, var, eps, C, scale, shift);
    NASB_CHECK_LAUNCH();
    return 0;
}

// 2*C doubles for the sums, then 2*C floats for the backward coefficients
extern "C" long long nasb_bn_stats_workspace(int C) { return (long long)C * (2 * 8 + 2 * 4); }

extern "C" int nasb_bn_stats(const NasbTensor *z, const float *gamma, const float *beta, float eps, float momentum,
                             float *running_mean, float *running_var, float *save_mean, float *save_rstd, float *scale,
                             float *shift, long long *num_batches_tracked, void *workspace, void *stream) {
    if (!z || !scale || !shift || !workspace || (z->dtype != NASB_F32 && z->dtype != NASB_BF16)) return NASB_ERR_BAD_ARG;
    long long P = npix(*z);
    int C = z->c;
    if (P == 0) return NASB_ERR_BAD_ARG;
    double *ws = (double *)workspace;
    cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, ST);
    if (e != cudaSuccess) return (int)e;
    dim3 grid;
    long long rows;
    int blocks;
    size_t smem;
    if (z->dtype == NASB_BF16 && vec_ok(*z, 8) && vec_reduce_cfg<8>(C, P, blocks, rows, smem)) {
        nasb::launch_pdl((bn_sums_vec_kernel<bf16, 8>), dim3(blocks), dim3(256), smem, (cudaStream_t)(ST), (const bf16 *)z->ptr, z->cstride, P, C, ws, rows);
    } else if (z->dtype == NASB_F32 && vec_ok(*z, 4) && vec_reduce_cfg<4>(C, P, blocks, rows, smem)) {
        nasb::launch_pdl((bn_sums_vec_kernel<float, 4>), dim3(blocks), dim3(256), smem, (cudaStream_t)(ST), (const float *)z->ptr, z->cstride, P, C, ws, rows);
    } else {
        slab_grid(P, C, grid, rows);
        if (z->dtype == NASB_BF16)
            nasb::launch_pdl((bn_sums_kernel<bf16>), dim3(grid), dim3(dim3(32, 8)), 0, (cudaStream_t)(ST), (const bf16 *)z->ptr, z->cstride, P, C, ws, rows);
        else
            nasb::launch_pdl((bn_sums_kernel<float>), dim3(grid), dim3(dim3(32, 8)), 0, (cudaStream_t)(ST), (const float *)z->ptr, z->cstride, P, C, ws, rows);
    }
    NASB_CHECK_LAUNCH();
    nasb::launch_pdl((bn_stats_finalize_kernel), dim3(cdiv(C, 128)), dim3(128), 0, (cudaStream_t)(ST), ws, P, C, gamma, beta, eps, momentum, running_mean, running_var,
                                                           save_mean, save_rstd, scale, shift, num_batches_tracked);
    NASB_CHECK_LAUNCH();
    return 0;
}

// finalize from externally accumulated fp64 sums (the tcgen05 pointwise kernel fuses the statistics into its epilogue)
extern "C" int nasb_bn_finalize(const double *sums, long long P, int C, const float *gamma, const float *beta, float eps,
                                float momentum, float *running_mean, float *running_var, float *save_mean, float *save_rstd,
                                float *scale, float *shift, long long *num_batches_tracked, void *stream) {
    if (!sums || !scale || !shift || P <= 0 || C <= 0) return NASB_ERR_BAD_ARG;
    nasb::launch_pdl((bn_stats_finalize_kernel), dim3(cdiv(C, 128)), dim3(128), 0, (cudaStream_t)(ST), sums, P, C, gamma, beta, eps, momentum, running_mean, running_var,
                                                           save_mean, save_rstd, scale, shift, num_batches_tracked);
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_affine_act(const NasbTensor *z, const float *scale, const float *shift, int act, const NasbTensor *y,
                               void *stream) {
    if (!z || !y || z->dtype != y->dtype || (z->dtype != NASB_F32 && z->dtype != NASB_BF16) || z->c != y->c ||
        npix(*z) != npix(*y))
        return NASB_ERR_BAD_ARG;
    long long P = npix(*z);
    if (P == 0) return 0;
    int C = z->c;
    {
        int blocks;
        if (z->dtype == NASB_BF16 && vec_ok(*z, 8) && vec_ok(*y, 8) && fixed_cfg(C, 8, P, blocks)) {
            // descending: z was just written front-to-back by the convolution, its tail is still in L2
            nasb::launch_pdl((affine_act_bf16_kernel<true, false>), dim3(blocks), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)z->ptr, z->cstride, scale, shift, act,
                                                                 (bf16 *)y->ptr, y->cstride, P, C, BnFin{}, nullptr, 0);
            NASB_CHECK_LAUNCH();
            return 0;
        }
        if (z->dtype == NASB_F32 && vec_ok(*z, 4) && vec_ok(*y, 4) && fixed_cfg(C, 4, P, blocks)) {
            nasb::launch_pdl((affine_act_fixed_kernel<float, 4>), dim3(blocks), dim3(256), 0, (cudaStream_t)(ST), (const float *)z->ptr, z->cstride, scale, shift, act,
                                                                      (float *)y->ptr, y->cstride, P, C);
            NASB_CHECK_LAUNCH();
            return 0;
        }
    }
    if (z->dtype == NASB_BF16) {
        if (vec_ok(*z, 8) && vec_ok(*y, 8))
            nasb::launch_pdl((affine_act_kernel<bf16, 8>), dim3(ew_grid(P * (C / 8))), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)z->ptr, z->cstride, scale, shift, act,
                                                                             (bf16 *)y->ptr, y->cstride, P, C);
        else
            nasb::launch_pdl((affine_act_kernel<bf16, 1>), dim3(ew_grid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)z->ptr, z->cstride, scale, shift, act,
                                                                       (bf16 *)y->ptr, y->cstride, P, C);
    } else {
        if (vec_ok(*z, 4) && vec_ok(*y, 4))
            nasb::launch_pdl((affine_act_kernel<float, 4>), dim3(ew_grid(P * (C / 4))), dim3(256), 0, (cudaStream_t)(ST), (const float *)z->ptr, z->cstride, scale, shift, act,
                                                                              (float *)y->ptr, y->cstride, P, C);
        else
            nasb::launch_pdl((affine_act_kernel<float, 1>), dim3(ew_grid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const float *)z->ptr, z->cstride, scale, shift, act,
                                                                        (float *)y->ptr, y->cstride, P, C);
    }
    NASB_CHECK_LAUNCH();
    return 0;
}

// nasb_bn_finalize followed by nasb_affine_act in one launch where the layout allows (bf16, 16-byte channel vectors);
// otherwise the two kernels run back to back.  Same results either way.
extern "C" int nasb_bn_finalize_affine_act(const double *sums, long long P, const NasbTensor *z, const float *gamma,
                                           const float *beta, float eps, float momentum, float *running_mean,
                                           float *running_var, float *save_mean, float *save_rstd, float *scale, float *shift,
                                           long long *num_batches_tracked, int act, const NasbTensor *res, const NasbTensor *y,
                                           void *stream) {
    if (!sums || !z || !y || !scale || !shift || P <= 0 || npix(*z) != P || npix(*y) != P || z->c != y->c || z->dtype != y->dtype)
        return NASB_ERR_BAD_ARG;
    if (res && (res->dtype != y->dtype || res->c != y->c || npix(*res) != P)) return NASB_ERR_BAD_ARG;
    const int C = z->c;
    int blocks;
    if (!(bn_safe() & 8) && z->dtype == NASB_BF16 && vec_ok(*z, 8) && vec_ok(*y, 8) && (!res || vec_ok(*res, 8)) &&
        fixed_cfg(C, 8, P, blocks)) {
        BnFin f{sums, P, gamma, beta, eps, momentum, running_mean, running_var, save_mean, save_rstd, scale, shift, num_batches_tracked};
        if (bn_coop(P, C))
            nasb::launch_pdl((affine_act_bf16_kernel<true, true>), dim3(blocks), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)z->ptr, z->cstride, nullptr, nullptr, act,
                                                                       (bf16 *)y->ptr, y->cstride, P, C, f,
                                                                       res ? (const bf16 *)res->ptr : nullptr, res ? res->cstride : 0);
        else
            nasb::launch_pdl((affine_act_bf16_kernel<true, false>), dim3(blocks), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)z->ptr, z->cstride, nullptr, nullptr, act,
                                                                        (bf16 *)y->ptr, y->cstride, P, C, f,
                                                                        res ? (const bf16 *)res->ptr : nullptr, res ? res->cstride : 0);
        NASB_CHECK_LAUNCH();
        return 0;
    }
    int rc = nasb_bn_finalize(sums, P, C, gamma, beta, eps, momentum, running_mean, running_var, save_mean, save_rstd, scale, shift,
                              num_batches_tracked, stream);
    if (rc) return rc;
    rc = nasb_affine_act(z, scale, shift, act, y, stream);
    if (rc || !res) return rc;
    return nasb_resize_axpby(y, nullptr, res, nullptr, 0, y, stream);  // y += res
}

// raw reductions of a gated data-gradient epilogue (S1 = sum g, S2 = sum g*z) -> the centred form the dz pass consumes
__global__ void bn_bwd_centre_kernel(const double *raw, const float *mu, const float *rs, int C, double *ws) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double s1 = raw[c], s2 = raw[C + c];
    ws[c] = s1;
    ws[C + c] = (double)rs[c] * (s2 - (double)mu[c] * s1);
}

// Training-mode BN backward when the two reductions already exist (NasbGate epilogue of the kernel that produced dy): only
// the dz pass runs (2 reads + 1 write instead of 4 reads + 1 write).  bf16 tensors of the packed layout only.
extern "C" int nasb_bn_bwd_from_sums(const NasbTensor *dy, const NasbTensor *z, int act, const float *scale, const float *shift,
                                     const float *save_mean, const float *save_rstd, const double *raw_sums, float *dgamma,
                                     float *dbeta, const NasbTensor *dz, void *workspace, void *stream) {
    if (!dy || !z || !dz || !scale || !shift || !save_mean || !save_rstd || !raw_sums || !workspace) return NASB_ERR_BAD_ARG;
    if (dy->dtype != NASB_BF16 || z->dtype != NASB_BF16 || dz->dtype != NASB_BF16 || z->c != dy->c || dz->c != dy->c ||
        npix(*z) != npix(*dy) || npix(*dz) != npix(*dy))
        return NASB_ERR_UNSUPPORTED;
    const long long P = npix(*dy);
    const int C = dy->c;
    if (P == 0) return 0;
    int dzblocks;
    if (!vec_ok(*dy, 8) || !vec_ok(*z, 8) || !vec_ok(*dz, 8) || !fixed_cfg(C, 8, P, dzblocks)) return NASB_ERR_UNSUPPORTED;
    double *ws = (double *)workspace;
    const float lo = act == NASB_ACT_NONE ? -INFINITY : 0.f, hi = act == NASB_ACT_RELU6 ? 6.f : INFINITY;
    nasb::launch_pdl((bn_bwd_centre_kernel), dim3(cdiv(C, 128)), dim3(128), 0, (cudaStream_t)(ST), raw_sums, save_mean, save_rstd, C, ws);
    NASB_CHECK_LAUNCH();
    if (bn_coop(P, C))
        nasb::launch_pdl((bn_bwd_dz_bf16_kernel<true>), dim3(dzblocks), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)dy->ptr, dy->cstride, (const bf16 *)z->ptr, z->cstride, ws,
                                                              scale, shift, save_mean, save_rstd, dgamma, dbeta, lo, hi,
                                                              (bf16 *)dz->ptr, dz->cstride, P, C);
    else
        nasb::launch_pdl((bn_bwd_dz_bf16_kernel<false>), dim3(dzblocks), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)dy->ptr, dy->cstride, (const bf16 *)z->ptr, z->cstride, ws,
                                                               scale, shift, save_mean, save_rstd, dgamma, dbeta, lo, hi,
                                                               (bf16 *)dz->ptr, dz->cstride, P, C);
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_bn_act_bwd(const NasbTensor *dy, const NasbTensor *y, const NasbTensor *z, int act, const float *gamma,
                               const float *beta, const float *scale, const float *shift, const float *save_mean,
                               const float *save_rstd, int training, float *dgamma, float *dbeta, const NasbTensor *dz,
                               void *workspace, void *stream) {
    if (!dy || !dz || !workspace) return NASB_ERR_BAD_ARG;
    if (training && (!z || !scale || !shift || !save_mean || !save_rstd || z->dtype != dy->dtype || z->c != dy->c ||
                     npix(*z) != npix(*dy)))
        return NASB_ERR_BAD_ARG;
    if (!training && !y) return NASB_ERR_BAD_ARG;
    if (training) y = z;  // the activation mask is recomputed from z, y is not read
    if (!training) z = nullptr;
    const void *zp = z ? z->ptr : nullptr;
    const int zcs = z ? z->cstride : 0;
    if (dy->dtype != y->dtype || dy->dtype != dz->dtype || (dy->dtype != NASB_F32 && dy->dtype != NASB_BF16))
        return NASB_ERR_BAD_ARG;
    if (dy->c != y->c || dy->c != dz->c || npix(*dy) != npix(*y) || npix(*dy) != npix(*dz)) return NASB_ERR_BAD_ARG;
    long long P = npix(*dy);
    if (P == 0) return 0;
    int C = dy->c;
    double *ws = (double *)workspace;
    float *coef = (float *)(ws + 2 * C);
    bool need_sums = training || dgamma || dbeta;
    const float lo = act == NASB_ACT_NONE ? -INFINITY : 0.f, hi = act == NASB_ACT_RELU6 ? 6.f : INFINITY;
    {
        // speed mode (bf16, training statistics): 4-pixel-in-flight packed kernels
        int blocks, dzblocks;
        long long rows;
        size_t smem;
        if (training && !(bn_safe() & 4) && dy->dtype == NASB_BF16 && vec_ok(*dy, 8) && vec_ok(*z, 8) && vec_ok(*dz, 8) &&
            vec_reduce_cfg<8>(C, P, blocks, rows, smem) && fixed_cfg(C, 8, P, dzblocks)) {
            cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, ST);
            if (e != cudaSuccess) return (int)e;
            nasb::launch_pdl((bn_bwd_sums_bf16_kernel), dim3(blocks), dim3(256), smem, (cudaStream_t)(ST), (const bf16 *)dy->ptr, dy->cstride, (const bf16 *)z->ptr, z->cstride,
                                                               scale, shift, save_mean, save_rstd, lo, hi, P, C, ws, rows);
            NASB_CHECK_LAUNCH();
            if (bn_coop(P, C))
                nasb::launch_pdl((bn_bwd_dz_bf16_kernel<true>), dim3(dzblocks), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)dy->ptr, dy->cstride, (const bf16 *)z->ptr, z->cstride,
                                                                      ws, scale, shift, save_mean, save_rstd, dgamma, dbeta, lo, hi,
                                                                      (bf16 *)dz->ptr, dz->cstride, P, C);
            else
                nasb::launch_pdl((bn_bwd_dz_bf16_kernel<false>), dim3(dzblocks), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)dy->ptr, dy->cstride, (const bf16 *)z->ptr, z->cstride,
                                                                       ws, scale, shift, save_mean, save_rstd, dgamma, dbeta, lo, hi,
                                                                       (bf16 *)dz->ptr, dz->cstride, P, C);
            NASB_CHECK_LAUNCH();
            return 0;
        }
    }
    if (need_sums) {
        cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, ST);
        if (e != cudaSuccess) return (int)e;
        dim3 grid;
        long long rows;
        int blocks;
        size_t smem;
        const NasbTensor *yz = training ? z : y;
        if (dy->dtype == NASB_BF16 && vec_ok(*dy, 8) && vec_ok(*yz, 8) && vec_reduce_cfg<8>(C, P, blocks, rows, smem)) {
            nasb::launch_pdl((bn_bwd_sums_vec_kernel<bf16, 8>), dim3(blocks), dim3(256), smem, (cudaStream_t)(ST), (const bf16 *)dy->ptr, dy->cstride, (const bf16 *)yz->ptr,
                                                                       yz->cstride, training, scale, shift, save_mean,
                                                                       save_rstd, gamma, beta, act, P, C, ws, rows);
        } else if (dy->dtype == NASB_F32 && vec_ok(*dy, 4) && vec_ok(*yz, 4) && vec_reduce_cfg<4>(C, P, blocks, rows, smem)) {
            nasb::launch_pdl((bn_bwd_sums_vec_kernel<float, 4>), dim3(blocks), dim3(256), smem, (cudaStream_t)(ST), (const float *)dy->ptr, dy->cstride,
                                                                        (const float *)yz->ptr, yz->cstride, training, scale,
                                                                        shift, save_mean, save_rstd, gamma, beta, act, P, C, ws,
                                                                        rows);
        } else {
        slab_grid(P, C, grid, rows);
        if (dy->dtype == NASB_BF16)
            nasb::launch_pdl((bn_bwd_sums_kernel<bf16>), dim3(grid), dim3(dim3(32, 8)), 0, (cudaStream_t)(ST), (const bf16 *)dy->ptr, dy->cstride, (const bf16 *)y->ptr,
                                                                   y->cstride, (const bf16 *)zp, zcs, scale, shift, save_mean, save_rstd, act,
                                                                   gamma, beta, P, C, ws, rows);
        else
            nasb::launch_pdl((bn_bwd_sums_kernel<float>), dim3(grid), dim3(dim3(32, 8)), 0, (cudaStream_t)(ST), (const float *)dy->ptr, dy->cstride, (const float *)y->ptr,
                                                                    y->cstride, (const float *)zp, zcs, scale, shift, save_mean, save_rstd, act,
                                                                    gamma, beta, P, C, ws, rows);
        }
        NASB_CHECK_LAUNCH();
        nasb::launch_pdl((bn_bwd_finalize_kernel), dim3(cdiv(C, 128)), dim3(128), 0, (cudaStream_t)(ST), ws, P, C, dgamma, dbeta, coef);
        NASB_CHECK_LAUNCH();
    }
    {
        int blocks;
        const NasbTensor *yz = training ? z : y;
        if (dy->dtype == NASB_BF16 && vec_ok(*dy, 8) && vec_ok(*yz, 8) && vec_ok(*dz, 8) && fixed_cfg(C, 8, P, blocks)) {
            nasb::launch_pdl((bn_bwd_dz_fixed_kernel<bf16, 8>), dim3(blocks), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)dy->ptr, dy->cstride, (const bf16 *)yz->ptr,
                                                                    yz->cstride, training, scale, shift, save_mean, save_rstd,
                                                                    coef, act, (bf16 *)dz->ptr, dz->cstride, P, C);
            NASB_CHECK_LAUNCH();
            return 0;
        }
        if (dy->dtype == NASB_F32 && vec_ok(*dy, 4) && vec_ok(*yz, 4) && vec_ok(*dz, 4) && fixed_cfg(C, 4, P, blocks)) {
            nasb::launch_pdl((bn_bwd_dz_fixed_kernel<float, 4>), dim3(blocks), dim3(256), 0, (cudaStream_t)(ST), (const float *)dy->ptr, dy->cstride, (const float *)yz->ptr,
                                                                     yz->cstride, training, scale, shift, save_mean, save_rstd,
                                                                     coef, act, (float *)dz->ptr, dz->cstride, P, C);
            NASB_CHECK_LAUNCH();
            return 0;
        }
    }
    if (dy->dtype == NASB_BF16) {
        if (vec_ok(*dy, 8) && vec_ok(*y, 8) && vec_ok(*dz, 8) && (!z || vec_ok(*z, 8)))
            nasb::launch_pdl((bn_bwd_dz_kernel<bf16, 8>), dim3(ew_grid(P * (C / 8))), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)dy->ptr, dy->cstride, (const bf16 *)y->ptr,
                                                                            y->cstride, (const bf16 *)zp, zcs, save_mean, save_rstd, act, scale, shift, coef,
                                                                            training, (bf16 *)dz->ptr, dz->cstride, P, C);
        else
            nasb::launch_pdl((bn_bwd_dz_kernel<bf16, 1>), dim3(ew_grid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const bf16 *)dy->ptr, dy->cstride, (const bf16 *)y->ptr,
                                                                      y->cstride, (const bf16 *)zp, zcs, save_mean, save_rstd, act, scale, shift, coef,
                                                                      training, (bf16 *)dz->ptr, dz->cstride, P, C);
    } else {
        if (vec_ok(*dy, 4) && vec_ok(*y, 4) && vec_ok(*dz, 4) && (!z || vec_ok(*z, 4)))
            nasb::launch_pdl((bn_bwd_dz_kernel<float, 4>), dim3(ew_grid(P * (C / 4))), dim3(256), 0, (cudaStream_t)(ST), (const float *)dy->ptr, dy->cstride,
                                                                             (const float *)y->ptr, y->cstride, (const float *)zp, zcs, save_mean,
                                                                             save_rstd, act, scale, shift, coef, training, (float *)dz->ptr,
                                                                             dz->cstride, P, C);
        else
            nasb::launch_pdl((bn_bwd_dz_kernel<float, 1>), dim3(ew_grid(P * C)), dim3(256), 0, (cudaStream_t)(ST), (const float *)dy->ptr, dy->cstride, (const float *)y->ptr,
                                                                       y->cstride, (const float *)zp, zcs, save_mean, save_rstd, act, scale, shift, coef,
                                                                       training, (float *)dz->ptr, dz->cstride, P, C);
    }
    NASB_CHECK_LAUNCH();
    return 0;
}
