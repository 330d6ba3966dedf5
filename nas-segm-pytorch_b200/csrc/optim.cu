// optim.cu -- multi-tensor kernels for the tail of a training iteration and the bf16 operand packs.
//
//   nasb_mt_grad_sumsq : per clip group, sum of squares of every gradient (fp64 cells) + Adam step counters += 1
//   nasb_mt_optim_step : global-norm clip (torch.nn.utils.clip_grad_norm_) + SGD / Adam update (torch.optim semantics)
//                        + Polyak average, one pass over every parameter
//   nasb_mt_pack_bf16  : every tensor-core weight operand of a model (pointwise forward / transposed, dense 3x3
//                        forward / flipped-transposed) re-packed from the fp32 master weights in one launch
//
// Reference: src/engine/trainer.py:163-169,258-272 (clip, step, Polyak), src/utils/solvers.py:35-52 (SGD / Adam).
// The ~490 parameter tensors of a candidate are tiny (median a few hundred elements), so the work is launch-bound as
// separate torch kernels (~1500 launches); here the tensor table travels BY VALUE in the kernel parameters (CUDA >= 12.1
// allows 32 KB), so that a captured CUDA graph bakes it in and eager mode needs no device-side table upload.
#include "common.cuh"

namespace nasb {

constexpr int MT_MAX = 320;        // tensors per launch (table is ~20 KB of kernel parameters)
constexpr int MT_CHUNK = 4096;     // elements per CTA
constexpr int MT_THREADS = 256;

struct MtTable {
    float *p[MT_MAX], *g[MT_MAX], *s1[MT_MAX], *s2[MT_MAX], *avg[MT_MAX], *step[MT_MAX];
    int numel[MT_MAX];
    int chunk0[MT_MAX + 1];  // prefix sum of chunk counts
    signed char group[MT_MAX], clip[MT_MAX];
    int n;
};

struct MtGroups {
    NasbOptGroup g[8];
    float max_norm[8];
};

__device__ __forceinline__ int mt_find(const int *chunk0, int n, int b) {
    int lo = 0, hi = n;  // largest t with chunk0[t] <= b
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (chunk0[mid] <= b) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(MT_THREADS) mt_sumsq_kernel(const __grid_constant__ MtTable T, double *cells) {
    pdl_sync();
    __shared__ double sm[MT_THREADS];
    const int t = mt_find(T.chunk0, T.n, blockIdx.x);
    const int chunk = blockIdx.x - T.chunk0[t];
    if (chunk == 0 && threadIdx.x == 0 && T.step[t] && T.g[t]) T.step[t][0] += 1.f;  // torch: state["step"] += 1 before the update
    const float *g = T.g[t];
    const int c = T.clip[t];
    if (!g || c < 0) return;
    const int n = T.numel[t], lo = chunk * MT_CHUNK, hi = min(n, lo + MT_CHUNK);
    double acc = 0.0;
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        for (int i = lo + threadIdx.x * 4; i < hi; i += MT_THREADS * 4) {
            if (i + 4 <= hi) {
                float4 v = *reinterpret_cast<const float4 *>(g + i);
                acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
            } else {
                for (int j = i; j < hi; ++j) acc += (double)g[j] * g[j];
            }
        }
    } else {
        for (int i = lo + threadIdx.x; i < hi; i += MT_THREADS) acc += (double)g[i] * g[i];
    }
    acc = block_sum<MT_THREADS>(acc, sm);
    if (threadIdx.x == 0) atomicAdd(&cells[c], acc);
}

// One element of the update; torch.optim semantics (foreach / single-tensor paths agree up to fp32 rounding):
//   SGD  (sgd.py)  : g += wd*p ; buf = first ? g : mom*buf + (1-damp)*g ; g = nesterov ? g + mom*buf : buf ; p -= lr*g
//   Adam (adam.py) : g += wd*p ; m = lerp(m, g, 1-b1) ; v = v*b2 + (1-b2)*g*g ; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
struct MtScalars {
    int kind, first, nesterov;
    float lr, b1, b2, eps, wd, coef, step_size, bc2_sqrt, decay;
};

__device__ __forceinline__ void mt_update(float &p, float &g, float &s1, float &s2, const MtScalars &h) {
    g *= h.coef;
    const float gc = g;  // the clipped gradient is what clip_grad_norm_ leaves in .grad
    float d = gc;
    if (h.wd != 0.f) d = fmaf(h.wd, p, d);
    if (h.kind == NASB_OPT_SGD) {
        if (h.b1 != 0.f) {
            s1 = h.first ? d : fmaf(h.b1, s1, (1.f - h.b2) * d);
            d = h.nesterov ? fmaf(h.b1, s1, d) : s1;
        }
        p = fmaf(-h.lr, d, p);
    } else {
        s1 = fmaf(d - s1, 1.f - h.b1, s1);
        s2 = fmaf((1.f - h.b2) * d, d, s2 * h.b2);
        const float denom = sqrtf(s2) / h.bc2_sqrt + h.eps;
        p = fmaf(-h.step_size, s1 / denom, p);
    }
    g = gc;
}

__global__ void __launch_bounds__(MT_THREADS) mt_step_kernel(const __grid_constant__ MtTable T, const __grid_constant__ MtGroups G,
                                                             const double *cells, float polyak_decay) {
    pdl_sync();
    const int t = mt_find(T.chunk0, T.n, blockIdx.x);
    const int chunk = blockIdx.x - T.chunk0[t];
    const int n = T.numel[t], lo = chunk * MT_CHUNK, hi = min(n, lo + MT_CHUNK);
    float *p = T.p[t], *g = T.g[t], *s1 = T.s1[t], *s2 = T.s2[t], *avg = T.avg[t];
    const int gi = T.group[t];
    MtScalars h;
    h.kind = NASB_OPT_NONE;
    h.decay = polyak_decay;
    h.coef = 1.f;
    if (g) {
        const int c = T.clip[t];
        if (c >= 0 && G.max_norm[c] > 0.f) {  // clip_grad_norm_: coef = min(1, max_norm / (total_norm + 1e-6))
            const float total = (float)sqrt(cells[c]);
            h.coef = fminf(G.max_norm[c] / (total + 1e-6f), 1.f);
        }
    }
    if (gi >= 0 && g) {
        const NasbOptGroup &q = G.g[gi];
        h.kind = q.kind, h.first = q.first, h.nesterov = q.nesterov;
        h.lr = q.lr, h.b1 = q.beta1, h.b2 = q.beta2, h.eps = q.eps, h.wd = q.weight_decay;
        h.step_size = h.lr, h.bc2_sqrt = 1.f;
        if (h.kind == NASB_OPT_ADAM) {
            const double st = (double)T.step[t][0];  // already incremented by mt_sumsq_kernel
            h.step_size = (float)((double)h.lr / (1.0 - pow((double)h.b1, st)));
            h.bc2_sqrt = (float)sqrt(1.0 - pow((double)h.b2, st));
        }
    } else if (g && h.coef != 1.f) {  // clipped but owned by no optimiser: clip_grad_norm_ still scales the gradient
        for (int j = lo + threadIdx.x; j < hi; j += MT_THREADS) g[j] *= h.coef;
    }
    const bool upd = h.kind != NASB_OPT_NONE;
    const bool has1 = upd && s1 != nullptr, has2 = upd && s2 != nullptr;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(s1) |
                       reinterpret_cast<uintptr_t>(s2) | reinterpret_cast<uintptr_t>(avg)) & 15) == 0;
    if (vec) {
        for (int i = lo + threadIdx.x * 4; i < hi; i += MT_THREADS * 4) {
            if (i + 4 <= hi) {
                float4 pv = *reinterpret_cast<float4 *>(p + i);
                float *pe = reinterpret_cast<float *>(&pv);
                if (upd) {
                    float4 gv = *reinterpret_cast<float4 *>(g + i), av = make_float4(0, 0, 0, 0), bv = make_float4(0, 0, 0, 0);
                    if (has1) av = *reinterpret_cast<float4 *>(s1 + i);
                    if (has2) bv = *reinterpret_cast<float4 *>(s2 + i);
                    float *ge = reinterpret_cast<float *>(&gv), *ae = reinterpret_cast<float *>(&av), *be = reinterpret_cast<float *>(&bv);
#pragma unroll
                    for (int j = 0; j < 4; ++j) mt_update(pe[j], ge[j], ae[j], be[j], h);
                    *reinterpret_cast<float4 *>(p + i) = pv;
                    *reinterpret_cast<float4 *>(g + i) = gv;
                    if (has1) *reinterpret_cast<float4 *>(s1 + i) = av;
                    if (has2) *reinterpret_cast<float4 *>(s2 + i) = bv;
                }
                if (avg) {  // Polyak: avg = avg*decay + (1-decay)*p_new (trainer.py:167-169)
                    float4 qv = *reinterpret_cast<float4 *>(avg + i);
                    float *qe = reinterpret_cast<float *>(&qv);
#pragma unroll
                    for (int j = 0; j < 4; ++j) qe[j] = fmaf(1.f - h.decay, pe[j], qe[j] * h.decay);
                    *reinterpret_cast<float4 *>(avg + i) = qv;
                }
            } else {
                for (int j = i; j < hi; ++j) {
                    float pj = p[j];
                    if (upd) {
                        float gj = g[j], a = has1 ? s1[j] : 0.f, b = has2 ? s2[j] : 0.f;
                        mt_update(pj, gj, a, b, h);
                        p[j] = pj, g[j] = gj;
                        if (has1) s1[j] = a;
                        if (has2) s2[j] = b;
                    }
                    if (avg) avg[j] = fmaf(1.f - h.decay, pj, avg[j] * h.decay);
                }
            }
        }
    } else {
        for (int j = lo + threadIdx.x; j < hi; j += MT_THREADS) {
            float pj = p[j];
            if (upd) {
                float gj = g[j], a = has1 ? s1[j] : 0.f, b = has2 ? s2[j] : 0.f;
                mt_update(pj, gj, a, b, h);
                p[j] = pj, g[j] = gj;
                if (has1) s1[j] = a;
                if (has2) s2[j] = b;
            }
            if (avg) avg[j] = fmaf(1.f - h.decay, pj, avg[j] * h.decay);
        }
    }
}

// ------------------------------------------------------------------------------------------------ operand packs
constexpr int PK_MAX = 448;
constexpr int PK_CHUNK = 2048;

struct PkTable {
    const float *src[PK_MAX];
    bf16 *dst[PK_MAX];
    int rows[PK_MAX], cols[PK_MAX];   // C_out, C_in of the fp32 weight
    int chunk0[PK_MAX + 1];
    signed char kind[PK_MAX];
    int n;
};

__global__ void __launch_bounds__(256) mt_pack_kernel(const __grid_constant__ PkTable T) {
    pdl_sync();
    const int t = mt_find(T.chunk0, T.n, blockIdx.x);
    const int lo = (blockIdx.x - T.chunk0[t]) * PK_CHUNK;
    const float *w = T.src[t];
    bf16 *out = T.dst[t];
    const int Co = T.rows[t], Ci = T.cols[t], kind = T.kind[t];
    if (kind == NASB_PACK_PW || kind == NASB_PACK_PW_T) {  // [R][Kp]: out[r][k] = w[r][k] or w[k][r]
        const int R = kind == NASB_PACK_PW ? Co : Ci, K = kind == NASB_PACK_PW ? Ci : Co, Kp = (K + 7) / 8 * 8;
        const int total = R * Kp, hi = min(total, lo + PK_CHUNK);
        for (int i = lo + threadIdx.x; i < hi; i += 256) {
            const int r = i / Kp, k = i - r * Kp;
            float v = 0.f;
            if (k < K) v = kind == NASB_PACK_PW ? w[(long long)r * Ci + k] : w[(long long)k * Ci + r];
            out[i] = __float2bfloat16_rn(v);
        }
    } else {  // dense 3x3: bf16 [9][Nr][Kp], same layout as pack_conv3_kernel (conv3_tcgen05.cu)
        const int mode = kind == NASB_PACK_C3 ? 0 : 1;
        const int N = mode == 0 ? Co : Ci, K = mode == 0 ? Ci : Co;
        const int Nr = (N + 63) / 64 * 64, Kp = (K + 7) / 8 * 8;
        const int total = 9 * Nr * Kp, hi = min(total, lo + PK_CHUNK);
        for (int i = lo + threadIdx.x; i < hi; i += 256) {
            const int k = i % Kp, tt = i / Kp, n = tt % Nr, tap = tt / Nr;
            float v = 0.f;
            if (mode == 0) {
                if (n < Co && k < Ci) v = w[((size_t)n * Ci + k) * 9 + tap];
            } else {
                if (n < Ci && k < Co) v = w[((size_t)k * Ci + n) * 9 + (8 - tap)];
            }
            out[i] = __float2bfloat16_rn(v);
        }
    }
}

static long long pack_elems(int kind, int Co, int Ci) {
    if (kind == NASB_PACK_PW) return (long long)Co * ((Ci + 7) / 8 * 8);
    if (kind == NASB_PACK_PW_T) return (long long)Ci * ((Co + 7) / 8 * 8);
    const int N = kind == NASB_PACK_C3 ? Co : Ci, K = kind == NASB_PACK_C3 ? Ci : Co;
    return 9LL * ((N + 63) / 64 * 64) * ((K + 7) / 8 * 8);
}

}  // namespace nasb

using namespace nasb;

static int mt_fill(MtTable &T, const NasbOptTensor *t, int n) {
    T.n = n;
    int c = 0;
    for (int i = 0; i < n; ++i) {
        if (!t[i].param || t[i].numel <= 0 || t[i].numel > 0x7fffffffLL || t[i].group >= 8 || t[i].clip >= 8) return -1;
        T.p[i] = t[i].param, T.g[i] = t[i].grad, T.s1[i] = t[i].state1, T.s2[i] = t[i].state2, T.avg[i] = t[i].avg;
        T.step[i] = t[i].step;
        T.numel[i] = (int)t[i].numel;
        T.group[i] = (signed char)t[i].group, T.clip[i] = (signed char)t[i].clip;
        T.chunk0[i] = c;
        c += cdiv(t[i].numel, MT_CHUNK);
    }
    T.chunk0[n] = c;
    return c;
}

extern "C" int nasb_mt_grad_sumsq(const NasbOptTensor *tensors, int n, double *cells, int n_cells, void *stream) {
    if (!tensors || n < 0 || !cells || n_cells < 1 || n_cells > 8) return NASB_ERR_BAD_ARG;
    cudaError_t e = cudaMemsetAsync(cells, 0, sizeof(double) * n_cells, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    static thread_local MtTable T;
    for (int i0 = 0; i0 < n; i0 += MT_MAX) {
        const int m = n - i0 < MT_MAX ? n - i0 : MT_MAX;
        const int blocks = mt_fill(T, tensors + i0, m);
        if (blocks < 0) return NASB_ERR_BAD_ARG;
        if (blocks == 0) continue;
        nasb::launch_pdl((mt_sumsq_kernel), dim3(blocks), dim3(MT_THREADS), 0, (cudaStream_t)((cudaStream_t)stream), T, cells);
        NASB_CHECK_LAUNCH();
    }
    return 0;
}

extern "C" int nasb_mt_optim_step(const NasbOptTensor *tensors, int n, const NasbOptGroup *groups, int n_groups,
                                  const float *max_norm, int n_cells, const double *cells, float polyak_decay, void *stream) {
    if (!tensors || n < 0 || n_groups < 0 || n_groups > 8 || n_cells < 0 || n_cells > 8 || (n_groups && !groups) ||
        (n_cells && (!max_norm || !cells)))
        return NASB_ERR_BAD_ARG;
    static thread_local MtTable T;
    MtGroups G;
    for (int i = 0; i < 8; ++i) {
        if (i < n_groups) G.g[i] = groups[i];
        else G.g[i] = NasbOptGroup{NASB_OPT_NONE, 0, 0, 0.f, 0.f, 0.f, 0.f, 0.f};
        G.max_norm[i] = i < n_cells ? max_norm[i] : 0.f;
    }
    for (int i0 = 0; i0 < n; i0 += MT_MAX) {
        const int m = n - i0 < MT_MAX ? n - i0 : MT_MAX;
        const int blocks = mt_fill(T, tensors + i0, m);
        if (blocks < 0) return NASB_ERR_BAD_ARG;
        for (int i = 0; i < m; ++i)
            if (T.group[i] >= n_groups || T.clip[i] >= n_cells) return NASB_ERR_BAD_ARG;
        if (blocks == 0) continue;
        nasb::launch_pdl((mt_step_kernel), dim3(blocks), dim3(MT_THREADS), 0, (cudaStream_t)((cudaStream_t)stream), T, G, cells, polyak_decay);
        NASB_CHECK_LAUNCH();
    }
    return 0;
}

extern "C" long long nasb_pack_elems(int kind, int Co, int Ci) {
    if (kind < NASB_PACK_PW || kind > NASB_PACK_C3_T || Co <= 0 || Ci <= 0) return -1;
    return pack_elems(kind, Co, Ci);
}

extern "C" int nasb_mt_pack_bf16(const NasbPackJob *jobs, int n, void *stream) {
    if (!jobs || n < 0) return NASB_ERR_BAD_ARG;
    static thread_local PkTable T;
    for (int i0 = 0; i0 < n; i0 += PK_MAX) {
        const int m = n - i0 < PK_MAX ? n - i0 : PK_MAX;
        int c = 0;
        for (int i = 0; i < m; ++i) {
            const NasbPackJob &j = jobs[i0 + i];
            if (!j.src || !j.dst || j.c_out <= 0 || j.c_in <= 0 || j.kind < NASB_PACK_PW || j.kind > NASB_PACK_C3_T)
                return NASB_ERR_BAD_ARG;
            T.src[i] = j.src, T.dst[i] = (bf16 *)j.dst, T.rows[i] = j.c_out, T.cols[i] = j.c_in, T.kind[i] = (signed char)j.kind;
            T.chunk0[i] = c;
            c += cdiv(pack_elems(j.kind, j.c_out, j.c_in), PK_CHUNK);
        }
        T.chunk0[m] = c;
        T.n = m;
        if (c == 0) continue;
        nasb::launch_pdl((mt_pack_kernel), dim3(c), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), T);
        NASB_CHECK_LAUNCH();
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ pointwise unit: BN backward without dz
// See include/nasb200.h (NasbGate / nasb_pw_bn_bwd_prepare).  All matrices here are tiny (c_out <= 1024, c_in <= 64):
// block b handles output channel rows; the c_in x c_in matrix M = W^T diag(A) W is reduced over c_out by block 0's threads.
namespace nasb {

struct PwBnP {
    const float *w;
    int co, ci;
    const float *scale, *mean, *rstd;
    const double *sums;
    double invP;
    const float *gx, *xx, *sx;
    float *dw, *dgamma, *dbeta;
    bf16 *pack_g, *pack_x;
    float *bias_row;
    float *coef;  // scratch [3][co]: s, A, B
};

__device__ __forceinline__ float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// per output channel: the constants of dz = s*g + A*z + B, and the BatchNorm parameter gradients
__global__ void __launch_bounds__(256) pw_bn_coef_kernel(const PwBnP p) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.co) return;
    const double S1 = p.sums[c], S2 = p.sums[p.co + c];
    const double mu = (double)p.mean[c], rs = (double)p.rstd[c];
    const double s2c = rs * (S2 - mu * S1);  // sum g*xhat
    const float k1 = (float)(S1 * p.invP), k2 = (float)(s2c * p.invP);
    const float s = p.scale[c];
    p.coef[c] = s;
    p.coef[p.co + c] = -s * k2 * (float)rs;
    p.coef[2 * p.co + c] = s * (k2 * (float)rs * (float)mu - k1);
    if (p.dgamma) p.dgamma[c] += (float)s2c;
    if (p.dbeta) p.dbeta[c] += (float)S1;
}

// blockIdx.y == 0: rows -- dW[c][i] += s gx + A (W_b xx)[c][i] + B sx[i]; pack_g[i][c] = bf16(s W[c][i])
// blockIdx.y == 1: M[o][i] = sum_c W_b[c][i] A[c] W_b[c][o] -> pack_x[o][Kp(ci)]; bias_row[o] = sum_c B[c] W_b[c][o]
__global__ void __launch_bounds__(256) pw_bn_mats_kernel(const PwBnP p) {
    pdl_sync();
    const int Kp = (p.co + 7) / 8 * 8, Kpi = (p.ci + 7) / 8 * 8;
    const float *cs = p.coef, *cA = p.coef + p.co, *cB = p.coef + 2 * p.co;
    if (blockIdx.y == 0) {
        for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p.co * p.ci; idx += gridDim.x * blockDim.x) {
            const int c = idx / p.ci, i = idx - c * p.ci;
            float wg = 0.f;  // (W_b . xx)[c][i]
            for (int k = 0; k < p.ci; ++k) wg = fmaf(bf16r(p.w[(size_t)c * p.ci + k]), p.xx[(size_t)k * p.ci + i], wg);
            p.dw[idx] += cs[c] * p.gx[idx] + cA[c] * wg + cB[c] * p.sx[i];
            p.pack_g[(size_t)i * Kp + c] = __float2bfloat16_rn(cs[c] * p.w[idx]);
        }
        for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p.ci * (Kp - p.co); idx += gridDim.x * blockDim.x) {
            const int i = idx / (Kp - p.co), k = p.co + idx % (Kp - p.co);  // zero the K padding of pack_g
            p.pack_g[(size_t)i * Kp + k] = __float2bfloat16_rn(0.f);
        }
    } else {
        for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p.ci * Kpi; idx += gridDim.x * blockDim.x) {
            const int o = idx / Kpi, i = idx - o * Kpi;
            float m = 0.f, br = 0.f;
            if (i < p.ci) {
                for (int c = 0; c < p.co; ++c) {
                    const float wo = bf16r(p.w[(size_t)c * p.ci + o]);
                    m = fmaf(bf16r(p.w[(size_t)c * p.ci + i]) * cA[c], wo, m);
                    if (i == 0) br = fmaf(cB[c], wo, br);
                }
            }
            p.pack_x[idx] = __float2bfloat16_rn(m);
            if (i == 0) p.bias_row[o] = br;
        }
    }
}

}  // namespace nasb

extern "C" long long nasb_pw_bn_bwd_scratch(int c_out) { return c_out > 0 ? 3LL * c_out * (long long)sizeof(float) : -1; }

extern "C" int nasb_pw_bn_bwd_prepare(const float *weight, int c_out, int c_in, const float *scale, const float *mean,
                                      const float *rstd, const double *sums, long long P, const float *gx, const float *xx,
                                      const float *sx, float *dweight, float *dgamma, float *dbeta, void *pack_g, void *pack_x,
                                      float *bias_row, void *scratch, void *stream) {
    if (!weight || !scale || !mean || !rstd || !sums || !gx || !xx || !sx || !dweight || !pack_g || !pack_x || !bias_row ||
        !scratch || c_out <= 0 || c_in <= 0 || P <= 0)
        return NASB_ERR_BAD_ARG;
    if (c_in > 256 || c_out > 4096) return NASB_ERR_UNSUPPORTED;
    PwBnP p{weight, c_out, c_in, scale, mean, rstd, sums, 1.0 / (double)P, gx, xx, sx, dweight, dgamma, dbeta,
            (bf16 *)pack_g, (bf16 *)pack_x, bias_row, (float *)scratch};
    nasb::launch_pdl((pw_bn_coef_kernel), dim3(cdiv(c_out, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), p);
    NASB_CHECK_LAUNCH();
    const int Kpi = (c_in + 7) / 8 * 8;
    long long work = (long long)c_out * c_in > (long long)c_in * Kpi ? (long long)c_out * c_in : (long long)c_in * Kpi;
    nasb::launch_pdl((pw_bn_mats_kernel), dim3(dim3(cdiv(work, 256), 2)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), p);
    NASB_CHECK_LAUNCH();
    return 0;
}
