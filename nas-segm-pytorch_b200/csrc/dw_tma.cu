// dw_tma.cu -- depthwise k x k convolution on TMA-staged shared-memory tiles (bf16 activations, sm_100a).
//
// The gather kernels of dwconv.cu re-read every input pixel k*k times through L1 with per-tap address arithmetic and
// bounds checks; on B200 they are instruction-bound (0.5-0.7 TB/s).  Here one CTA owns a TH x TW patch of output pixels
// times CC channels: a single 4-D TMA box (channels, x, y, image) brings the input patch INCLUDING its halo into shared
// memory -- out-of-image coordinates are zero-filled by the hardware, which is exactly the convolution's zero padding --
// and the threads then work out of shared memory with no global address arithmetic at all:
//   * forward / stride-1 data gradient: each thread produces a strip of 4 consecutive output pixels x 8 channels
//     (16-byte LDS per tap, fp32 accumulation, fused BN-fold / activation epilogue, 16-byte coalesced stores);
//   * weight gradient: thread = (channel vector, tap) pair walking the patch; partial sums stay in registers across the
//     CTA's patches and are flushed once with fp32 atomics.
// HBM traffic = input once (+ halo from L2) + output once.
#include <cuda.h>

#include "common.cuh"

namespace nasb {

__device__ __forceinline__ uint32_t dsmem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dmbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dsmem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void dmbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dsmem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dmbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "DW_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DW_DONE;\n\t"
        "bra DW_WAIT;\n\t"
        "DW_DONE:\n\t"
        "}" ::"r"(dsmem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            dsmem_u32(dst)),
        "l"((uint64_t)map), "r"(dsmem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

struct DwT {
    int N, IH, IW, OH, OW, C;     // forward-conv geometry of the tensor being READ (IH,IW) and WRITTEN (OH,OW)
    int CC, nchunks;              // channels per CTA (multiple of 8, <= 64), C / CC
    int TH, TW, ITH, ITW;         // output patch, input patch (with halo)
    int stride, dil, pad;         // as seen by this launch (the data gradient passes pad' = dil*(k-1) - pad)
    int flip;                     // 1: use the spatially flipped kernel (data gradient)
    int tiles_x, tiles_y;
    const float *w;               // [C][k*k]
    const float *scale, *shift;
    int act;
    bf16 *out;
    int out_cs;
    double *stats;                // optional [2][C]: sum / sum of squares of the stored output (training-mode BN fused)
};

constexpr int DW_P = 4;  // output pixels per strip

// Persistent: CTA (x, chunk) walks the patches x, x + gridDim.x, ... of its channel chunk with two shared-memory buffers; the
// TMA box of the next patch is in flight while the current one is consumed.
template <int K>
__global__ void __launch_bounds__(256) dw_tile_kernel(const __grid_constant__ CUtensorMap map_x, const DwT p) {
    extern __shared__ __align__(1024) uint8_t dsm[];
    uint8_t *base = (uint8_t *)(((uintptr_t)dsm + 127) & ~(uintptr_t)127);
    const size_t tile_bytes = (size_t)p.ITH * p.ITW * p.CC * 2;
    const size_t tile_stride = (tile_bytes + 127) & ~(size_t)127;
    bf16 *tiles[2] = {reinterpret_cast<bf16 *>(base), reinterpret_cast<bf16 *>(base + tile_stride)};  // [ITH][ITW][CC] x 2
    float *wsm = reinterpret_cast<float *>(base + 2 * tile_stride);                                   // [K*K][CC]
    uint64_t *bar = reinterpret_cast<uint64_t *>(wsm + K * K * p.CC);                                 // [2]

    const int tid = threadIdx.x;
    const int chunk = blockIdx.y, c_base = chunk * p.CC;
    const int tiles_img = p.tiles_x * p.tiles_y, total = tiles_img * p.N;
    auto origin = [&](int t, int &n, int &oy0, int &ox0) {
        n = t / tiles_img;
        const int r = t - n * tiles_img, ty = r / p.tiles_x;
        oy0 = ty * p.TH;
        ox0 = (r - ty * p.tiles_x) * p.TW;
    };
    auto issue = [&](int t, int buf) {  // thread 0
        int n, oy0, ox0;
        origin(t, n, oy0, ox0);
        dmbar_expect_tx(&bar[buf], (uint32_t)tile_bytes);
        tma_load_4d(tiles[buf], &map_x, &bar[buf], c_base, ox0 * p.stride - p.pad, oy0 * p.stride - p.pad, n);
    };
    if (tid == 0) {
        dmbar_init(&bar[0], 1);
        dmbar_init(&bar[1], 1);
    }
    for (int i = tid; i < K * K * p.CC; i += blockDim.x) {
        int tap = i / p.CC, c = i - tap * p.CC;
        int src_tap = p.flip ? (K * K - 1 - tap) : tap;
        wsm[i] = p.w[(size_t)(c_base + c) * K * K + src_tap];
    }
    __syncthreads();  // barrier init + weights visible
    if (tid == 0 && (int)blockIdx.x < total) issue(blockIdx.x, 0);

    const int CVn = p.CC / 8;
    const int strips_x = p.TW / DW_P;
    const int items = p.TH * strips_x * CVn;
    const bool epi = p.scale || p.shift || p.act;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int buf = it & 1;
        if (tid == 0 && t + (int)gridDim.x < total) issue(t + gridDim.x, buf ^ 1);  // buffer buf^1 was released by the last sync
        int n, oy0, ox0;
        origin(t, n, oy0, ox0);
        dmbar_wait(&bar[buf], (it >> 1) & 1);
        const bf16 *tile = tiles[buf];
        for (int item = tid; item < items; item += blockDim.x) {
            const int cv = item % CVn;
            const int sidx = item / CVn;
            const int sx = sidx % strips_x, sy = sidx / strips_x;
            const int oy = oy0 + sy, oxs = ox0 + sx * DW_P;
            if (oy >= p.OH || oxs >= p.OW) continue;
            float acc[DW_P][8];
#pragma unroll
            for (int q = 0; q < DW_P; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                const int iy = sy * p.stride + ky * p.dil;
                const bf16 *rowp = tile + ((size_t)iy * p.ITW) * p.CC + cv * 8;
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    const float4 w0 = *reinterpret_cast<const float4 *>(wsm + (ky * K + kx) * p.CC + cv * 8);
                    const float4 w1 = *reinterpret_cast<const float4 *>(wsm + (ky * K + kx) * p.CC + cv * 8 + 4);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int q = 0; q < DW_P; ++q) {
                        const int ix = (sx * DW_P + q) * p.stride + kx * p.dil;
                        float v[8];
                        load_vec<bf16, 8>(rowp + (size_t)ix * p.CC, v);
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[q][j] = fmaf(v[j], wv[j], acc[q][j]);
                    }
                }
            }
            const int c0 = c_base + cv * 8;
            bf16 *orow = p.out + (((size_t)n * p.OH + oy) * p.OW + oxs) * p.out_cs + c0;
            if (epi) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float sc = p.scale ? p.scale[c0 + j] : 1.f, sh = p.shift ? p.shift[c0 + j] : 0.f;
#pragma unroll
                    for (int q = 0; q < DW_P; ++q) acc[q][j] = apply_act(acc[q][j] * sc + sh, p.act);
                }
            }
#pragma unroll
            for (int q = 0; q < DW_P; ++q) {
                if (oxs + q >= p.OW) break;
                store_vec<bf16, 8>(orow + (size_t)q * p.out_cs, acc[q]);
            }
            if (p.stats) {  // optional fused statistics of the stored values (off by default, see __init__.py)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int q = 0; q < DW_P; ++q)
                        if (oxs + q < p.OW) {
                            const float v = __bfloat162float(__float2bfloat16_rn(acc[q][j]));
                            s1 += v;
                            s2 = fmaf(v, v, s2);
                        }
                    atomicAdd(&p.stats[c0 + j], (double)s1);
                    atomicAdd(&p.stats[p.C + c0 + j], (double)s2);
                }
            }
        }
        __syncthreads();  // everyone is done with tiles[buf]; it may be refilled two iterations from now
    }
}

// ---- weight gradient on tiles: thread = (tap, channel vector, pixel lane); persistent over the CTA's patches
struct DwW {
    int N, IH, IW, OH, OW, C;
    int CC, nchunks;
    int TH, TW, ITH, ITW;
    int stride, dil, pad;
    int tiles_x, tiles_y;
    float *dw;  // [C][k*k]
};

template <int K>
__global__ void __launch_bounds__(256) dw_wgrad_tile_kernel(const __grid_constant__ CUtensorMap map_x,
                                                            const __grid_constant__ CUtensorMap map_dz, const DwW p) {
    extern __shared__ __align__(1024) uint8_t dsm[];
    uint8_t *base = (uint8_t *)(((uintptr_t)dsm + 127) & ~(uintptr_t)127);
    const size_t xt_bytes = (size_t)p.ITH * p.ITW * p.CC * 2, zt_bytes = (size_t)p.TH * p.TW * p.CC * 2;
    const size_t xs = (xt_bytes + 127) & ~(size_t)127, zs = (zt_bytes + 127) & ~(size_t)127;
    bf16 *xt[2] = {reinterpret_cast<bf16 *>(base), reinterpret_cast<bf16 *>(base + xs + zs)};            // [ITH][ITW][CC]
    bf16 *zt[2] = {reinterpret_cast<bf16 *>(base + xs), reinterpret_cast<bf16 *>(base + 2 * xs + zs)};   // [TH][TW][CC]
    uint64_t *bar = reinterpret_cast<uint64_t *>(base + 2 * (xs + zs));                                  // [2]

    const int tid = threadIdx.x;
    const int chunk = blockIdx.y, c_base = chunk * p.CC;
    const int CVn = p.CC / 8;
    const int pairs = K * K * CVn;          // (tap, cv) pairs
    const int PLn = blockDim.x / pairs;     // pixel lanes per pair (>= 1 guaranteed by the host)
    const int pair = tid % pairs, pl = tid / pairs;
    const int tap = pair / CVn, cv = pair - tap * CVn;
    const int ky = tap / K, kx = tap - ky * K;
    const bool active = pl < PLn;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;

    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int total = tiles_per_img * p.N;
    auto issue = [&](int t, int buf) {  // thread 0
        const int n = t / tiles_per_img, r = t - n * tiles_per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        const int oy0 = ty * p.TH, ox0 = tx * p.TW;
        dmbar_expect_tx(&bar[buf], (uint32_t)(xt_bytes + zt_bytes));
        tma_load_4d(xt[buf], &map_x, &bar[buf], c_base, ox0 * p.stride - p.pad, oy0 * p.stride - p.pad, n);
        tma_load_4d(zt[buf], &map_dz, &bar[buf], c_base, ox0, oy0, n);
    };
    if (tid == 0) {
        dmbar_init(&bar[0], 1);
        dmbar_init(&bar[1], 1);
    }
    __syncthreads();
    if (tid == 0 && (int)blockIdx.x < total) issue(blockIdx.x, 0);
    uint32_t it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int buf = it & 1;
        if (tid == 0 && t + (int)gridDim.x < total) issue(t + gridDim.x, buf ^ 1);
        dmbar_wait(&bar[buf], (it >> 1) & 1);
        if (active) {
            // out-of-image dz pixels are zero-filled by TMA, so the whole patch can be walked unconditionally
            const bf16 *zb = zt[buf] + cv * 8;
            const bf16 *xb = xt[buf] + ((size_t)(ky * p.dil) * p.ITW + kx * p.dil) * p.CC + cv * 8;
            for (int sy = 0; sy < p.TH; ++sy) {
                const bf16 *zr = zb + (size_t)sy * p.TW * p.CC;
                const bf16 *xr = xb + (size_t)(sy * p.stride) * p.ITW * p.CC;
                for (int sx = pl; sx < p.TW; sx += PLn) {
                    float g[8], v[8];
                    load_vec<bf16, 8>(zr + (size_t)sx * p.CC, g);
                    load_vec<bf16, 8>(xr + (size_t)(sx * p.stride) * p.CC, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(g[j], v[j], acc[j]);
                }
            }
        }
        __syncthreads();  // everyone done with buffer `buf` before it is refilled
    }
    if (active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&p.dw[(size_t)(c_base + cv * 8 + j) * K * K + tap], acc[j]);
    }
}

// ---- strided data gradient on tiles: dx[iy,ix] = sum_{taps with (iy+pad-ky*dil) % s == 0 ...} dz[(iy+pad-ky*dil)/s, ...] * w[ky,kx]
// One CTA owns a TH x TW patch of dx pixels; the dz patch that can contribute is staged by one TMA box (zero fill outside).
struct DwG {
    int N, IH, IW, OH, OW, C;     // dx is IH x IW, dz is OH x OW
    int CC, nchunks;
    int TH, TW, ZTH, ZTW;         // dx patch, dz patch
    int stride, dil, pad;
    int tiles_x, tiles_y;
    const float *w;
    bf16 *dx;
    int dx_cs;
};

__device__ __forceinline__ int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <int K>
__global__ void __launch_bounds__(256) dw_dgrad_strided_tile_kernel(const __grid_constant__ CUtensorMap map_dz, const DwG p) {
    extern __shared__ __align__(1024) uint8_t dsm[];
    uint8_t *base = (uint8_t *)(((uintptr_t)dsm + 127) & ~(uintptr_t)127);
    bf16 *zt = reinterpret_cast<bf16 *>(base);  // [ZTH][ZTW][CC]
    const size_t zt_bytes = (size_t)p.ZTH * p.ZTW * p.CC * 2;
    float *wsm = reinterpret_cast<float *>(base + ((zt_bytes + 127) & ~(size_t)127));
    uint64_t *bar = reinterpret_cast<uint64_t *>(wsm + K * K * p.CC);
    const int tid = threadIdx.x;
    const int chunk = blockIdx.y, n = blockIdx.z, c_base = chunk * p.CC;
    const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
    const int iy0 = ty * p.TH, ix0 = tx * p.TW;
    // first dz row / column that any pixel of this patch can touch: min over taps of (i + pad - k*dil) / s
    const int zy0 = floordiv(iy0 + p.pad - (K - 1) * p.dil, p.stride), zx0 = floordiv(ix0 + p.pad - (K - 1) * p.dil, p.stride);
    if (tid == 0) {
        dmbar_init(bar, 1);
        dmbar_expect_tx(bar, (uint32_t)zt_bytes);
        tma_load_4d(zt, &map_dz, bar, c_base, zx0, zy0, n);
    }
    for (int i = tid; i < K * K * p.CC; i += blockDim.x) {
        int tap = i / p.CC, c = i - tap * p.CC;
        wsm[i] = p.w[(size_t)(c_base + c) * K * K + tap];
    }
    __syncthreads();
    dmbar_wait(bar, 0);
    const int CVn = p.CC / 8;
    const int items = p.TH * p.TW * CVn;
    for (int it = tid; it < items; it += blockDim.x) {
        const int cv = it % CVn, px = it / CVn;
        const int sy = px / p.TW, sx = px - sy * p.TW;
        const int iy = iy0 + sy, ix = ix0 + sx;
        if (iy >= p.IH || ix >= p.IW) continue;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            const int ty_ = iy + p.pad - ky * p.dil;
            if (ty_ < 0 || ty_ % p.stride) continue;
            const int zy = ty_ / p.stride - zy0;
            if (zy < 0 || zy >= p.ZTH) continue;
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const int tx_ = ix + p.pad - kx * p.dil;
                if (tx_ < 0 || tx_ % p.stride) continue;
                const int zx = tx_ / p.stride - zx0;
                if (zx < 0 || zx >= p.ZTW) continue;
                float v[8];
                load_vec<bf16, 8>(zt + ((size_t)zy * p.ZTW + zx) * p.CC + cv * 8, v);
                const float *wt = wsm + (ky * K + kx) * p.CC + cv * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j], wt[j], acc[j]);
            }
        }
        store_vec<bf16, 8>(p.dx + (((size_t)n * p.IH + iy) * p.IW + ix) * p.dx_cs + c_base + cv * 8, acc);
    }
}

typedef CUresult (*DwEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static DwEncodeFn dw_get_encode() {
    static DwEncodeFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (DwEncodeFn)ptr;
    }
    return fn;
}

// 4-D map (C, W, H, N) over an NHWC bf16 tensor, box (cc, bw, bh, 1), no swizzle
static bool make_map4(CUtensorMap *m, const NasbTensor *t, int cc, int bw, int bh) {
    DwEncodeFn enc = dw_get_encode();
    if (!enc || bw > 256 || bh > 256 || cc > 256) return false;
    cuuint64_t dims[4] = {(cuuint64_t)t->c, (cuuint64_t)t->w, (cuuint64_t)t->h, (cuuint64_t)t->n};
    cuuint64_t strides[3] = {(cuuint64_t)t->cstride * 2, (cuuint64_t)t->w * t->cstride * 2, (cuuint64_t)t->h * t->w * t->cstride * 2};
    cuuint32_t box[4] = {(cuuint32_t)cc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// channel chunk: the largest multiple of 8 that divides C and is <= 64
static int pick_cc(int C) {
    for (int cc = 64; cc >= 8; cc -= 8)
        if (C % cc == 0) return cc;
    return 0;
}

struct TilePlan {
    int CC, TH, TW, ITH, ITW;
    size_t smem;
};

static bool plan_tiles(int C, int ks, int stride, int dil, size_t extra_per_cc, size_t budget, TilePlan &pl) {
    int cc = pick_cc(C);
    if (!cc) return false;
    const int th_opts[3] = {8, 4, 2}, tw = stride == 1 ? 32 : 16;
    for (int cci = cc; cci >= 8; cci -= 8) {
        if (C % cci) continue;
        for (int i = 0; i < 3; ++i) {
            int th = th_opts[i];
            int ith = (th - 1) * stride + (ks - 1) * dil + 1, itw = (tw - 1) * stride + (ks - 1) * dil + 1;
            size_t bytes = (size_t)ith * itw * cci * 2 + extra_per_cc * cci + 1024;
            if (bytes <= budget && ith <= 256 && itw <= 256) {
                pl = {cci, th, tw, ith, itw, bytes};
                return true;
            }
        }
    }
    return false;
}

}  // namespace nasb

using namespace nasb;

// Forward (mode 0) or stride-1 data gradient (mode 1: x = dz, out = dx, flipped kernel).  Returns NASB_ERR_UNSUPPORTED for
// configurations the tile path does not cover (the caller then uses the gather kernels of dwconv.cu).
extern "C" int nasb_dwconv_tile(const NasbTensor *x, const float *weight, int ks, int stride, int dil, int pad, int mode,
                                const float *out_scale, const float *out_shift, int act, const NasbTensor *out, double *stats,
                                void *stream) {
    if (!x || !out || !weight) return NASB_ERR_BAD_ARG;
    if (x->dtype != NASB_BF16 || out->dtype != NASB_BF16 || x->c != out->c || x->n != out->n) return NASB_ERR_UNSUPPORTED;
    if ((ks != 3 && ks != 5) || !vec_ok(*x, 8) || !vec_ok(*out, 8)) return NASB_ERR_UNSUPPORTED;
    if (mode == 1 && stride != 1) return NASB_ERR_UNSUPPORTED;
    if (x->n > 65535) return NASB_ERR_UNSUPPORTED;
    const int epad = mode == 1 ? dil * (ks - 1) - pad : pad;
    if (epad < 0) return NASB_ERR_UNSUPPORTED;
    {
        int eh = (x->h + 2 * epad - dil * (ks - 1) - 1) / stride + 1, ew = (x->w + 2 * epad - dil * (ks - 1) - 1) / stride + 1;
        if (eh != out->h || ew != out->w) return NASB_ERR_BAD_ARG;
    }
    TilePlan pl;
    if (!plan_tiles(x->c, ks, stride, dil, (size_t)ks * ks * 4, 48 * 1024, pl)) return NASB_ERR_UNSUPPORTED;  // x2 buffers
    if (npix(*out) == 0) return 0;
    DwT p{};
    p.N = x->n;
    p.IH = x->h;
    p.IW = x->w;
    p.OH = out->h;
    p.OW = out->w;
    p.C = x->c;
    p.CC = pl.CC;
    p.nchunks = x->c / pl.CC;
    p.TH = pl.TH;
    p.TW = pl.TW;
    p.ITH = pl.ITH;
    p.ITW = pl.ITW;
    p.stride = stride;
    p.dil = dil;
    p.pad = epad;
    p.flip = mode == 1 ? 1 : 0;
    p.tiles_x = cdiv(out->w, pl.TW);
    p.tiles_y = cdiv(out->h, pl.TH);
    p.w = weight;
    p.scale = out_scale;
    p.shift = out_shift;
    p.act = act;
    p.out = (bf16 *)out->ptr;
    p.out_cs = out->cstride;
    p.stats = stats;
    CUtensorMap mx;
    if (!make_map4(&mx, x, pl.CC, pl.ITW, pl.ITH)) return NASB_ERR_UNSUPPORTED;
    const size_t tile_b = ((size_t)pl.ITH * pl.ITW * pl.CC * 2 + 127) & ~(size_t)127;
    size_t smem = 2 * tile_b + (size_t)ks * ks * pl.CC * 4 + 16 + 256;
    long long total = (long long)p.tiles_x * p.tiles_y * x->n;
    int per_sm = (int)((200 * 1024) / smem);
    if (per_sm > 3) per_sm = 3;
    if (per_sm < 1) per_sm = 1;
    long long gx = (long long)NASB_SM_COUNT * per_sm / p.nchunks;
    if (gx < 1) gx = 1;
    if (gx > total) gx = total;
    dim3 grid((unsigned)gx, p.nchunks);
    static bool cfg3 = false, cfg5 = false;
    if (ks == 3) {
        if (!cfg3) {
            if (cudaFuncSetAttribute(dw_tile_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
                return NASB_ERR_UNSUPPORTED;
            cfg3 = true;
        }
        dw_tile_kernel<3><<<grid, 256, smem, (cudaStream_t)stream>>>(mx, p);
    } else {
        if (!cfg5) {
            if (cudaFuncSetAttribute(dw_tile_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
                return NASB_ERR_UNSUPPORTED;
            cfg5 = true;
        }
        dw_tile_kernel<5><<<grid, 256, smem, (cudaStream_t)stream>>>(mx, p);
    }
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_dwconv_wgrad_tile(const NasbTensor *x, const NasbTensor *dz, int ks, int stride, int dil, int pad,
                                      float *dweight, void *stream) {
    if (!x || !dz || !dweight) return NASB_ERR_BAD_ARG;
    if (x->dtype != NASB_BF16 || dz->dtype != NASB_BF16 || x->c != dz->c || x->n != dz->n) return NASB_ERR_UNSUPPORTED;
    if ((ks != 3 && ks != 5) || !vec_ok(*x, 8) || !vec_ok(*dz, 8)) return NASB_ERR_UNSUPPORTED;
    TilePlan pl;
    // extra per channel: the dz patch (TH*TW <= 8*32 pixels) * 2 bytes
    if (!plan_tiles(x->c, ks, stride, dil, (size_t)8 * 32 * 2, 48 * 1024, pl)) return NASB_ERR_UNSUPPORTED;  // x2 buffers
    const int pairs = ks * ks * (pl.CC / 8);
    if (pairs > 256) return NASB_ERR_UNSUPPORTED;
    if (npix(*dz) == 0) return 0;
    DwW p{};
    p.N = x->n;
    p.IH = x->h;
    p.IW = x->w;
    p.OH = dz->h;
    p.OW = dz->w;
    p.C = x->c;
    p.CC = pl.CC;
    p.nchunks = x->c / pl.CC;
    p.TH = pl.TH;
    p.TW = pl.TW;
    p.ITH = pl.ITH;
    p.ITW = pl.ITW;
    p.stride = stride;
    p.dil = dil;
    p.pad = pad;
    p.tiles_x = cdiv(dz->w, pl.TW);
    p.tiles_y = cdiv(dz->h, pl.TH);
    p.dw = dweight;
    CUtensorMap mx, mz;
    if (!make_map4(&mx, x, pl.CC, pl.ITW, pl.ITH) || !make_map4(&mz, dz, pl.CC, pl.TW, pl.TH)) return NASB_ERR_UNSUPPORTED;
    size_t smem = 2 * ((((size_t)pl.ITH * pl.ITW * pl.CC * 2 + 127) & ~(size_t)127) + (((size_t)pl.TH * pl.TW * pl.CC * 2 + 127) & ~(size_t)127)) + 16 + 256;
    long long total = (long long)p.tiles_x * p.tiles_y * x->n;
    int per_sm = (int)((200 * 1024) / smem);
    if (per_sm > 3) per_sm = 3;
    if (per_sm < 1) per_sm = 1;
    long long gx = (long long)NASB_SM_COUNT * per_sm / p.nchunks;
    if (gx < 1) gx = 1;
    if (gx > total) gx = total;
    dim3 grid((unsigned)gx, p.nchunks);
    static bool cfg3 = false, cfg5 = false;
    if (ks == 3) {
        if (!cfg3) {
            if (cudaFuncSetAttribute(dw_wgrad_tile_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
                return NASB_ERR_UNSUPPORTED;
            cfg3 = true;
        }
        dw_wgrad_tile_kernel<3><<<grid, 256, smem, (cudaStream_t)stream>>>(mx, mz, p);
    } else {
        if (!cfg5) {
            if (cudaFuncSetAttribute(dw_wgrad_tile_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
                return NASB_ERR_UNSUPPORTED;
            cfg5 = true;
        }
        dw_wgrad_tile_kernel<5><<<grid, 256, smem, (cudaStream_t)stream>>>(mx, mz, p);
    }
    NASB_CHECK_LAUNCH();
    return 0;
}

// Strided (stride >= 2) data gradient on tiles; stride-1 goes through nasb_dwconv_tile(mode 1).
extern "C" int nasb_dwconv_dgrad_strided_tile(const NasbTensor *dz, const float *weight, int ks, int stride, int dil, int pad,
                                              const NasbTensor *dx, void *stream) {
    if (!dz || !dx || !weight) return NASB_ERR_BAD_ARG;
    if (dz->dtype != NASB_BF16 || dx->dtype != NASB_BF16 || dz->c != dx->c || dz->n != dx->n) return NASB_ERR_UNSUPPORTED;
    if ((ks != 3 && ks != 5) || stride < 2 || !vec_ok(*dz, 8) || !vec_ok(*dx, 8) || dx->n > 65535) return NASB_ERR_UNSUPPORTED;
    int cc = pick_cc(dx->c);
    if (!cc) return NASB_ERR_UNSUPPORTED;
    const int TH = 16, TW = 32;
    DwG p{};
    // dz rows touched by TH dx rows: ((TH-1) + (ks-1)*dil) / stride + 2
    p.ZTH = (TH - 1 + (ks - 1) * dil) / stride + 2;
    p.ZTW = (TW - 1 + (ks - 1) * dil) / stride + 2;
    while (cc >= 8) {
        if (dx->c % cc == 0 && (size_t)p.ZTH * p.ZTW * cc * 2 + (size_t)ks * ks * cc * 4 + 1024 <= 72 * 1024) break;
        cc -= 8;
    }
    if (cc < 8) return NASB_ERR_UNSUPPORTED;
    if (npix(*dx) == 0) return 0;
    p.N = dx->n;
    p.IH = dx->h;
    p.IW = dx->w;
    p.OH = dz->h;
    p.OW = dz->w;
    p.C = dx->c;
    p.CC = cc;
    p.nchunks = dx->c / cc;
    p.TH = TH;
    p.TW = TW;
    p.stride = stride;
    p.dil = dil;
    p.pad = pad;
    p.tiles_x = cdiv(dx->w, TW);
    p.tiles_y = cdiv(dx->h, TH);
    p.w = weight;
    p.dx = (bf16 *)dx->ptr;
    p.dx_cs = dx->cstride;
    CUtensorMap mz;
    if (!make_map4(&mz, dz, cc, p.ZTW, p.ZTH)) return NASB_ERR_UNSUPPORTED;
    size_t smem = (size_t)p.ZTH * p.ZTW * cc * 2 + (size_t)ks * ks * cc * 4 + 1024 + 256;
    dim3 grid(p.tiles_x * p.tiles_y, p.nchunks, dx->n);
    if (ks == 3) {
        static bool cfg = false;
        if (!cfg) {
            if (cudaFuncSetAttribute(dw_dgrad_strided_tile_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
                return NASB_ERR_UNSUPPORTED;
            cfg = true;
        }
        dw_dgrad_strided_tile_kernel<3><<<grid, 256, smem, (cudaStream_t)stream>>>(mz, p);
    } else {
        static bool cfg = false;
        if (!cfg) {
            if (cudaFuncSetAttribute(dw_dgrad_strided_tile_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
                return NASB_ERR_UNSUPPORTED;
            cfg = true;
        }
        dw_dgrad_strided_tile_kernel<5><<<grid, 256, smem, (cudaStream_t)stream>>>(mz, p);
    }
    NASB_CHECK_LAUNCH();
    return 0;
}
