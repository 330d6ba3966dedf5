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
    // STATS == 2 (data gradient feeding a conv -> BN(train) -> act unit, whose pre-BN output is gz): the stored value is
    // g = dx gated by lo < gz*g_scale + g_shift < hi (the activation's derivative mask) and stats receives
    // [sum g, sum g*gz] -- the two reductions of that unit's BatchNorm backward, which then needs no pass of its own.
    const bf16 *gz;
    int gz_cs;
    const float *g_scale, *g_shift;
    float g_lo, g_hi;
};

constexpr int DW_P = 4;  // output pixels per strip (a vertical strip: 4 consecutive rows of one column)

// explicit shared-space 16-byte loads on 32-bit shared addresses (a generic pointer derived through uintptr_t arithmetic
// makes the compiler emit generic LD instead of LDS)
__device__ __forceinline__ uint4 lds16(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ void ldw8(uint32_t a, float2 (&w)[4]) {
    const uint4 lo = lds16(a), hi = lds16(a + 16);
    w[0] = make_float2(__uint_as_float(lo.x), __uint_as_float(lo.y));
    w[1] = make_float2(__uint_as_float(lo.z), __uint_as_float(lo.w));
    w[2] = make_float2(__uint_as_float(hi.x), __uint_as_float(hi.y));
    w[3] = make_float2(__uint_as_float(hi.z), __uint_as_float(hi.w));
}

// Persistent: CTA (x, chunk) walks the patches x, x + gridDim.x, ... of its channel chunk with two shared-memory buffers; the
// TMA box of the next patch is in flight while the current one is consumed.
//
// Thread = (channel vector cv, lane): cv = tid % CVn is fixed for the thread's whole life (so the optional training-BN
// statistics stay in registers until the end), lane = tid / CVn walks the patch's vertical strips -- DW_P consecutive
// output rows of one column.  Consecutive lanes own consecutive columns: shared-memory reads and global stores of a warp
// are contiguous.  With dilation 1 (template S = stride 1 or 2) every input pixel of the strip's column window is loaded
// and converted once and feeds all the taps / output rows it belongs to out of registers; the weights of the current
// kernel column sit in registers.  S == 0 is the general (dilated) path: one load per tap.
// All arithmetic is packed fp32 (fma.rn.f32x2), accumulation in fp32.
template <int K, int S, int STATS>
__global__ void __launch_bounds__(256, 2) dw_tile_kernel(const __grid_constant__ CUtensorMap map_x, const DwT p) {
    pdl_sync();
    extern __shared__ __align__(1024) uint8_t dsm[];
    uint8_t *base = (uint8_t *)(((uintptr_t)dsm + 127) & ~(uintptr_t)127);
    const size_t tile_bytes = (size_t)p.ITH * p.ITW * p.CC * 2;
    const size_t tile_stride = (tile_bytes + 127) & ~(size_t)127;
    bf16 *tiles[2] = {reinterpret_cast<bf16 *>(base), reinterpret_cast<bf16 *>(base + tile_stride)};  // [ITH][ITW][CC] x 2
    float *wsm = reinterpret_cast<float *>(base + 2 * tile_stride);                                   // [K*K][CC]
    float *ssm = wsm + K * K * p.CC;                                                                  // [2][CC] (STATS)
    uint64_t *bar = reinterpret_cast<uint64_t *>(ssm + 2 * p.CC);                                     // [2]

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
    if (STATS)
        for (int i = tid; i < 2 * p.CC; i += blockDim.x) ssm[i] = 0.f;
    __syncthreads();  // barrier init + weights visible
    if (tid == 0 && (int)blockIdx.x < total) issue(blockIdx.x, 0);

    const int CVn = p.CC / 8;
    const int cv = tid % CVn, lane = tid / CVn, L = blockDim.x / CVn;  // host: blockDim.x is a multiple of CVn
    const int nstrips = (p.TH / DW_P) * p.TW;
    const int c0 = c_base + cv * 8;
    const bool epi = p.scale || p.shift || p.act;
    float2 st1[4], st2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) st1[j] = st2[j] = make_float2(0.f, 0.f);
    const uint32_t wcv = dsmem_u32(wsm) + cv * 32;  // this thread's 8 weights of tap 0; next tap: + CC*4 bytes
    const int CC = p.CC, ITW = p.ITW;
    const uint32_t pxb = (uint32_t)CC * 2, rowb = (uint32_t)ITW * pxb, tapb = (uint32_t)CC * 4;  // byte pitches
    const uint32_t tile_a[2] = {dsmem_u32(tiles[0]) + cv * 16, dsmem_u32(tiles[1]) + cv * 16};

    uint32_t it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int buf = it & 1;
        if (tid == 0 && t + (int)gridDim.x < total) issue(t + gridDim.x, buf ^ 1);  // buffer buf^1 was released by the last sync
        int n, oy0, ox0;
        origin(t, n, oy0, ox0);
        dmbar_wait(&bar[buf], (it >> 1) & 1);
        const uint32_t tile = buf ? tile_a[1] : tile_a[0];
        for (int s = lane; s < nstrips; s += L) {
            const int sx = s % p.TW, sy = (s / p.TW) * DW_P;
            const int ox = ox0 + sx, oy = oy0 + sy;
            if (ox >= p.OW || oy >= p.OH) continue;
            float2 acc[DW_P][4];
#pragma unroll
            for (int q = 0; q < DW_P; ++q)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[q][j] = make_float2(0.f, 0.f);
            if constexpr (S > 0) {
                // column window: input rows sy*S .. sy*S + (DW_P-1)*S + K-1 of column sx*S + kx
                constexpr int ROWS = (DW_P - 1) * S + K;
#pragma unroll 1
                for (int kx = 0; kx < K; ++kx) {
                    float2 wc[K][4];
#pragma unroll
                    for (int ky = 0; ky < K; ++ky) ldw8(wcv + (uint32_t)(ky * K + kx) * tapb, wc[ky]);
                    const uint32_t col = tile + (uint32_t)(sy * S) * rowb + (uint32_t)(sx * S + kx) * pxb;
#pragma unroll
                    for (int i = 0; i < ROWS; ++i) {
                        float2 v[4];
                        cvt8(lds16(col + (uint32_t)i * rowb), v);
#pragma unroll
                        for (int q = 0; q < DW_P; ++q) {
                            const int ky = i - q * S;
                            if (ky >= 0 && ky < K) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) acc[q][j] = ffma2(v[j], wc[ky][j], acc[q][j]);
                            }
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int ky = 0; ky < K; ++ky) {
#pragma unroll 1
                    for (int kx = 0; kx < K; ++kx) {
                        float2 w[4];
                        ldw8(wcv + (uint32_t)(ky * K + kx) * tapb, w);
                        const uint32_t tp = tile + (uint32_t)(sy * p.stride + ky * p.dil) * rowb + (uint32_t)(sx * p.stride + kx * p.dil) * pxb;
#pragma unroll
                        for (int q = 0; q < DW_P; ++q) {
                            float2 v[4];
                            cvt8(lds16(tp + (uint32_t)(q * p.stride) * rowb), v);
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[q][j] = ffma2(v[j], w[j], acc[q][j]);
                        }
                    }
                }
            }
            bf16 *op = p.out + (((size_t)n * p.OH + oy) * p.OW + ox) * p.out_cs + c0;
            uint4 zr[DW_P];
            if constexpr (STATS == 2) {  // all gate operands of the strip in flight before the first one is needed
                const bf16 *zp = p.gz + (((size_t)n * p.OH + oy) * p.OW + ox) * p.gz_cs + c0;
#pragma unroll
                for (int q = 0; q < DW_P; ++q)
                    zr[q] = oy + q < p.OH ? __ldg(reinterpret_cast<const uint4 *>(zp + (size_t)q * p.OW * p.gz_cs)) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int q = 0; q < DW_P; ++q) {
                if (oy + q >= p.OH) break;
                if (epi) {  // folded BN / activation (eval mode): constants come from L1
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 sc = p.scale ? *reinterpret_cast<const float2 *>(p.scale + c0 + 2 * j) : make_float2(1.f, 1.f);
                        const float2 sh = p.shift ? *reinterpret_cast<const float2 *>(p.shift + c0 + 2 * j) : make_float2(0.f, 0.f);
                        const float2 e = ffma2(acc[q][j], sc, sh);
                        acc[q][j] = make_float2(apply_act(e.x, p.act), apply_act(e.y, p.act));
                    }
                }
                float2 zv[4];
                if constexpr (STATS == 2) {
                    cvt8(zr[q], zv);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 gs = *reinterpret_cast<const float2 *>(p.g_scale + c0 + 2 * j);
                        const float2 gb = *reinterpret_cast<const float2 *>(p.g_shift + c0 + 2 * j);
                        const float2 pre = ffma2(zv[j], gs, gb);
                        if (!(pre.x > p.g_lo && pre.x < p.g_hi)) acc[q][j].x = 0.f;
                        if (!(pre.y > p.g_lo && pre.y < p.g_hi)) acc[q][j].y = 0.f;
                    }
                }
                uint4 o;
                o.x = pack_bf16x2(acc[q][0].x, acc[q][0].y);
                o.y = pack_bf16x2(acc[q][1].x, acc[q][1].y);
                o.z = pack_bf16x2(acc[q][2].x, acc[q][2].y);
                o.w = pack_bf16x2(acc[q][3].x, acc[q][3].y);
                *reinterpret_cast<uint4 *>(op + (size_t)q * p.OW * p.out_cs) = o;
                if (STATS) {  // statistics of the values as stored (bf16-rounded)
                    float2 r[4];
                    cvt8(o, r);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        st1[j].x += r[j].x;
                        st1[j].y += r[j].y;
                        if constexpr (STATS == 2) st2[j] = ffma2(r[j], zv[j], st2[j]);
                        else st2[j] = ffma2(r[j], r[j], st2[j]);
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with tiles[buf]; it may be refilled two iterations from now
    }
    if (STATS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&ssm[cv * 8 + 2 * j], st1[j].x);
            atomicAdd(&ssm[cv * 8 + 2 * j + 1], st1[j].y);
            atomicAdd(&ssm[CC + cv * 8 + 2 * j], st2[j].x);
            atomicAdd(&ssm[CC + cv * 8 + 2 * j + 1], st2[j].y);
        }
        __syncthreads();
        for (int i = tid; i < CC; i += blockDim.x) {
            atomicAdd(&p.stats[c_base + i], (double)ssm[i]);
            atomicAdd(&p.stats[p.C + c_base + i], (double)ssm[CC + i]);
        }
    }
}

// ---- weight gradient on tiles.  Thread = (channel vector, kernel-column group, lane); lanes walk vertical strips of DW_P
// dz rows.  KC = kernel columns per thread (1: one kernel column, K taps in registers).
// dz and x patches arrive by TMA (zero fill outside the image => no bounds checks); partial sums stay in registers across
// the CTA's patches, are merged through shared-memory atomics and flushed once with fp32 global atomics.
struct DwW {
    int N, IH, IW, OH, OW, C;
    int CC, nchunks;
    int TH, TW, ITH, ITW;
    int stride, dil, pad;
    int tiles_x, tiles_y;
    int lanes;  // lanes per (cv, column group); blockDim.x = lanes * CVn * (K / KC)
    float *dw;  // [C][k*k]
};

template <int K, int S, int KC>
__global__ void __launch_bounds__(256, 2) dw_wgrad_tile_kernel(const __grid_constant__ CUtensorMap map_x,
                                                               const __grid_constant__ CUtensorMap map_dz, const DwW p) {
    pdl_sync();
    extern __shared__ __align__(1024) uint8_t dsm[];
    uint8_t *base = (uint8_t *)(((uintptr_t)dsm + 127) & ~(uintptr_t)127);
    const size_t xt_bytes = (size_t)p.ITH * p.ITW * p.CC * 2, zt_bytes = (size_t)p.TH * p.TW * p.CC * 2;
    const size_t xs = (xt_bytes + 127) & ~(size_t)127, zs = (zt_bytes + 127) & ~(size_t)127;
    bf16 *xt[2] = {reinterpret_cast<bf16 *>(base), reinterpret_cast<bf16 *>(base + xs + zs)};            // [ITH][ITW][CC]
    bf16 *zt[2] = {reinterpret_cast<bf16 *>(base + xs), reinterpret_cast<bf16 *>(base + 2 * xs + zs)};   // [TH][TW][CC]
    float *dsum = reinterpret_cast<float *>(base + 2 * (xs + zs));                                       // [CC][K*K]
    uint64_t *bar = reinterpret_cast<uint64_t *>(dsum + p.CC * K * K);                                   // [2]

    const int tid = threadIdx.x;
    const int chunk = blockIdx.y, c_base = chunk * p.CC;
    const int CVn = p.CC / 8;
    constexpr int NG = K / KC;                 // column groups
    const int G = CVn * NG;
    const int grp = tid % G, lane = tid / G;
    const int cv = grp % CVn, kx0 = (grp / CVn) * KC;
    const int L = p.lanes;
    const int CC = p.CC, ITW = p.ITW, TW = p.TW;
    const uint32_t pxb = (uint32_t)CC * 2, rowb = (uint32_t)ITW * pxb, zrowb = (uint32_t)TW * pxb;  // byte pitches
    float2 acc[KC * K][4];  // [kxi * K + ky]
#pragma unroll
    for (int i = 0; i < KC * K; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);

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
    for (int i = tid; i < CC * K * K; i += blockDim.x) dsum[i] = 0.f;
    __syncthreads();
    if (tid == 0 && (int)blockIdx.x < total) issue(blockIdx.x, 0);
    const int nstrips = (p.TH / DW_P) * TW;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int buf = it & 1;
        if (tid == 0 && t + (int)gridDim.x < total) issue(t + gridDim.x, buf ^ 1);
        dmbar_wait(&bar[buf], (it >> 1) & 1);
        const uint32_t zb = dsmem_u32(buf ? zt[1] : zt[0]) + cv * 16;
        const uint32_t xb = dsmem_u32(buf ? xt[1] : xt[0]) + cv * 16;
        for (int s = lane; s < nstrips; s += L) {
            const int sx = s % TW, sy = (s / TW) * DW_P;
            float2 g[DW_P][4];
#pragma unroll
            for (int q = 0; q < DW_P; ++q) cvt8(lds16(zb + (uint32_t)(sy + q) * zrowb + (uint32_t)sx * pxb), g[q]);
            if constexpr (S > 0) {
                constexpr int ROWS = (DW_P - 1) * S + K;
#pragma unroll
                for (int kxi = 0; kxi < KC; ++kxi) {
                    const uint32_t col = xb + (uint32_t)(sy * S) * rowb + (uint32_t)(sx * S + kx0 + kxi) * pxb;
#pragma unroll
                    for (int i = 0; i < ROWS; ++i) {
                        float2 v[4];
                        cvt8(lds16(col + (uint32_t)i * rowb), v);
#pragma unroll
                        for (int q = 0; q < DW_P; ++q) {
                            const int ky = i - q * S;
                            if (ky >= 0 && ky < K) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) acc[kxi * K + ky][j] = ffma2(g[q][j], v[j], acc[kxi * K + ky][j]);
                            }
                        }
                    }
                }
            } else {
#pragma unroll
                for (int kxi = 0; kxi < KC; ++kxi) {
#pragma unroll
                    for (int ky = 0; ky < K; ++ky) {
                        const uint32_t tp = xb + (uint32_t)(sy * p.stride + ky * p.dil) * rowb + (uint32_t)(sx * p.stride + (kx0 + kxi) * p.dil) * pxb;
#pragma unroll
                        for (int q = 0; q < DW_P; ++q) {
                            float2 v[4];
                            cvt8(lds16(tp + (uint32_t)(q * p.stride) * rowb), v);
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[kxi * K + ky][j] = ffma2(g[q][j], v[j], acc[kxi * K + ky][j]);
                        }
                    }
                }
            }
        }
        __syncthreads();  // everyone done with buffer `buf` before it is refilled
    }
#pragma unroll
    for (int kxi = 0; kxi < KC; ++kxi)
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            const int tap = ky * K + kx0 + kxi;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(&dsum[(cv * 8 + 2 * j) * K * K + tap], acc[kxi * K + ky][j].x);
                atomicAdd(&dsum[(cv * 8 + 2 * j + 1) * K * K + tap], acc[kxi * K + ky][j].y);
            }
        }
    __syncthreads();
    for (int i = tid; i < CC * K * K; i += blockDim.x) atomicAdd(&p.dw[(size_t)c_base * K * K + i], dsum[i]);
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
    pdl_sync();
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

// ---- stride-2 3x3 (pad 1, dilation 1) data gradient, the MobileNet-v2 down-sampling case.  A 2x2 quad of dx pixels
// (rows 2a, 2a+1; columns 2b, 2b+1) depends on the 2x2 dz neighbourhood (a..a+1, b..b+1) only:
//   dx[2a  ,2b  ] = z00 w11                         dx[2a  ,2b+1] = z01 w10 + z00 w12
//   dx[2a+1,2b  ] = z10 w01 + z00 w21               dx[2a+1,2b+1] = z11 w00 + z10 w02 + z01 w20 + z00 w22
// so no tap needs a divisibility test.  Same persistent double-buffered TMA structure as dw_tile_kernel; thread = (channel
// vector, lane), a lane walks vertical strips of QS quads of one quad column and slides its two dz rows down.
struct DwQ {
    int N, IH, IW, C;        // dx is IH x IW
    int CC, nchunks;
    int TH, TW, ZTH, ZTW;    // dx patch (multiples of 2*QS and 2), dz patch = TH/2+1 x TW/2+1
    int tiles_x, tiles_y;
    const float *w;
    bf16 *dx;
    int dx_cs;
    // GATE: see DwT (gated data gradient + the producing unit's BatchNorm-backward reductions)
    const bf16 *gz;
    int gz_cs;
    const float *g_scale, *g_shift;
    float g_lo, g_hi;
    double *stats;
};
constexpr int DQ_QS = 4;  // quads per strip

template <bool GATE>
__global__ void __launch_bounds__(256, 2) dw_dgrad_s2k3_kernel(const __grid_constant__ CUtensorMap map_dz, const DwQ p) {
    pdl_sync();
    extern __shared__ __align__(1024) uint8_t dsm[];
    uint8_t *base = (uint8_t *)(((uintptr_t)dsm + 127) & ~(uintptr_t)127);
    const size_t tile_bytes = (size_t)p.ZTH * p.ZTW * p.CC * 2;
    const size_t tile_stride = (tile_bytes + 127) & ~(size_t)127;
    bf16 *tiles[2] = {reinterpret_cast<bf16 *>(base), reinterpret_cast<bf16 *>(base + tile_stride)};
    float *wsm = reinterpret_cast<float *>(base + 2 * tile_stride);  // [9][CC]
    float *ssm = wsm + 9 * p.CC;                                      // [2][CC] (GATE)
    uint64_t *bar = reinterpret_cast<uint64_t *>(ssm + 2 * p.CC);
    const int tid = threadIdx.x;
    const int chunk = blockIdx.y, c_base = chunk * p.CC;
    const int tiles_img = p.tiles_x * p.tiles_y, total = tiles_img * p.N;
    if (GATE)
        for (int i = tid; i < 2 * p.CC; i += blockDim.x) ssm[i] = 0.f;
    float2 st1[4], st2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) st1[j] = st2[j] = make_float2(0.f, 0.f);
    auto origin = [&](int t, int &n, int &iy0, int &ix0) {
        n = t / tiles_img;
        const int r = t - n * tiles_img, ty = r / p.tiles_x;
        iy0 = ty * p.TH;
        ix0 = (r - ty * p.tiles_x) * p.TW;
    };
    auto issue = [&](int t, int buf) {  // thread 0
        int n, iy0, ix0;
        origin(t, n, iy0, ix0);
        dmbar_expect_tx(&bar[buf], (uint32_t)tile_bytes);
        tma_load_4d(tiles[buf], &map_dz, &bar[buf], c_base, ix0 / 2, iy0 / 2, n);
    };
    if (tid == 0) {
        dmbar_init(&bar[0], 1);
        dmbar_init(&bar[1], 1);
    }
    for (int i = tid; i < 9 * p.CC; i += blockDim.x) {
        int tap = i / p.CC, c = i - tap * p.CC;
        wsm[i] = p.w[(size_t)(c_base + c) * 9 + tap];
    }
    __syncthreads();
    if (tid == 0 && (int)blockIdx.x < total) issue(blockIdx.x, 0);

    const int CVn = p.CC / 8;
    const int cv = tid % CVn, lane = tid / CVn, L = blockDim.x / CVn;
    const int qcols = p.TW / 2, nstrips = (p.TH / (2 * DQ_QS)) * qcols;
    const int c0 = c_base + cv * 8;
    const uint32_t pxb = (uint32_t)p.CC * 2, rowb = (uint32_t)p.ZTW * pxb, tapb = (uint32_t)p.CC * 4;
    const uint32_t wcv = dsmem_u32(wsm) + cv * 32;
    const uint32_t tile_a[2] = {dsmem_u32(tiles[0]) + cv * 16, dsmem_u32(tiles[1]) + cv * 16};
    const size_t orow = (size_t)p.IW * p.dx_cs;
    const size_t zrow = (size_t)p.IW * p.gz_cs;

    uint32_t it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int buf = it & 1;
        if (tid == 0 && t + (int)gridDim.x < total) issue(t + gridDim.x, buf ^ 1);
        int n, iy0, ix0;
        origin(t, n, iy0, ix0);
        dmbar_wait(&bar[buf], (it >> 1) & 1);
        const uint32_t tile = buf ? tile_a[1] : tile_a[0];
        for (int s = lane; s < nstrips; s += L) {
            const int qb = s % qcols, qa0 = (s / qcols) * DQ_QS;
            const int ix = ix0 + 2 * qb, iyb = iy0 + 2 * qa0;
            if (ix >= p.IW || iyb >= p.IH) continue;
            const bool x1ok = ix + 1 < p.IW;
            float2 z0[2][4], z1[2][4];  // dz rows a and a+1, columns b and b+1
            const uint32_t zp = tile + (uint32_t)qa0 * rowb + (uint32_t)qb * pxb;
            cvt8(lds16(zp), z0[0]);
            cvt8(lds16(zp + pxb), z0[1]);
            bf16 *op = p.dx + (((size_t)n * p.IH + iyb) * p.IW + ix) * p.dx_cs + c0;
            const bf16 *gp = GATE ? p.gz + (((size_t)n * p.IH + iyb) * p.IW + ix) * p.gz_cs + c0 : nullptr;
            constexpr int UNR = GATE ? 2 : DQ_QS;  // the gated epilogue needs some of the registers four unrolled quads would take
#pragma unroll UNR
            for (int q = 0; q < DQ_QS; ++q) {
                const int iy = iyb + 2 * q;
                if (iy >= p.IH) break;
                uint4 zq[4];
                if (GATE) {  // the quad's gate operands: in flight while the taps are accumulated
                    const bf16 *g0 = gp + (size_t)(2 * q) * zrow;
                    const bool y1ok = iy + 1 < p.IH;
                    zq[0] = __ldg(reinterpret_cast<const uint4 *>(g0));
                    zq[1] = x1ok ? __ldg(reinterpret_cast<const uint4 *>(g0 + p.gz_cs)) : make_uint4(0, 0, 0, 0);
                    zq[2] = y1ok ? __ldg(reinterpret_cast<const uint4 *>(g0 + zrow)) : make_uint4(0, 0, 0, 0);
                    zq[3] = (y1ok && x1ok) ? __ldg(reinterpret_cast<const uint4 *>(g0 + zrow + p.gz_cs)) : make_uint4(0, 0, 0, 0);
                }
                cvt8(lds16(zp + (uint32_t)(q + 1) * rowb), z1[0]);
                cvt8(lds16(zp + (uint32_t)(q + 1) * rowb + pxb), z1[1]);
                float2 w[4], o00[4], o01[4], o10[4], o11[4];
                const float2 zero = make_float2(0.f, 0.f);
                ldw8(wcv + 4 * tapb, w);  // w11
#pragma unroll
                for (int j = 0; j < 4; ++j) o00[j] = ffma2(z0[0][j], w[j], zero);
                ldw8(wcv + 3 * tapb, w);  // w10
#pragma unroll
                for (int j = 0; j < 4; ++j) o01[j] = ffma2(z0[1][j], w[j], zero);
                ldw8(wcv + 5 * tapb, w);  // w12
#pragma unroll
                for (int j = 0; j < 4; ++j) o01[j] = ffma2(z0[0][j], w[j], o01[j]);
                ldw8(wcv + 1 * tapb, w);  // w01
#pragma unroll
                for (int j = 0; j < 4; ++j) o10[j] = ffma2(z1[0][j], w[j], zero);
                ldw8(wcv + 7 * tapb, w);  // w21
#pragma unroll
                for (int j = 0; j < 4; ++j) o10[j] = ffma2(z0[0][j], w[j], o10[j]);
                ldw8(wcv + 0 * tapb, w);  // w00
#pragma unroll
                for (int j = 0; j < 4; ++j) o11[j] = ffma2(z1[1][j], w[j], zero);
                ldw8(wcv + 2 * tapb, w);  // w02
#pragma unroll
                for (int j = 0; j < 4; ++j) o11[j] = ffma2(z1[0][j], w[j], o11[j]);
                ldw8(wcv + 6 * tapb, w);  // w20
#pragma unroll
                for (int j = 0; j < 4; ++j) o11[j] = ffma2(z0[1][j], w[j], o11[j]);
                ldw8(wcv + 8 * tapb, w);  // w22
#pragma unroll
                for (int j = 0; j < 4; ++j) o11[j] = ffma2(z0[0][j], w[j], o11[j]);
                auto st = [&](bf16 *dst, float2 (&o)[4], const uint4 &zraw) {
                    float2 zv[4];
                    if (GATE) {
                        cvt8(zraw, zv);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {  // per-channel constants from L1 (registers are the scarce resource here)
                            const float2 pre = ffma2(zv[j], *reinterpret_cast<const float2 *>(p.g_scale + c0 + 2 * j),
                                                     *reinterpret_cast<const float2 *>(p.g_shift + c0 + 2 * j));
                            if (!(pre.x > p.g_lo && pre.x < p.g_hi)) o[j].x = 0.f;
                            if (!(pre.y > p.g_lo && pre.y < p.g_hi)) o[j].y = 0.f;
                        }
                    }
                    const uint4 ov = make_uint4(pack_bf16x2(o[0].x, o[0].y), pack_bf16x2(o[1].x, o[1].y),
                                                pack_bf16x2(o[2].x, o[2].y), pack_bf16x2(o[3].x, o[3].y));
                    *reinterpret_cast<uint4 *>(dst) = ov;
                    if (GATE) {  // reductions over the values as stored (bf16-rounded)
                        float2 r[4];
                        cvt8(ov, r);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            st1[j].x += r[j].x;
                            st1[j].y += r[j].y;
                            st2[j] = ffma2(r[j], zv[j], st2[j]);
                        }
                    }
                };
                bf16 *r0 = op + (size_t)(2 * q) * orow;
                st(r0, o00, zq[0]);
                if (x1ok) st(r0 + p.dx_cs, o01, zq[1]);
                if (iy + 1 < p.IH) {
                    st(r0 + orow, o10, zq[2]);
                    if (x1ok) st(r0 + orow + p.dx_cs, o11, zq[3]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    z0[0][j] = z1[0][j];
                    z0[1][j] = z1[1][j];
                }
            }
        }
        __syncthreads();
    }
    if (GATE) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&ssm[cv * 8 + 2 * j], st1[j].x);
            atomicAdd(&ssm[cv * 8 + 2 * j + 1], st1[j].y);
            atomicAdd(&ssm[p.CC + cv * 8 + 2 * j], st2[j].x);
            atomicAdd(&ssm[p.CC + cv * 8 + 2 * j + 1], st2[j].y);
        }
        __syncthreads();
        for (int i = tid; i < p.CC; i += blockDim.x) {
            atomicAdd(&p.stats[c_base + i], (double)ssm[i]);
            atomicAdd(&p.stats[p.C + c_base + i], (double)ssm[p.CC + i]);
        }
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

// Dilated kernels (sep_conv_3x3_dil3, sep_conv_5x5_dil6 of the search space): the halo is 2*dil*(k-1)/2 pixels per side, so the
// 8 x 32 patch of the dense path re-reads the input 7 times through TMA (k = 5, dil = 6: 32 x 56 input pixels for 256 outputs,
// 8 channels per box).  Here the patch may be up to 32 rows and use one CTA's whole shared memory (two buffers of <= 104 KB):
// the plan minimises input pixels per output pixel, counting boxes narrower than a 32-byte sector as 1.5 times as expensive.
static bool plan_tiles_dilated(int C, int ks, int dil, size_t extra_per_cc, TilePlan &pl) {
    const size_t budget = 104 * 1024;
    const int tw = 32;
    double best = 0.0;
    bool found = false;
    for (int cci = 64; cci >= 8; cci -= 8) {
        if (C % cci) continue;
        for (int th = 32; th >= 4; th >>= 1) {
            const int ith = th - 1 + (ks - 1) * dil + 1, itw = tw - 1 + (ks - 1) * dil + 1;
            const size_t bytes = (size_t)ith * itw * cci * 2 + extra_per_cc * cci + 1024;
            if (bytes > budget || ith > 256 || itw > 256) continue;
            const double cost = (double)ith * itw / ((double)th * tw) * (cci >= 16 ? 1.0 : 1.5);
            if (!found || cost < best - 1e-9) {
                best = cost;
                pl = {cci, th, tw, ith, itw, bytes};
                found = true;
            }
        }
    }
    return found;
}

static bool plan_tiles(int C, int ks, int stride, int dil, size_t extra_per_cc, size_t budget, TilePlan &pl) {
    int cc = pick_cc(C);
    if (!cc) return false;
    const int th_opts[2] = {8, 4}, tw = stride == 1 ? 32 : 16;  // multiples of DW_P
    for (int cci = cc; cci >= 8; cci -= 8) {
        if (C % cci) continue;
        for (int i = 0; i < 2; ++i) {
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
static int dw_dgrad_strided(const NasbTensor *dz, const float *weight, int ks, int stride, int dil, int pad, const NasbTensor *dx,
                            const NasbGate *gate, void *stream);

static int dw_tile_launch(const NasbTensor *x, const float *weight, int ks, int stride, int dil, int pad, int mode,
                          const float *out_scale, const float *out_shift, int act, const NasbTensor *out, double *stats,
                          const NasbGate *gate, void *stream) {
    if (!x || !out || !weight) return NASB_ERR_BAD_ARG;
    if (gate) {
        if (mode != 1 || !gate->z || !gate->sums || !gate->scale || !gate->shift || stats) return NASB_ERR_BAD_ARG;
        if (gate->z->dtype != NASB_BF16 || gate->z->c != out->c || npix(*gate->z) != npix(*out) || gate->z->h != out->h ||
            !vec_ok(*gate->z, 8))
            return NASB_ERR_UNSUPPORTED;
    }
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
    const bool big_halo = dil > 1 && stride == 1;
    if (big_halo ? !plan_tiles_dilated(x->c, ks, dil, (size_t)ks * ks * 4, pl)
                 : !plan_tiles(x->c, ks, stride, dil, (size_t)ks * ks * 4, 48 * 1024, pl))
        return NASB_ERR_UNSUPPORTED;  // x2 buffers
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
    if (gate) {
        p.stats = gate->sums;
        p.gz = (const bf16 *)gate->z->ptr;
        p.gz_cs = gate->z->cstride;
        p.g_scale = gate->scale;
        p.g_shift = gate->shift;
        p.g_lo = gate->act == NASB_ACT_NONE ? -INFINITY : 0.f;
        p.g_hi = gate->act == NASB_ACT_RELU6 ? 6.f : INFINITY;
    }
    CUtensorMap mx;
    if (!make_map4(&mx, x, pl.CC, pl.ITW, pl.ITH)) return NASB_ERR_UNSUPPORTED;
    const size_t tile_b = ((size_t)pl.ITH * pl.ITW * pl.CC * 2 + 127) & ~(size_t)127;
    size_t smem = 2 * tile_b + (size_t)ks * ks * pl.CC * 4 + (size_t)2 * pl.CC * 4 + 16 + 256;
    const int CVn = pl.CC / 8, nstrips = (pl.TH / DW_P) * pl.TW;
    int L = big_halo ? 128 : 64;  // one CTA per SM with the large dilated patches: more lanes instead of more CTAs
    while (L * CVn > 256 || L > nstrips) L >>= 1;
    const int threads = L * CVn;
    typedef void (*Kern)(const CUtensorMap, const DwT);
    // [k][s][stats]: stats 2 = gated data gradient (stride 1 only, so the s = 2 slots repeat the general kernel)
    static const Kern kerns[2][3][3] = {
        {{dw_tile_kernel<3, 0, 0>, dw_tile_kernel<3, 0, 1>, dw_tile_kernel<3, 0, 2>},
         {dw_tile_kernel<3, 1, 0>, dw_tile_kernel<3, 1, 1>, dw_tile_kernel<3, 1, 2>},
         {dw_tile_kernel<3, 2, 0>, dw_tile_kernel<3, 2, 1>, dw_tile_kernel<3, 0, 2>}},
        {{dw_tile_kernel<5, 0, 0>, dw_tile_kernel<5, 0, 1>, dw_tile_kernel<5, 0, 2>},
         {dw_tile_kernel<5, 1, 0>, dw_tile_kernel<5, 1, 1>, dw_tile_kernel<5, 1, 2>},
         {dw_tile_kernel<5, 2, 0>, dw_tile_kernel<5, 2, 1>, dw_tile_kernel<5, 0, 2>}}};
    static bool cfg[2][3][3] = {};
    const int ki = ks == 3 ? 0 : 1, si = (dil == 1 && stride <= 2) ? stride : 0, ti = gate ? 2 : (stats ? 1 : 0);
    Kern kern = kerns[ki][si][ti];
    if (!cfg[ki][si][ti]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, si == 0 ? 216 * 1024 : 100 * 1024) != cudaSuccess)
            return NASB_ERR_UNSUPPORTED;
        cfg[ki][si][ti] = true;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (per_sm > 6) per_sm = 6;
    long long total = (long long)p.tiles_x * p.tiles_y * x->n;
    long long gx = (long long)NASB_SM_COUNT * per_sm / p.nchunks;
    if (gx < 1) gx = 1;
    if (gx > total) gx = total;
    dim3 grid((unsigned)gx, p.nchunks);
    nasb::launch_pdl((kern), dim3(grid), dim3(threads), smem, (cudaStream_t)((cudaStream_t)stream), mx, p);
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
    // (the large dilated patches of the forward kernel are not used here: with K column groups per channel vector the CTA has
    // 160 threads, and one such CTA per SM hides less latency than the three small ones)
    if (!plan_tiles(x->c, ks, stride, dil, (size_t)8 * 32 * 2, 48 * 1024, pl)) return NASB_ERR_UNSUPPORTED;  // x2 buffers
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
    size_t smem = 2 * ((((size_t)pl.ITH * pl.ITW * pl.CC * 2 + 127) & ~(size_t)127) + (((size_t)pl.TH * pl.TW * pl.CC * 2 + 127) & ~(size_t)127)) +
                  (size_t)pl.CC * ks * ks * 4 + 16 + 256;
    const int G = (pl.CC / 8) * ks, nstrips = (pl.TH / DW_P) * pl.TW;
    int L = 64;
    while (L * G > 256 || L > nstrips) L >>= 1;
    if (L < 1) return NASB_ERR_UNSUPPORTED;
    p.lanes = L;
    const int threads = L * G;
    typedef void (*Kern)(const CUtensorMap, const CUtensorMap, const DwW);
    static const Kern kerns[2][3] = {{dw_wgrad_tile_kernel<3, 0, 1>, dw_wgrad_tile_kernel<3, 1, 1>, dw_wgrad_tile_kernel<3, 2, 1>},
                                     {dw_wgrad_tile_kernel<5, 0, 1>, dw_wgrad_tile_kernel<5, 1, 1>, dw_wgrad_tile_kernel<5, 2, 1>}};
    static bool cfg[2][3] = {};
    const int ki = ks == 3 ? 0 : 1, si = (dil == 1 && stride <= 2) ? stride : 0;
    Kern kern = kerns[ki][si];
    if (!cfg[ki][si]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024) != cudaSuccess)
            return NASB_ERR_UNSUPPORTED;
        cfg[ki][si] = true;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (per_sm > 6) per_sm = 6;
    long long total = (long long)p.tiles_x * p.tiles_y * x->n;
    long long gx = (long long)NASB_SM_COUNT * per_sm / p.nchunks;
    if (gx < 1) gx = 1;
    if (gx > total) gx = total;
    dim3 grid((unsigned)gx, p.nchunks);
    nasb::launch_pdl((kern), dim3(grid), dim3(threads), smem, (cudaStream_t)((cudaStream_t)stream), mx, mz, p);
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_dwconv_tile(const NasbTensor *x, const float *weight, int ks, int stride, int dil, int pad, int mode,
                                const float *out_scale, const float *out_shift, int act, const NasbTensor *out, double *stats,
                                void *stream) {
    return dw_tile_launch(x, weight, ks, stride, dil, pad, mode, out_scale, out_shift, act, out, stats, nullptr, stream);
}

// Stride-1 data gradient whose epilogue gates dx with the activation mask of the unit that produced the conv input and
// accumulates that unit's two BatchNorm-backward reductions (see DwT / NasbGate).
extern "C" int nasb_dwconv_dgrad_gated(const NasbTensor *dz, const float *weight, int ks, int stride, int dil, int pad,
                                       const NasbGate *gate, const NasbTensor *dx, void *stream) {
    if (!gate) return NASB_ERR_BAD_ARG;
    if (stride == 1) return dw_tile_launch(dz, weight, ks, 1, dil, pad, 1, nullptr, nullptr, NASB_ACT_NONE, dx, nullptr, gate, stream);
    return dw_dgrad_strided(dz, weight, ks, stride, dil, pad, dx, gate, stream);
}

// Strided (stride >= 2) data gradient on tiles; stride-1 goes through nasb_dwconv_tile(mode 1).
extern "C" int nasb_dwconv_dgrad_strided_tile(const NasbTensor *dz, const float *weight, int ks, int stride, int dil, int pad,
                                              const NasbTensor *dx, void *stream) {
    return dw_dgrad_strided(dz, weight, ks, stride, dil, pad, dx, nullptr, stream);
}

static int dw_dgrad_strided(const NasbTensor *dz, const float *weight, int ks, int stride, int dil, int pad, const NasbTensor *dx,
                            const NasbGate *gate, void *stream) {
    if (!dz || !dx || !weight) return NASB_ERR_BAD_ARG;
    if (gate) {
        if (!gate->z || !gate->sums || !gate->scale || !gate->shift) return NASB_ERR_BAD_ARG;
        if (gate->z->dtype != NASB_BF16 || gate->z->c != dx->c || npix(*gate->z) != npix(*dx) || gate->z->h != dx->h ||
            !vec_ok(*gate->z, 8))
            return NASB_ERR_UNSUPPORTED;
    }
    if (dz->dtype != NASB_BF16 || dx->dtype != NASB_BF16 || dz->c != dx->c || dz->n != dx->n) return NASB_ERR_UNSUPPORTED;
    if ((ks != 3 && ks != 5) || stride < 2 || !vec_ok(*dz, 8) || !vec_ok(*dx, 8) || dx->n > 65535) return NASB_ERR_UNSUPPORTED;
    int cc = pick_cc(dx->c);
    if (!cc) return NASB_ERR_UNSUPPORTED;
    if (ks == 3 && stride == 2 && dil == 1 && pad == 1 && dz->h == (dx->h - 1) / 2 + 1 && dz->w == (dx->w - 1) / 2 + 1) {
        if (npix(*dx) == 0) return 0;
        DwQ q{};
        q.N = dx->n;
        q.IH = dx->h;
        q.IW = dx->w;
        q.C = dx->c;
        q.TH = 16;
        q.TW = 64;
        q.ZTH = q.TH / 2 + 1;
        q.ZTW = q.TW / 2 + 1;
        while (cc >= 8 && (dx->c % cc || (size_t)q.ZTH * q.ZTW * cc * 2 > 40 * 1024)) cc -= 8;
        if (cc < 8) return NASB_ERR_UNSUPPORTED;
        q.CC = cc;
        q.nchunks = dx->c / cc;
        q.tiles_x = cdiv(dx->w, q.TW);
        q.tiles_y = cdiv(dx->h, q.TH);
        q.w = weight;
        q.dx = (bf16 *)dx->ptr;
        q.dx_cs = dx->cstride;
        CUtensorMap mq;
        if (!make_map4(&mq, dz, cc, q.ZTW, q.ZTH)) return NASB_ERR_UNSUPPORTED;
        const size_t tile_b = ((size_t)q.ZTH * q.ZTW * cc * 2 + 127) & ~(size_t)127;
        const size_t smem = 2 * tile_b + (size_t)9 * cc * 4 + (size_t)2 * cc * 4 + 16 + 256;
        const int CVn = cc / 8, nstrips = (q.TH / (2 * DQ_QS)) * (q.TW / 2);
        int L = 64;
        while (L * CVn > 256 || L > nstrips) L >>= 1;
        const int threads = L * CVn;
        if (gate) {
            q.gz = (const bf16 *)gate->z->ptr;
            q.gz_cs = gate->z->cstride;
            q.g_scale = gate->scale;
            q.g_shift = gate->shift;
            q.g_lo = gate->act == NASB_ACT_NONE ? -INFINITY : 0.f;
            q.g_hi = gate->act == NASB_ACT_RELU6 ? 6.f : INFINITY;
            q.stats = gate->sums;
        }
        typedef void (*KernQ)(const CUtensorMap, const DwQ);
        KernQ kq = gate ? (KernQ)dw_dgrad_s2k3_kernel<true> : (KernQ)dw_dgrad_s2k3_kernel<false>;
        static bool cfgq[2] = {false, false};
        if (!cfgq[gate ? 1 : 0]) {
            if (cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
                return NASB_ERR_UNSUPPORTED;
            cfgq[gate ? 1 : 0] = true;
        }
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kq, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        if (per_sm > 6) per_sm = 6;
        long long total = (long long)q.tiles_x * q.tiles_y * dx->n;
        long long gx = (long long)NASB_SM_COUNT * per_sm / q.nchunks;
        if (gx < 1) gx = 1;
        if (gx > total) gx = total;
        nasb::launch_pdl((kq), dim3(dim3((unsigned)gx, q.nchunks)), dim3(threads), smem, (cudaStream_t)((cudaStream_t)stream), mq, q);
        NASB_CHECK_LAUNCH();
        return 0;
    }
    if (gate) return NASB_ERR_UNSUPPORTED;  // the general strided kernel has no gated epilogue
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
        nasb::launch_pdl((dw_dgrad_strided_tile_kernel<3>), dim3(grid), dim3(256), smem, (cudaStream_t)((cudaStream_t)stream), mz, p);
    } else {
        static bool cfg = false;
        if (!cfg) {
            if (cudaFuncSetAttribute(dw_dgrad_strided_tile_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
                return NASB_ERR_UNSUPPORTED;
            cfg = true;
        }
        nasb::launch_pdl((dw_dgrad_strided_tile_kernel<5>), dim3(grid), dim3(256), smem, (cudaStream_t)((cudaStream_t)stream), mz, p);
    }
    NASB_CHECK_LAUNCH();
    return 0;
}
