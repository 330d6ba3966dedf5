// sep_tcgen05.cu -- the fused separable primitive (inference): depthwise k x k  ->  [folded BN + activation]  ->  pointwise
// 1x1 on the tensor cores  ->  folded BN + activation (+ residual), ONE kernel, the depthwise output never leaves the SM.
//
//   reference: SepConv's  nn.Conv2d(groups=C) -> nn.Conv2d(1x1) -> BatchNorm2d -> ReLU       (src/nn/layer_factory.py:241-256)
//              InvertedResidual's  dw 3x3 -> BN -> ReLU6 -> 1x1 project -> BN (+ x)          (src/nn/layer_factory.py:141-158)
//
// One CTA (256 threads) owns an 8 x 16 patch of output pixels (= the 128 rows of one UMMA tile) and all output channels
// (C_out <= 64).  Per 64-channel block of the depthwise tensor:
//   * TMA (4-D box, hardware zero fill = the convolution's padding) stages the input patch WITH ITS HALO, double buffered;
//   * the threads compute the depthwise convolution out of shared memory in packed fp32 (thread = 8 channels x a vertical
//     strip of 4 pixels, the register-window scheme of dw_tile_kernel), apply the folded BN / activation that sits between
//     the two convolutions, and write the bf16 result straight into a 128-byte-swizzled K-major A-operand tile;
//   * one thread issues tcgen05.mma (M = 128, N = C_out, K = 64) against the TMA-staged weight block, accumulating in TMEM,
//     while all threads already compute the next channel block into the other A tile (tcgen05.commit releases the tile).
// Epilogue: tcgen05.ld -> scale / shift / activation / residual -> bf16 -> swizzled tile -> one 4-D TMA store (tails clipped).
// HBM traffic = x once (+ halo from L2) + out once: the depthwise tensor's write and re-read (2 x its size) are gone.
#include "tc_common.cuh"

namespace nasb {

constexpr int SEP_TH = 8, SEP_TW = 16, SEP_THREADS = 256;

struct SepP {
    int NI, H, W, C, N;        // images, spatial size (stride 1: output = input size), depthwise channels, output channels
    int nkb;                   // 64-channel blocks of C
    int tiles_x, tiles_y;
    int pad;
    const float *dw_w;         // [C][K*K]
    const float *m_scale, *m_shift;  // between the convolutions (NULL: identity)
    int m_act;
    const float *o_scale, *o_shift;
    int o_act;
    const bf16 *res;
    int res_cs;
};

__device__ __forceinline__ uint4 sep_lds16(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ void sep_sts16(uint32_t a, const uint4 &v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sep_ldw8(uint32_t a, float2 (&w)[4]) {
    const uint4 lo = sep_lds16(a), hi = sep_lds16(a + 16);
    w[0] = make_float2(__uint_as_float(lo.x), __uint_as_float(lo.y));
    w[1] = make_float2(__uint_as_float(lo.z), __uint_as_float(lo.w));
    w[2] = make_float2(__uint_as_float(hi.x), __uint_as_float(hi.y));
    w[3] = make_float2(__uint_as_float(hi.z), __uint_as_float(hi.w));
}

template <int K>
__global__ void __launch_bounds__(SEP_THREADS) sep_tc_kernel(const __grid_constant__ CUtensorMap map_x,
                                                              const __grid_constant__ CUtensorMap map_b,
                                                              const __grid_constant__ CUtensorMap map_o, const SepP p) {
    pdl_sync();
    constexpr int ITH = SEP_TH + K - 1, ITW = SEP_TW + K - 1, KK = K * K;
    constexpr uint32_t XS_BYTES = ((uint32_t)ITH * ITW * 128 + 1023) / 1024 * 1024;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sB = smem;                                      // nkb x [64 x 128 B]   pointwise weights, SWIZZLE_128B
    uint8_t *sA = sB + (size_t)p.nkb * 64 * 128;             // 2 x [128 x 128 B]    depthwise output = A operand (also the out tile)
    uint8_t *xs = sA + 2 * 128 * 128;                        // 2 x XS_BYTES         input patch with halo [ITH][ITW][64], no swizzle
    float *wsm = (float *)(xs + 2 * XS_BYTES);               // [nkb][KK][64]        depthwise weights
    float *s_ms = wsm + (size_t)p.nkb * KK * 64;             // [2][nkb*64]          scale / shift between the convolutions
    float *s_os = s_ms + 2 * p.nkb * 64;                     // [2][64]              output scale / shift
    uint64_t *bar_b = (uint64_t *)(s_os + 128);
    uint64_t *bar_x = bar_b + 1;    // [2]
    uint64_t *bar_m = bar_x + 2;    // [2]
    uint32_t *s_tmem = (uint32_t *)(bar_m + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total = p.tiles_x * p.tiles_y * p.NI;
    const int npb = (p.N + 15) / 16 * 16;

    if (tid == 0) {
        mbar_init(bar_b, 1);
        mbar_init(&bar_x[0], 1);
        mbar_init(&bar_x[1], 1);
        mbar_init(&bar_m[0], 1);
        mbar_init(&bar_m[1], 1);
        fence_barrier_init();
    }
    for (int i = tid; i < p.nkb * KK * 64; i += SEP_THREADS) {
        const int kb = i / (KK * 64), r = i - kb * KK * 64, tap = r >> 6, c = kb * 64 + (r & 63);
        wsm[i] = c < p.C ? p.dw_w[(size_t)c * KK + tap] : 0.f;
    }
    for (int i = tid; i < p.nkb * 64; i += SEP_THREADS) {  // channels beyond C: scale = shift = 0, so act(0) = 0 lands in A
        s_ms[i] = i < p.C ? (p.m_scale ? p.m_scale[i] : 1.f) : 0.f;
        s_ms[p.nkb * 64 + i] = (i < p.C && p.m_shift) ? p.m_shift[i] : 0.f;
    }
    for (int i = tid; i < 64; i += SEP_THREADS) {
        s_os[i] = (p.o_scale && i < p.N) ? p.o_scale[i] : 1.f;
        s_os[64 + i] = (p.o_shift && i < p.N) ? p.o_shift[i] : 0.f;
    }
    if (warp == 0) tmem_alloc(s_tmem, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    auto origin = [&](int t, int &n, int &oy0, int &ox0) {
        const int per = p.tiles_x * p.tiles_y;
        n = t / per;
        const int r = t - n * per, ty = r / p.tiles_x;
        oy0 = ty * SEP_TH;
        ox0 = (r - ty * p.tiles_x) * SEP_TW;
    };
    auto issue_x = [&](int t, int kb, int stage) {  // thread 0
        int n, oy0, ox0;
        origin(t, n, oy0, ox0);
        mbar_expect_tx(&bar_x[stage], (uint32_t)ITH * ITW * 128);
        tma_load_4d_sw(xs + (size_t)stage * XS_BYTES, &map_x, &bar_x[stage], kb * 64, ox0 - p.pad, oy0 - p.pad, n);
    };
    if (tid == 0 && (int)blockIdx.x < total) {
        mbar_expect_tx(bar_b, (uint32_t)(p.nkb * 64 * 128));
        for (int kb = 0; kb < p.nkb; ++kb) tma_load_2d(sB + (size_t)kb * 64 * 128, &map_b, bar_b, kb * 64, 0);
        issue_x(blockIdx.x, 0, 0);
    }

    // depthwise role: 8 channels (cv) x the vertical strip of 4 output pixels (rows sy..sy+3 of column sx)
    const int cv = tid & 7, strip = tid >> 3, sx = strip & 15, sy = (strip >> 4) * 4;
    const uint32_t xs_a[2] = {smem_u32(xs) + cv * 16, smem_u32(xs + XS_BYTES) + cv * 16};
    const uint32_t sA_a[2] = {smem_u32(sA), smem_u32(sA + 128 * 128)};
    const uint32_t w_a = smem_u32(wsm) + cv * 32;
    constexpr uint32_t PXB = 128, ROWB = (uint32_t)ITW * 128;
    const uint32_t idesc = make_idesc_bf16(npb);
    // epilogue role: TMEM lane group = warp & 3, column half = warp >> 2
    const int erow = (warp & 3) * 32 + lane, ecol0 = (warp >> 2) * 32;
    const bool affine = p.o_scale || p.o_shift || p.o_act != NASB_ACT_NONE;

    uint32_t g = 0;  // running index of (tile, channel block) items of this CTA: stage = g & 1
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int n, oy0, ox0;
        origin(t, n, oy0, ox0);
        for (int kb = 0; kb < p.nkb; ++kb, ++g) {
            const int stage = g & 1;
            if (tid == 0) {  // next input patch into the other buffer (released by the __syncthreads that ended item g-1)
                if (kb + 1 < p.nkb) issue_x(t, kb + 1, stage ^ 1);
                else if (t + (int)gridDim.x < total) issue_x(t + gridDim.x, 0, stage ^ 1);
            }
            mbar_wait(&bar_x[stage], (g >> 1) & 1);
            if (g >= 2) mbar_wait(&bar_m[stage], ((g - 2) >> 1) & 1);  // the MMAs that read this A tile two items ago are done
            tc_fence_after();
            // ---- depthwise convolution of this channel block: column windows in registers, packed fp32
            float2 acc[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[q][j] = make_float2(0.f, 0.f);
            const uint32_t wk = w_a + (uint32_t)kb * KK * 256;
            constexpr int ROWS = 3 + K;
#pragma unroll 1
            for (int kx = 0; kx < K; ++kx) {
                float2 wc[K][4];
#pragma unroll
                for (int ky = 0; ky < K; ++ky) sep_ldw8(wk + (uint32_t)(ky * K + kx) * 256, wc[ky]);
                const uint32_t col = xs_a[stage] + (uint32_t)sy * ROWB + (uint32_t)(sx + kx) * PXB;
#pragma unroll
                for (int i = 0; i < ROWS; ++i) {
                    float2 v[4];
                    cvt8(sep_lds16(col + (uint32_t)i * ROWB), v);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int ky = i - q;
                        if (ky >= 0 && ky < K) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[q][j] = ffma2(v[j], wc[ky][j], acc[q][j]);
                        }
                    }
                }
            }
            // ---- folded BN / activation between the convolutions, bf16, into the swizzled K-major A tile
            float2 ms[4], mb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ms[j] = *reinterpret_cast<const float2 *>(s_ms + kb * 64 + cv * 8 + 2 * j);
                mb[j] = *reinterpret_cast<const float2 *>(s_ms + p.nkb * 64 + kb * 64 + cv * 8 + 2 * j);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 e = ffma2(acc[q][j], ms[j], mb[j]);
                    o[j] = pack_bf16x2(apply_act(e.x, p.m_act), apply_act(e.y, p.m_act));
                }
                const int m = (sy + q) * SEP_TW + sx;
                sep_sts16(sA_a[stage] + (uint32_t)m * 128 + (uint32_t)((cv ^ (m & 7)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
            }
            fence_proxy_async();  // the generic-proxy writes above must be visible to the tensor core's async proxy
            __syncthreads();
            if (tid == 0) {
                if (g == 0) mbar_wait(bar_b, 0);
                tc_fence_after();
                const uint64_t ad0 = make_desc_sw128(sA_a[stage]), bd0 = make_desc_sw128(smem_u32(sB + (size_t)kb * 64 * 128));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)  // K steps of 32 bytes = 2 units of the descriptor's start-address field
                    umma_f16(tmem_base, ad0 + (uint64_t)(2 * ks), bd0 + (uint64_t)(2 * ks), idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                umma_commit(&bar_m[stage]);
            }
        }
        // ---- tile epilogue: all MMAs of the tile are complete when the last commit arrives
        const uint32_t gl = g - 1;
        mbar_wait(&bar_m[gl & 1], (gl >> 1) & 1);
        tc_fence_after();
        uint8_t *sO = sA;  // A tile 0 is free now: it becomes the output tile
        if (ecol0 < npb) {
            const int py = erow >> 4, px = erow & 15;
            const bool pix_ok = oy0 + py < p.H && ox0 + px < p.W;
            const long long pix = ((long long)n * p.H + oy0 + py) * p.W + ox0 + px;
#pragma unroll 1
            for (int c0 = ecol0; c0 < ecol0 + 32 && c0 < npb; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, v);
                if (affine) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j] * s_os[c0 + j] + s_os[64 + c0 + j], p.o_act);
                }
                if (p.res && pix_ok) {
                    const bf16 *rp = p.res + pix * p.res_cs + c0;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < p.N) v[j] += __bfloat162float(rp[j]);
                }
                uint4 q0, q1;
                q0.x = pack_bf16x2(v[0], v[1]);
                q0.y = pack_bf16x2(v[2], v[3]);
                q0.z = pack_bf16x2(v[4], v[5]);
                q0.w = pack_bf16x2(v[6], v[7]);
                q1.x = pack_bf16x2(v[8], v[9]);
                q1.y = pack_bf16x2(v[10], v[11]);
                q1.z = pack_bf16x2(v[12], v[13]);
                q1.w = pack_bf16x2(v[14], v[15]);
                uint8_t *orow = sO + (size_t)erow * 128;
                const int ch = c0 >> 3;
                *reinterpret_cast<uint4 *>(orow + ((ch ^ (erow & 7)) << 4)) = q0;
                *reinterpret_cast<uint4 *>(orow + (((ch + 1) ^ (erow & 7)) << 4)) = q1;
            }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();  // tile complete in shared memory; every thread is done with TMEM
        if (tid == 0) {
            tma_store_4d(&map_o, sO, 0, ox0, oy0, n);
            tma_store_commit();
            tma_store_wait_read();  // A tile 0 is written again by the next tile's first channel block
        }
        __syncthreads();
    }
    if (tid == 0) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// 4-D map (C, W, H, N) over an NHWC bf16 tensor, box (64, bw, bh, 1), NO swizzle (the depthwise stage reads it with plain
// 16-byte shared loads; channels beyond C and pixels outside the image are zero-filled)
static bool sep_make_map_x(CUtensorMap *m, const NasbTensor *t, int bw, int bh) {
    EncodeTiledFn enc = tc_get_encode();
    if (!enc || bw > 256 || bh > 256) return false;
    cuuint64_t dims[4] = {(cuuint64_t)t->c, (cuuint64_t)t->w, (cuuint64_t)t->h, (cuuint64_t)t->n};
    cuuint64_t strides[3] = {(cuuint64_t)t->cstride * 2, (cuuint64_t)t->w * t->cstride * 2, (cuuint64_t)t->h * t->w * t->cstride * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static size_t sep_smem(int nkb, int K) {
    const size_t xsb = ((size_t)(SEP_TH + K - 1) * (SEP_TW + K - 1) * 128 + 1023) / 1024 * 1024;
    return (size_t)nkb * 64 * 128 + 2 * 128 * 128 + 2 * xsb + (size_t)nkb * K * K * 64 * 4 + (size_t)2 * nkb * 64 * 4 + 128 * 4 + 128 + 1024;
}

}  // namespace nasb

using namespace nasb;

extern "C" int nasb_sepconv_tc_supported(int C, int N, int ks, int stride, int dil, int pad) {
    if ((ks != 3 && ks != 5) || stride != 1 || dil != 1 || pad != (ks - 1) / 2) return 0;
    if (C < 8 || (C % 8) || N < 8 || (N % 8) || N > 64 || C > 1024) return 0;
    // The depthwise stage works on whole 64-channel blocks: a partly filled last block is computed in full.  Measured on a
    // B200 (profiles/r2_sepconv_kbench.txt): -12 % vs the two kernels at C = 192 and C = 64, +16 % at C = 144 (33 % padding),
    // +45...65 % at C = 32 (100 % padding) -- so the fused kernel takes the shapes with at most 15 % padding.
    const int nkb = (C + 63) / 64;
    if ((nkb * 64 - C) * 100 > 15 * C) return 0;
    return sep_smem(nkb, ks) <= 200 * 1024 ? 1 : 0;
}

extern "C" int nasb_sepconv_tc_fwd(const NasbTensor *x, const float *dw_weight, int ks, int stride, int dil, int pad,
                                   const float *mid_scale, const float *mid_shift, int mid_act, const void *wpack, int N,
                                   const float *out_scale, const float *out_shift, int out_act, const NasbTensor *res,
                                   const NasbTensor *out, void *stream) {
    if (!x || !out || !dw_weight || !wpack) return NASB_ERR_BAD_ARG;
    if (x->dtype != NASB_BF16 || out->dtype != NASB_BF16 || out->c != N || x->n != out->n || x->h != out->h || x->w != out->w)
        return NASB_ERR_UNSUPPORTED;
    if (!nasb_sepconv_tc_supported(x->c, N, ks, stride, dil, pad) || !vec_ok(*x, 8) || !vec_ok(*out, 8)) return NASB_ERR_UNSUPPORTED;
    if (res && (res->dtype != NASB_BF16 || res->c != N || npix(*res) != npix(*out))) return NASB_ERR_BAD_ARG;
    if (npix(*x) == 0) return 0;
    if (x->n > 65535) return NASB_ERR_UNSUPPORTED;
    SepP p{};
    p.NI = x->n, p.H = x->h, p.W = x->w, p.C = x->c, p.N = N;
    p.nkb = (x->c + 63) / 64;
    p.tiles_x = cdiv(x->w, SEP_TW), p.tiles_y = cdiv(x->h, SEP_TH);
    p.pad = pad;
    p.dw_w = dw_weight;
    p.m_scale = mid_scale, p.m_shift = mid_shift, p.m_act = mid_act;
    p.o_scale = out_scale, p.o_shift = out_shift, p.o_act = out_act;
    p.res = res ? (const bf16 *)res->ptr : nullptr;
    p.res_cs = res ? res->cstride : 0;
    const int Kp = (x->c + 7) / 8 * 8;
    CUtensorMap mx, mb, mo;
    if (!sep_make_map_x(&mx, x, SEP_TW + ks - 1, SEP_TH + ks - 1)) return NASB_ERR_UNSUPPORTED;
    if (!tc_make_map2(&mb, wpack, (uint64_t)Kp, (uint64_t)N, (uint64_t)Kp, 64)) return NASB_ERR_UNSUPPORTED;
    if (!tc_make_map4(&mo, out, SEP_TW, SEP_TH)) return NASB_ERR_UNSUPPORTED;
    const size_t smem = sep_smem(p.nkb, ks);
    typedef void (*Kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const SepP);
    Kern kern = ks == 3 ? (Kern)sep_tc_kernel<3> : (Kern)sep_tc_kernel<5>;
    static bool cfg[2] = {false, false};
    if (!cfg[ks == 3 ? 0 : 1]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 2048) != cudaSuccess)
            return NASB_ERR_UNSUPPORTED;
        cfg[ks == 3 ? 0 : 1] = true;
    }
    int per_sm = (int)((220 * 1024) / smem);
    if (per_sm > 3) per_sm = 3;
    if (per_sm < 1) per_sm = 1;
    long long total = (long long)p.tiles_x * p.tiles_y * x->n;
    long long grid = (long long)NASB_SM_COUNT * per_sm;
    if (grid > total) grid = total;
    nasb::launch_pdl((kern), dim3((int)grid), dim3(SEP_THREADS), smem, (cudaStream_t)((cudaStream_t)stream), mx, mb, mo, p);
    NASB_CHECK_LAUNCH();
    return 0;
}
