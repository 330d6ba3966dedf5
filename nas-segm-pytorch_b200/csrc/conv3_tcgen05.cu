// conv3_tcgen05.cu -- dense 3x3 convolution (stride 1, any dilation) as an implicit GEMM on the 5th-generation tensor
// cores.  The nine taps are nine shifted views of the same NHWC tensor: a 4-D TMA box (64 channels, TW, TH, 1) started at
// (x0 + dx, y0 + dy) lands in shared memory as 128 consecutive 128-byte rows -- exactly the K-major SWIZZLE_128B operand
// tile tcgen05.mma wants -- and pixels outside the image are zero-filled by the hardware, which IS the convolution's zero
// padding.  So no im2col buffer exists anywhere: HBM sees the input once (halo re-reads hit L2) and the output once.
//   forward / data gradient : D[128 pixels, 64 c_out] += A_tap[128, 64] * B_tap[64, 64]^T over (tap, C_in block), TMEM
//                             accumulator, ring of TMA stages with look-ahead across patches, fused epilogue
//                             (BN fold / bias / activation, optional BN statistics), bf16 TMA store or fp32 direct store.
//                             The data gradient is the same kernel on dz with the transposed + flipped weight pack.
//   weight gradient         : dW[:, :, tap] = x_tap^T dz.  Two taps share one M=128 instruction (rows 0-63 = tap 2j,
//                             rows 64-127 = tap 2j+1, MN-major descriptors), 5 TMEM accumulators live across all the
//                             CTA's patches, one atomic flush at the end.
#include <stdlib.h>

#include "tc_common.cuh"

namespace nasb {

constexpr int C3_THREADS = 128;   // weight-gradient kernel
constexpr int C3F_THREADS = 320;  // forward / data-gradient kernel: warps 0-7 epilogue (two groups), warp 8 TMA producer, warp 9 MMA issuer
constexpr int C3_TILE = 128;      // pixels per patch
constexpr int C3_NB = 64;         // output channels per CTA

struct C3Params {
    int NI, H, W;        // images, height, width (input == output size: stride 1, pad == dil*(3-1)/2 .. general pad below)
    int K, N;            // C_in, C_out of this GEMM
    int nkb, nnb;        // K blocks of 64, N blocks of 64
    int dil, pad;
    int TH, TW, tiles_x, tiles_y;
    int stages;
    int halo;            // 1: row-halo staging (TH == 1): three (TW + 2*dil)-pixel row boxes per (patch, K block) feed all nine taps
    int rowbuf;          // bytes between the row buffers of one halo stage (multiple of 1024)
    int nr;              // rows per tap in the packed weight (N rounded up to 64)
    int brows;           // rows per tap kept in shared memory (32 when N <= 32, else 64)
    const float *scale, *shift;
    int act;
    void *out;           // fp32 output when out_f32, else the TMA map is used
    int out_cs, out_f32;
    int f32_staged;      // fp32 output with out_cs == N: the patch is a contiguous span per image row, written from a staging tile
    double *stats;
};

__device__ __forceinline__ void c3_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void c3_epi_barrier(int group) {
    if (group == 0)
        asm volatile("bar.sync 1, 128;" ::: "memory");
    else
        asm volatile("bar.sync 2, 128;" ::: "memory");
}

// Warp-specialised: the producer streams (patch, tap / row-halo, K block) items through a ring of `stages` shared-memory
// slots, the MMA issuer accumulates patch i into TMEM accumulator i & 1, and epilogue group i & 1 (four warps each, its own
// accumulator, output tile and named barrier) drains it.  Two groups because ONE warp per SM sub-partition cannot hide its own
// ALU latency: ncu showed the single-group epilogue 69 % busy at ~0.25 instructions per clock, the MMA warp waiting on it.
// Barriers: full[s] (tx bytes, producer -> MMA), done[s] (commit, MMA -> producer), acc_full[a] (commit, MMA -> epilogue),
// acc_empty[a] (128 arrivals, epilogue -> MMA).  Completion k of each barrier requires the waiter of completion k-1 to have
// passed (item g+S waits done of item g, committed after the MMA's wait on full of item g; patch i+2 waits acc_empty of
// patch i, arrived after the epilogue's wait on acc_full of patch i), so a parity wait is never overtaken by two phases.
__global__ void __launch_bounds__(C3F_THREADS) c3_tc_kernel(const __grid_constant__ CUtensorMap map_x,
                                                            const __grid_constant__ CUtensorMap map_b,
                                                            const __grid_constant__ CUtensorMap map_o, const C3Params p) {
    pdl_sync();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int nb_items = 9 * p.nkb;
    const size_t b_tile = (size_t)p.brows * 128;
    uint8_t *sB = smem;                                        // 9*nkb x [brows x 128 B]
    uint8_t *sA = sB + (size_t)nb_items * b_tile;              // stages x stage_bytes
    const size_t stage_bytes = p.halo ? (size_t)3 * p.rowbuf : (size_t)C3_TILE * 128;
    uint8_t *sO = sA + (size_t)p.stages * stage_bytes;          // per epilogue group: bf16 [128 x 128 B] or staged fp32 [128 x N]
    const size_t so_tile = p.out_f32 ? (p.f32_staged ? (((size_t)C3_TILE * p.N * 4 + 1023) & ~(size_t)1023) : 0)
                                     : (size_t)C3_TILE * 128;
    float *s_scale = (float *)(sO + 2 * so_tile);
    float *s_shift = s_scale + C3_NB;
    float *s_sum = s_shift + C3_NB;
    float *s_sq = s_sum + C3_NB;
    uint64_t *bar_b = (uint64_t *)(s_sq + C3_NB);
    uint64_t *acc_full = bar_b + 1;    // [2]
    uint64_t *acc_empty = acc_full + 2;  // [2]
    uint64_t *full = acc_empty + 2;    // [stages <= 8]
    uint64_t *done = full + 8;         // [stages <= 8]
    uint32_t *s_tmem = (uint32_t *)(done + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = (int)blockIdx.x % p.nnb, n0 = nb * C3_NB;
    const int nblk = p.N - n0 < C3_NB ? p.N - n0 : C3_NB;
    const int npb = (nblk + 15) / 16 * 16;
    const int tiles_img = p.tiles_x * p.tiles_y, total_patches = tiles_img * p.NI;
    const int patch0 = (int)blockIdx.x / p.nnb, pstride = (int)gridDim.x / p.nnb;
    const int my_patches = patch0 < total_patches ? (total_patches - 1 - patch0) / pstride + 1 : 0;
    const uint32_t acc_cols = npb > 32 ? 64u : 32u;
    const int S = p.stages;

    if (tid == 0) {
        mbar_init(bar_b, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 128);
        }
        for (int i = 0; i < S; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&done[i], 1);
        }
        fence_barrier_init();
    }
    for (int i = tid; i < C3_NB; i += C3F_THREADS) {
        s_scale[i] = (p.scale && i < nblk) ? p.scale[n0 + i] : 1.f;
        s_shift[i] = (p.shift && i < nblk) ? p.shift[n0 + i] : 0.f;
        s_sum[i] = 0.f;
        s_sq[i] = 0.f;
    }
    if (warp == 9) tmem_alloc(s_tmem, 2 * acc_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    // patch index -> (image, y0, x0)
    auto patch_origin = [&](int pi, int &n, int &y0, int &x0) {
        const int t = patch0 + pi * pstride;
        n = t / tiles_img;
        const int r = t - n * tiles_img;
        const int ty = r / p.tiles_x;
        y0 = ty * p.TH;
        x0 = (r - ty * p.tiles_x) * p.TW;
    };
    // pipeline item: (patch, tap, K block) in tap mode, (patch, K block) in row-halo mode
    const int items_pp = p.halo ? p.nkb : nb_items;

    if (warp == 8) {
        if (lane == 0 && my_patches > 0) {  // ---- producer
            mbar_expect_tx(bar_b, (uint32_t)(nb_items * b_tile));
            for (int j = 0; j < nb_items; ++j) {
                const int tap = j / p.nkb, kb = j - tap * p.nkb;
                tma_load_2d(sB + (size_t)j * b_tile, &map_b, bar_b, kb * 64, tap * p.nr + n0);
            }
            int g = 0, s = 0;
            uint32_t done_par = 0;  // parity the producer waits for on done[s]: completion (pass - 1) of the slot, pass >= 1
            for (int pi = 0; pi < my_patches; ++pi) {
                int n, y0, x0;
                patch_origin(pi, n, y0, x0);
                for (int j = 0; j < items_pp; ++j, ++g) {
                    if (g >= S) mbar_wait(&done[s], done_par);
                    if (p.halo) {
                        // rows y0 - pad + {0, dil, 2 dil}, pixels x0 - pad .. x0 - pad + TW + 2 dil - 1: tap (ty, tx) is the 128
                        // consecutive 128-byte rows of row buffer ty that start at pixel tx*dil (both TMA and UMMA swizzle on
                        // absolute address bits)
                        const uint32_t row_bytes = (uint32_t)(p.TW + 2 * p.dil) * 128;
                        mbar_expect_tx(&full[s], 3 * row_bytes);
                        for (int ty = 0; ty < 3; ++ty)
                            tma_load_4d_sw(sA + (size_t)s * stage_bytes + (size_t)ty * p.rowbuf, &map_x, &full[s], j * 64,
                                           x0 - p.pad, y0 - p.pad + ty * p.dil, n);
                    } else {
                        const int tap = j / p.nkb, kb = j - tap * p.nkb;
                        mbar_expect_tx(&full[s], C3_TILE * 128);
                        tma_load_4d_sw(sA + (size_t)s * C3_TILE * 128, &map_x, &full[s], kb * 64,
                                       x0 - p.pad + (tap % 3) * p.dil, y0 - p.pad + (tap / 3) * p.dil, n);
                    }
                    if (++s == S) {
                        s = 0;
                        if (g >= S) done_par ^= 1;
                    }
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0 && my_patches > 0) {  // ---- MMA issuer
            // Descriptors are built ONCE: a SWIZZLE_128B K-major descriptor differs between operand tiles only in its 14-bit
            // start-address field (bytes >> 4), so every MMA's descriptors are `base + constant` (the first version rebuilt
            // both descriptors from byte addresses for each of the 36 MMAs of a patch: ~110 dependent uniform-datapath
            // instructions per tap, 3 us per patch -- ncu showed this single thread as the kernel's critical path).
            const uint32_t idesc = make_idesc_bf16(npb);
            const uint64_t a_desc0 = make_desc_sw128(smem_u32(sA)), b_desc0 = make_desc_sw128(smem_u32(sB));
            const uint32_t stage16 = (uint32_t)(stage_bytes >> 4), btile16 = (uint32_t)(b_tile >> 4);
            const uint32_t row16 = (uint32_t)p.rowbuf >> 4, col16 = (uint32_t)p.dil * 8;  // halo: row buffer / tap column step
            mbar_wait(bar_b, 0);
            int g = 0, s = 0;
            uint32_t full_par = 0;  // parity of full[s] for the current pass over the ring
            for (int pi = 0; pi < my_patches; ++pi) {
                const int a = pi & 1;
                const uint32_t acc = tmem_base + (uint32_t)a * acc_cols;
                if (pi >= 2) mbar_wait(&acc_empty[a], (uint32_t)((pi >> 1) - 1) & 1);
                for (int j = 0; j < items_pp; ++j, ++g) {
                    mbar_wait(&full[s], full_par);
                    tc_fence_after();
                    const int kb = p.halo ? j : j % p.nkb;
                    const int krem = p.K - kb * 64;
                    const int ksteps = krem >= 64 ? 4 : (krem + 15) / 16;
                    const uint64_t a_st = a_desc0 + (uint64_t)((uint32_t)s * stage16);
                    if (p.halo) {
                        const uint64_t b_kb = b_desc0 + (uint64_t)((uint32_t)kb * btile16);
                        const uint32_t b_tap = (uint32_t)p.nkb * btile16;
                        if (ksteps == 4) {
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const uint64_t ad = a_st + (uint64_t)((uint32_t)(tap / 3) * row16 + (uint32_t)(tap % 3) * col16);
                                const uint64_t bd = b_kb + (uint64_t)((uint32_t)tap * b_tap);
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    umma_f16(acc, ad + 2 * ks, bd + 2 * ks, idesc, (j > 0 || tap > 0 || ks > 0) ? 1u : 0u);
                            }
                        } else {
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const uint64_t ad = a_st + (uint64_t)((uint32_t)(tap / 3) * row16 + (uint32_t)(tap % 3) * col16);
                                const uint64_t bd = b_kb + (uint64_t)((uint32_t)tap * b_tap);
                                for (int ks = 0; ks < ksteps; ++ks)
                                    umma_f16(acc, ad + 2 * ks, bd + 2 * ks, idesc, (j > 0 || tap > 0 || ks > 0) ? 1u : 0u);
                            }
                        }
                    } else {
                        const uint64_t bd = b_desc0 + (uint64_t)((uint32_t)j * btile16);
                        for (int ks = 0; ks < ksteps; ++ks) umma_f16(acc, a_st + 2 * ks, bd + 2 * ks, idesc, (j > 0 || ks > 0) ? 1u : 0u);
                    }
                    umma_commit(&done[s]);  // the slot may be refilled once these MMAs have read it
                    if (++s == S) {
                        s = 0;
                        full_par ^= 1;
                    }
                }
                umma_commit(&acc_full[a]);
            }
        }
    } else {
        // ---- epilogue group g = warps 4g..4g+3 (128 threads): patches i = g, g+2, ...; thread = TMEM lane = pixel of the patch
        const int grp = warp >> 2, gt = tid & 127;
        const int row = (warp & 3) * 32 + lane;
        uint8_t *sOt = sO + (size_t)grp * so_tile;
        float *sF = reinterpret_cast<float *>(sOt);
        const float hi = p.act == NASB_ACT_RELU6 ? 6.f : INFINITY;
        const uint32_t acc = tmem_base + (((uint32_t)(warp & 3) * 32) << 16) + (uint32_t)grp * acc_cols;
        for (int pi = grp; pi < my_patches; pi += 2) {
            int n, y0, x0;
            patch_origin(pi, n, y0, x0);
            mbar_wait(&acc_full[grp], (uint32_t)(pi >> 1) & 1);
            tc_fence_after();
            if (!p.out_f32) {
                if (gt == 0 && pi >= 2) tma_store_wait_read();  // this group's previous bulk store has finished reading the tile
                c3_epi_barrier(grp);
            } else if (p.f32_staged) {
                c3_epi_barrier(grp);  // the copy-out of this group's previous patch is complete
            }
            const int py = y0 + row / p.TW, px = x0 + row % p.TW;
            const bool row_ok = py < p.H && px < p.W;
#pragma unroll 1
            for (int c0 = 0; c0 < npb; c0 += 16) {
                float v[16];
                tmem_ld16(acc + (uint32_t)c0, v);
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 sc = *reinterpret_cast<const float4 *>(s_scale + c0 + j);
                    const float4 sh = *reinterpret_cast<const float4 *>(s_shift + c0 + j);
                    v[j] = fmaf(v[j], sc.x, sh.x);
                    v[j + 1] = fmaf(v[j + 1], sc.y, sh.y);
                    v[j + 2] = fmaf(v[j + 2], sc.z, sh.z);
                    v[j + 3] = fmaf(v[j + 3], sc.w, sh.w);
                }
                if (p.act != NASB_ACT_NONE) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fminf(fmaxf(v[j], 0.f), hi);
                }
                if (p.stats) {
                    float q[16], q2[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        q[j] = row_ok ? (p.out_f32 ? v[j] : __bfloat162float(__float2bfloat16_rn(v[j]))) : 0.f;
                        q2[j] = q[j] * q[j];
                    }
                    int col;
                    float t1 = warp_colsum16(q, lane, col), t2 = warp_colsum16(q2, lane, col);
                    if (!(lane & 1) && c0 + col < nblk) {
                        atomicAdd(&s_sum[c0 + col], t1);
                        atomicAdd(&s_sq[c0 + col], t2);
                    }
                }
                if (p.out_f32) {
                    if (p.f32_staged) {
                        float *o = sF + (size_t)row * nblk + c0;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < nblk) o[j] = v[j];
                    } else if (row_ok) {
                        float *o = reinterpret_cast<float *>(p.out) + (((size_t)n * p.H + py) * p.W + px) * p.out_cs + n0 + c0;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < nblk) o[j] = v[j];
                    }
                } else {
                    const int ch = c0 >> 3;
                    uint8_t *orow = sOt + (size_t)row * 128;
                    uint4 q0, q1;
                    q0.x = pack_bf16x2(v[0], v[1]);
                    q0.y = pack_bf16x2(v[2], v[3]);
                    q0.z = pack_bf16x2(v[4], v[5]);
                    q0.w = pack_bf16x2(v[6], v[7]);
                    q1.x = pack_bf16x2(v[8], v[9]);
                    q1.y = pack_bf16x2(v[10], v[11]);
                    q1.z = pack_bf16x2(v[12], v[13]);
                    q1.w = pack_bf16x2(v[14], v[15]);
                    *reinterpret_cast<uint4 *>(orow + (((ch) ^ (row & 7)) << 4)) = q0;
                    *reinterpret_cast<uint4 *>(orow + (((ch + 1) ^ (row & 7)) << 4)) = q1;
                }
            }
            tc_fence_before();
            c3_arrive(&acc_empty[grp]);  // this thread is done with the accumulator
            if (!p.out_f32) {
                fence_proxy_async();
                c3_epi_barrier(grp);  // patch complete in shared memory
                if (gt == 0) {
                    tma_store_4d(&map_o, sOt, n0, x0, y0, n);
                    tma_store_commit();
                }
            } else if (p.f32_staged) {
                c3_epi_barrier(grp);
                // out_cs == N: each image row of the patch is ONE contiguous span of (valid pixels x N) floats
                const int vw = p.W - x0 < p.TW ? p.W - x0 : p.TW;
                for (int ty = 0; ty < p.TH && y0 + ty < p.H; ++ty) {
                    const float *src = sF + (size_t)ty * p.TW * nblk;
                    float *dst = reinterpret_cast<float *>(p.out) + (((size_t)n * p.H + y0 + ty) * p.W + x0) * nblk;
                    const int cnt = vw * nblk;
                    if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
                        const int c4 = cnt >> 2;
                        for (int e = gt; e < c4; e += 128)
                            reinterpret_cast<float4 *>(dst)[e] = reinterpret_cast<const float4 *>(src)[e];
                        for (int e = (c4 << 2) + gt; e < cnt; e += 128) dst[e] = src[e];
                    } else {
                        for (int e = gt; e < cnt; e += 128) dst[e] = src[e];
                    }
                }
            }
        }
        if (gt == 0) tma_store_wait_all();
    }
    if (p.stats) {  // shared-memory partials of both groups -> global fp64 sums
        __syncthreads();
        for (int c = tid; c < nblk; c += C3F_THREADS) {
            atomicAdd(&p.stats[n0 + c], (double)s_sum[c]);
            atomicAdd(&p.stats[p.N + n0 + c], (double)s_sq[c]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, 2 * acc_cols);
}

// ------------------------------------------------------------------------------------------------ weight gradient
struct C3WParams {
    int NI, H, W;
    int Ci, Co;          // this launch's channel blocks (<= 64 each)
    int npb;             // Co rounded to 16
    int dil, pad;
    int TH, TW, tiles_x, tiles_y;
    int tmem_cols;
    float *dw;           // full weight gradient [Co_total][Ci_total][3][3]
    int ci_total, co0, ci0;
};

__global__ void __launch_bounds__(C3_THREADS) c3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                 const __grid_constant__ CUtensorMap map_dz, const C3WParams p) {
    pdl_sync();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int S = 2, PAIRS = 5;
    uint8_t *sX = smem;                              // S x [2 taps x 128 x 128 B]
    uint8_t *sZ = sX + (size_t)S * 2 * C3_TILE * 128;  // 2 x [128 x 128 B]
    uint64_t *full = (uint64_t *)(sZ + (size_t)2 * C3_TILE * 128);
    uint64_t *done = full + S;
    uint64_t *final_bar = done + S;
    uint32_t *s_tmem = (uint32_t *)(final_bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_img = p.tiles_x * p.tiles_y, total_patches = tiles_img * p.NI;
    const int my_patches = (int)blockIdx.x < total_patches ? (total_patches - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&done[i], 1);
        }
        mbar_init(final_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(s_tmem, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (tid == 0 && my_patches > 0) {
        const uint32_t idesc = make_idesc_bf16(p.npb) | (1u << 15) | (1u << 16);  // A, B MN-major
        const int total_items = my_patches * PAIRS;
        // producer and consumer cursors of the ONE thread that drives both (item = (patch, tap pair)); plain counters instead
        // of 64-bit divisions, descriptors as base + constant (start-address field in 16-byte units)
        int g = 0, g_s = 0, g_pi = 0, g_j = 0;
        uint32_t g_par = 0;  // parity the producer waits for on done[g_s] (passes >= 1)
        int gn = 0, gy0 = 0, gx0 = 0;
        auto origin = [&](int pi, int &n, int &y0, int &x0) {
            const int t = (int)blockIdx.x + pi * (int)gridDim.x;
            n = t / tiles_img;
            const int r = t - n * tiles_img, ty = r / p.tiles_x;
            y0 = ty * p.TH;
            x0 = (r - ty * p.tiles_x) * p.TW;
        };
        origin(0, gn, gy0, gx0);
        auto issue = [&]() {
            if (g >= S) mbar_wait(&done[g_s], g_par);
            const int ntaps = g_j == PAIRS - 1 ? 1 : 2;
            mbar_expect_tx(&full[g_s], (uint32_t)((ntaps + (g_j == 0 ? 1 : 0)) * C3_TILE * 128));
            for (int q = 0; q < ntaps; ++q) {
                const int tap = 2 * g_j + q;
                tma_load_4d_sw(sX + ((size_t)g_s * 2 + q) * C3_TILE * 128, &map_x, &full[g_s], 0, gx0 - p.pad + (tap % 3) * p.dil,
                               gy0 - p.pad + (tap / 3) * p.dil, gn);
            }
            if (g_j == 0) tma_load_4d_sw(sZ + (size_t)(g_pi & 1) * C3_TILE * 128, &map_dz, &full[g_s], 0, gx0, gy0, gn);
            ++g;
            if (++g_s == S) {
                g_s = 0;
                if (g > S) g_par ^= 1;
            }
            if (++g_j == PAIRS) {
                g_j = 0;
                ++g_pi;
                if (g_pi < my_patches) origin(g_pi, gn, gy0, gx0);
            }
        };
        while (g < total_items && g < S) issue();
        const uint64_t xd0 = make_desc_mn_sw128(smem_u32(sX), C3_TILE * 128), zd0 = make_desc_mn_sw128(smem_u32(sZ), C3_TILE * 128);
        constexpr uint32_t SLOT16 = (2 * C3_TILE * 128) >> 4, ZBUF16 = (C3_TILE * 128) >> 4, KS16 = 2048 >> 4;
        int s = 0, pi = 0, j = 0;
        uint32_t par = 0;
        for (int c = 0; c < total_items; ++c) {
            if (c >= 1 && g < total_items) issue();
            mbar_wait(&full[s], par);
            tc_fence_after();
            const uint64_t ad0 = xd0 + (uint64_t)((uint32_t)s * SLOT16), bd0 = zd0 + (uint64_t)((uint32_t)(pi & 1) * ZBUF16);
            const uint32_t acc = tmem_base + (uint32_t)(j * p.npb);
#pragma unroll
            for (int ks = 0; ks < C3_TILE / 16; ++ks)
                umma_f16(acc, ad0 + (uint64_t)(ks * KS16), bd0 + (uint64_t)(ks * KS16), idesc, (pi > 0 || ks > 0) ? 1u : 0u);
            umma_commit(&done[s]);
            if (++s == S) {
                s = 0;
                par ^= 1;
            }
            if (++j == PAIRS) {
                j = 0;
                ++pi;
            }
        }
        umma_commit(final_bar);  // completes exactly once, after every MMA of this CTA
    }
    if (my_patches > 0) {
        mbar_wait(final_bar, 0);
        tc_fence_after();
        const int row = warp * 32 + lane;
        const int ci = row & 63, half = row >> 6;
        for (int j = 0; j < PAIRS; ++j) {
            const int tap = 2 * j + half;
#pragma unroll 1
            for (int c0 = 0; c0 < p.npb; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * p.npb + c0), v);
                if (tap < 9 && ci < p.Ci) {
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (c0 + q < p.Co)
                            atomicAdd(&p.dw[((size_t)(p.co0 + c0 + q) * p.ci_total + (p.ci0 + ci)) * 9 + tap], v[q]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// weight [Co][Ci][3][3] fp32 -> bf16 [9][Nr][Kp]
//   mode 0 (forward)      : row n = co, col k = ci, tap unchanged
//   mode 1 (data gradient): row n = ci, col k = co, tap flipped (8 - tap)
__global__ void pack_conv3_kernel(const float *w, int Co, int Ci, int mode, bf16 *out, int Nr, int Kp) {
    pdl_sync();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * Nr * Kp) return;
    int k = i % Kp, t = i / Kp, n = t % Nr, tap = t / Nr;
    float v = 0.f;
    if (mode == 0) {
        if (n < Co && k < Ci) v = w[((size_t)n * Ci + k) * 9 + tap];
    } else {
        if (n < Ci && k < Co) v = w[((size_t)k * Ci + n) * 9 + (8 - tap)];
    }
    out[i] = __float2bfloat16_rn(v);
}

// 128-pixel patch TH x TW.  Rows of at least 64 pixels use one image row per patch (TH == 1): that enables the row-halo ring,
// which stages 3 row boxes per patch instead of 9 tap boxes -- worth more than the edge pixels a 128-wide patch wastes on
// e.g. W = 88.  Narrower images take the shape that wastes the fewest edge pixels (W = 44: 8 x 16), ties to the wider patch.
static void pick_patch(int H, int W, int &TH, int &TW) {
    if (W >= 64) {
        TH = 1;
        TW = 128;
        return;
    }
    long long best = -1;
    for (int tw = 64; tw >= 16; tw >>= 1) {
        const int th = C3_TILE / tw;
        const long long area = (long long)cdiv(W, tw) * tw * ((long long)cdiv(H, th) * th);
        if (best < 0 || area < best) {
            best = area;
            TH = th;
            TW = tw;
        }
    }
}

// shared memory of the forward kernel: weight taps + ring + output staging + constants/barriers + alignment slack
static size_t c3_smem(int nkb, int stages, size_t stage_bytes = (size_t)C3_TILE * 128, int brows = C3_NB,
                      size_t so_bytes = (size_t)2 * C3_TILE * 128) {
    return (size_t)9 * nkb * brows * 128 + (size_t)stages * stage_bytes + so_bytes + 4 * C3_NB * 4 + 256 + 1024;
}

}  // namespace nasb

using namespace nasb;

extern "C" long long nasb_pack_conv3_elems(int Co, int Ci, int mode) {
    int N = mode == 0 ? Co : Ci, K = mode == 0 ? Ci : Co;
    return 9LL * ((N + 63) / 64 * 64) * ((K + 7) / 8 * 8);
}

extern "C" int nasb_pack_conv3_bf16(const float *w, int Co, int Ci, int mode, void *out, void *stream) {
    if (!w || !out || Co <= 0 || Ci <= 0) return NASB_ERR_BAD_ARG;
    int N = mode == 0 ? Co : Ci, K = mode == 0 ? Ci : Co;
    int Nr = (N + 63) / 64 * 64, Kp = (K + 7) / 8 * 8;
    nasb::launch_pdl((pack_conv3_kernel), dim3(cdiv(9LL * Nr * Kp, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), w, Co, Ci, mode, (bf16 *)out, Nr, Kp);
    NASB_CHECK_LAUNCH();
    return 0;
}

// K = channels of x, N = channels of out.  Requires stride 1 and out spatial size == x spatial size (pad == dil).
extern "C" int nasb_conv3_tc_supported(int K, int N) {
    if (K < 1 || N < 1 || N > 4096) return 0;
    int nkb = (K + 63) / 64;
    return c3_smem(nkb, 2) <= 216 * 1024 ? 1 : 0;
}

extern "C" int nasb_conv3_tc_fwd(const NasbTensor *x, const void *wpack, int N, int dil, int pad, const float *scale,
                                 const float *shift, int act, const NasbTensor *out, double *stats, void *stream) {
    if (!x || !out || !wpack) return NASB_ERR_BAD_ARG;
    if (x->dtype != NASB_BF16 || (out->dtype != NASB_BF16 && out->dtype != NASB_F32) || out->c != N) return NASB_ERR_UNSUPPORTED;
    if (x->n != out->n || x->h != out->h || x->w != out->w || x->h + 2 * pad - 2 * dil != out->h) return NASB_ERR_UNSUPPORTED;
    // TMA needs a 16-byte aligned base and a 16-byte multiple pixel pitch; the channel COUNT may be anything
    if (((uintptr_t)x->ptr & 15) || (x->cstride % 8) || !nasb_conv3_tc_supported(x->c, N)) return NASB_ERR_UNSUPPORTED;
    if (out->dtype == NASB_BF16 && (((uintptr_t)out->ptr & 15) || (out->cstride % 8))) return NASB_ERR_UNSUPPORTED;
    if (npix(*x) == 0) return 0;
    C3Params p{};
    p.NI = x->n;
    p.H = x->h;
    p.W = x->w;
    p.K = x->c;
    p.N = N;
    p.nkb = (p.K + 63) / 64;
    p.nnb = (N + C3_NB - 1) / C3_NB;
    p.dil = dil;
    p.pad = pad;
    pick_patch(p.H, p.W, p.TH, p.TW);
    p.tiles_x = cdiv(p.W, p.TW);
    p.tiles_y = cdiv(p.H, p.TH);
    // Row-halo staging when a patch is one image row (TH == 1): 3 boxes per (patch, K block) instead of 9 -- the TMA unit
    // moves ~one 16 KB box per L2 round trip per SM, which is what bounded the tap-per-box pipeline (12 us per patch).
    static int halo_on = -1;
    if (halo_on < 0) halo_on = getenv("NASB_C3_HALO") ? atoi(getenv("NASB_C3_HALO")) : 1;
    p.halo = (halo_on && p.TH == 1 && p.TW + 2 * dil <= 256) ? 1 : 0;
    p.rowbuf = (int)((((size_t)(p.TW + 2 * dil) * 128) + 1023) / 1024 * 1024);
    p.out_f32 = out->dtype == NASB_F32 ? 1 : 0;
    p.f32_staged = (p.out_f32 && out->cstride == N && N <= C3_NB) ? 1 : 0;
    p.brows = N <= 32 ? 32 : C3_NB;
    // two output tiles (one per epilogue group); preference: staged fp32 output before direct stores, row-halo ring before
    // tap ring; each combination needs >= 2 ring stages to fit
    size_t stage_bytes = 0, so_bytes = 0;
    bool placed = false;
    const int want_halo = p.halo;
    for (int staged = p.f32_staged; staged >= 0 && !placed; --staged) {
        so_bytes = p.out_f32 ? (staged ? 2 * (((size_t)C3_TILE * N * 4 + 1023) & ~(size_t)1023) : 0) : (size_t)2 * C3_TILE * 128;
        for (int use_halo = want_halo; use_halo >= 0 && !placed; --use_halo) {
            stage_bytes = use_halo ? (size_t)3 * p.rowbuf : (size_t)C3_TILE * 128;
            if (c3_smem(p.nkb, 2, stage_bytes, p.brows, so_bytes) > 216 * 1024) continue;
            p.halo = use_halo;
            p.f32_staged = staged;
            p.stages = use_halo ? 3 : 8;
            while (p.stages > 2 && c3_smem(p.nkb, p.stages, stage_bytes, p.brows, so_bytes) > 216 * 1024) --p.stages;
            placed = true;
        }
    }
    if (!placed) return NASB_ERR_UNSUPPORTED;
    p.nr = (N + 63) / 64 * 64;
    p.scale = scale;
    p.shift = shift;
    p.act = act;
    p.out = out->ptr;
    p.out_cs = out->cstride;
    p.stats = stats;
    int Kp = (p.K + 7) / 8 * 8;
    CUtensorMap mx, mb, mo;
    if (!tc_make_map4(&mx, x, p.halo ? p.TW + 2 * dil : p.TW, p.TH)) return NASB_ERR_UNSUPPORTED;
    if (!tc_make_map2(&mb, wpack, (uint64_t)Kp, (uint64_t)9 * p.nr, (uint64_t)Kp, (uint32_t)p.brows)) return NASB_ERR_UNSUPPORTED;
    if (p.out_f32) {
        mo = mx;  // unused
    } else if (!tc_make_map4(&mo, out, p.TW, p.TH)) {
        return NASB_ERR_UNSUPPORTED;
    }
    size_t smem = c3_smem(p.nkb, p.stages, stage_bytes, p.brows, so_bytes);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(c3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess)
            return NASB_ERR_UNSUPPORTED;
        configured = true;
    }
    long long total = (long long)p.tiles_x * p.tiles_y * p.NI;
    int per_sm = (int)((220 * 1024) / smem);
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)NASB_SM_COUNT * per_sm / p.nnb * p.nnb;
    if (grid < p.nnb) grid = p.nnb;
    if (grid > total * p.nnb) grid = total * p.nnb;
    nasb::launch_pdl((c3_tc_kernel), dim3((int)grid), dim3(C3F_THREADS), smem, (cudaStream_t)((cudaStream_t)stream), mx, mb, mo, p);
    NASB_CHECK_LAUNCH();
    return 0;
}

// dweight[co][ci][tap] += sum_pixels dz[., co] * x[. + offset(tap), ci]   (fp32 [C_out][C_in][3][3])
extern "C" int nasb_conv3_tc_wgrad(const NasbTensor *x, const NasbTensor *dz, int dil, int pad, float *dweight, void *stream) {
    if (!x || !dz || !dweight) return NASB_ERR_BAD_ARG;
    if (x->dtype != NASB_BF16 || dz->dtype != NASB_BF16) return NASB_ERR_UNSUPPORTED;
    if (x->n != dz->n || x->h != dz->h || x->w != dz->w) return NASB_ERR_UNSUPPORTED;
    if (((uintptr_t)x->ptr & 15) || (x->cstride % 8) || ((uintptr_t)dz->ptr & 15) || (dz->cstride % 8)) return NASB_ERR_UNSUPPORTED;
    if (npix(*x) == 0) return 0;
    C3WParams p{};
    p.NI = x->n;
    p.H = x->h;
    p.W = x->w;
    p.dil = dil;
    p.pad = pad;
    pick_patch(p.H, p.W, p.TH, p.TW);
    p.tiles_x = cdiv(p.W, p.TW);
    p.tiles_y = cdiv(p.H, p.TH);
    p.dw = dweight;
    p.ci_total = x->c;
    size_t smem = (size_t)2 * 2 * C3_TILE * 128 + (size_t)2 * C3_TILE * 128 + 128 + 1024;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(c3_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 2048) != cudaSuccess)
            return NASB_ERR_UNSUPPORTED;
        configured = true;
    }
    long long total = (long long)p.tiles_x * p.tiles_y * p.NI;
    for (int co0 = 0; co0 < dz->c; co0 += 64) {
        for (int ci0 = 0; ci0 < x->c; ci0 += 64) {
            NasbTensor xs = *x, zs = *dz;
            xs.ptr = (bf16 *)x->ptr + ci0;
            xs.c = x->c - ci0 < 64 ? x->c - ci0 : 64;
            zs.ptr = (bf16 *)dz->ptr + co0;
            zs.c = dz->c - co0 < 64 ? dz->c - co0 : 64;
            if (((uintptr_t)xs.ptr & 15) || ((uintptr_t)zs.ptr & 15)) return NASB_ERR_UNSUPPORTED;
            p.Ci = xs.c;
            p.Co = zs.c;
            p.npb = (p.Co + 15) / 16 * 16;
            int cols = 32;
            while (cols < 5 * p.npb) cols <<= 1;
            p.tmem_cols = cols;
            p.co0 = co0;
            p.ci0 = ci0;
            CUtensorMap mx, mz;
            if (!tc_make_map4(&mx, &xs, p.TW, p.TH) || !tc_make_map4(&mz, &zs, p.TW, p.TH)) return NASB_ERR_UNSUPPORTED;
            int per_sm = 512 / cols;
            if (per_sm > 2) per_sm = 2;
            long long grid = (long long)NASB_SM_COUNT * per_sm;
            if (grid > total) grid = total;
            nasb::launch_pdl((c3_wgrad_tc_kernel), dim3((int)grid), dim3(C3_THREADS), smem, (cudaStream_t)((cudaStream_t)stream), mx, mz, p);
            NASB_CHECK_LAUNCH();
        }
    }
    return 0;
}
