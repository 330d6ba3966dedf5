// pw_tcgen05.cu -- pointwise (1x1) convolution as a bf16 GEMM on the 5th-generation tensor cores (sm_100a):
//
//     out[m, n] = act( scale[n] * sum_k A[m, k] * B[n, k] + shift[n] ) (+ res[m, n])          m = pixel, n = C_out
//
//  * A ([pixels, C_in] NHWC activations, K-major) and B (bf16-packed weights [C_out, C_in], K-major) are staged into
//    128-byte-swizzled shared memory by TMA (cp.async.bulk.tensor, zero fill for the K / N / M tails);
//  * tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = C_out rounded to 16, K = 16 per instruction) issued by one
//    thread, fp32 accumulator in TMEM;
//  * the four warps read the accumulator with tcgen05.ld (one TMEM lane = one pixel row per thread), apply the folded
//    BN / bias / activation / residual epilogue, write the bf16 tile to swizzled shared memory and one thread stores it
//    with TMA (tails clipped by the hardware);
//  * optional fused training-mode BatchNorm statistics: per-channel sum / sum-of-squares of the stored tile are
//    accumulated per CTA and flushed once with fp64 atomics (removes the separate statistics pass over z).
//
// The path is HBM-bound (C_in, C_out <= 256): the kernel is deliberately simple -- every CTA loops over its 128-pixel
// tiles (load -> mma -> epilogue, serial per CTA) and latency is hidden by 2-5 co-resident CTAs per SM, each with its
// own TMEM columns.  The same kernel computes the data gradient (A = dz, B = W^T packed by nasb_pack_weight_bf16).
#include <stdlib.h>

#include "tc_common.cuh"

namespace nasb {

struct PwParams {
    int M, K, N;     // pixels, C_in, C_out
    int nkb, nnb;    // K blocks of 64, N blocks of 64 (one N block per CTA)
    int kring;       // > 0: warp-specialised kernel streams (A, B) K blocks through a ring of `kring` stages (large C_in)
    const float *scale, *shift;
    int act;
    const bf16 *res;
    int res_cs;
    double *stats;  // [2][N] (sum, sumsq) or null
    // gate != 0 (data gradient feeding a conv -> BN(train) -> act unit whose pre-BN output is read through map_z): the
    // stored value is g = out gated by g_lo < z*g_scale + g_shift < g_hi and stats receives [sum g, sum g*z] (NasbGate)
    int gate;
    const float *g_scale, *g_shift;
    float g_lo, g_hi;
};

constexpr int TC_THREADS = 128;
constexpr int TILE_M = 128;
constexpr int TILE_N = 64;

// CTA c works on output-channel block nb = c % nnb (64 channels, weights loaded once) and on the 128-pixel tiles
// c / nnb, c / nnb + gridDim.x / nnb, ...  Splitting N keeps the per-CTA footprint small (8 KB of weights per K block,
// a 16 KB output tile, 64 TMEM columns), so 4-5 CTAs are resident per SM and their load / MMA / epilogue / store phases
// overlap; the A tile of a pixel block is re-read by the nnb CTAs that share it out of L2.
__global__ void __launch_bounds__(TC_THREADS) pw_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                           const __grid_constant__ CUtensorMap map_b,
                                                           const __grid_constant__ CUtensorMap map_o,
                                                           const __grid_constant__ CUtensorMap map_z, const PwParams p) {
    pdl_sync();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sB = smem;                                    // nkb x [64 x 128 B]
    uint8_t *sA = sB + (size_t)p.nkb * TILE_N * 128;       // nkb x [128 x 128 B]
    uint8_t *sO = sA + (size_t)p.nkb * TILE_M * 128;       // [128 x 128 B]
    uint8_t *sZ = sO + (size_t)TILE_M * 128;               // [128 x 128 B] gate operand (only when p.gate)
    float *s_scale = (float *)(sZ + (p.gate ? (size_t)TILE_M * 128 : 0));
    float *s_shift = s_scale + TILE_N;
    float *s_sum = s_shift + TILE_N;
    float *s_sq = s_sum + TILE_N;
    uint64_t *bar_b = (uint64_t *)(s_sq + TILE_N);
    uint64_t *bar_a = bar_b + 1;
    uint64_t *bar_mma = bar_a + 1;
    uint64_t *bar_z = bar_mma + 1;
    uint32_t *s_tmem = (uint32_t *)(bar_z + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (p.M + TILE_M - 1) / TILE_M;
    const int nb = (int)blockIdx.x % p.nnb, n0 = nb * TILE_N;
    const int nblk = p.N - n0 < TILE_N ? p.N - n0 : TILE_N;   // valid channels of this block
    const int npb = (nblk + 15) / 16 * 16;                     // MMA N
    const int tile0 = (int)blockIdx.x / p.nnb, tstride = (int)gridDim.x / p.nnb;
    const uint32_t tmem_cols = npb > 32 ? 64u : 32u;

    if (tid == 0) {
        mbar_init(bar_b, 1);
        mbar_init(bar_a, 1);
        mbar_init(bar_mma, 1);
        mbar_init(bar_z, 1);
        fence_barrier_init();
    }
    for (int i = tid; i < TILE_N; i += TC_THREADS) {
        // gated data gradient: the epilogue constants are the GATE's (the accumulator itself goes out unscaled)
        s_scale[i] = p.gate ? (i < nblk ? p.g_scale[n0 + i] : 0.f) : ((p.scale && i < nblk) ? p.scale[n0 + i] : 1.f);
        s_shift[i] = p.gate ? (i < nblk ? p.g_shift[n0 + i] : 0.f) : ((p.shift && i < nblk) ? p.shift[n0 + i] : 0.f);
        s_sum[i] = 0.f;
        s_sq[i] = 0.f;
    }
    if (warp == 0) tmem_alloc(s_tmem, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    auto load_a = [&](int tile) {
        mbar_expect_tx(bar_a, (uint32_t)(p.nkb * TILE_M * 128));
        for (int kb = 0; kb < p.nkb; ++kb) tma_load_2d(sA + (size_t)kb * TILE_M * 128, &map_a, bar_a, kb * 64, tile * TILE_M);
    };
    if (tid == 0 && tile0 < ntiles) {  // this block's weights (once) and the first pixel tile
        mbar_expect_tx(bar_b, (uint32_t)(p.nkb * TILE_N * 128));
        for (int kb = 0; kb < p.nkb; ++kb) tma_load_2d(sB + (size_t)kb * TILE_N * 128, &map_b, bar_b, kb * 64, n0);
        load_a(tile0);
    }
    const uint32_t idesc = make_idesc_bf16(npb);
    const int row = warp * 32 + lane;  // TMEM lane == row of the tile owned by this thread
    const bool affine = !p.gate && (p.scale || p.shift || p.act != NASB_ACT_NONE);
    const int st_ch = tid & 7, st_rg = tid >> 3;  // statistics: this thread's 8-channel chunk and first row
    const bool st_on = st_ch * 8 < nblk;
    float2 st1[4], st2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) st1[j] = st2[j] = make_float2(0.f, 0.f);

    uint32_t it = 0;
    for (int tile = tile0; tile < ntiles; tile += tstride, ++it) {
        const uint32_t parity = it & 1;
        const int m0 = tile * TILE_M;
        if (tid == 0) {
            if (p.gate) {  // sZ was released by the __syncthreads that ended the previous tile
                mbar_expect_tx(bar_z, (uint32_t)(TILE_M * 128));
                tma_load_2d(sZ, &map_z, bar_z, n0, m0);
            }
            if (it == 0) mbar_wait(bar_b, 0);
            mbar_wait(bar_a, parity);
            tc_fence_after();
            const int ksteps = (p.K + 15) / 16;
            // descriptors differ only in the start-address field (16-byte units): base + constant per K step
            const uint64_t ad0 = make_desc_sw128(smem_u32(sA)), bd0 = make_desc_sw128(smem_u32(sB));
            for (int ks = 0; ks < ksteps; ++ks) {
                const uint32_t kb = (uint32_t)ks >> 2, kin = (uint32_t)ks & 3;  // 4 K-steps of 16 elements (32 B) per swizzle row
                umma_f16(tmem_base, ad0 + (uint64_t)(kb * (TILE_M * 128 / 16) + kin * 2),
                         bd0 + (uint64_t)(kb * (TILE_N * 128 / 16) + kin * 2), idesc, ks > 0 ? 1u : 0u);
            }
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, parity);
        tc_fence_after();
        // the A tile has been consumed: prefetch the next one underneath this tile's epilogue
        if (tid == 0 && tile + tstride < ntiles) load_a(tile + tstride);

        // ---- epilogue: TMEM -> registers -> (BN fold / bias, activation, residual) -> bf16 -> swizzled smem
        const long long m = (long long)m0 + row;
        const bool row_ok = m < p.M;
        uint8_t *orow = sO + (size_t)row * 128;
        if (p.gate) mbar_wait(bar_z, parity);
#pragma unroll 1
        for (int c0 = 0; c0 < npb; c0 += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            if (affine) {  // uniform: raw accumulator goes out untouched for training-mode z and for data gradients
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j] * s_scale[c0 + j] + s_shift[c0 + j], p.act);
            }
            if (p.gate) {  // activation mask of the unit that produced this GEMM's "input": pre-activation = z*gs + gb
                const uint8_t *zrow = sZ + (size_t)row * 128;
                float2 zv[8];
                cvt8(*reinterpret_cast<const uint4 *>(zrow + ((((c0 >> 3)) ^ (row & 7)) << 4)), *reinterpret_cast<float2(*)[4]>(&zv[0]));
                cvt8(*reinterpret_cast<const uint4 *>(zrow + ((((c0 >> 3) + 1) ^ (row & 7)) << 4)), *reinterpret_cast<float2(*)[4]>(&zv[4]));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float pa = zv[j].x * s_scale[c0 + 2 * j] + s_shift[c0 + 2 * j];
                    const float pb = zv[j].y * s_scale[c0 + 2 * j + 1] + s_shift[c0 + 2 * j + 1];
                    if (!(pa > p.g_lo && pa < p.g_hi)) v[2 * j] = 0.f;
                    if (!(pb > p.g_lo && pb < p.g_hi)) v[2 * j + 1] = 0.f;
                }
            }
            if (p.res && row_ok) {
                const bf16 *rp = p.res + m * p.res_cs + n0 + c0;
                if (c0 + 16 <= nblk) {
                    float r[8];
                    load_vec<bf16, 8>(rp, r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] += r[j];
                    load_vec<bf16, 8>(rp + 8, r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[8 + j] += r[j];
                } else {
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < nblk) v[j] += __bfloat162float(rp[j]);
                }
            }
            if (p.stats && !row_ok) {  // rows beyond M must not count in the statistics read back from the tile
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
            }
            const int ch = c0 >> 3;  // first 16-byte chunk inside the 128-byte row
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
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();  // tile complete in smem; every thread is done with TMEM
        if (tid == 0) {
            tma_store_2d(&map_o, sO, n0, m0);
            tma_store_commit();
        }
        if (p.stats && st_on) {
            // statistics of the tile AS STORED (bf16), read back from shared memory while the bulk store drains: thread =
            // (16-byte channel chunk, row group), 8 rows per tile, partials stay in registers until the CTA is done
#pragma unroll
            for (int k = 0; k < TILE_M / 16; ++k) {
                const int r = st_rg + 16 * k;
                float2 q[4], zq[4];
                cvt8(*reinterpret_cast<const uint4 *>(sO + (size_t)r * 128 + ((st_ch ^ (r & 7)) << 4)), q);
                if (p.gate) cvt8(*reinterpret_cast<const uint4 *>(sZ + (size_t)r * 128 + ((st_ch ^ (r & 7)) << 4)), zq);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    st1[j].x += q[j].x;
                    st1[j].y += q[j].y;
                    st2[j] = ffma2(q[j], p.gate ? zq[j] : q[j], st2[j]);
                }
            }
        }
        if (tid == 0) tma_store_wait_read();  // smem tile may be overwritten once the bulk store has read it
        __syncthreads();
    }
    if (tid == 0) tma_store_wait_all();
    if (p.stats) {
        // merge the four row groups that share a chunk inside each warp (lanes c, c+8, c+16, c+24), then the warps
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int off = 8; off <= 16; off <<= 1) {
                st1[j].x += __shfl_xor_sync(0xffffffffu, st1[j].x, off);
                st1[j].y += __shfl_xor_sync(0xffffffffu, st1[j].y, off);
                st2[j].x += __shfl_xor_sync(0xffffffffu, st2[j].x, off);
                st2[j].y += __shfl_xor_sync(0xffffffffu, st2[j].y, off);
            }
        }
        if (lane < 8 && st_on) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(&s_sum[st_ch * 8 + 2 * j], st1[j].x);
                atomicAdd(&s_sum[st_ch * 8 + 2 * j + 1], st1[j].y);
                atomicAdd(&s_sq[st_ch * 8 + 2 * j], st2[j].x);
                atomicAdd(&s_sq[st_ch * 8 + 2 * j + 1], st2[j].y);
            }
        }
        __syncthreads();
        for (int c = tid; c < nblk; c += TC_THREADS) {
            atomicAdd(&p.stats[n0 + c], (double)s_sum[c]);
            atomicAdd(&p.stats[p.N + n0 + c], (double)s_sq[c]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}


// ------------------------------------------------------------------------------------------------ warp-specialised variant
// Validated on a B200 in round 2 (full GPU suite + kbench with NASB_PW_WS=1; chosen by nasb_pw_tc_fwd for M <= 2^20 pixels).
// Same math, same epilogue, same statistics as pw_tc_kernel; what changes is the schedule:
// pw_tc_kernel runs load -> MMA -> epilogue -> store serially per CTA and relies on 4-6 co-resident CTAs per SM for overlap
// (small layers: 2-3 tiles per CTA, 11-17 us for 4-33 MB).  Here one CTA pipelines its own tiles:
//   warp 4 (one lane)  producer : TMA of the weight block once, then the A tiles through a ring of WS_SA stages
//   warp 5 (one lane)  MMA      : tcgen05.mma into one of TWO TMEM accumulators; tcgen05.commit releases the A stage and
//                                 hands the accumulator to the epilogue
//   warps 0-3          epilogue : tcgen05.ld (warp w owns TMEM lanes 32w..32w+31), math, bf16 tile into one of two shared
//                                 tiles, TMA store, statistics read-back -- while the MMA warp already fills the other
//                                 accumulator and the producer loads two tiles ahead.
// Barriers (k = use count of the slot): a_full[s] (tx bytes, producer -> MMA), a_empty[s] (commit, MMA -> producer),
// acc_full[a] (commit, MMA -> epilogue), acc_empty[a] (128 arrivals, epilogue -> MMA).  Every k-th completion of a barrier
// requires the waiter of completion k-1 to have passed (producer item i+SA waits a_empty of item i, whose commit follows the
// MMA's wait on a_full of item i; MMA item i+2 waits acc_empty of item i, which the epilogue arrives on after its wait on
// acc_full of item i), so a parity wait can never be overtaken by two phases.
constexpr int WS_THREADS = 192;
constexpr int WS_SA = 2;
constexpr int WS_RING_MAX = 8;                                    // K-ring mode: at most 8 stages
constexpr int WS_RING_STAGE = TILE_M * 128 + TILE_N * 128;       // one A K-block + one B K-block (24 KB)

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

__global__ void __launch_bounds__(WS_THREADS) pw_tc_ws_kernel(const __grid_constant__ CUtensorMap map_a,
                                                              const __grid_constant__ CUtensorMap map_b,
                                                              const __grid_constant__ CUtensorMap map_o, const PwParams p) {
    pdl_sync();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sB = smem;                                               // nkb x [64 x 128 B]
    uint8_t *sA = sB + (size_t)p.nkb * TILE_N * 128;                  // WS_SA x nkb x [128 x 128 B]
    // K-ring mode (C_in too large for resident weights + whole A tiles: MobileNet-v2's 960-channel layers): the same shared
    // memory holds `kring` stages of [A K-block 128 x 128 B | B K-block 64 x 128 B]; weights are re-read per tile out of L2
    uint8_t *sO = p.kring ? smem + (size_t)p.kring * WS_RING_STAGE : sA + (size_t)WS_SA * p.nkb * TILE_M * 128;  // 2 x [128 x 128 B]
    float *s_scale = (float *)(sO + (size_t)2 * TILE_M * 128);
    float *s_shift = s_scale + TILE_N;
    float *s_sum = s_shift + TILE_N;
    float *s_sq = s_sum + TILE_N;
    uint64_t *b_full = (uint64_t *)(s_sq + TILE_N);
    uint64_t *a_full = b_full + 1;          // [WS_SA]
    uint64_t *a_empty = a_full + WS_SA;     // [WS_SA]
    uint64_t *acc_full = a_empty + WS_SA;   // [2]
    uint64_t *acc_empty = acc_full + 2;     // [2]
    uint64_t *r_full = acc_empty + 2;       // [WS_RING_MAX] K-ring mode
    uint64_t *r_empty = r_full + WS_RING_MAX;
    uint32_t *s_tmem = (uint32_t *)(r_empty + WS_RING_MAX);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (p.M + TILE_M - 1) / TILE_M;
    const int nb = (int)blockIdx.x % p.nnb, n0 = nb * TILE_N;
    const int nblk = p.N - n0 < TILE_N ? p.N - n0 : TILE_N;
    const int npb = (nblk + 15) / 16 * 16;
    const int tile0 = (int)blockIdx.x / p.nnb, tstride = (int)gridDim.x / p.nnb;
    const int my_n = tile0 < ntiles ? (ntiles - 1 - tile0) / tstride + 1 : 0;
    const int RS = p.kring;

    if (tid == 0) {
        mbar_init(b_full, 1);
        for (int i = 0; i < WS_SA; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 128);
        }
        for (int i = 0; i < WS_RING_MAX; ++i) {
            mbar_init(&r_full[i], 1);
            mbar_init(&r_empty[i], 1);
        }
        fence_barrier_init();
    }
    for (int i = tid; i < TILE_N; i += WS_THREADS) {
        s_scale[i] = (p.scale && i < nblk) ? p.scale[n0 + i] : 1.f;
        s_shift[i] = (p.shift && i < nblk) ? p.shift[n0 + i] : 0.f;
        s_sum[i] = 0.f;
        s_sq[i] = 0.f;
    }
    if (warp == 5) tmem_alloc(s_tmem, 128);  // two 64-column accumulators
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 4) {
        if (lane == 0 && my_n > 0 && RS > 0) {  // ---- producer, K-ring mode: one (A, B) K block per ring stage
            int g = 0, s = 0;
            uint32_t par = 0;  // parity the producer waits for on r_empty[s] (passes >= 1 over the ring)
            for (int i = 0; i < my_n; ++i) {
                const int tile = tile0 + i * tstride;
                for (int kb = 0; kb < p.nkb; ++kb) {
                    if (g >= RS) mbar_wait(&r_empty[s], par);
                    uint8_t *st = smem + (size_t)s * WS_RING_STAGE;
                    mbar_expect_tx(&r_full[s], (uint32_t)WS_RING_STAGE);
                    tma_load_2d(st, &map_a, &r_full[s], kb * 64, tile * TILE_M);
                    tma_load_2d(st + TILE_M * 128, &map_b, &r_full[s], kb * 64, n0);
                    ++g;
                    if (++s == RS) {
                        s = 0;
                        if (g > RS) par ^= 1;
                    }
                }
            }
        } else if (lane == 0 && my_n > 0) {  // ---- producer
            mbar_expect_tx(b_full, (uint32_t)(p.nkb * TILE_N * 128));
            for (int kb = 0; kb < p.nkb; ++kb) tma_load_2d(sB + (size_t)kb * TILE_N * 128, &map_b, b_full, kb * 64, n0);
            for (int i = 0; i < my_n; ++i) {
                const int s = i % WS_SA, tile = tile0 + i * tstride;
                if (i >= WS_SA) mbar_wait(&a_empty[s], (uint32_t)((i / WS_SA) - 1) & 1);
                mbar_expect_tx(&a_full[s], (uint32_t)(p.nkb * TILE_M * 128));
                for (int kb = 0; kb < p.nkb; ++kb)
                    tma_load_2d(sA + ((size_t)s * p.nkb + kb) * TILE_M * 128, &map_a, &a_full[s], kb * 64, tile * TILE_M);
            }
        }
    } else if (warp == 5) {
        if (lane == 0 && my_n > 0 && RS > 0) {  // ---- MMA issuer, K-ring mode
            const uint32_t idesc = make_idesc_bf16(npb);
            const uint64_t ring_desc0 = make_desc_sw128(smem_u32(smem));
            int s = 0;
            uint32_t par = 0;
            for (int i = 0; i < my_n; ++i) {
                const int a = i & 1;
                const uint32_t acc = tmem_base + (uint32_t)a * 64;
                if (i >= 2) mbar_wait(&acc_empty[a], (uint32_t)((i >> 1) - 1) & 1);
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(&r_full[s], par);
                    tc_fence_after();
                    const int krem = p.K - kb * 64;
                    const int ksteps = krem >= 64 ? 4 : (krem + 15) / 16;
                    const uint64_t ad0 = ring_desc0 + (uint64_t)((uint32_t)s * (WS_RING_STAGE / 16));
                    const uint64_t bd0 = ad0 + (uint64_t)(TILE_M * 128 / 16);
                    for (int ks = 0; ks < ksteps; ++ks)
                        umma_f16(acc, ad0 + (uint64_t)(2 * ks), bd0 + (uint64_t)(2 * ks), idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                    umma_commit(&r_empty[s]);  // the stage may be refilled once these MMAs have read it
                    if (++s == RS) {
                        s = 0;
                        par ^= 1;
                    }
                }
                umma_commit(&acc_full[a]);
            }
        } else if (lane == 0 && my_n > 0) {  // ---- MMA issuer
            const uint32_t idesc = make_idesc_bf16(npb);
            const int ksteps = (p.K + 15) / 16;
            // descriptors differ only in the start-address field (16-byte units): base + constant per stage / K step
            const uint64_t a_desc0 = make_desc_sw128(smem_u32(sA)), b_desc0 = make_desc_sw128(smem_u32(sB));
            mbar_wait(b_full, 0);
            for (int i = 0; i < my_n; ++i) {
                const int s = i % WS_SA, a = i & 1;
                if (i >= 2) mbar_wait(&acc_empty[a], (uint32_t)((i >> 1) - 1) & 1);
                mbar_wait(&a_full[s], (uint32_t)(i / WS_SA) & 1);
                tc_fence_after();
                const uint64_t ad0 = a_desc0 + (uint64_t)((uint32_t)(s * p.nkb) * (TILE_M * 128 / 16));
                const uint32_t acc = tmem_base + (uint32_t)a * 64;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t kb = (uint32_t)ks >> 2, kin = (uint32_t)ks & 3;
                    umma_f16(acc, ad0 + (uint64_t)(kb * (TILE_M * 128 / 16) + kin * 2),
                             b_desc0 + (uint64_t)(kb * (TILE_N * 128 / 16) + kin * 2), idesc, ks > 0 ? 1u : 0u);
                }
                umma_commit(&a_empty[s]);   // the A stage may be refilled once these MMAs have read it
                umma_commit(&acc_full[a]);  // ... and the accumulator is complete
            }
        }
    } else {
        // ---- epilogue warps 0..3 (128 threads): thread = TMEM lane = pixel row of the tile
        const int row = warp * 32 + lane;
        const bool affine = p.scale || p.shift || p.act != NASB_ACT_NONE;
        const int st_ch = tid & 7, st_rg = tid >> 3;
        const bool st_on = st_ch * 8 < nblk;
        float2 st1[4], st2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) st1[j] = st2[j] = make_float2(0.f, 0.f);
        for (int i = 0; i < my_n; ++i) {
            const int a = i & 1, tile = tile0 + i * tstride, m0 = tile * TILE_M;
            uint8_t *sOt = sO + (size_t)a * TILE_M * 128;
            mbar_wait(&acc_full[a], (uint32_t)(i >> 1) & 1);
            tc_fence_after();
            if (tid == 0 && i >= 2) tma_store_wait_read1();  // the bulk store of tile i-2 has finished reading this shared tile
            epi_barrier();
            const long long m = (long long)m0 + row;
            const bool row_ok = m < p.M;
            uint8_t *orow = sOt + (size_t)row * 128;
#pragma unroll 1
            for (int c0 = 0; c0 < npb; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * 64 + c0), v);
                if (affine) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j] * s_scale[c0 + j] + s_shift[c0 + j], p.act);
                }
                if (p.res && row_ok) {
                    const bf16 *rp = p.res + m * p.res_cs + n0 + c0;
                    if (c0 + 16 <= nblk) {
                        float r[8];
                        load_vec<bf16, 8>(rp, r);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] += r[j];
                        load_vec<bf16, 8>(rp + 8, r);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[8 + j] += r[j];
                    } else {
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < nblk) v[j] += __bfloat162float(rp[j]);
                    }
                }
                if (p.stats && !row_ok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                }
                const int ch = c0 >> 3;
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
            tc_fence_before();
            mbar_arrive(&acc_empty[a]);  // this thread is done with accumulator a
            fence_proxy_async();
            epi_barrier();               // tile complete in shared memory
            if (tid == 0) {
                tma_store_2d(&map_o, sOt, n0, m0);
                tma_store_commit();
            }
            if (p.stats && st_on) {
#pragma unroll
                for (int k = 0; k < TILE_M / 16; ++k) {
                    const int r = st_rg + 16 * k;
                    float2 q[4];
                    cvt8(*reinterpret_cast<const uint4 *>(sOt + (size_t)r * 128 + ((st_ch ^ (r & 7)) << 4)), q);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        st1[j].x += q[j].x;
                        st1[j].y += q[j].y;
                        st2[j] = ffma2(q[j], q[j], st2[j]);
                    }
                }
            }
        }
        if (tid == 0) tma_store_wait_all();
        if (p.stats) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int off = 8; off <= 16; off <<= 1) {
                    st1[j].x += __shfl_xor_sync(0xffffffffu, st1[j].x, off);
                    st1[j].y += __shfl_xor_sync(0xffffffffu, st1[j].y, off);
                    st2[j].x += __shfl_xor_sync(0xffffffffu, st2[j].x, off);
                    st2[j].y += __shfl_xor_sync(0xffffffffu, st2[j].y, off);
                }
            }
            if (lane < 8 && st_on) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    atomicAdd(&s_sum[st_ch * 8 + 2 * j], st1[j].x);
                    atomicAdd(&s_sum[st_ch * 8 + 2 * j + 1], st1[j].y);
                    atomicAdd(&s_sq[st_ch * 8 + 2 * j], st2[j].x);
                    atomicAdd(&s_sq[st_ch * 8 + 2 * j + 1], st2[j].y);
                }
            }
            epi_barrier();
            for (int c = tid; c < nblk; c += 128) {
                atomicAdd(&p.stats[n0 + c], (double)s_sum[c]);
                atomicAdd(&p.stats[p.N + n0 + c], (double)s_sq[c]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[co][ci] += sum_m dz[m][co] * x[m][ci].  Both operands are "MN-major" for this GEMM (the reduction index is the
// pixel, the contiguous index is the channel), so the very same TMA boxes (64 channels x 128 pixels, 128-byte swizzle)
// feed tcgen05.mma through MN-major descriptors: LBO = distance between 64-channel blocks, SBO = distance between
// 8-pixel groups, 16 pixels (2 groups) per instruction.  A CTA accumulates its pixel chunks in TMEM (two-stage TMA ring),
// then adds its 128 x C_in fp32 tile to dW with atomics.
struct WgParams {
    int M, Co, Ci;   // pixels, total channels of dz and of x
    int nco;         // 128-channel blocks of dz; blockIdx.y = (x block of 256 channels) * nco + (dz block)
    int tmem_cols;   // accumulator columns of the widest x block
    float *dw;       // [Co][Ci] fp32
};

constexpr int WG_STAGE_A = 2 * TILE_M * 128;  // two 64-channel blocks of dz

__global__ void __launch_bounds__(TC_THREADS) pw_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dz,
                                                                 const __grid_constant__ CUtensorMap map_x, const WgParams p) {
    pdl_sync();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    // Co <= 64: only one 64-channel dz block is staged; the descriptor's second block (LBO) then aliases the first x block and
    // produces accumulator rows 64..127 that the epilogue never reads
    // this CTA's block of the weight gradient: 128 output channels x 256 input channels (one launch covers all blocks, so the
    // 960-channel layers of MobileNet-v2 do not become eight serial latency-bound launches)
    const int co0 = ((int)blockIdx.y % p.nco) * 128, ci0 = ((int)blockIdx.y / p.nco) * 256;
    const int Co = p.Co - co0 < 128 ? p.Co - co0 : 128, Ci = p.Ci - ci0 < 256 ? p.Ci - ci0 : 256;
    const int npad = (Ci + 15) / 16 * 16, nbb = (Ci + 63) / 64;
    float *dw = p.dw + (size_t)co0 * p.Ci + ci0;
    const int a_bytes = Co > 64 ? WG_STAGE_A : TILE_M * 128;
    const int stage_bytes = a_bytes + nbb * TILE_M * 128;
    uint64_t *bars = (uint64_t *)(smem + 2 * (size_t)stage_bytes);  // full[2], done[2], final
    uint32_t *s_tmem = (uint32_t *)(bars + 5);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = (p.M + TILE_M - 1) / TILE_M;
    const int my_n = blockIdx.x < nchunks ? (nchunks - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(s_tmem, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (tid == 0 && my_n > 0) {
        uint32_t idesc = make_idesc_bf16(npad) | (1u << 15) | (1u << 16);  // A and B MN-major
        const uint64_t wg_desc0 = make_desc_mn_sw128(smem_u32(smem), TILE_M * 128);  // stage / K-step descriptors = base + constant
        auto load = [&](int i) {
            const int s = i & 1;
            uint8_t *st = smem + (size_t)s * stage_bytes;
            const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * TILE_M;
            mbar_expect_tx(&bars[s], (uint32_t)stage_bytes);
            tma_load_2d(st, &map_dz, &bars[s], co0, m0);
            if (Co > 64) tma_load_2d(st + TILE_M * 128, &map_dz, &bars[s], co0 + 64, m0);
            for (int b = 0; b < nbb; ++b) tma_load_2d(st + a_bytes + (size_t)b * TILE_M * 128, &map_x, &bars[s], ci0 + b * 64, m0);
        };
        load(0);
        for (int i = 0; i < my_n; ++i) {
            const int s = i & 1;
            if (i + 1 < my_n) {
                if (i >= 1) mbar_wait(&bars[2 + ((i + 1) & 1)], (uint32_t)(((i + 1) >> 1) - 1) & 1);  // stage free again
                load(i + 1);
            }
            mbar_wait(&bars[s], (uint32_t)(i >> 1) & 1);
            tc_fence_after();
            const uint64_t ad0 = wg_desc0 + (uint64_t)((uint32_t)(s * stage_bytes) >> 4), bd0 = ad0 + (uint64_t)((uint32_t)a_bytes >> 4);
#pragma unroll
            for (int ks = 0; ks < TILE_M / 16; ++ks)
                umma_f16(tmem_base, ad0 + (uint64_t)(ks * (2048 / 16)), bd0 + (uint64_t)(ks * (2048 / 16)), idesc,
                         (i > 0 || ks > 0) ? 1u : 0u);
            umma_commit(&bars[2 + s]);
        }
        umma_commit(&bars[4]);  // completes exactly once, after every MMA of this CTA
    }
    if (my_n > 0) {
        // NOT done[last]: a parity wait on a barrier that cycles through many phases can be satisfied by an earlier phase
        mbar_wait(&bars[4], 0);
        tc_fence_after();
        const int co = warp * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < npad; c0 += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            if (co < Co) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < Ci) atomicAdd(&dw[(size_t)co * p.Ci + c0 + j], v[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// weight [rows][cols] fp32 (row-major) -> bf16 [R][Kp]:  transpose=0: out[r][k] = w[r][k] ; transpose=1: out[r][k] = w[k][r]
__device__ __forceinline__ void pack_one(const float *w, int rows, int cols, int transpose, bf16 *out, int R, int Kp, int i) {
    if (i >= R * Kp) return;
    int r = i / Kp, k = i - r * Kp;
    float v = 0.f;
    if (!transpose) {
        if (r < rows && k < cols) v = w[(long long)r * cols + k];
    } else {
        if (k < rows && r < cols) v = w[(long long)k * cols + r];
    }
    out[i] = __float2bfloat16_rn(v);
}
__global__ void pack_weight_kernel(const float *w, int rows, int cols, int transpose, bf16 *out, int R, int Kp) {
    pdl_sync();
    pack_one(w, rows, cols, transpose, out, R, Kp, blockIdx.x * blockDim.x + threadIdx.x);
}
}  // namespace nasb

using namespace nasb;

extern "C" int nasb_pack_weight_bf16(const float *w, int rows, int cols, int transpose, void *out, void *stream) {
    if (!w || !out || rows <= 0 || cols <= 0) return NASB_ERR_BAD_ARG;
    int R = transpose ? cols : rows, K = transpose ? rows : cols;
    int Kp = (K + 7) / 8 * 8;
    nasb::launch_pdl((pack_weight_kernel), dim3(cdiv((long long)R * Kp, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), w, rows, cols, transpose, (bf16 *)out, R, Kp);
    NASB_CHECK_LAUNCH();
    return 0;
}

// Shapes the tensor-core path accepts: bf16 in/out, 8-aligned channels and pitches, C_in / C_out small enough for the
// whole weight matrix plus one A tile and one output tile to fit in shared memory.
static size_t pw_smem_bytes(int nkb) {
    return (size_t)nkb * TILE_N * 128 + (size_t)nkb * TILE_M * 128 + (size_t)TILE_M * 128 + 4 * TILE_N * 4 + 64 + 1024 + 64;
}

extern "C" int nasb_pw_tc_supported(int K, int N) {
    if (K < 8 || N < 8 || (K % 8) || (N % 8) || N > 4096) return 0;
    int nkb = (K + 63) / 64;
    // beyond 7 K blocks the weights + an A tile no longer fit: the warp-specialised kernel streams K blocks through a ring
    return (pw_smem_bytes(nkb) <= 200 * 1024 || K <= 4096) ? 1 : 0;
}

static int pw_tc_launch(const NasbTensor *x, const void *wpack, int N, const float *scale, const float *shift, int act,
                        const NasbTensor *res, const NasbTensor *out, double *stats, const NasbGate *gate, void *stream) {
    if (!x || !out || !wpack) return NASB_ERR_BAD_ARG;
    if (gate) {
        if (!gate->z || !gate->sums || !gate->scale || !gate->shift || stats || res || scale || shift || act != NASB_ACT_NONE)
            return NASB_ERR_BAD_ARG;
        if (gate->z->dtype != NASB_BF16 || gate->z->c != N || npix(*gate->z) != npix(*out) || !vec_ok(*gate->z, 8))
            return NASB_ERR_UNSUPPORTED;
    }
    if (x->dtype != NASB_BF16 || out->dtype != NASB_BF16 || out->c != N || npix(*x) != npix(*out)) return NASB_ERR_BAD_ARG;
    if (!vec_ok(*x, 8) || !vec_ok(*out, 8) || !nasb_pw_tc_supported(x->c, N)) return NASB_ERR_UNSUPPORTED;
    if (res && (res->dtype != NASB_BF16 || res->c != N || npix(*res) != npix(*out) || !vec_ok(*res, 8))) return NASB_ERR_BAD_ARG;
    long long M = npix(*x);
    if (M == 0) return 0;
    if (M > 0x7fffffffLL) return NASB_ERR_UNSUPPORTED;
    PwParams p{};
    p.M = (int)M;
    p.K = x->c;
    p.N = N;
    p.nkb = (p.K + 63) / 64;
    p.nnb = (N + TILE_N - 1) / TILE_N;
    p.scale = scale;
    p.shift = shift;
    p.act = act;
    p.res = res ? (const bf16 *)res->ptr : nullptr;
    p.res_cs = res ? res->cstride : 0;
    p.stats = stats;
    if (gate) {
        p.gate = 1;
        p.stats = gate->sums;
        p.g_scale = gate->scale;
        p.g_shift = gate->shift;
        p.g_lo = gate->act == NASB_ACT_NONE ? -INFINITY : 0.f;
        p.g_hi = gate->act == NASB_ACT_RELU6 ? 6.f : INFINITY;
    }
    int Kp = (p.K + 7) / 8 * 8;
    CUtensorMap ma, mb, mo, mz;
    if (!tc_make_map2(&ma, x->ptr, (uint64_t)p.K, (uint64_t)M, (uint64_t)x->cstride, TILE_M)) return NASB_ERR_UNSUPPORTED;
    if (!tc_make_map2(&mb, wpack, (uint64_t)Kp, (uint64_t)N, (uint64_t)Kp, TILE_N)) return NASB_ERR_UNSUPPORTED;
    if (!tc_make_map2(&mo, out->ptr, (uint64_t)N, (uint64_t)M, (uint64_t)out->cstride, TILE_M)) return NASB_ERR_UNSUPPORTED;
    mz = mo;
    if (gate && !tc_make_map2(&mz, gate->z->ptr, (uint64_t)N, (uint64_t)M, (uint64_t)gate->z->cstride, TILE_M))
        return NASB_ERR_UNSUPPORTED;
    // Schedule choice (measured on a B200, profiles/r2_switches_kbench.txt): the warp-specialised kernel wins 5-30 % up to
    // 2^20 pixels (8 x 256 x 512) and loses 5-20 % on the 512 x 1024 maps, where its second output tile costs occupancy.
    // NASB_PW_WS=0 / 1 forces one schedule (kernel comparisons in tools/kbench.py).
    static int ws_mode = -1;
    if (ws_mode < 0) ws_mode = getenv("NASB_PW_WS") ? atoi(getenv("NASB_PW_WS")) : 2;
    int ntiles = (int)((M + TILE_M - 1) / TILE_M);
    const bool big_k = pw_smem_bytes(p.nkb) > 200 * 1024;  // K-ring mode of the warp-specialised kernel
    if (big_k && gate) return NASB_ERR_UNSUPPORTED;
    if (!gate && (big_k || ws_mode == 1 || (ws_mode == 2 && M <= (1LL << 20)))) {  // warp-specialised schedule (see pw_tc_ws_kernel)
        size_t smem_ws = (size_t)p.nkb * TILE_N * 128 + (size_t)WS_SA * p.nkb * TILE_M * 128 + (size_t)2 * TILE_M * 128 +
                         4 * TILE_N * 4 + 256 + 1024;
        if (big_k || smem_ws > 200 * 1024) {
            p.kring = p.nkb > 7 ? 6 : 0;
            if (p.kring) smem_ws = (size_t)p.kring * WS_RING_STAGE + (size_t)2 * TILE_M * 128 + 4 * TILE_N * 4 + 256 + 1024;
        }
        if (smem_ws <= 200 * 1024) {
            static bool configured_ws = false;
            if (!configured_ws) {
                cudaError_t e = cudaFuncSetAttribute(pw_tc_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024 + 2048));
                if (e != cudaSuccess) return (int)e;
                configured_ws = true;
            }
            int per_sm = (int)((220 * 1024) / smem_ws);
            if (per_sm > 4) per_sm = 4;  // 128 TMEM columns per CTA
            if (per_sm < 1) per_sm = 1;
            long long grid = (long long)NASB_SM_COUNT * per_sm / p.nnb * p.nnb;
            if (grid < p.nnb) grid = p.nnb;
            if (grid > (long long)ntiles * p.nnb) grid = (long long)ntiles * p.nnb;
            nasb::launch_pdl((pw_tc_ws_kernel), dim3((int)grid), dim3(WS_THREADS), smem_ws, (cudaStream_t)((cudaStream_t)stream), ma, mb, mo, p);
            NASB_CHECK_LAUNCH();
            return 0;
        }
    }
    size_t smem = pw_smem_bytes(p.nkb) + (gate ? (size_t)TILE_M * 128 : 0);
    if (smem > 200 * 1024) return NASB_ERR_UNSUPPORTED;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(pw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024 + 2048));
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    int per_sm = (int)((220 * 1024) / smem);
    if (per_sm > 6) per_sm = 6;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)NASB_SM_COUNT * per_sm / p.nnb * p.nnb;  // a multiple of the N blocks
    if (grid < p.nnb) grid = p.nnb;
    if (grid > (long long)ntiles * p.nnb) grid = (long long)ntiles * p.nnb;
    nasb::launch_pdl((pw_tc_kernel), dim3((int)grid), dim3(TC_THREADS), smem, (cudaStream_t)((cudaStream_t)stream), ma, mb, mo, mz, p);
    NASB_CHECK_LAUNCH();
    return 0;
}

extern "C" int nasb_pw_tc_fwd(const NasbTensor *x, const void *wpack, int N, const float *scale, const float *shift, int act,
                              const NasbTensor *res, const NasbTensor *out, double *stats, void *stream) {
    return pw_tc_launch(x, wpack, N, scale, shift, act, res, out, stats, nullptr, stream);
}

// Data gradient dx = dz . W (transposed pack) whose epilogue gates dx with the activation mask of the unit that produced the
// convolution's input and accumulates that unit's BatchNorm-backward reductions (NasbGate).
extern "C" int nasb_pw_tc_dgrad_gated(const NasbTensor *dz, const void *wpack_t, int N, const NasbGate *gate, const NasbTensor *dx,
                                      void *stream) {
    if (!gate) return NASB_ERR_BAD_ARG;
    return pw_tc_launch(dz, wpack_t, N, nullptr, nullptr, NASB_ACT_NONE, nullptr, dx, nullptr, gate, stream);
}

extern "C" int nasb_pw_tc_wgrad_supported(int Co, int Ci) {
    // any multiple of 8 up to 4096: the launcher walks 128-channel blocks of dz and 256-channel blocks of x
    if (Co < 8 || Ci < 8 || (Co % 8) || (Ci % 8) || Ci > 4096 || Co > 4096) return 0;
    return 1;
}

// dweight[co][ci] += sum over pixels dz[.,co] * x[.,ci]   (1x1 convolution, fp32 [C_out][C_in] layout)
extern "C" int nasb_pw_tc_wgrad(const NasbTensor *x, const NasbTensor *dz, float *dweight, void *stream) {
    if (!x || !dz || !dweight) return NASB_ERR_BAD_ARG;
    if (x->dtype != NASB_BF16 || dz->dtype != NASB_BF16 || npix(*x) != npix(*dz)) return NASB_ERR_BAD_ARG;
    if (!vec_ok(*x, 8) || !vec_ok(*dz, 8) || !nasb_pw_tc_wgrad_supported(dz->c, x->c)) return NASB_ERR_UNSUPPORTED;
    long long M = npix(*x);
    if (M == 0) return 0;
    if (M > 0x7fffffffLL) return NASB_ERR_UNSUPPORTED;
    const int Ci = x->c, Co = dz->c;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(pw_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024 + 2048));
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    int nchunks = (int)((M + TILE_M - 1) / TILE_M);
    WgParams p{};
    p.M = (int)M;
    p.Co = Co;
    p.Ci = Ci;
    p.nco = (Co + 127) / 128;
    const int nci = (Ci + 255) / 256, nblocks = p.nco * nci;
    const int ci_max = Ci < 256 ? Ci : 256, co_max = Co < 128 ? Co : 128;
    int cols = 32;
    while (cols < (ci_max + 15) / 16 * 16) cols <<= 1;
    p.tmem_cols = cols;
    p.dw = dweight;
    CUtensorMap mx, mdz;
    if (!tc_make_map2(&mx, x->ptr, (uint64_t)Ci, (uint64_t)M, (uint64_t)x->cstride, TILE_M)) return NASB_ERR_UNSUPPORTED;
    if (!tc_make_map2(&mdz, dz->ptr, (uint64_t)Co, (uint64_t)M, (uint64_t)dz->cstride, TILE_M)) return NASB_ERR_UNSUPPORTED;
    const size_t smem = 2 * ((size_t)(co_max > 64 ? WG_STAGE_A : TILE_M * 128) + (size_t)((ci_max + 63) / 64) * TILE_M * 128) + 64 + 1024;
    int per_sm = (int)((220 * 1024) / smem);
    if (per_sm > 4) per_sm = 4;
    if (per_sm * p.tmem_cols > 512) per_sm = 512 / p.tmem_cols;
    if (per_sm < 1) per_sm = 1;
    // every CTA ends with (its block of) Co x Ci atomics onto the same addresses: at least 16 pixel chunks per CTA, but >= 32
    // CTAs per launch -- measured on B200: 2048 chunks 49 -> 27 us, 128 chunks 24 -> 14 us, large M unchanged.  The CTAs are
    // shared out over the channel blocks.
    int want = nchunks / 16 < 32 ? 32 : nchunks / 16;
    if (want > NASB_SM_COUNT * per_sm) want = NASB_SM_COUNT * per_sm;
    int gx = want / nblocks;
    if (gx > nchunks) gx = nchunks;
    if (gx < 1) gx = 1;
    nasb::launch_pdl((pw_wgrad_tc_kernel), dim3(gx, nblocks), dim3(TC_THREADS), smem, (cudaStream_t)((cudaStream_t)stream), mdz, mx, p);
    NASB_CHECK_LAUNCH();
    return 0;
}
