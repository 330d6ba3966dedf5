// tc_common.cuh -- PTX wrappers shared by the tcgen05 / TMA kernels (sm_100a): mbarrier, TMA bulk tensor copies,
// TMEM allocation, tcgen05.mma / commit / ld, shared-memory matrix descriptors and the kind::f16 instruction descriptor.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace nasb {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread = lane = row)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row atoms of 1024 B (SBO), version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address [0,14)
    d |= (uint64_t)1 << 16;                   // leading byte offset (ignored for swizzled K-major) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset [32,46)
    d |= (uint64_t)1 << 46;                   // descriptor version [46,48)
    d |= (uint64_t)2 << 61;                   // SWIZZLE_128B [61,64)
    return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc_bf16(int n) {
    uint32_t d = 0;
    d |= 1u << 4;                    // c_format = F32
    d |= 1u << 7;                    // a_format = BF16
    d |= 1u << 10;                   // b_format = BF16
    d |= (uint32_t)(n >> 3) << 17;   // n_dim
    d |= (uint32_t)(128 >> 4) << 24; // m_dim
    return d;
}


__device__ __forceinline__ void tma_load_4d_sw(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn tc_get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D bf16 map over [rows][inner] with row pitch `pitch_elems`; box = 64 x box_rows, 128-byte swizzle
static inline bool tc_make_map2(CUtensorMap *m, const void *ptr, uint64_t inner, uint64_t rows, uint64_t pitch_elems, uint32_t box_rows) {
    EncodeTiledFn enc = tc_get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {pitch_elems * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 4-D bf16 map (C, W, H, N) over an NHWC tensor; box (64 channels, bw, bh, 1), 128-byte swizzle: the box lands in shared
// memory as bw*bh consecutive 128-byte rows (x fastest) == a K-major / MN-major SWIZZLE_128B UMMA operand tile
static inline bool tc_make_map4(CUtensorMap *m, const NasbTensor *t, int bw, int bh) {
    EncodeTiledFn enc = tc_get_encode();
    if (!enc || bw > 256 || bh > 256) return false;
    cuuint64_t dims[4] = {(cuuint64_t)t->c, (cuuint64_t)t->w, (cuuint64_t)t->h, (cuuint64_t)t->n};
    cuuint64_t strides[3] = {(cuuint64_t)t->cstride * 2, (cuuint64_t)t->w * t->cstride * 2, (cuuint64_t)t->h * t->w * t->cstride * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// MN-major, 128-byte-swizzled operand tile (reduction index = row of the tile): LBO = next 64-element MN block,
// SBO = next group of 8 K (rows)
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;  // leading byte offset: next 64-element MN block
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: next group of 8 K (pixels)
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}


// Sum over the 32 lanes (= 32 pixel rows) of 16 per-lane column values with a transposing butterfly: 16 shuffles instead
// of 80.  Returns, in every lane, the total of column `col` where col = 8*b4 + 4*b3 + 2*b2 + b1 of the lane index.
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane, int &col) {
    float w8[8], w4[4], w2[2];
    bool hi = lane & 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float send = hi ? v[j] : v[j + 8], keep = hi ? v[j + 8] : v[j];
        w8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    hi = lane & 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float send = hi ? w8[j] : w8[j + 4], keep = hi ? w8[j + 4] : w8[j];
        w4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    hi = lane & 4;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float send = hi ? w4[j] : w4[j + 2], keep = hi ? w4[j + 2] : w4[j];
        w2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    hi = lane & 2;
    float send = hi ? w2[0] : w2[1], keep = hi ? w2[1] : w2[0];
    float w1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
    col = ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
    return w1;
}


}  // namespace nasb
