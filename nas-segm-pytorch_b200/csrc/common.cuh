// common.cuh -- shared device helpers for libnasb200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nasb200.h"

#define NASB_CHECK_LAUNCH()                      \
    do {                                         \
        cudaError_t e__ = cudaGetLastError();    \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

#define NASB_SM_COUNT 148

#include <stdlib.h>
#include <utility>

namespace nasb {

// ---- programmatic dependent launch (PDL).  Every kernel of the library starts with pdl_sync(): it lets the NEXT kernel of
// the stream be scheduled as soon as all CTAs of this one have started (its CTAs then sit in griddepcontrol.wait), and waits
// itself until the previous kernel has completed and its writes are visible.  Nothing is read before that wait, so the
// data dependency is unchanged; what disappears is the launch latency and ramp-up between the ~700 (training iteration) /
// ~1300 (NAS task-0 iteration) dependent launches of a captured graph.  NASB_PDL=0 launches without the attribute.
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
inline bool pdl_enabled() {
    static int v = -1;
    if (v < 0) v = getenv("NASB_PDL") ? atoi(getenv("NASB_PDL")) : 1;
    return v != 0;
}
// A kernel launched with the attribute becomes resident while its predecessor drains: its CTAs hold shared memory and
// register-file space without working, which costs more than the launch gap it hides once the grid is large (the slots are
// taken from kernels of the concurrent weight-gradient / branch streams, and a capped persistent grid gets packed onto the SMs
// that happened to free first).  So only grids of at most NASB_PDL_MAX_CTAS CTAs ask for it.
inline long long pdl_max_ctas() {
    static long long v = -1;
    if (v < 0) v = getenv("NASB_PDL_MAX_CTAS") ? atoll(getenv("NASB_PDL_MAX_CTAS")) : 1184;
    return v;
}
template <typename... KA, typename... A>
inline void launch_pdl(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl_enabled() && (long long)grid.x * grid.y * grid.z <= pdl_max_ctas()) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);
}

typedef __nv_bfloat16 bf16;

__host__ __device__ inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- element <-> float
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---- V-wide vector access (V elements of T, V*sizeof(T) <= 16 bytes, address must be aligned to that)
template <typename T, int V>
struct Vec {
    T v[V];
};

template <typename T, int V>
__device__ __forceinline__ void load_vec(const T *p, float (&out)[V]) {
    if constexpr (V == 1) {
        out[0] = to_f(p[0]);
    } else if constexpr (sizeof(T) * V == 16) {
        uint4 r = *reinterpret_cast<const uint4 *>(p);
        const T *e = reinterpret_cast<const T *>(&r);
#pragma unroll
        for (int i = 0; i < V; ++i) out[i] = to_f(e[i]);
    } else if constexpr (sizeof(T) * V == 8) {
        uint2 r = *reinterpret_cast<const uint2 *>(p);
        const T *e = reinterpret_cast<const T *>(&r);
#pragma unroll
        for (int i = 0; i < V; ++i) out[i] = to_f(e[i]);
    } else if constexpr (sizeof(T) * V == 4) {
        uint32_t r = *reinterpret_cast<const uint32_t *>(p);
        const T *e = reinterpret_cast<const T *>(&r);
#pragma unroll
        for (int i = 0; i < V; ++i) out[i] = to_f(e[i]);
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) out[i] = to_f(p[i]);
    }
}

template <typename T, int V>
__device__ __forceinline__ void store_vec(T *p, const float (&in)[V]) {
    if constexpr (V == 1) {
        p[0] = from_f<T>(in[0]);
    } else if constexpr (sizeof(T) * V == 16) {
        uint4 r;
        T *e = reinterpret_cast<T *>(&r);
#pragma unroll
        for (int i = 0; i < V; ++i) e[i] = from_f<T>(in[i]);
        *reinterpret_cast<uint4 *>(p) = r;
    } else if constexpr (sizeof(T) * V == 8) {
        uint2 r;
        T *e = reinterpret_cast<T *>(&r);
#pragma unroll
        for (int i = 0; i < V; ++i) e[i] = from_f<T>(in[i]);
        *reinterpret_cast<uint2 *>(p) = r;
    } else if constexpr (sizeof(T) * V == 4) {
        uint32_t r;
        T *e = reinterpret_cast<T *>(&r);
#pragma unroll
        for (int i = 0; i < V; ++i) e[i] = from_f<T>(in[i]);
        *reinterpret_cast<uint32_t *>(p) = r;
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) p[i] = from_f<T>(in[i]);
    }
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == NASB_ACT_RELU) return fmaxf(v, 0.f);
    if (act == NASB_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    return v;
}
// derivative mask evaluated on the activation OUTPUT y (y>0 for relu, 0<y<6 for relu6)
__device__ __forceinline__ float act_mask(float y, int act) {
    if (act == NASB_ACT_RELU) return y > 0.f ? 1.f : 0.f;
    if (act == NASB_ACT_RELU6) return (y > 0.f && y < 6.f) ? 1.f : 0.f;
    return 1.f;
}

// Is a tensor addressable with V-wide vectors along channels?
inline bool vec_ok(const NasbTensor &t, int V) {
    size_t esz = t.dtype == NASB_BF16 ? 2 : 4;
    return (t.c % V == 0) && (t.cstride % V == 0) && ((reinterpret_cast<uintptr_t>(t.ptr) % (esz * V)) == 0);
}
inline long long npix(const NasbTensor &t) { return (long long)t.n * t.h * t.w; }

// PyTorch's bilinear source index, align_corners=False (UpSampleBilinear2d.cu / UpSample.h:
// area_pixel_compute_source_index): src = scale*(dst+0.5)-0.5 clamped below at 0.
struct Lerp {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ Lerp lerp_coord(int dst, float scale, int in_size) {
    float src = scale * (dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    Lerp r;
    r.i0 = (int)src;
    if (r.i0 > in_size - 1) r.i0 = in_size - 1;
    r.i1 = r.i0 + ((r.i0 < in_size - 1) ? 1 : 0);
    r.l1 = src - (float)r.i0;
    r.l0 = 1.f - r.l1;
    return r;
}

// ---- packed helpers: 8 bf16 channels (one 16-byte LDS) -> four float2, and the Blackwell packed fp32 FMA (FFMA2)
__device__ __forceinline__ void cvt8(const uint4 &r, float2 (&v)[4]) {
    v[0] = make_float2(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u));
    v[1] = make_float2(__uint_as_float(r.y << 16), __uint_as_float(r.y & 0xffff0000u));
    v[2] = make_float2(__uint_as_float(r.z << 16), __uint_as_float(r.z & 0xffff0000u));
    v[3] = make_float2(__uint_as_float(r.w << 16), __uint_as_float(r.w & 0xffff0000u));
}
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) {
    uint64_t ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&t);
}

// block-wide sum of doubles through shared memory; result valid in thread 0.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double *sm) {
    int t = threadIdx.x;
    sm[t] = v;
    __syncthreads();
#pragma unroll
    for (int s = NT / 2; s > 0; s >>= 1) {
        if (t < s) sm[t] += sm[t + s];
        __syncthreads();
    }
    double r = sm[0];
    __syncthreads();
    return r;
}

}  // namespace nasb
