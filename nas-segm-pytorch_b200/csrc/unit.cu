// unit.cu -- one C-ABI call per conv -> BN(eval) -> activation (-> residual) unit for inference.
//
// The reference runs nn.Conv2d -> nn.BatchNorm2d.eval() -> nn.ReLU(6) as three modules (src/nn/layer_factory.py:94-114,
// 125-158, 225-265).  The training path of this library composes the same unit from several entry points in Python
// (functional._ConvUnit: BN fold, operand pack, kernel choice, fall-backs) because the backward pass needs the pieces.
// Inference needs none of that, and at batch 1 the ~300 calls of an arch0 forward are bound by host dispatch (~28 us per
// call through Python), not by the GPU: here the fold, the bf16 operand pack and the kernel choice happen behind ONE call.
// No new device code: this file only sequences kernels of the other translation units through their C entry points.
#include "common.cuh"

namespace nasb {

// ---- every unit of a model prepared in ONE launch (BN fold + bf16 operand pack), for scopes in which the weights are constant
constexpr int UP_MAX = 160;     // units per launch (by-value table, ~13 KB of kernel parameters)
constexpr int UP_CHUNK = 1024;  // elements per CTA

struct UpTable {
    const float *w[UP_MAX], *gamma[UP_MAX], *beta[UP_MAX], *mean[UP_MAX], *var[UP_MAX];
    float *ss[UP_MAX];
    bf16 *pack[UP_MAX];
    float eps[UP_MAX];
    int co[UP_MAX], ci[UP_MAX];
    int chunk0[UP_MAX + 1];
    signed char kind[UP_MAX];  // -1 none, NASB_PACK_PW, NASB_PACK_C3
    int n;
};

__global__ void __launch_bounds__(256) units_prepare_kernel(const __grid_constant__ UpTable T) {
    pdl_sync();
    int lo = 0, hi = T.n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (T.chunk0[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
    }
    const int u = lo, chunk = (int)blockIdx.x - T.chunk0[u];
    const int co = T.co[u], ci = T.ci[u];
    if (chunk == 0 && T.ss[u]) {  // nasb_bn_fold
        for (int c = threadIdx.x; c < co; c += blockDim.x) {
            const float g = T.gamma[u] ? T.gamma[u][c] : 1.f, b = T.beta[u] ? T.beta[u][c] : 0.f;
            const float sc = g / sqrtf(T.var[u][c] + T.eps[u]);  // bit-identical to bn_fold_kernel
            T.ss[u][c] = sc;
            T.ss[u][co + c] = b - T.mean[u][c] * sc;
        }
    }
    const int kind = T.kind[u];
    if (kind < 0) return;
    const float *w = T.w[u];
    bf16 *out = T.pack[u];
    const int Kp = (ci + 7) / 8 * 8;
    const int e0 = chunk * UP_CHUNK;
    if (kind == NASB_PACK_PW) {
        const int total = co * Kp, e1 = min(total, e0 + UP_CHUNK);
        for (int i = e0 + threadIdx.x; i < e1; i += blockDim.x) {
            const int r = i / Kp, k = i - r * Kp;
            out[i] = __float2bfloat16_rn(k < ci ? w[(long long)r * ci + k] : 0.f);
        }
    } else {  // NASB_PACK_C3, forward: [9][Nr][Kp]
        const int Nr = (co + 63) / 64 * 64, total = 9 * Nr * Kp, e1 = min(total, e0 + UP_CHUNK);
        for (int i = e0 + threadIdx.x; i < e1; i += blockDim.x) {
            const int k = i % Kp, t = i / Kp, n = t % Nr, tap = t / Nr;
            out[i] = __float2bfloat16_rn((n < co && k < ci) ? w[((size_t)n * ci + k) * 9 + tap] : 0.f);
        }
    }
}

}  // namespace nasb

using namespace nasb;

// which operand nasb_conv_unit_infer would pack for this unit (-1: none)
static int unit_pack_kind(const NasbConvUnit *u, int cin, int flags) {
    if (!(flags & NASB_UNIT_TENSOR_CORES) || u->dw || u->in_relu) return -1;
    if (u->ks == 1 && u->stride == 1 && u->pad == 0 && nasb_pw_tc_supported(cin, u->c_out)) return NASB_PACK_PW;
    if (u->ks == 3 && u->stride == 1 && u->pad == u->dil && nasb_conv3_tc_supported(cin, u->c_out)) return NASB_PACK_C3;
    return -1;
}

extern "C" int nasb_conv_units_prepare(const NasbConvUnit *const *units, const int *c_in, void *const *scratch, int n, int flags,
                                       void *stream) {
    if (n < 0 || (n && (!units || !c_in || !scratch))) return NASB_ERR_BAD_ARG;
    static thread_local UpTable T;
    for (int i0 = 0; i0 < n; i0 += UP_MAX) {
        const int m = n - i0 < UP_MAX ? n - i0 : UP_MAX;
        int c = 0;
        for (int i = 0; i < m; ++i) {
            const NasbConvUnit *u = units[i0 + i];
            if (!u || !u->weight || u->c_out <= 0 || c_in[i0 + i] <= 0 || !scratch[i0 + i]) return NASB_ERR_BAD_ARG;
            const int co = u->c_out, ci = c_in[i0 + i];
            const bool bn = u->running_mean != nullptr;
            if (bn && !u->running_var) return NASB_ERR_BAD_ARG;
            char *sp = (char *)scratch[i0 + i];
            const long long ss_bytes = (2LL * co * (long long)sizeof(float) + 255) / 256 * 256;
            T.w[i] = u->weight, T.gamma[i] = u->gamma, T.beta[i] = u->beta, T.mean[i] = u->running_mean, T.var[i] = u->running_var;
            T.ss[i] = bn ? (float *)sp : nullptr;
            T.pack[i] = (bf16 *)(sp + ss_bytes);
            T.eps[i] = u->eps, T.co[i] = co, T.ci[i] = ci;
            const int kind = unit_pack_kind(u, ci, flags);
            T.kind[i] = (signed char)kind;
            long long elems = kind == NASB_PACK_PW ? (long long)co * ((ci + 7) / 8 * 8) : (kind == NASB_PACK_C3 ? nasb_pack_conv3_elems(co, ci, 0) : 0);
            T.chunk0[i] = c;
            int chunks = cdiv(elems, UP_CHUNK);
            c += chunks < 1 ? 1 : chunks;  // at least one CTA per unit (the BN fold)
        }
        T.chunk0[m] = c;
        T.n = m;
        if (c == 0) continue;
        nasb::launch_pdl((units_prepare_kernel), dim3(c), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), T);
        NASB_CHECK_LAUNCH();
    }
    return 0;
}

extern "C" long long nasb_conv_unit_scratch(int c_out, int c_in) {
    if (c_out <= 0 || c_in <= 0) return -1;
    long long pack = nasb_pack_conv3_elems(c_out, c_in, 0);  // the largest operand this unit can need
    long long pw = (long long)c_out * ((c_in + 7) / 8 * 8);
    if (pw > pack) pack = pw;
    return 256 + 2LL * c_out * (long long)sizeof(float) + 256 + pack * 2;
}

extern "C" int nasb_conv_unit_infer(const NasbTensor *x, const NasbConvUnit *u, const NasbTensor *res, const NasbTensor *out,
                                    void *scratch, long long scratch_bytes, int flags, void *stream) {
    if (!x || !u || !out || !u->weight || u->c_out <= 0) return NASB_ERR_BAD_ARG;
    const int cout = u->c_out, cin = x->c;
    const float *scale = nullptr, *shift = u->bias;
    char *sp = (char *)scratch;
    const bool prepared = (flags & NASB_UNIT_PREPARED) != 0;  // fold and pack were done by nasb_conv_units_prepare
    if (u->gamma || u->running_mean) {  // eval-mode BatchNorm folded to (scale, shift)
        if (!u->running_mean || !u->running_var || !scratch || scratch_bytes < 2LL * cout * (long long)sizeof(float)) return NASB_ERR_BAD_ARG;
        float *ss = (float *)sp;
        if (!prepared) {
            int rc = nasb_bn_fold(u->gamma, u->beta, u->running_mean, u->running_var, u->eps, cout, ss, ss + cout, stream);
            if (rc) return rc;
        }
        scale = ss, shift = ss + cout;
    }
    const long long ss_bytes = (2LL * cout * (long long)sizeof(float) + 255) / 256 * 256;
    void *pack = sp ? sp + ss_bytes : nullptr;
    const long long pack_bytes = scratch_bytes - ss_bytes;
    const bool tc = (flags & NASB_UNIT_TENSOR_CORES) != 0, tiles = (flags & NASB_UNIT_TMA_TILES) != 0;
    if (u->dw) {
        if (res) return NASB_ERR_BAD_ARG;
        if (tiles && !u->in_relu && x->dtype == NASB_BF16) {
            int rc = nasb_dwconv_tile(x, u->weight, u->ks, u->stride, u->dil, u->pad, 0, scale, shift, u->act, out, nullptr, stream);
            if (rc != NASB_ERR_UNSUPPORTED) return rc;
        }
        return nasb_dwconv_fwd(x, u->weight, u->ks, u->stride, u->dil, u->pad, u->in_relu, scale, shift, u->act, out, stream);
    }
    if (tc && !u->in_relu && x->dtype == NASB_BF16 && pack) {
        if (u->ks == 1 && u->stride == 1 && u->pad == 0 && out->dtype == NASB_BF16 && vec_ok(*x, 8) && vec_ok(*out, 8) &&
            (!res || (res->dtype == NASB_BF16 && vec_ok(*res, 8))) && nasb_pw_tc_supported(cin, cout) &&
            pack_bytes >= 2LL * cout * ((cin + 7) / 8 * 8)) {
            int rc = prepared ? 0 : nasb_pack_weight_bf16(u->weight, cout, cin, 0, pack, stream);
            if (rc) return rc;
            rc = nasb_pw_tc_fwd(x, pack, cout, scale, shift, u->act, res, out, nullptr, stream);
            if (rc != NASB_ERR_UNSUPPORTED) return rc;
        }
        if (u->ks == 3 && u->stride == 1 && u->pad == u->dil && !res && vec_ok(*x, 8) && nasb_conv3_tc_supported(cin, cout) &&
            pack_bytes >= 2 * nasb_pack_conv3_elems(cout, cin, 0)) {
            int rc = prepared ? 0 : nasb_pack_conv3_bf16(u->weight, cout, cin, 0, pack, stream);
            if (rc) return rc;
            rc = nasb_conv3_tc_fwd(x, pack, cout, u->dil, u->pad, scale, shift, u->act, out, nullptr, stream);
            if (rc != NASB_ERR_UNSUPPORTED) return rc;
        }
    }
    return nasb_conv_fwd(x, nullptr, u->weight, u->ks, u->stride, u->dil, u->pad, nullptr, nullptr, u->in_relu, scale, shift,
                         u->act, res, out, stream);
}

// Depthwise unit followed by a pointwise unit (SepConv repeat, InvertedResidual's depthwise + projection) as ONE kernel
// (csrc/sep_tcgen05.cu) when the shapes allow; NASB_ERR_UNSUPPORTED otherwise (the caller then runs the two units one after
// the other).  scratch_* as for nasb_conv_unit_infer; with NASB_UNIT_PREPARED both must have been prepared.
extern "C" int nasb_sep_unit_infer(const NasbTensor *x, const NasbConvUnit *udw, void *scratch_dw, long long scratch_dw_bytes,
                                   const NasbConvUnit *upw, void *scratch_pw, long long scratch_pw_bytes, const NasbTensor *res,
                                   const NasbTensor *out, int flags, void *stream) {
    if (!x || !udw || !upw || !out || !udw->weight || !upw->weight) return NASB_ERR_BAD_ARG;
    if (!(flags & NASB_UNIT_TENSOR_CORES) || !udw->dw || upw->dw || udw->in_relu || upw->in_relu || udw->bias || upw->bias ||
        upw->ks != 1 || upw->stride != 1 || upw->pad != 0 || udw->c_out != x->c || x->dtype != NASB_BF16 || out->dtype != NASB_BF16)
        return NASB_ERR_UNSUPPORTED;
    const int C = x->c, N = upw->c_out;
    if (!nasb_sepconv_tc_supported(C, N, udw->ks, udw->stride, udw->dil, udw->pad)) return NASB_ERR_UNSUPPORTED;
    const bool prepared = (flags & NASB_UNIT_PREPARED) != 0;
    const float *ms = nullptr, *mb = nullptr, *os = nullptr, *ob = nullptr;
    if (udw->running_mean) {
        if (!udw->running_var || !scratch_dw || scratch_dw_bytes < 2LL * C * (long long)sizeof(float)) return NASB_ERR_BAD_ARG;
        float *ss = (float *)scratch_dw;
        if (!prepared) {
            int rc = nasb_bn_fold(udw->gamma, udw->beta, udw->running_mean, udw->running_var, udw->eps, C, ss, ss + C, stream);
            if (rc) return rc;
        }
        ms = ss, mb = ss + C;
    }
    const long long ss_bytes = (2LL * N * (long long)sizeof(float) + 255) / 256 * 256;
    if (!scratch_pw || scratch_pw_bytes < ss_bytes + 2LL * N * ((C + 7) / 8 * 8)) return NASB_ERR_BAD_ARG;
    if (upw->running_mean) {
        if (!upw->running_var) return NASB_ERR_BAD_ARG;
        float *ss = (float *)scratch_pw;
        if (!prepared) {
            int rc = nasb_bn_fold(upw->gamma, upw->beta, upw->running_mean, upw->running_var, upw->eps, N, ss, ss + N, stream);
            if (rc) return rc;
        }
        os = ss, ob = ss + N;
    }
    void *pack = (char *)scratch_pw + ss_bytes;
    if (!prepared) {
        int rc = nasb_pack_weight_bf16(upw->weight, N, C, 0, pack, stream);
        if (rc) return rc;
    }
    return nasb_sepconv_tc_fwd(x, udw->weight, udw->ks, udw->stride, udw->dil, udw->pad, ms, mb, udw->act, pack, N, os, ob, upw->act,
                               res, out, stream);
}
