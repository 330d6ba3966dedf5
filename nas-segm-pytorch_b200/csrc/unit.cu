// unit.cu -- one C-ABI call per conv -> BN(eval) -> activation (-> residual) unit for inference.
//
// The reference runs nn.Conv2d -> nn.BatchNorm2d.eval() -> nn.ReLU(6) as three modules (src/nn/layer_factory.py:94-114,
// 125-158, 225-265).  The training path of this library composes the same unit from several entry points in Python
// (functional._ConvUnit: BN fold, operand pack, kernel choice, fall-backs) because the backward pass needs the pieces.
// Inference needs none of that, and at batch 1 the ~300 calls of an arch0 forward are bound by host dispatch (~28 us per
// call through Python), not by the GPU: here the fold, the bf16 operand pack and the kernel choice happen behind ONE call.
// No new device code: this file only sequences kernels of the other translation units through their C entry points.
#include "common.cuh"

using namespace nasb;

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
    if (u->gamma || u->running_mean) {  // eval-mode BatchNorm folded to (scale, shift)
        if (!u->running_mean || !u->running_var || !scratch || scratch_bytes < 2LL * cout * (long long)sizeof(float)) return NASB_ERR_BAD_ARG;
        float *ss = (float *)sp;
        int rc = nasb_bn_fold(u->gamma, u->beta, u->running_mean, u->running_var, u->eps, cout, ss, ss + cout, stream);
        if (rc) return rc;
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
            int rc = nasb_pack_weight_bf16(u->weight, cout, cin, 0, pack, stream);
            if (rc) return rc;
            rc = nasb_pw_tc_fwd(x, pack, cout, scale, shift, u->act, res, out, nullptr, stream);
            if (rc != NASB_ERR_UNSUPPORTED) return rc;
        }
        if (u->ks == 3 && u->stride == 1 && u->pad == u->dil && !res && vec_ok(*x, 8) && nasb_conv3_tc_supported(cin, cout) &&
            pack_bytes >= 2 * nasb_pack_conv3_elems(cout, cin, 0)) {
            int rc = nasb_pack_conv3_bf16(u->weight, cout, cin, 0, pack, stream);
            if (rc) return rc;
            rc = nasb_conv3_tc_fwd(x, pack, cout, u->dil, u->pad, scale, shift, u->act, out, nullptr, stream);
            if (rc != NASB_ERR_UNSUPPORTED) return rc;
        }
    }
    return nasb_conv_fwd(x, nullptr, u->weight, u->ks, u->stride, u->dil, u->pad, nullptr, nullptr, u->in_relu, scale, shift,
                         u->act, res, out, stream);
}
