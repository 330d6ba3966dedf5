/* nasb200.h -- C ABI of libnasb200.so: the B200 (sm_100a) kernels behind the NAS inner-loop hot path of
 * DrSleep/nas-segm-pytorch (reference paths below are relative to the reference repo root).
 *
 * The reference has no FFI: its "plugin API" is the Python op/cell registry (src/nn/layer_factory.py:27-91),
 * the decoder/encoder modules (src/nn/micro_decoders.py, src/nn/encoders.py), the engine functions
 * (src/engine/trainer.py:17,78,179; src/engine/inference.py:18) and the Cython metric
 * (src/helpers/miou_utils.pyx:7,32,59).  The host mirror of those (nas-segm-pytorch_b200/{nn,engine,helpers})
 * crosses exactly this boundary; every entry point says which reference code it replaces.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless marked host.  Pointers are borrowed; nothing is allocated.
 *  - Activations are NHWC ("channels-last"): element (n,y,x,c) of tensor t is at
 *    t.ptr[((n*h + y)*w + x)*cstride + c]; cstride >= c lets a tensor be a channel slice of a wider buffer
 *    (this is how torch.cat along channels is realised without a copy).
 *  - dtype: NASB_F32 or NASB_BF16 (fp32 accumulation everywhere).  NASB_F32_NCHW marks the planar fp32 image
 *    handed to the encoder stem (the layout the reference's DataLoader produces).
 *  - Parameters (conv weights, BN vectors, biases) are always fp32 in the reference's own layouts
 *    (conv weight [C_out][C_in/groups][kH][kW]); gradients are fp32 and are ACCUMULATED (+=) into the buffer.
 *  - `stream` is a cudaStream_t passed as void*.
 *  - Return value: 0 on success, otherwise a cudaError_t value or NASB_ERR_*.  The Python shim raises
 *    RuntimeError for any non-zero value (reference error convention: src/helpers/utils.py:172-187).
 */
#ifndef NASB200_H_
#define NASB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NASB_F32 0
#define NASB_BF16 1
#define NASB_F32_NCHW 2

#define NASB_ACT_NONE 0
#define NASB_ACT_RELU 1
#define NASB_ACT_RELU6 2

#define NASB_POOL_MAX 0
#define NASB_POOL_AVG 1

#define NASB_ERR_UNSUPPORTED 10001
#define NASB_ERR_BAD_ARG 10002

typedef struct NasbTensor {
    void *ptr;
    int32_t n, h, w, c;
    int32_t cstride;
    int32_t dtype;
} NasbTensor;

/* library / build identification: returns a static string "nasb200 <version> sm_100a" */
const char *nasb_version(void);

/* ---------------------------------------------------------------------------------------------------------
 * Convolution as implicit GEMM (dense k x k, k in {1,3}; any stride / dilation / padding).
 * Replaces nn.Conv2d(+BatchNorm2d eval +ReLU/ReLU6 +residual) chains: layer_factory.py:7-24 (conv3x3/conv1x1),
 * :56-75 (registry conv ops), :94-114 (conv_bn, conv_bn_relu, conv_bn_relu6), :125-158 (InvertedResidual
 * pointwise convs), :316-335 (Adapt), :369-382 (ConcatReduce: BN->ReLU->1x1 over a channel concat),
 * micro_decoders.py:44-45,218-224 (adapt / pre_clf / conv_clf / aux_clf).
 *
 *   out = act( out_scale[co] * sum_{tap,ci} pro(x)[.., ci] * W[co][ci][tap] + out_shift[co] ) (+ res)
 *   pro(x) = in_relu ? relu(x*in_scale[ci]+in_shift[ci]) : x*in_scale+in_shift      (only if in_scale != NULL)
 * x is the channel concatenation of x0 and (optional) x1, which share n,h,w.
 * out_scale / out_shift may be NULL (identity / zero); a conv bias is passed as out_shift with out_scale NULL.
 * -------------------------------------------------------------------------------------------------------*/
int nasb_conv_fwd(const NasbTensor *x0, const NasbTensor *x1, const float *weight, int ks, int stride, int dil,
                  int pad, const float *in_scale, const float *in_shift, int in_relu, const float *out_scale,
                  const float *out_shift, int act, const NasbTensor *res, const NasbTensor *out, void *stream);

/* dx = conv^T(dz): gradient w.r.t. the (concatenated) conv input; dx1 may be NULL.  dx0/dx1 are fully written. */
int nasb_conv_dgrad(const NasbTensor *dz, const float *weight, int ks, int stride, int dil, int pad,
                    const NasbTensor *dx0, const NasbTensor *dx1, void *stream);

/* dweight[co][ci][tap] += sum_pixels dz[.., co] * pro(x)[.., ci];  dweight is fp32 in the reference layout. */
int nasb_conv_wgrad(const NasbTensor *x0, const NasbTensor *x1, const float *in_scale, const float *in_shift,
                    int in_relu, const NasbTensor *dz, int ks, int stride, int dil, int pad, float *dweight,
                    void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Dense 3x3 convolution (stride 1, any dilation, output size == input size) as an implicit GEMM on the tensor cores:
 * nine shifted 4-D TMA boxes per pixel patch (hardware zero fill = zero padding), tcgen05.mma, TMEM accumulation.
 * Registry ops conv3x3 / conv3x3_dil3 / conv3x3_dil12 (layer_factory.py:61-75) and the classifier heads
 * (micro_decoders.py:218-224,224).
 * nasb_pack_conv3_bf16 : fp32 [C_out][C_in][3][3] -> bf16 [9][Nr][Kp]; mode 0 forward, mode 1 data gradient
 *                        (transposed + spatially flipped).  nasb_pack_conv3_elems gives the element count.
 * nasb_conv3_tc_fwd    : x bf16 (any channel count, 16-byte pixel pitch) -> out bf16 or fp32, fused epilogue and
 *                        optional BN statistics; the data gradient is the same call on dz with the mode-1 pack and
 *                        pad' = 2*dil - pad.
 * nasb_conv3_tc_wgrad  : dweight[co][ci][tap] += sum_pixels dz[.,co] * x[.+offset(tap),ci].
 * -------------------------------------------------------------------------------------------------------*/
long long nasb_pack_conv3_elems(int Co, int Ci, int mode);
int nasb_pack_conv3_bf16(const float *w, int Co, int Ci, int mode, void *out, void *stream);
int nasb_conv3_tc_supported(int K, int N);
int nasb_conv3_tc_fwd(const NasbTensor *x, const void *wpack, int N, int dil, int pad, const float *scale,
                      const float *shift, int act, const NasbTensor *out, double *stats, void *stream);
int nasb_conv3_tc_wgrad(const NasbTensor *x, const NasbTensor *dz, int dil, int pad, float *dweight, void *stream);

/* Encoder stem (encoders.py:38): 3x3 convolution of the planar fp32 image (NASB_F32_NCHW, 3 channels) into 32 NHWC
 * channels with the folded-BN / activation epilogue, and its weight gradient.  NASB_ERR_UNSUPPORTED for any other shape
 * (the generic implicit GEMM then runs). */
int nasb_stem_fwd(const NasbTensor *img, const float *weight, int ks, int stride, int dil, int pad, const float *out_scale,
                  const float *out_shift, int act, const NasbTensor *out, void *stream);
int nasb_stem_wgrad(const NasbTensor *img, const NasbTensor *dz, int ks, int stride, int dil, int pad, float *dweight,
                    void *stream);
/* Speed mode: the same stem as a tensor-core GEMM.  nasb_stem_im2col writes the bf16 patch matrix [N, OH, OW, 32]
 * (k = ci*9 + ky*3 + kx, zero for k >= 27); forward = nasb_pw_tc_fwd with the [32][27] weight packed to K = 32,
 * weight gradient = nasb_pw_tc_wgrad on the same matrix (columns 27..31 of the result are zero). */
int nasb_stem_im2col(const NasbTensor *img, int ks, int stride, int dil, int pad, const NasbTensor *out, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Tensor-core path of the pointwise (1x1, stride 1) convolution: bf16 activations, TMA-staged 128-byte-swizzled
 * operands, tcgen05.mma (kind::f16, M=128) with the fp32 accumulator in TMEM, fused epilogue, TMA store.
 * nasb_pack_weight_bf16 : fp32 [rows][cols] -> bf16 [R][Kp] (K padded to 8): transpose=0 for the forward
 *                         (R=C_out, K=C_in), transpose=1 for the data gradient (R=C_in, K=C_out).
 * nasb_pw_tc_supported  : 1 if a (K=C_in, N=C_out) pair fits the kernel (8-aligned, K and N <= 4096; up to 448 input
 *                         channels the weights stay resident in shared memory, beyond that K blocks of A and B
 *                         stream through a ring -- MobileNet-v2's 960-channel layers).
 * nasb_pw_tc_fwd        : out = act(scale*x.W^T + shift) (+res); x/out/res bf16.  stats (optional) -> fp64
 *                         [2][N] sum / sum-of-squares of the stored output, ACCUMULATED (training-mode BN fused
 *                         into the epilogue; finish with nasb_bn_finalize).  The data gradient is the same call
 *                         with x = dz and the transposed pack.
 * -------------------------------------------------------------------------------------------------------*/
int nasb_pack_weight_bf16(const float *w, int rows, int cols, int transpose, void *out, void *stream);
int nasb_pw_tc_supported(int K, int N);
int nasb_pw_tc_wgrad_supported(int Co, int Ci);
/* dweight[co][ci] += sum_pixels dz[.,co]*x[.,ci] on the tensor cores (MN-major operands, TMEM accumulation). */
int nasb_pw_tc_wgrad(const NasbTensor *x, const NasbTensor *dz, float *dweight, void *stream);
int nasb_pw_tc_fwd(const NasbTensor *x, const void *wpack, int N, const float *scale, const float *shift, int act,
                   const NasbTensor *res, const NasbTensor *out, double *stats, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Depthwise k x k convolution (groups == channels), k in {3,5,7}, any stride/dilation/padding, with optional
 * folded BN + activation epilogue.  Replaces the depthwise nn.Conv2d of SepConv / DilConv
 * (layer_factory.py:198-265) and of InvertedResidual (:141-151).   weight: [C][1][k][k].
 * -------------------------------------------------------------------------------------------------------*/
int nasb_dwconv_fwd(const NasbTensor *x, const float *weight, int ks, int stride, int dil, int pad, int in_relu,
                    const float *out_scale, const float *out_shift, int act, const NasbTensor *out, void *stream);
int nasb_dwconv_dgrad(const NasbTensor *dz, const float *weight, int ks, int stride, int dil, int pad,
                      const NasbTensor *dx, void *stream);
int nasb_dwconv_wgrad(const NasbTensor *x, int in_relu, const NasbTensor *dz, int ks, int stride, int dil, int pad,
                      float *dweight, void *stream);
/* TMA-tiled variants (bf16, k in {3,5}): one 4-D TMA box stages the input patch with its halo (zero padding = the
 * hardware's out-of-bounds fill), compute runs out of shared memory.  mode 0 = forward (+ folded BN / activation),
 * mode 1 = stride-1 data gradient (x = dz, out = dx); stats (optional) accumulates fp64 [2][C] sum / sum of squares of the
 * stored output (training-mode BN statistics fused, finish with nasb_bn_finalize).  Return NASB_ERR_UNSUPPORTED for shapes outside their envelope;
 * the host then uses the gather kernels above. */
int nasb_dwconv_tile(const NasbTensor *x, const float *weight, int ks, int stride, int dil, int pad, int mode,
                     const float *out_scale, const float *out_shift, int act, const NasbTensor *out, double *stats,
                     void *stream);
int nasb_dwconv_dgrad_strided_tile(const NasbTensor *dz, const float *weight, int ks, int stride, int dil, int pad,
                                   const NasbTensor *dx, void *stream);
int nasb_dwconv_wgrad_tile(const NasbTensor *x, const NasbTensor *dz, int ks, int stride, int dil, int pad,
                           float *dweight, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Training-mode BatchNorm backward folded into its neighbours (InvertedResidual, layer_factory.py:125-158: 1x1 expand ->
 * BN -> ReLU6 -> depthwise 3x3 -> ...).  For a unit  z = conv(x); y = act(BN_train(z))  the gradient is
 *     dz = s*g + A*z + B,   g = dy * act'(y),  per-channel s, A, B from the two reductions  S1 = sum g, S2 = sum g*z.
 * (1) The kernel that PRODUCES dy -- the depthwise data gradient of the consumer -- gates its output with act'(y) and
 *     accumulates S1, S2 in its epilogue (NasbGate): the reductions cost one extra read of z instead of a pass over dy and z.
 * (2) For a pointwise (1x1) unit z = W x, both consumers of dz are linear in (g, z) and z is linear in x, so dz is never
 *     formed:   dW = diag(s) (g^T x) + diag(A) W (x^T x) + B (sum x)^T        dx = g (diag(s) W) + x (W^T diag(A) W) + B^T W
 *     nasb_pw_bn_bwd_prepare turns S1, S2 and the small matrices g^T x, x^T x, sum x (nasb_pw_tc_wgrad / nasb_channel_sum)
 *     into dW, dgamma, dbeta and the two bf16 operands + bias row of the dx GEMMs (nasb_pw_tc_fwd on g, and on x with the
 *     first result as residual).  BatchNorm's own backward pass over the large tensors disappears.
 * -------------------------------------------------------------------------------------------------------*/
typedef struct NasbGate {
    const NasbTensor *z;         /* pre-BN output of the unit whose input gradient is being produced (geometry of dx) */
    const float *scale, *shift;  /* that unit's folded BN constants: activation input = z*scale + shift */
    int32_t act;                 /* its activation (NASB_ACT_*) */
    int32_t reserved;
    double *sums;                /* fp64 [2][C], ACCUMULATED: S1 = sum g, S2 = sum g*z over the stored (bf16) g */
} NasbGate;
/* dx = gate(dwconv^T(dz)) for stride 1 (tile kernel) and the stride-2 3x3 case; NASB_ERR_UNSUPPORTED otherwise. */
int nasb_dwconv_dgrad_gated(const NasbTensor *dz, const float *weight, int ks, int stride, int dil, int pad,
                            const NasbGate *gate, const NasbTensor *dx, void *stream);
/* weight [c_out][c_in] fp32; scale/mean/rstd = the unit's BN constants (scale = gamma*rstd); sums = S1,S2; P pixels;
 * gx = g^T x [c_out][c_in], xx = x^T x [c_in][c_in], sx = sum x [c_in] (fp32).
 * Outputs: dweight += (see above), dgamma += rstd*(S2 - mean*S1), dbeta += S1 (either may be NULL);
 * pack_g = bf16 [c_in][Kp(c_out)] operand of dx = g.(diag(s)W); pack_x = bf16 [c_in][Kp(c_in)] operand of x.(W^T diag(A) W);
 * bias_row = fp32 [c_in] = B^T W.  W enters the z terms as the bf16 values the forward GEMM multiplied with. */
/* dx = gate(dz . W) on the tensor cores (the pointwise data gradient of the consumer; wpack_t = the transposed operand). */
int nasb_pw_tc_dgrad_gated(const NasbTensor *dz, const void *wpack_t, int N, const NasbGate *gate, const NasbTensor *dx,
                           void *stream);
/* dz pass only, from the reductions a NasbGate epilogue accumulated (any conv type; bf16).  dgamma / dbeta are accumulated.
 * workspace >= nasb_bn_stats_workspace(C) bytes.  NASB_ERR_UNSUPPORTED for layouts outside the packed bf16 kernels. */
int nasb_bn_bwd_from_sums(const NasbTensor *dy, const NasbTensor *z, int act, const float *scale, const float *shift,
                          const float *save_mean, const float *save_rstd, const double *raw_sums, float *dgamma, float *dbeta,
                          const NasbTensor *dz, void *workspace, void *stream);
int nasb_pw_bn_bwd_prepare(const float *weight, int c_out, int c_in, const float *scale, const float *mean, const float *rstd,
                           const double *sums, long long P, const float *gx, const float *xx, const float *sx,
                           float *dweight, float *dgamma, float *dbeta, void *pack_g, void *pack_x, float *bias_row,
                           void *scratch, void *stream);
long long nasb_pw_bn_bwd_scratch(int c_out); /* bytes of `scratch` */

/* ---------------------------------------------------------------------------------------------------------
 * BatchNorm2d pieces (eps 1e-5, momentum 0.1 in the reference; both are arguments here).
 * nasb_bn_fold      : running stats -> per-channel (scale, shift) for the fused eval-mode epilogues.
 * nasb_bn_stats     : training mode: batch mean / biased var of z over (n,h,w); updates running stats
 *                     (unbiased var, momentum) exactly like nn.BatchNorm2d.train(); also emits (scale, shift).
 *                     workspace: >= nasb_bn_stats_workspace(C) bytes.
 * nasb_affine_act   : y = act(z*scale[c]+shift[c])                       (in place allowed)
 * nasb_bn_act_bwd   : backward of y = act(gamma*xhat+beta):
 *                       g = dy * act'(y); dbeta += sum g; dgamma += sum g*xhat
 *                       xhat = (z-save_mean)*save_rstd in training (z = saved conv output), and is rebuilt from
 *                       y as (y-beta)/gamma in eval mode (only unmasked pixels matter there).  dz may alias dy.
 *                       In training the activation mask is recomputed from z (scale/shift of the forward), so y
 *                       is not read at all (y may be NULL): 2 reads for the sums, 2 reads + 1 write for dz.
 *                       eval : dz = g * scale
 *                       train: dz = scale * (g - mean(g) - xhat*mean(g*xhat))
 *                     workspace: >= nasb_bn_stats_workspace(C) bytes.
 * -------------------------------------------------------------------------------------------------------*/
int nasb_bn_fold(const float *gamma, const float *beta, const float *mean, const float *var, float eps, int C,
                 float *scale, float *shift, void *stream);
long long nasb_bn_stats_workspace(int C);
int nasb_bn_stats(const NasbTensor *z, const float *gamma, const float *beta, float eps, float momentum,
                  float *running_mean, float *running_var, float *save_mean, float *save_rstd, float *scale,
                  float *shift, long long *num_batches_tracked, void *workspace, void *stream);
int nasb_bn_finalize(const double *sums, long long P, int C, const float *gamma, const float *beta, float eps,
                     float momentum, float *running_mean, float *running_var, float *save_mean, float *save_rstd,
                     float *scale, float *shift, long long *num_batches_tracked, void *stream);
int nasb_affine_act(const NasbTensor *z, const float *scale, const float *shift, int act, const NasbTensor *y,
                    void *stream);
/* nasb_bn_finalize + nasb_affine_act as ONE launch (training mode, statistics accumulated by the convolution kernel):
 * every thread derives scale/shift of its channels from the fp64 sums, block 0 publishes them with the saved / running
 * statistics and the step counter; res (optional) is added after the activation (residual blocks: the block output is
 * written directly, y = act(BN(z)) is never materialised -- the backward pass works from z).  Falls back to separate
 * kernels for layouts the fused pass does not cover. */
int nasb_bn_finalize_affine_act(const double *sums, long long P, const NasbTensor *z, const float *gamma, const float *beta,
                                float eps, float momentum, float *running_mean, float *running_var, float *save_mean,
                                float *save_rstd, float *scale, float *shift, long long *num_batches_tracked, int act,
                                const NasbTensor *res, const NasbTensor *y, void *stream);
int nasb_bn_act_bwd(const NasbTensor *dy, const NasbTensor *y, const NasbTensor *z, int act, const float *gamma,
                    const float *beta, const float *scale, const float *shift, const float *save_mean,
                    const float *save_rstd, int training, float *dgamma, float *dbeta, const NasbTensor *dz,
                    void *workspace, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * 3x3 pooling, padding 1 (layer_factory.py:161-178).  Max pooling optionally records the arg-max tap
 * (uint8, same NHW C layout, dense) for the backward pass; avg pooling is count_include_pad=False.
 * -------------------------------------------------------------------------------------------------------*/
int nasb_pool3x3_fwd(const NasbTensor *x, int mode, int stride, const NasbTensor *out, uint8_t *argmax,
                     void *stream);
int nasb_pool3x3_bwd(const NasbTensor *dy, int mode, int stride, const uint8_t *argmax, const NasbTensor *dx,
                     void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Bilinear resize, align_corners=False, up or down (layer_factory.py:338-350, micro_decoders.py:11-25,47-50,
 * trainer.py:141-143,153-155, inference.py:58-60), fused with the aggregation that follows it:
 *    out = [relu]( sa[c] * resize(x) + sb[c] * y )   (sa, sb NULL = 1;  y NULL = no second operand; relu = the F.relu
 *    the decoders apply to the collected concat, micro_decoders.py:251,395)
 * When x already has out's size the resize is the identity.  Covers AggregateCell's sum
 * (micro_decoders.py:51), ParamSum (layer_factory.py:353-366) and writing resized maps into concat slices.
 * nasb_resize_bwd : dx = resize^T(sa * dz)   (deterministic gather form)
 * nasb_axpby_bwd_params : dsa[c] += sum dz*resize(x), dsb[c] += sum dz*y  (ParamSum's a / b gradients)
 * -------------------------------------------------------------------------------------------------------*/
int nasb_resize_axpby(const NasbTensor *x, const float *sa, const NasbTensor *y, const float *sb, int relu,
                      const NasbTensor *out, void *stream);
int nasb_resize_bwd(const NasbTensor *dz, const float *sa, const NasbTensor *dx, void *stream);
int nasb_axpby_bwd_params(const NasbTensor *dz, const NasbTensor *x, const NasbTensor *y, float *dsa, float *dsb,
                          void *workspace, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Small data-movement ops.
 * nasb_scale_copy   : out = s[c] * x (s NULL = 1), dtype conversion allowed (concat-slice writes, dY = b*dz)
 * nasb_channel_tile : Skip / Zero (layer_factory.py:268-297): out[.., r*C+c] = scale * x[.., c] with spatial
 *                     subsampling by `stride` (Zero honours stride, Skip passes 1); scale 0 gives exact zeros.
 * nasb_channel_tile_bwd : dx[.., c] = scale * sum_r dz[.., r*C+c] scattered to the strided positions.
 * nasb_spatial_mean : GAPConv1x1's mean over H then W (layer_factory.py:189): out[n,c] fp32.
 * nasb_spatial_bcast: out[n,y,x,c] = s * v[n,c]  (bilinear resize from a 1x1 map is a broadcast; also GAP bwd)
 * nasb_spatial_sum  : out[n,c] = sum_{y,x} x[n,y,x,c]   (fp32; backward of the broadcast)
 * nasb_channel_sum  : out[c] += sum over all pixels x[..,c]  (conv bias gradient)
 * -------------------------------------------------------------------------------------------------------*/
int nasb_scale_copy(const NasbTensor *x, const float *s, int relu, const NasbTensor *out, void *stream);
int nasb_channel_tile(const NasbTensor *x, int stride, float scale, const NasbTensor *out, void *stream);
int nasb_channel_tile_bwd(const NasbTensor *dz, int stride, float scale, const NasbTensor *dx, void *stream);
int nasb_spatial_mean(const NasbTensor *x, float *out_nc, void *stream);
int nasb_spatial_bcast(const NasbTensor *v_nc, float s, const NasbTensor *out, void *stream);
int nasb_spatial_sum(const NasbTensor *x, float *out_nc, void *stream);
int nasb_channel_sum(const NasbTensor *x, float *out_c, void *workspace, void *stream);
int nasb_relu_bwd(const NasbTensor *dy, const NasbTensor *y, const NasbTensor *dx, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Losses (trainer.py:144-158, main_search.py:435,458).
 * nasb_ce_fwd : LogSoftmax(dim=C) + NLLLoss2d(ignore_index, mean over non-ignored pixels).
 *               target: int64 [n,h,w].  out2[0] = loss (fp32), out2[1] = number of valid pixels (fp32).
 *               workspace >= nasb_loss_workspace() bytes.
 * nasb_ce_bwd : dlogits = gscale * (softmax - onehot) / n_valid   (0 at ignored pixels); n_valid read from out2[1].
 * nasb_mse_*  : nn.MSELoss() (mean over all elements) between x and y.
 * nasb_berhu_*: reverse Huber (defined by this repo): valid = target > valid_min.
 * -------------------------------------------------------------------------------------------------------*/
long long nasb_loss_workspace(void);
int nasb_ce_fwd(const NasbTensor *logits, const int64_t *target, int ignore_index, float *out2, void *workspace,
                void *stream);
int nasb_ce_bwd(const NasbTensor *logits, const int64_t *target, int ignore_index, const float *out2,
                const float *gscale_dev, const NasbTensor *dlogits, void *stream);
int nasb_mse_fwd(const NasbTensor *x, const NasbTensor *y, float *out1, void *workspace, void *stream);
int nasb_mse_bwd(const NasbTensor *x, const NasbTensor *y, const float *gscale_dev, const NasbTensor *dx,
                 void *stream);
int nasb_berhu_fwd(const NasbTensor *pred, const NasbTensor *target, float valid_min, float *out3, void *workspace,
                   void *stream);
int nasb_berhu_bwd(const NasbTensor *pred, const NasbTensor *target, float valid_min, const float *out3,
                   const float *gscale_dev, const NasbTensor *dpred, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Metric: the CUDA replacement of the Cython module src/helpers/miou_utils.pyx.
 * nasb_confmat_labels : fast_cm (:7-30): cm[gt[i]][pred[i]] += 1 for i < n (rows = ground truth), int64,
 *                       ACCUMULATES into cm (caller zeroes).  Entries with gt >= n_classes or
 *                       pred >= n_classes are skipped (the reference caller pre-masks gt < n_classes,
 *                       inference.py:65, so the mask can be fused here).
 * nasb_confmat_logits : inference.py:58-66 fused: bilinear-upsample logits [n,h,w,C] to the label size,
 *                       arg-max over classes (first maximal index), uint8 labels gt [n,H,W], mask
 *                       gt < n_classes, accumulate.  Removes the full-logits device->host copy.
 * nasb_ius_accs       : compute_iu / compute_ius_accs (:32-90): float64 IoU, int64 n_pixels, float64 acc;
 *                       32-bit unsigned accumulators exactly like the reference; default value 2.
 * -------------------------------------------------------------------------------------------------------*/
int nasb_confmat_labels(const uint8_t *pred, const uint8_t *gt, long long n, int n_classes, long long *cm,
                        void *stream);
int nasb_confmat_logits(const NasbTensor *logits, const uint8_t *gt, int H, int W, int n_classes, long long *cm,
                        void *stream);
int nasb_ius_accs(const long long *cm, int n_classes, double *iu, long long *n_pixels, double *accs, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Tail of a training iteration as multi-tensor kernels (scope row f2): torch.nn.utils.clip_grad_norm_ +
 * optimiser.step() + the Polyak average of trainer.py:163-169,258-272 over the ~490 small parameter tensors of a
 * candidate (optimisers built by src/utils/solvers.py:35-52: torch.optim.SGD / torch.optim.Adam).
 * `tensors` / `groups` / `max_norm` are HOST arrays (the table travels by value in the kernel parameters, so a captured
 * CUDA graph bakes it in); every pointer inside a NasbOptTensor is a DEVICE pointer to contiguous fp32.
 *   NasbOptTensor : param; grad (NULL = no update, Polyak only); state1 = SGD momentum buffer / Adam exp_avg;
 *                   state2 = Adam exp_avg_sq; avg = Polyak average (NULL = none); step = Adam's per-parameter step
 *                   counter (fp32 scalar on the device, torch's state["step"]); group = index into groups (-1 = none);
 *                   clip = index of the gradient-norm cell this tensor belongs to (-1 = not clipped).
 *   NasbOptGroup  : kind NASB_OPT_SGD  : beta1 = momentum, beta2 = dampening, first = 1 on the step that creates the
 *                                        momentum buffers (buf = grad), nesterov
 *                   kind NASB_OPT_ADAM : beta1, beta2, eps (amsgrad / maximize / decoupled decay are not supported)
 * nasb_mt_grad_sumsq : cells[c] = sum over the tensors of clip cell c of grad^2 (fp64; cells are zeroed first), and
 *                      step += 1 for every tensor with a step counter.  Must precede nasb_mt_optim_step.
 * nasb_mt_optim_step : grad *= min(1, max_norm[c] / (sqrt(cells[c]) + 1e-6)) (written back, like clip_grad_norm_);
 *                      the optimiser update; avg = avg*polyak_decay + (1 - polyak_decay)*param.
 * nasb_sumsq         : out[0] += sum x^2 over one flat buffer.
 * -------------------------------------------------------------------------------------------------------*/
#define NASB_OPT_NONE 0
#define NASB_OPT_SGD 1
#define NASB_OPT_ADAM 2
typedef struct NasbOptTensor {
    float *param, *grad, *state1, *state2, *avg, *step;
    long long numel;
    int32_t group, clip;
} NasbOptTensor;
typedef struct NasbOptGroup {
    int32_t kind, first, nesterov;
    float lr, beta1, beta2, eps, weight_decay;
} NasbOptGroup;
int nasb_mt_grad_sumsq(const NasbOptTensor *tensors, int n, double *cells, int n_cells, void *stream);
int nasb_mt_optim_step(const NasbOptTensor *tensors, int n, const NasbOptGroup *groups, int n_groups,
                       const float *max_norm, int n_cells, const double *cells, float polyak_decay, void *stream);
int nasb_sumsq(const float *x, long long n, float *out1, void *stream);

/* Every tensor-core weight operand of a model re-packed from the fp32 master weights in ONE launch (the per-call
 * nasb_pack_weight_bf16 / nasb_pack_conv3_bf16 launches were 135 per arch0 iteration).  jobs is a HOST array.
 *   kind NASB_PACK_PW / _PW_T : nasb_pack_weight_bf16 with transpose 0 / 1 (src [c_out][c_in])
 *   kind NASB_PACK_C3 / _C3_T : nasb_pack_conv3_bf16 with mode 0 / 1       (src [c_out][c_in][3][3])
 * nasb_pack_elems gives the bf16 element count of one packed operand. */
#define NASB_PACK_PW 0
#define NASB_PACK_PW_T 1
#define NASB_PACK_C3 2
#define NASB_PACK_C3_T 3
typedef struct NasbPackJob {
    const float *src;
    void *dst;
    int32_t c_out, c_in, kind, reserved;
} NasbPackJob;
long long nasb_pack_elems(int kind, int c_out, int c_in);
int nasb_mt_pack_bf16(const NasbPackJob *jobs, int n, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Inference: one call per nn.Conv2d -> nn.BatchNorm2d.eval() -> nn.ReLU / nn.ReLU6 (-> + residual) unit
 * (layer_factory.py:94-114,125-158,225-265 with BN in eval mode).  Folds the running statistics, packs the bf16
 * tensor-core operand when the shape allows and picks the kernel (tcgen05 pointwise / dense 3x3, TMA-tiled depthwise,
 * general implicit GEMM) -- what the Python host does with several calls on the training path.
 *   u->gamma / beta may be NULL (affine=False); running_mean == NULL means "no BatchNorm" (bias is then the shift).
 *   scratch: >= nasb_conv_unit_scratch(c_out, c_in) bytes of device memory owned by the caller, rewritten by every call.
 *   flags  : NASB_UNIT_TENSOR_CORES | NASB_UNIT_TMA_TILES enable the specialised kernels.
 * -------------------------------------------------------------------------------------------------------*/
#define NASB_UNIT_TENSOR_CORES 1
#define NASB_UNIT_TMA_TILES 2
#define NASB_UNIT_PREPARED 4 /* scratch already holds the folded constants and the packed operand (nasb_conv_units_prepare) */
typedef struct NasbConvUnit {
    const float *weight;                                      /* [c_out][c_in / groups][ks][ks] */
    const float *gamma, *beta, *running_mean, *running_var;   /* BatchNorm2d (eval) or NULL */
    const float *bias;                                        /* conv bias (only without BatchNorm) or NULL */
    float eps;
    int32_t c_out, ks, stride, dil, pad, dw, in_relu, act;
} NasbConvUnit;
long long nasb_conv_unit_scratch(int c_out, int c_in);
int nasb_conv_unit_infer(const NasbTensor *x, const NasbConvUnit *u, const NasbTensor *res, const NasbTensor *out,
                         void *scratch, long long scratch_bytes, int flags, void *stream);
/* BN fold + operand pack of n units in ONE launch (units / c_in / scratch are HOST arrays; scratch[i] as above).  For regions
 * in which the weights do not change -- a validate() loop, a captured inference graph -- this replaces two small launches per
 * unit and call (200 of the ~360 launches of an arch0 forward); nasb_conv_unit_infer is then called with NASB_UNIT_PREPARED. */
int nasb_conv_units_prepare(const NasbConvUnit *const *units, const int *c_in, void *const *scratch, int n, int flags,
                            void *stream);

/* The fused separable primitive (inference): depthwise k x k (k in {3,5}, stride 1, dilation 1, "same" padding) -> folded BN
 * / activation -> pointwise 1x1 (C_out <= 64) on the tensor cores -> folded BN / activation (+ residual) in ONE kernel:
 * TMA halo tile -> depthwise in registers -> bf16 A operand written into SWIZZLE_128B shared memory -> tcgen05.mma (TMEM
 * accumulator) -> epilogue -> TMA store.  The depthwise tensor is never written to HBM.
 * SepConv's dw -> 1x1 -> BN -> ReLU (layer_factory.py:241-256), InvertedResidual's dw -> BN -> ReLU6 -> 1x1 -> BN (+x) (:141-158).
 * nasb_sepconv_tc_fwd : the kernel (mid_* = folded constants between the convolutions, wpack = nasb_pack_weight_bf16 of the
 *                       pointwise weight); nasb_sep_unit_infer: the same behind two NasbConvUnit blocks (fold + pack inside). */
int nasb_sepconv_tc_supported(int C, int N, int ks, int stride, int dil, int pad);
int nasb_sepconv_tc_fwd(const NasbTensor *x, const float *dw_weight, int ks, int stride, int dil, int pad, const float *mid_scale,
                        const float *mid_shift, int mid_act, const void *wpack, int N, const float *out_scale,
                        const float *out_shift, int out_act, const NasbTensor *res, const NasbTensor *out, void *stream);
int nasb_sep_unit_infer(const NasbTensor *x, const NasbConvUnit *udw, void *scratch_dw, long long scratch_dw_bytes,
                        const NasbConvUnit *upw, void *scratch_pw, long long scratch_pw_bytes, const NasbTensor *res,
                        const NasbTensor *out, int flags, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Row f3: the per-sample transform chain of the data loaders as ONE kernel over a batch of raw uint8 images
 * (reference: src/data/datasets.py:135-165 ResizeScale, :168-189 RandomMirror, :91-116 RandomCrop, :64-88 CentralCrop,
 * :192-208 Normalise, :211-220 ToTensor; composed in src/data/loaders.py:43-64).  The HOST draws the random parameters in
 * the reference's order and fills one NasbAugSample per image; the device computes cv2.resize(INTER_CUBIC) of the image /
 * cv2.resize(INTER_NEAREST) of the mask at factor `scale`, the optional horizontal flip, the crop window
 * [top, top+out_h) x [left, left+out_w) and (norm_scale * v - mean) / std in float64 rounded to float32 -- bit-exact with
 * OpenCV's own 8-bit resize (what cv2 computes with cv2.ipp.setUseIPP(False); oracle/augment_oracle.py).
 * image: device uint8 [h][w][3]; mask: device uint8 [h][w]; rh, rw = round(h*scale), round(w*scale) (half to even).
 * samples is a HOST array; out_image [n][3][out_h][out_w] float32 and out_mask [n][out_h][out_w] uint8 are device buffers. */
typedef struct NasbAugSample {
    const uint8_t *image;
    const uint8_t *mask;
    int32_t h, w, rh, rw;
    double scale;
    int32_t top, left, mirror, reserved;
} NasbAugSample;
int nasb_augment_batch(const NasbAugSample *samples, int n, int out_h, int out_w, double norm_scale, const double *mean,
                       const double *stdv, float *out_image, uint8_t *out_mask, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NASB200_H_ */
