/* phiseg_sm100.h -- C-ABI of libphiseg_sm100.so: hand-written sm_100a CUDA kernels for the PHiSeg hot path.
 *
 * The reference (baumgach/PHiSeg-code) has no FFI: its op layer is TensorFlow-1.12 library calls made from
 * tfwrapper/layers.py, tfwrapper/normalisation.py and phiseg/phiseg_model.py.  Each entry point below names the
 * reference call site (file:line under /root/reference) whose device work it replaces.  Conventions:
 *   - every tensor is NHWC, described by phs_tensor {ptr,N,H,W,C,ld,dtype}; ld = pixel pitch in ELEMENTS
 *     (ld >= C lets a tensor be a channel slice of a wider buffer: zero-copy tf.concat);
 *   - dtype 0 = float32, 1 = bfloat16; filters are HWIO float32 masters (tfwrapper/layers.py:115) plus bf16
 *     shadows in the two K-major layouts the tensor-core kernels read (see phs_weight_prep);
 *   - the caller owns all memory (device pointers, e.g. torch tensors' data_ptr()); nothing is allocated,
 *     no implicit synchronisation, all work is enqueued on `stream` (a cudaStream_t);
 *   - return value 0 = ok, <0 = argument error, >0 = cudaError_t; phs_last_error() gives the message.
 */
#ifndef PHISEG_SM100_H
#define PHISEG_SM100_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PHS_F32 0
#define PHS_BF16 1
#define PHS_F64 2 /* only the resident images of phs_augment_batch (the reference stores np.float = float64 in its HDF5) */

#define PHS_NORM_BN_TRAIN 0 /* tfwrapper/normalisation.py:145-163, is_training=True  */
#define PHS_NORM_BN_INFER 1 /* same, is_training=False (moving statistics)            */
#define PHS_NORM_GN 2       /* tfwrapper/normalisation.py:17-36                       */

#define PHS_IMPL_SIMT 0 /* fp32-accumulate CUDA-core kernels (parity mode, odd shapes)            */
#define PHS_IMPL_TC 1   /* tcgen05 implicit GEMM, TMA-staged, TMEM accumulators (bf16 operands)  */

typedef struct {
  void* ptr;
  int32_t N, H, W, C;
  int32_t ld;    /* pixel pitch in elements */
  int32_t dtype; /* PHS_F32 | PHS_BF16 */
} phs_tensor;

int phs_version(void);
int phs_arch(void); /* 100: built for sm_100a */
const char* phs_last_error(void);
/* CRC32C of a HOST buffer (continuing from crc; 0 to start): the checksum of TensorFlow checkpoint bundles, which
 * tfwrapper/checkpoint.py reads and writes in place of tf.train.Saver (phiseg_model.py:144-148,505-525). */
unsigned int phs_crc32c(const void* data, size_t n, unsigned int crc);
/* 1 if the current device can run the tcgen05 kernels (compute capability 10.x), else 0 */
int phs_device_ok(void);

/* ---- convolution: tf.nn.conv2d(x, W, [1,1,1,1], "SAME") + tf.nn.bias_add (tfwrapper/layers.py:123,132) ------
 * ksize in {1,3}.  dgrad != 0 computes the input gradient instead (dx = conv(dy, rot180(W)^T)), i.e. the
 * Conv2DBackpropInput that optimizer.minimize (phiseg/phiseg_model.py:141) adds.  accumulate != 0: y += result.
 * SIMT: w = float32 HWIO master.  TC: w = bf16 shadow from phs_weight_prep (fwd layout, or dgrad layout when dgrad). */
int phs_conv2d(const phs_tensor* x, const void* w, const float* bias, const phs_tensor* y, int ksize, int dgrad,
               int accumulate, int impl, void* stream);
/* forward convolution (tensor-core path only) that also accumulates, from the fp32 accumulators, the per-(sample,
 * channel) sum and sum of squares the following batch_norm / group_norm2D needs: stats[N][C][2] must be zeroed.
 * Every statistics / reduction buffer of this library (stats, sums) is DOUBLE: the kernels add fp32 partial sums into
 * them with fp64 atomics, whose result does not depend on the arrival order in practice, so a step is reproducible
 * run to run (fp32 atomics were not, and bf16 roundings downstream amplified the difference). */
int phs_conv2d_stats(const phs_tensor* x, const void* w, const float* bias, const phs_tensor* y, int ksize,
                     double* stats, void* stream);
/* Same, for callers that clear the statistics of MANY layers with one fill (the launch-program engine keeps them in one
 * arena): adds onto stats, which the caller must have zeroed, and launches no memset of its own.  stats holds
 * (N + 1) * C * 2 doubles here: the per-sample sums [N][C][2] followed by the batch totals [C][2] that batch norm needs
 * (phs_norm_act_fwd_stats reads them). */
int phs_conv2d_stats_acc(const phs_tensor* x, const void* w, const float* bias, const phs_tensor* y, int ksize,
                         double* stats, void* stream);
/* The conv -> norm -> ReLU -> conv fusion (the composite tfwrapper/layers.py:123-135 builds, followed by the next
 * layers.conv2D): a 3x3 forward convolution whose input is the RAW output yprev of the previous convolution.  batch_norm /
 * group_norm2D + ReLU (tfwrapper/normalisation.py:17-36,145-163, layers.py:134-135) are applied to the activation tile in
 * shared memory between the TMA load and the tensor-core MMAs, so the normalised activation is never written to or read
 * from HBM (phs_norm_act_fwd_stats and its two tensor passes disappear; zero padding stays exact: halo pixels outside the
 * image are left at zero).  `pre` describes the producer layer's normalisation; the kernel derives mean / rstd from its
 * statistics exactly like phs_norm_act_fwd_stats (same bits), writes them to pre->mean / pre->rstd [N][C] for the backward
 * kernels (may be NULL) and updates the batch-norm moving averages (training mode).  stats: NULL, or this layer's own fused
 * statistics with the phs_conv2d_stats_acc contract (pre-zeroed, (N + 1) * Cout * 2 doubles).  Returns -3 (nothing
 * launched) when the layer is not eligible - ask phs_conv2d_pre_plan first. */
typedef struct phs_norm_pre {
  const double* stats;       /* producer's statistics, phs_conv2d_stats_acc layout ((N + 1) * C * 2); unused for BN_INFER */
  int32_t mode;              /* PHS_NORM_BN_TRAIN | PHS_NORM_BN_INFER | PHS_NORM_GN */
  float eps, decay;
  float* moving_mean;        /* [C] batch norm only (updated in training mode, read in inference mode) */
  float* moving_var;
  float* mean;               /* [N][C] out, may be NULL */
  float* rstd;
  const float* gamma;        /* [C] */
  const float* beta;
  int32_t relu;
} phs_norm_pre;
int phs_conv2d_pre(const phs_tensor* yprev, const phs_norm_pre* pre, const void* w, const float* bias, const phs_tensor* y,
                   double* stats, void* stream);
/* Inference-mode batch norm folded into the convolution (sampling / validation graphs, phiseg_model.py:61-109,537-549):
 * a = act(gamma * (conv(x) + bias - moving_mean) * rsqrt(moving_var + eps) + beta) in ONE launch, the affine map + ReLU
 * applied to the fp32 accumulators in the epilogue (phs_norm_finalize + phs_norm_act_fwd and the raw convolution output
 * disappear).  post->mode must be PHS_NORM_BN_INFER; post->stats / mean / rstd are unused.  Any tensor-core shape
 * (Cin % 32 == 0, Cout % 16 == 0, bf16 input), ksize 1 or 3. */
int phs_conv2d_post(const phs_tensor* x, const void* w, const float* bias, const phs_norm_pre* post, const phs_tensor* a,
                    int ksize, void* stream);
/* Host-only: 1 if phs_conv2d_pre takes the layer (plan[12] as phs_conv_halo_plan), 0 if not. */
int phs_conv2d_pre_plan(const phs_tensor* x, const phs_tensor* y, int with_stats, int* plan);
/* Host-only introspection (no device work, callable without a GPU): the launch geometry the halo-tile tcgen05 kernel
 * would use for a 3x3 layer.  plan[12] = {CTAs per SM, S, halo stages, filter stages, filter resident, staging group,
 * accumulator stages, TMEM columns, dynamic shared memory bytes, grid, tiles, BK}; accumulate bit 1 = statistics buffer
 * pre-zeroed.  Returns 1 if that kernel takes the layer, 0 if another kernel does. */
int phs_conv_halo_plan(const phs_tensor* x, const phs_tensor* y, int accumulate, int with_stats, int* plan);
/* Same for the halo-tile filter-gradient kernel.  plan[12] = {resident CTAs per SM, filter rows per CTA, accumulators per
 * filter row, output channels per CTA, input-channel blocks, output-channel blocks, pipeline stages, TMEM columns,
 * dynamic shared memory bytes, work items (grid.x), brick splits (grid.y), bricks per split}. */
int phs_wgrad_halo_plan(const phs_tensor* x, const phs_tensor* dy, int* plan);
/* Conv2DBackpropFilter: dw[kh][kw][ci][co] (+)= sum x[.,h+kh-p,w+kw-p,ci]*dy[.,h,w,co]; db (+)= sum dy (may be NULL).
 * dw/db are float32 in the HWIO master layout.  The TC variant accumulates with atomics: zero or reuse dw first. */
int phs_conv2d_wgrad(const phs_tensor* x, const phs_tensor* dy, float* dw, float* db, int ksize, int accumulate,
                     int impl, void* stream);

/* ---- normalisation + activation -------------------------------------------------------------------------- */
/* per-(sample,channel) sum and sum of squares of y: stats[N][C][2] (overwritten). */
int phs_chan_stats(const phs_tensor* y, double* stats, void* stream);
/* stats -> mean[N][C], rstd[N][C] for batch_norm (train: batch statistics + moving-average update with decay,
 * Bessel-corrected variance; infer: moving statistics) or group_norm2D (groups of C/max(2,C/16) channels). */
int phs_norm_finalize(const double* stats, int N, int HW, int C, int mode, float eps, float decay, float* moving_mean,
                      float* moving_var, float* mean, float* rstd, void* stream);
/* phs_norm_finalize + phs_norm_act_fwd in one launch, for training-mode batch_norm and for group_norm2D: mean / rstd are
 * derived inside the kernel from the phs_conv2d_stats_acc layout ((N + 1) * C * 2 doubles), written to mean/rstd[N][C] for
 * the backward kernels, and the batch-norm moving averages are updated (decay; may be NULL for group norm). */
int phs_norm_act_fwd_stats(const phs_tensor* y, const double* stats, int mode, float eps, float decay, float* moving_mean,
                           float* moving_var, float* mean, float* rstd, const float* gamma, const float* beta, int relu,
                           const phs_tensor* a, void* stream);
/* a = act(gamma*(y-mean)*rstd + beta); relu != 0 applies tf.nn.relu (tfwrapper/layers.py:134-135). */
int phs_norm_act_fwd(const phs_tensor* y, const float* mean, const float* rstd, const float* gamma, const float* beta,
                     int relu, const phs_tensor* a, void* stream);
/* backward of the above, three launches: sums[N][C][2] = (sum g*mask, sum g*mask*xhat) ... */
int phs_norm_bwd_reduce(const phs_tensor* g, const phs_tensor* y, const float* mean, const float* rstd,
                        const float* gamma, const float* beta, int relu, double* sums, void* stream);
/* Same, and the activation a = act(norm(y)) is re-materialised on the way (one extra tensor write): the backward half of
 * phs_conv2d_pre - the forward pass never wrote a, the consumer's filter gradient (phs_conv2d_wgrad) reads it.  Same bits
 * as phs_norm_act_fwd. */
int phs_norm_bwd_reduce_remat(const phs_tensor* g, const phs_tensor* y, const float* mean, const float* rstd,
                              const float* gamma, const float* beta, int relu, double* sums, const phs_tensor* a,
                              void* stream);
/* batch_norm (training) only, two launches instead of three: only the batch totals matter, so the reduction adds straight
 * into totals[C][2] (doubles, zeroed by the caller - the engine clears one arena per step, no memset on the chain) ... */
int phs_norm_bwd_reduce_bn(const phs_tensor* g, const phs_tensor* y, const float* mean, const float* rstd,
                           const float* gamma, const float* beta, int relu, double* totals, void* stream);
/* ... and the apply kernel derives its two coefficients per channel from the totals itself (same expressions and
 * rounding as phs_norm_bwd_finalize) and writes (+)= dgamma / dbeta (may be NULL).  The convolution in front of a batch
 * norm has no bias (tfwrapper/layers.py:126-128), so there is no dbias. */
int phs_norm_bwd_apply_bn(const phs_tensor* g, const phs_tensor* y, const float* mean, const float* rstd,
                          const float* gamma, const float* beta, int relu, const double* totals, const phs_tensor* dy,
                          float* dgamma, float* dbeta, int accumulate, void* stream);
/* ... dgamma/dbeta (+= when accumulate) and per-(n,c) coefficients coef[N][C][2]; dbias (may be NULL) is the
 * gradient of the conv bias that precedes the norm, derived analytically from the forward statistics ... */
int phs_norm_bwd_finalize(const double* sums, const double* stats, const float* mean, const float* rstd,
                          const float* gamma, int N, int HW, int C, int mode, float* coef, float* dgamma, float* dbeta,
                          float* dbias, int accumulate, void* stream);
/* ... dy = rstd*(g*mask*gamma - m1 - xhat*m2). */
int phs_norm_bwd_apply(const phs_tensor* g, const phs_tensor* y, const float* mean, const float* rstd,
                       const float* gamma, const float* beta, int relu, const float* coef, const phs_tensor* dy,
                       void* stream);

/* ---- resampling ------------------------------------------------------------------------------------------ */
/* tf.nn.avg_pool 2x2/2 (tfwrapper/layers.py:44-54) and its adjoint */
int phs_avgpool2_fwd(const phs_tensor* x, const phs_tensor* y, void* stream);
int phs_avgpool2_bwd(const phs_tensor* dy, const phs_tensor* dx, int accumulate, void* stream);
/* TF1 legacy bilinear x2, align_corners=False (tfwrapper/layers.py:336-345) and its adjoint */
int phs_upsample2_fwd(const phs_tensor* x, const phs_tensor* y, void* stream);
int phs_upsample2_bwd(const phs_tensor* dy, const phs_tensor* dx, int accumulate, void* stream);

/* ---- latent heads: softplus, reparameterisation, KL (posteriors.py:105-108,125-128; phiseg_model.py:210-226) */
/* All tensors float32 [N, hw, zd] flattened (count = N*hw*zd; per_sample = hw*zd).  sp_* are pre-softplus sigma
 * maps.  gap != 0: ProbUNet form, mu/sigma are means over the hw positions (posteriors.py:41-45) and outputs are [N,zd].
 * z = mu_q + sigma_q*eps (use_prior_z: z = mu_p + sigma_p*eps, priors.py:100).  kl_out (may be NULL) += kl_scale * sum KL. */
int phs_latent_fwd(const float* mu_q, const float* sp_q, const float* mu_p, const float* sp_p, const float* eps, int N,
                   int hw, int zd, int gap, int use_prior_z, float* mu_q_out, float* sigma_q, float* mu_p_out,
                   float* sigma_p, float* z, float* kl_out, float kl_scale, void* stream);
/* gradients wrt the four head maps given dz and the KL weight (kl_scale = KL_w * 4^l / B). */
int phs_latent_bwd(const float* dz, const float* mu_q, const float* sp_q, const float* sigma_q, const float* mu_p,
                   const float* sp_p, const float* sigma_p, const float* eps, int N, int hw, int zd, int gap,
                   float kl_scale, float* d_mu_q, float* d_sp_q, float* d_mu_p, float* d_sp_p, void* stream);

/* ---- multi-scale residual cross-entropy (phiseg_model.py:229-262, likelihoods.py:218-221) ------------------ */
/* logits[l]: float32 [N, H>>l, W>>l, nlabels] native-resolution head outputs, l = 0..L-1; labels uint8 [N,H,W].
 * loss_out[l] += scale * sum over pixels of xent(onehot, s_accum[l]) with s_accum[l] = sum_{i>=l} NN-upsample(logits[i]).
 * dlogits[l] (may be NULL => forward only) receive d(sum_l loss_l)/d logits[l] (buffers must be zeroed by the caller). */
int phs_xent_multiscale(const float* const* logits, float* const* dlogits, const uint8_t* labels, int N, int H, int W,
                        int nlabels, int L, float scale, float* loss_out, void* stream);
/* s_out = sum_l NN-upsample(logits[l]) (phiseg_model.py:304-311); optional softmax / running sum / argmax outputs.
 * rep >= 1: the N rows are rep samples of N/rep images (sample-major, row s*(N/rep) + b); softmax_accum is then
 * [N/rep, H, W, nlabels] and receives the sum over the samples of every image (predict, phiseg_model.py:344-349). */
int phs_aggregate_logits(const float* const* logits, int N, int H, int W, int nlabels, int L, int rep, float* s_out,
                         float* softmax_out, float* softmax_accum, int64_t* argmax_out, void* stream);

/* ---- optimizer (phiseg_model.py:134-141): tf.train.AdamOptimizer, TF "epsilon-hat" form -------------------- */
/* lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t); when lr_t_dev != NULL the step size is read from device memory instead
 * (lets a captured CUDA graph of the whole training step be replayed with a new learning rate). */
int phs_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr_t, const float* lr_t_dev,
                  float beta1, float beta2, float eps, float grad_scale, void* stream);
/* tf.train.MomentumOptimizer(momentum, use_nesterov=True) */
int phs_momentum_step(float* p, const float* g, float* acc, int64_t n, float lr, const float* lr_dev, float momentum,
                      float grad_scale, void* stream);
/* bf16 shadows of every conv filter for the tensor-core kernels.  table: int64[nconv][7] =
 * {src_off (floats into master), fwd_off, dgrad_off (bf16 elements into shadow; < 0: no dgrad layout), taps, cin, cout,
 *  kpitch (row pitch of the forward layout in elements; 0 = taps*cin; larger = rows zero padded by the caller)}.
 * fwd layout  [cout][tap*cin + ci]; dgrad layout [cin][(taps-1-tap)*cout + co]. */
int phs_weight_prep(const float* master, void* shadow, const int64_t* table, int nconv, void* stream);
/* Low halves of the same shadows: shadow_lo = bf16(w - bf16(w)), same table / layouts.  Together with phs_split_bf16 this
 * gives the fp32-accurate tensor-core mode ('parity_tc'): x*w = x_hi*w_hi + x_lo*w_hi + x_hi*w_lo (+ O(2^-17)), three bf16
 * tcgen05 passes accumulated in fp32 through the accumulate flag of phs_conv2d / phs_conv2d_wgrad, which meets the 1e-3
 * logit contract of the reference's fp32 path (tf.nn.conv2d, tfwrapper/layers.py:123) on the tensor cores. */
int phs_weight_prep_lo(const float* master, void* shadow_lo, const int64_t* table, int nconv, void* stream);
/* hi = bf16(src), lo = bf16(src - hi): src float32, hi / lo bfloat16 tensors of the same shape */
int phs_split_bf16(const phs_tensor* src, const phs_tensor* hi, const phs_tensor* lo, void* stream);

/* ---- small helpers --------------------------------------------------------------------------------------- */
/* dst[.., c_off + c] = src[.., c] with dtype conversion (strided channel-slice copy).  dst->N may be a multiple of
 * src->N: dst[n] = src[n % src->N] (the per-image part of a sampling pass, e.g. the prior's encoder pyramid, is computed
 * once per image and tiled over the samples drawn for it, phiseg_model.py:337-353). */
int phs_copy_cast(const phs_tensor* src, const phs_tensor* dst, void* stream);
/* out[.., tap*Cin + ci] = x[.. shifted by tap .., ci] (zero outside the image), other channels 0; out bf16 with
 * C >= 9*Cin.  Rewrites the 3x3 convolution of a 1..7-channel network input (posteriors.py:87, priors.py:80,
 * likelihoods.py:110) as a 1x1 convolution over <= 64 channels for the tensor-core kernels. */
int phs_im2col3x3(const phs_tensor* x, const phs_tensor* out, void* stream);
/* posterior input tf.concat([x, one_hot(s) - 0.5], -1) (phiseg_model.py:29, posteriors.py:87) */
int phs_posterior_input(const float* x, const uint8_t* s, int N, int H, int W, int Cx, int nlabels,
                        const phs_tensor* out, void* stream);
/* ProbUNet: tile z [N,zd] over H x W into a channel slice (likelihoods.py:147-151) and the adjoint reduction */
int phs_broadcast_z(const float* z, const phs_tensor* out, void* stream);
int phs_broadcast_z_bwd(const phs_tensor* g, float* dz, int accumulate, void* stream);
int phs_fill_f32(float* p, int64_t n, float v, void* stream);
/* dst (+)= src over float buffers */
int phs_axpy_f32(float* dst, const float* src, int64_t n, float alpha, void* stream);
/* out[0] += scale * sum(src^2): the tf.nn.l2_loss terms of add_weight_decay (phiseg_model.py:290-300) */
int phs_sumsq_f32(const float* src, int64_t n, float scale, float* out, void* stream);
/* add_weight_decay (phiseg_model.py:290-300: weight_decay_weight * sum over the 'weight_variables' collection of
 * tf.nn.l2_loss) over all filters in one launch.  segs[nseg][2] (device) = (offset, count) of each filter inside the flat
 * fp32 parameter buffer; loss_out[0] += 0.5 * wd * sum W^2 (may be NULL); grads (same layout as params; may be NULL, e.g.
 * for validation losses) += wd * W. */
int phs_weight_decay(const float* params, float* grads, const int64_t* segs, int nseg, float wd, float* loss_out,
                     void* stream);
/* first-maximum argmax over the label axis of [npix, nlabels] (np.argmax in predict, phiseg_model.py:351-353) */
/* ---- validation metrics (phiseg_model.py:558-640; utils.py:103-118 ncc, :270-320 generalised_energy_distance,
 * :323-362 variance_ncc_dist; medpy jc / dc) ------------------------------------------------------------------
 * masks_a [Ka][npix], masks_b [Kb][npix]: label masks, uint8 (elem_size 1) or int64 (8: the argmax output).  For every
 * pair (i, j) and label l: inter[i][j][l] = |{a_i == l} & {b_j == l}|; count_a[i][l], count_b[j][l] = label histograms
 * (may be NULL).  IoU (jc) = inter / (ca + cb - inter), Dice (dc) = 2 inter / (ca + cb): the host finishes GED / Dice
 * from these few integers. */
int phs_pairwise_label_stats(const void* masks_a, int elem_size_a, int Ka, const void* masks_b, int elem_size_b, int Kb,
                             int64_t npix, int nlabels, int* inter, int* count_a, int* count_b, void* stream);
/* softmax [N][npix][nlabels] (samples), gt uint8 [M][npix] (annotations).  e_ss[npix] = mean_i xent(mean_seg, s_i),
 * e_sy[M][npix] = mean_i xent(onehot(gt_j), s_i) with xent(t, s) = -sum_l t_l log(s_l + 1e-8); sums[M][5] (double) =
 * (sum a, sum a^2, sum v, sum v^2, sum a*v) over the pixels for a = e_ss, v = e_sy[j]: ncc_j = cov(a, v) / (std a std v). */
int phs_ncc_maps(const float* softmax, const uint8_t* gt, int N, int M, int64_t npix, int nlabels, float* e_ss, float* e_sy,
                 double* sums, void* stream);
int phs_argmax_f32(const float* src, int64_t npix, int nlabels, int64_t* out, void* stream);
/* Uncertainty maps of phiseg_model.py:378-475 without stacking the samples on the host.  logits [S][B][npix][nlabels]
 * (one batched sampling pass, row = s*B + b), gt uint8 [B][npix] or NULL.  Per image pixel, acc (doubles, cleared by the
 * caller, [B*npix][nl + nl(nl+1)/2 + 1]) += (sum_s v_c | sum_s v_i v_j for i <= j | sum_s cross entropy of the logits
 * against gt), where v = softmax(logits) (kind 0: s_out_eval_sm) or clip(logits, lo, hi) (kind 1: :390). */
int phs_sample_moments(const float* logits, const uint8_t* gt, int S, int B, int64_t npix, int nlabels, int kind, float lo,
                       float hi, double* acc, void* stream);
/* ... and the maps from `count` accumulated samples (every output may be NULL): mean_arg = argmax_c of the mean (:470),
 * std_mean = mean_c np.std (:467-468), var_sum = trace of the covariance of the first nlabels - drop_last classes (the
 * eigenvalue sum of :393-402), cov_det = det(np.cov) of all classes (:423-428), err = mean cross entropy (:444-446,472). */
int phs_sample_maps(const double* acc, int64_t total_pix, int nlabels, int count, int drop_last, int64_t* mean_arg,
                    float* std_mean, float* var_sum, float* cov_det, float* err, void* stream);

/* ---- input pipeline (data/batch_provider.py:43-67,131-272; utils.py:18-37) --------------------------------------------
 * One output image of a batch: which resident image, which annotator's mask, and the random augmentation the host drew
 * for it in the reference's np.random order.  minv = the inverted 2x3 matrix cv2.warpAffine works with (from
 * cv2.getRotationMatrix2D((W/2, H/2), angle, 1)); crop/px/py = side and origin of the square crop that cv2.resize
 * stretches back to H x W (rows from py, columns from px, as batch_provider.py:219 indexes them). */
#define PHS_AUG_ROTATE 1
#define PHS_AUG_SCALE 2
#define PHS_AUG_FLIPLR 4
#define PHS_AUG_FLIPUD 8
typedef struct phs_aug_params {
  int32_t src;   /* index into the resident images / labels */
  int32_t annot; /* annotator plane of the labels (_select_random_label, :124-130) */
  int32_t flags; /* PHS_AUG_* */
  int32_t crop, px, py;
  double minv[6];
} phs_aug_params;
/* x_out[B][H][W] float32, s_out[B][H][W] uint8 from the resident data set images[N][H][W] (float32 or float64) and
 * labels[N][H][W][annotators] uint8 (labels and s_out may both be NULL), params[B] in DEVICE memory.  Applies, per image:
 * rotation (bilinear warpAffine, constant border 0), crop + bilinear resize, flips - label masks as one-hot planes with
 * np.argmax (nlabels <= 4, the reference's one-hot branch) - restating OpenCV's fixed-point coordinates, coefficient
 * types and summation order.  One launch per batch, nothing intermediate is stored. */
int phs_augment_batch(const void* images, int image_dtype, const uint8_t* labels, int H, int W, int annotators, int nlabels,
                      const phs_aug_params* params, int B, float* x_out, uint8_t* s_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
