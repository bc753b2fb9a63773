/*
 * b200r.h -- C-ABI of libb200robust.so: the B200 (sm_100a) implementation of RobustART's
 * noise-injection + classification-eval hot path.
 *
 * Nothing like this exists in the reference (it is 100 % Python); each entry point names the
 * reference Python function it replaces (path:line relative to the reference checkout).
 *
 * Conventions
 *   - every function returns 0 on success or a negative B200R_E* code; b200r_last_error() gives
 *     a thread-local human-readable message. Nothing throws, nothing owns caller memory.
 *   - all pointers are DEVICE pointers unless the parameter name ends in _host.
 *   - every launch takes an explicit CUDA stream (a CUstream / cudaStream_t handle passed as
 *     void*; NULL = legacy default stream). No function synchronises unless documented.
 *   - images are uint8 NHWC [n,h,w,3] (RobustART/noise/utils/add_noise_utils.py:27-31) or
 *     float32 NCHW [n,3,h,w] in [0,1] (prototype/prototype/solver/benchmark_eval_adv.py:229-232).
 */
#ifndef B200R_H_
#define B200R_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200R_OK 0
#define B200R_EINVAL (-1)   /* bad argument (shape, id, severity, null pointer) */
#define B200R_ECUDA (-2)    /* CUDA runtime / driver error, see b200r_last_error() */
#define B200R_ENOSPC (-3)   /* workspace too small */
#define B200R_ENOTSUP (-4)  /* valid request the library does not implement */

typedef void* b200r_stream_t;

const char* b200r_last_error(void);
int b200r_version(void);
/* number of SMs of the current device (148 on B200); grids are sized from it. */
int b200r_sm_count(int* sms);

/* ------------------------------------------------------------------------------------------
 * ImageNet-C corruptions.  ids follow corruption_tuple
 * (RobustART/noise/utils/imagenet_c/__init__.py:5-8).
 * ------------------------------------------------------------------------------------------ */
enum b200r_corruption {
  B200R_GAUSSIAN_NOISE = 0,    /* corruptions.py:122-126 */
  B200R_SHOT_NOISE = 1,        /* corruptions.py:129-133 */
  B200R_IMPULSE_NOISE = 2,     /* corruptions.py:136-140 */
  B200R_DEFOCUS_BLUR = 3,      /* corruptions.py:187-198 */
  B200R_GLASS_BLUR = 4,        /* corruptions.py:169-184 */
  B200R_MOTION_BLUR = 5,       /* corruptions.py:201-216 */
  B200R_ZOOM_BLUR = 6,         /* corruptions.py:219-232 */
  B200R_SNOW = 7,              /* corruptions.py:265-290 */
  B200R_FROST = 8,             /* corruptions.py:244-262 */
  B200R_FOG = 9,               /* corruptions.py:235-241 */
  B200R_BRIGHTNESS = 10,       /* corruptions.py:353-361 */
  B200R_CONTRAST = 11,         /* corruptions.py:345-350 */
  B200R_ELASTIC_TRANSFORM = 12,/* corruptions.py:395-424 */
  B200R_PIXELATE = 13,         /* corruptions.py:385-391 */
  B200R_JPEG_COMPRESSION = 14, /* corruptions.py:375-382 */
  B200R_SPECKLE_NOISE = 15,    /* corruptions.py:143-147 */
  B200R_GAUSSIAN_BLUR = 16,    /* corruptions.py:162-166 */
  B200R_SPATTER = 17,          /* corruptions.py:293-342 */
  B200R_SATURATE = 18,         /* corruptions.py:364-372 */
  B200R_NUM_CORRUPTIONS = 19
};

/* Scratch bytes b200r_corrupt_u8 needs for (id, severity, n, h, w). May be 0. */
int b200r_corrupt_workspace_bytes(int corruption_id, int severity, int n, int h, int w,
                                  size_t* bytes);

/* Number of float32 values of externally supplied randomness (`ext_noise`) the corruption
 * consumes for a batch of n images, 0 if the corruption is deterministic.  The layout per
 * corruption is documented in DESIGN.md ("shared-draw parity mode"). */
int b200r_corrupt_ext_noise_count(int corruption_id, int severity, int n, int h, int w,
                                  size_t* count);

/* corrupt(x, severity, corruption_number=id) for a whole batch
 * (replaces the per-image loop add_noise_utils.py:27-31 + imagenet_c/__init__.py:13-35,
 * including the final np.uint8 truncation).  in/out: uint8 NHWC; out may alias in.
 * RNG: counter-based Philox4x32-10 keyed by `seed`; image i of this call uses stream
 * `image_offset + i`, so results do not depend on batch split or world size.
 * ext_noise: NULL (device RNG) or b200r_corrupt_ext_noise_count() floats of caller-drawn
 * randomness, used instead of the device RNG (parity against the CPU oracle). */
int b200r_corrupt_u8(int corruption_id, int severity, const uint8_t* in, uint8_t* out, int n, int h,
                     int w, uint64_t seed, uint64_t image_offset, const float* ext_noise,
                     void* workspace, size_t workspace_bytes, b200r_stream_t stream);

/* Host-only: the quantile table the device-RNG gaussian / speckle kernels draw their normals from, unit scale, as
 * z[row * 64 + stratum] (256 rows x 64 strata): a random byte picks the row, lane and loop iteration pick the stratum.  Lets a test
 * check the distribution the kernel produces (corruptions.py:122-126,143-147 draw np.random.normal).  No GPU needed. */
int b200r_normal_strata_table(double* z16384);

/* Optional assets: frost textures (corruptions.py:250-260 reads frost{1..6}.{png,jpg}; the files
 * are not in the reference repo).  rgb: device uint8 [th,tw,3]; slot in [0,6). */
int b200r_set_frost_texture(int slot, const uint8_t* rgb, int th, int tw);

/* ------------------------------------------------------------------------------------------
 * Layout / normalisation (ToTensor+Normalize, imagenet_dataloader.py:74-80;
 * normalize(x)/normalize(x,"inv"), benchmark_eval_adv.py:33-46).
 * mean/std: host float[3].
 * ------------------------------------------------------------------------------------------ */
int b200r_u8nhwc_to_f32nchw(const uint8_t* in, float* out, int n, int h, int w,
                            const float* mean_host, const float* std_host, b200r_stream_t stream);
/* out = (x - mean)/std (inverse = 0), x*std + mean (inverse = 1), or x/std (inverse = 2: the
 * input-gradient of normalisation), NCHW float32 */
int b200r_normalize_f32nchw(const float* in, float* out, int n, int h, int w,
                            const float* mean_host, const float* std_host, int inverse,
                            b200r_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Gradient-attack inner steps on float32 NCHW [n, chw] in [0,1].
 * foolbox 3.3.1 gradient_descent_base.py (third-party, pinned requirements.txt:13) via
 * RobustART/noise/utils/adv/attack.py:20-33; in-repo restatement
 * prototype/prototype/solver/adv_cls_solver_train_pgd_new.py:84-103.
 * ------------------------------------------------------------------------------------------ */
/* x = x0 + U(-eps,eps), clipped to [0,1] when clip01 != 0 (foolbox clips, imfgsm_attack.py:73-74
 * does not); u: optional caller-drawn uniforms in [0,1) [n*chw] */
int b200r_random_start_linf(const float* x0, float* x, size_t n, size_t chw, float eps,
                            uint64_t seed, uint64_t image_offset, const float* u, int clip01,
                            b200r_stream_t stream);
/* L2 random start (foolbox L2ProjectedGradientDescentAttack.get_random_start -> uniform_l2_n_balls): x = clip01(x0 + eps * u), u uniform
 * in the unit chw-ball = the first chw of chw + 1 normals over their norm.  Philox4x32-10 keyed by `seed`, sample i of this call draws
 * stream image_offset + i (a sharded / re-batched attack reproduces the same starts). */
int b200r_random_start_l2(const float* x0, float* x, size_t n, size_t chw, float eps, uint64_t seed,
                          uint64_t image_offset, b200r_stream_t stream);
/* x = clip01(x0 + clip(x + alpha*sign(g) - x0, -eps, eps)); x updated in place. */
int b200r_pgd_step_linf(float* x, const float* g, const float* x0, size_t n, size_t chw,
                        float alpha, float eps, b200r_stream_t stream);
/* x = x + alpha*g/max(||g||_2,1e-12); d = x - x0; x = clip01(x0 + d*min(1, eps/||d||_2)) per
 * sample.  workspace: 2*n floats.  Images of up to 8*512*10*4 values (3 x 233 x 233) run as ONE launch (a cluster of 8 CTAs per
 * image, fixed-order reductions: bit-reproducible, workspace untouched); larger ones as three kernels with atomics. */
int b200r_pgd_step_l2(float* x, const float* g, const float* x0, size_t n, size_t chw, float alpha,
                      float eps, float* workspace, b200r_stream_t stream);
/* MI-FGSM step (Attacks/imfgsm_attack.py:85-90): m = decay*m + g/mean|g| (per sample);
 * x = clip01(x0 + clip(x + step*sign(m) - x0, -eps, eps)).  workspace: n floats (same one-launch / fallback rule as pgd_step_l2). */
int b200r_mim_step_linf(float* x, float* momentum, const float* g, const float* x0, size_t n,
                        size_t chw, float step, float eps, float decay, float* workspace,
                        b200r_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Loss / metrics on logits [n, classes] float32.
 * ------------------------------------------------------------------------------------------ */
/* softmax cross-entropy: loss[i] (nullable), dlogits = (softmax - onehot)*grad_scale (nullable)
 * (F.cross_entropy; foolbox sums, MIM means -> grad_scale = 1 or 1/n) */
int b200r_ce_loss_grad(const float* logits, const int64_t* labels, float* loss, float* dlogits,
                       int n, int classes, float grad_scale, b200r_stream_t stream);
/* scores = softmax(logits) (cls_solver.py:420) */
int b200r_softmax(const float* logits, float* scores, int n, int classes, b200r_stream_t stream);
/* counters[0] += #top1 hits, [1] += #top5 hits, [2] += n ; pred (nullable) = argmax
 * (misc.py:441-455 accuracy(); imagenet_evaluator.py:49-67).  Tie-break: lowest index wins. */
int b200r_topk_count(const float* logits, const int64_t* labels, int n, int classes,
                     int64_t* counters, int64_t* pred, b200r_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * AutoAttack device pieces (vendored fra31/auto-attack: .../Attacks/autoattack/autopgd_base.py,
 * square.py).
 * ------------------------------------------------------------------------------------------ */
/* APGD-Linf update with momentum (autopgd_base.py:332-338); x_adv_old <- x_adv, x_adv <- update;
 * step: device float[n] per-sample step sizes; a = 0.75 (1.0 on the first iteration) */
int b200r_apgd_step_linf(float* x_adv, float* x_adv_old, const float* g, const float* x0,
                         const float* step, size_t n, size_t chw, float eps, float a,
                         b200r_stream_t stream);
/* DLR loss (+ gradient w.r.t. logits): targets == NULL -> untargeted (autopgd_base.py:198-204),
 * else targeted (:599-604).  loss, dlogits nullable. */
int b200r_dlr_loss_grad(const float* logits, const int64_t* labels, const int64_t* targets, float* loss,
                        float* dlogits, int n, int classes, b200r_stream_t stream);
/* Square-attack proposal (square.py:246-254): paste 2*eps*sign[c] on the s x s window at (vh, vw) of
 * x_best, project onto the eps-ball around x0, clip to [0,1] */
int b200r_square_propose_linf(const float* x_best, const float* x0, float* out, int n, int c, int h,
                              int w, int vh, int vw, int s, const float* signs_host, float eps,
                              b200r_stream_t stream);
/* dst[i, :] = src[i, :] where mask[i] != 0 (masked accept of Square / best-point bookkeeping) */
int b200r_masked_rows_copy(float* dst, const float* src, const uint8_t* mask, size_t n, size_t chw,
                           b200r_stream_t stream);

/* FAB's projection_linf (Attacks/autoattack/fab_projections.py:7-59): row r of t is projected onto the hyperplane
 * <w[r], x> = b[r] intersected with the box [0,1]^dim, minimising the Linf norm of the step; d receives the step.  Sort-free
 * (Newton on the per-row threshold, double accumulation).  dmax (nullable): device float[rows] = max |d| per row (the a0 of
 * fab_base.py:200); passes (nullable): device int[rows] = number of passes each row took (diagnostics). */
int b200r_fab_projection_linf(const float* t, const float* w, const float* b, float* d, float* dmax, int rows,
                              int dim, int* passes, b200r_stream_t stream);
/* FAB's update (fab_base.py:200-232): alpha = min(max(a1 / (a1 + a2), 0), alpha_max) with a = max(dmax, 1e-8);
 * x1 <- clamp((x1 + eta d1) (1 - alpha) + (x0 + eta d2) alpha, 0, 1), in place. */
int b200r_fab_combine_linf(float* x1, const float* d1, const float* x0, const float* d2, const float* dmax1,
                           const float* dmax2, int rows, int dim, float eta, float alpha_max, b200r_stream_t stream);
/* APGD-L1's L1_projection (Attacks/autoattack/autopgd_base.py:19-83): x = centre, y = perturbation, delta such that
 * ||y + delta||_1 <= eps and 0 <= x + y + delta <= 1.  Sort-free (safeguarded Newton on the threshold). */
int b200r_l1_projection(const float* x, const float* y, float* delta, int rows, int dim, float eps,
                        b200r_stream_t stream);
/* One PGD-L1 step of ART's ProjectedGradientDescentPyTorch(norm=1) (called from adv/attack.py:44-49; ART itself is not
 * vendored by the reference): x <- x0 + scale * (clip(x + eps_step g / (||g||_1 + 1e-7), 0, 1) - x0),
 * scale = min(1, eps / (||.||_1 + 1e-7)); in place, one CTA per sample. */
int b200r_pgd_step_l1(float* x, const float* g, const float* x0, int rows, int dim, float eps_step, float eps,
                      b200r_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dense contractions on the tcgen05 tensor cores ("split" activations: every fp32 tensor
 * is stored as two fp16 planes hi = fp16(v), lo = fp16(v - hi), 22 significant bits; a product uses
 * hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM).
 * ------------------------------------------------------------------------------------------ */
#define B200R_PASSES_F16 16
enum b200r_act { B200R_ACT_NONE = 0, B200R_ACT_RELU = 1, B200R_ACT_RELU6 = 2,
                 B200R_ACT_GELU_TANH = 3, B200R_ACT_GELU_ERF = 4, B200R_ACT_SWISH = 5,
                 B200R_ACT_TANH = 6, B200R_ACT_SIGMOID = 7 };

/* split a float32 tensor into hi/lo fp16 planes: planes[0:count] = hi = rn16(v), planes[count:2count] = lo = rn16(v - hi) */
int b200r_split_f32(const float* in, uint16_t* planes, size_t count, b200r_stream_t stream);
/* the same of scale * v (the loss scale of a gradient pass) */
int b200r_split_f32_scaled(const float* in, uint16_t* planes, size_t count, float scale, b200r_stream_t stream);
int b200r_merge_f32(const uint16_t* planes, float* out, size_t count, b200r_stream_t stream);

/* Implicit-GEMM convolution, NHWC, replaces nn.Conv2d(bias=False)+BatchNorm2d(eval)+act(+residual)
 * (resnet_official.py:71-140,330-346).
 *   x      : split planes of [n, h, w, cin]           (cin % 8 == 0; K tails are zero-filled by TMA)
 *   wgt    : split planes of [cout, kh, kw, cin]      (any cout; cout % 16 == 0 takes the vector epilogue)
 *   scale, bias : float32 [cout] (folded BN; nullable = 1 / 0)
 *   res    : split planes of [n, ho, wo, cout] added before the activation (nullable)
 *   y      : split planes of [n, ho, wo, cout] (nullable if y_f32 given)
 *   y_f32  : float32 [n, ho, wo, cout] (nullable)
 *   passes : 3 = hi*hi+hi*lo+lo*hi (fp32-faithful), 1 = hi*hi only (plain fp16 on split storage),
 *            B200R_PASSES_F16 = every tensor argument is ONE plane of IEEE fp16 (x, wgt, res, y are then
 *            [1][...] instead of [2][...]): one MMA per product, fp32 accumulation -- the mantissa of the TF32
 *            path the reference's own GPU convolutions take by default (torch.backends.cudnn.allow_tf32).
 *            TF32-class accuracy: 2.6e-2 .. 4.7e-2 max logit error at realistic logit magnitude (DESIGN.md
 *            section 5) -- NOT within the 1e-3 tolerance; passes = 3 is.
 *            Accepted by conv2d / linear / stem_conv7x7 / conv2d_dgrad; the elementwise layers have _f16 twins.
 */
int b200r_conv2d_nhwc(const uint16_t* x, const uint16_t* wgt, const float* scale, const float* bias,
                      const uint16_t* res, uint16_t* y, float* y_f32, int n, int h, int w, int cin,
                      int cout, int kh, int kw, int stride, int pad, int act, int passes,
                      b200r_stream_t stream);

/* y[m, nout] = act(x[m,k] . wgt[nout,k]^T * scale + bias (+res)) -- nn.Linear (fc heads, ViT/Mixer
 * MLPs).  k % 8 == 0; nout arbitrary (tiles are masked). */
int b200r_linear(const uint16_t* x, const uint16_t* wgt, const float* scale, const float* bias,
                 const uint16_t* res, uint16_t* y, float* y_f32, int m, int k, int nout, int act,
                 int passes, b200r_stream_t stream);

/* The Linear a gradient pass keeps (vision_transformer.py:62-75 / mlp_mixer.py MLP blocks, where autograd saves the
 * pre-activation): one launch writes pre = x . wgt^T + bias to `pre` and act(pre) to `y`, both as split planes
 * [2][m][nout], bit-identical to b200r_linear(act = none) followed by b200r_act_planes.  Split precision only (three
 * passes); act must be one of the smooth activations (gelu / swish / tanh / sigmoid / relu6); nout % 8 == 0. */
int b200r_linear_keep_pre(const uint16_t* x, const uint16_t* wgt, const float* bias, uint16_t* y, uint16_t* pre,
                          int m, int k, int nout, int act, b200r_stream_t stream);

/* 7x7/s2 stem patches (resnet_official.py:221-224): u8 NHWC image -> normalised, im2col'd split
 * planes [n*ho*wo, kpad] with kpad = 192: column = ky*24 + kx*3 + c (kx < 7); each ky run is padded to 8
 * taps, so columns ky*24+21..23 and 168..191 are zero. */
int b200r_stem_im2col_u8(const uint8_t* img, uint16_t* planes, int n, int h, int w,
                         const float* mean_host, const float* std_host, b200r_stream_t stream);
/* same from float32 NCHW in [0,1] (the attack path) */
int b200r_stem_im2col_f32(const float* img, uint16_t* planes, int n, int h, int w,
                          const float* mean_host, const float* std_host, b200r_stream_t stream);

/* Fused stem: conv1 7x7/s2/p3 (3 -> 64) + BN + act straight from the raw uint8 NHWC image
 * (resnet_official.py:221-226,331-333 after ToTensor+Normalize, imagenet_dataloader.py:78-79): the patch
 * gather, x/255, (x-mean)/std and the hi/lo split happen while the operand tile is written to shared
 * memory; nothing but the image is read from HBM.  wgt: split planes [64, 192] in the im2col column order
 * above (column = ky*24 + kx*3 + c; the 45 padding columns MUST be zero); y: split planes [n, h/2, w/2, 64]. */
int b200r_stem_conv7x7_u8(const uint8_t* img, const uint16_t* wgt, const float* scale, const float* bias,
                          uint16_t* y, int n, int h, int w, const float* mean_host,
                          const float* std_host, int act, int passes, b200r_stream_t stream);

/* Whole ResNet stem in one launch, fp16 mode only: conv1 7x7/s2/p3 + BN + ReLU + MaxPool2d(3,2,1) from the raw uint8
 * NHWC image (resnet_official.py:221-227,330-334 after ToTensor+Normalize).  Implicit im2col through overlapping
 * no-swizzle UMMA descriptors over a staged row buffer, one N = 256/192 MMA per input row and K step over a ring of 8
 * TMEM accumulators, BN bias on the tensor core, pooling in the epilogue (csrc/stem_pool_sm100.cu); the 112x112
 * activation never reaches HBM.
 *   wgt  : ONE fp16 plane [64, 192] in the column order of b200r_stem_im2col_u8, BN scale ALREADY FOLDED IN
 *          (w * gamma / sqrt(var + eps), rounded to fp16 once)
 *   bias : float32 [64] folded BN bias (nullable); enters as an fp16 hi/lo pair, ~2^-22 relative
 *   y    : ONE fp16 plane [n, h/4, w/4, 64]
 * Normalisation is fp16(fma(byte, 1/(255 std), -mean/std)) in fp32.  Requires h % 4 == 0, w % 8 == 0, w <= 248. */
int b200r_stem_pool_u8_f16(const uint8_t* img, const uint16_t* wgt, const float* bias,
                           uint16_t* y, int n, int h, int w, const float* mean_host,
                           const float* std_host, b200r_stream_t stream);

/* Split-precision (fp32-faithful) twin of the one-launch stem: same schedule, an EXACT one-plane A operand (pixel / 256 as
 * fp16, fourth channel = 1 on real pixels, 0 in the padding) against the weights as an fp16 hi/lo pair -- two MMAs per product
 * instead of three -- fp32 pooling, and the hi/lo split of the pooled value on the way out.
 * b200r_stem_pool_split_prepare (HOST function, double arithmetic) folds x/255, 1/std, the BN scale, -mean/std, the BN bias
 * and a range-centring power of two into the operand:
 *   conv1_w      : host float32 [64, 3, 7, 7] (resnet_official.py:221)
 *   bn_scale/bias: host float32 [64], gamma / sqrt(var + eps) and beta - mean * scale (nullable = 1 / 0)
 *   f32_input    : 0 = the kernel stages byte / 256 (b200r_stem_pool_u8_split), 1 = a float image in [0, 1] (b200r_stem_pool_f32_split)
 *   planes_host  : host uint16 [2][64][224] out (hi plane, lo plane; row = [ky][t = kx + 1][R, G, B, 1]); upload as is
 *   out_scale    : 2^-k, to hand to b200r_stem_pool_u8_split
 * b200r_stem_pool_u8_split: img uint8 NHWC [n, h, w, 3]; wplanes = the uploaded planes; y = split planes [2][n, h/4, w/4, 64].
 * Same geometry limits as the fp16 twin. */
int b200r_stem_pool_split_prepare(const float* conv1_w, const float* bn_scale, const float* bn_bias,
                                  const float* mean_host, const float* std_host, int f32_input,
                                  uint16_t* planes_host, float* out_scale);
int b200r_stem_pool_u8_split(const uint8_t* img, const uint16_t* wplanes, float out_scale, uint16_t* y,
                             int n, int h, int w, b200r_stream_t stream);
/* The same from a float32 NCHW image in [0, 1] (the attack loops' iterate) for a SAVED forward: the pixel enters as an fp16 hi/lo pair
 * (three MMAs per product; prepare the operand with f32_input = 1: the colour weights then meet x itself, not byte / 256), and next
 * to the pooled planes the kernel writes the pool's arg-max codes [n, h/4, w/4, 64] (uint8: window position ky*3+kx of the first
 * maximum, 0xF where the maximum is not positive = the stem ReLU's backward folded in) for b200r_maxpool3x3s2_bwd_codes_hi -- the
 * 112 x 112 activation is never written and never needed by the gradient pass. */
int b200r_stem_pool_f32_split(const float* x01, const uint16_t* wplanes, float out_scale, uint16_t* y, uint8_t* codes,
                              int n, int h, int w, b200r_stream_t stream);

/* same, from a float32 NCHW image in [0,1] (the attack loops' iterate): (x - mean) / std and the hi/lo split happen
 * in the operand producer, with the arithmetic of b200r_stem_im2col_f32 */
int b200r_stem_conv7x7_f32(const float* img, const uint16_t* wgt, const float* scale, const float* bias,
                           uint16_t* y, int n, int h, int w, const float* mean_host,
                           const float* std_host, int act, int passes, b200r_stream_t stream);

/* Image.resize of Pillow, bit-exact, on uint8 NHWC batches, with a crop window of the resized image folded in
 * (csrc/resize.cu).  Replaces ImageTransfer.image_resize for the `pil-*` resize types
 * (RobustART/noise/utils/imagenet_s_gen.py:19-26,120-141: resize to (s*8/7)^2, centre crop s) and torchvision's
 * Resize + CenterCrop of the eval transform (imagenet_dataloader.py:74-80).
 *   in  : [n, hin, win, 3]       out : [n, ch, cw, 3] = resized[oy0 : oy0 + ch, ox0 : ox0 + cw]
 *   filter : B200R_RESIZE_* = PIL.Image.{NEAREST, BOX, BILINEAR, HAMMING, BICUBIC, LANCZOS}
 *   workspace : b200r_resize_workspace_bytes() bytes (the uint8 image between the horizontal and the vertical pass)
 * The first call per (device, in size, out size, filter) builds and uploads a coefficient table (blocking). */
enum b200r_resize_filter { B200R_RESIZE_NEAREST = 0, B200R_RESIZE_BOX = 1, B200R_RESIZE_BILINEAR = 2,
                           B200R_RESIZE_HAMMING = 3, B200R_RESIZE_BICUBIC = 4, B200R_RESIZE_LANCZOS = 5 };
int b200r_resize_workspace_bytes(int n, int hin, int win, int hout, int wout, int filter, int oy0, int ox0,
                                 int ch, int cw, size_t* bytes);
int b200r_resize_u8(const uint8_t* in, uint8_t* out, int n, int hin, int win, int hout, int wout, int filter,
                    int oy0, int ox0, int ch, int cw, void* workspace, size_t ws_bytes, b200r_stream_t stream);

/* cv2.resize(INTER_NEAREST | INTER_LINEAR | INTER_AREA | INTER_LANCZOS4) for uint8 NHWC batches, bit for bit, and INTER_CUBIC as
 * the float32 cubic the opencv-python wheels compute through IPP (1 LSB on <= 1e-4 of the pixels) -- the five `opencv-*` resize types of RobustART/noise/utils/imagenet_s_gen.py:28-34,120-148), cropped to the window [oy0, oy0+ch) x [ox0, ox0+cw) of the
 * hout x wout result: out is [n, ch, cw, 3].  interpolation uses cv2's own constants.  No workspace; cubic / lanczos4
 * upload a cached weight table on the first call per geometry (blocking). */
enum b200r_cv_interpolation { B200R_CV_INTER_NEAREST = 0, B200R_CV_INTER_LINEAR = 1, B200R_CV_INTER_CUBIC = 2, B200R_CV_INTER_AREA = 3,
                              B200R_CV_INTER_LANCZOS4 = 4 };
int b200r_resize_cv_u8(const uint8_t* in, uint8_t* out, int n, int hin, int win, int hout, int wout, int interpolation,
                       int oy0, int ox0, int ch, int cw, b200r_stream_t stream);

/* Direct 3x3/s2/p1 convolution from the image (3 -> cout <= 64) + folded BN + activation, fp32 on CUDA cores: the stems
 * of MobileNetV2 (mobilenet_v2.py:130) and EfficientNet-B0 (efficientnet.py:429-433), ToTensor + Normalize on the fly.
 * img: uint8 NHWC [n,h,w,3] or float32 NCHW in [0,1]; wgt: float32 [cout][27], column = (ky*3 + kx)*3 + c;
 * y: split planes [n, ho, wo, cout]. */
int b200r_image_stem3x3s2_u8(const uint8_t* img, const float* wgt, const float* scale, const float* bias, uint16_t* y,
                             int n, int h, int w, int cout, int act, const float* mean_host, const float* std_host,
                             b200r_stream_t stream);
int b200r_image_stem3x3s2_f32(const float* img, const float* wgt, const float* scale, const float* bias, uint16_t* y,
                              int n, int h, int w, int cout, int act, const float* mean_host, const float* std_host,
                              b200r_stream_t stream);

/* MaxPool2d(3, 2, 1) on split planes NHWC (resnet_official.py:227) */
int b200r_maxpool3x3s2_nhwc(const uint16_t* x, uint16_t* y, int n, int h, int w, int c,
                            b200r_stream_t stream);
/* AdaptiveAvgPool2d(1) on split planes NHWC -> split planes [n, c] (resnet_official.py:238) */
int b200r_global_avgpool_nhwc(const uint16_t* x, uint16_t* y, int n, int hw, int c,
                              b200r_stream_t stream);

/* fp16 single-plane twins of the layers above (activations produced with passes = B200R_PASSES_F16) */
int b200r_f32_to_f16(const float* in, uint16_t* out, size_t count, float scale, b200r_stream_t stream);
int b200r_f16_to_f32(const uint16_t* in, float* out, size_t count, float scale, b200r_stream_t stream);
int b200r_maxpool3x3s2_nhwc_f16(const uint16_t* x, uint16_t* y, int n, int h, int w, int c,
                                b200r_stream_t stream);
int b200r_global_avgpool_nhwc_f16(const uint16_t* x, uint16_t* y, int n, int hw, int c,
                                  b200r_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Input-gradient pass (d loss / d image) of the convolutional models: what the attack loops call once
 * per step (foolbox value_and_grad in attack.py:20-33; autopgd_base.py:371-376; imfgsm_attack.py:77-84).
 * Only INPUT gradients exist on this path -- no weight gradients.  The contractions are b200r_conv2d_nhwc /
 * b200r_linear on transposed (and for 3x3: spatially flipped) weights; a stride-2 convolution's dgrad is
 * dilate2 followed by the stride-1 convolution.  Everything stays in split planes.
 * ------------------------------------------------------------------------------------------ */
/* Input gradient of a stride-1 convolution with the ReLU backward of the layer below fused into the epilogue:
 *   dx = [mask > 0] * ( conv(dy, wgt_t) + res )
 *   dy    : split planes [n, h, w, cdy]          (gradient w.r.t. the convolution's output; zero-dilated for stride 2)
 *   wgt_t : split planes [cdx, kh, kw, cdy]      = forward weight [cdy, kh, kw, cdx] transposed and flipped in (kh, kw)
 *   res   : split planes [n, h, w, cdx], nullable (the gradient arriving over the block's identity path)
 *   mask  : the HI plane of the post-ReLU activation the gradient flows into, [n, h, w, cdx] fp16, nullable
 *   pad   : (k-1)/2 (the forward convolutions on this path are 'same'-padded)
 * Same kernel, tiling and passes as b200r_conv2d_nhwc. */
int b200r_conv2d_dgrad_nhwc(const uint16_t* dy, const uint16_t* wgt_t, const uint16_t* res,
                            const uint16_t* mask, uint16_t* dx, int n, int h, int w, int cdy, int cdx,
                            int kh, int kw, int pad, int passes, b200r_stream_t stream);
/* Input gradient of a 3x3 / stride 2 / pad 1 convolution (resnet_official.py:112 conv2 of the first block of a stage) without the
 * zero-dilated gradient: dx[2i+a, 2j+b] only meets the taps of the flipped kernel whose parity matches, so each of the four parity
 * classes (a, b) is a (1+a) x (1+b)-tap stride-1 convolution of dy, written through a strided tensor map into its quarter of dx --
 * a quarter of the tensor-core work of dilate + 3x3 and no dilated tensor in HBM.
 *   dy   : planes [n, ho, wo, cdy]          dx, res, mask : planes [n, 2 ho, 2 wo, cdx] (res / mask nullable, as b200r_conv2d_dgrad_nhwc)
 *   wab  : planes [cdx, 1+a, 1+b, cdy] = wgt_t[:, S_a, S_b, :] with S_0 = {1}, S_1 = {0, 2}, wgt_t the transposed + flipped weight
 *          of b200r_conv2d_dgrad_nhwc */
int b200r_conv2d_dgrad3x3s2_nhwc(const uint16_t* dy, const uint16_t* w00, const uint16_t* w01, const uint16_t* w10,
                                 const uint16_t* w11, const uint16_t* res, const uint16_t* mask, uint16_t* dx, int n,
                                 int ho, int wo, int cdy, int cdx, int passes, b200r_stream_t stream);

/* Input gradient of a 1x1 / stride-2 convolution (the ResNet downsample branch, resnet_official.py:255-262), ACCUMULATED in place
 * into a gradient that already holds the main branch: dx[:, 2i, 2j, :] = [mask > 0] * (dx[:, 2i, 2j, :] + dy[:, i, j, :] . wgt_t^T),
 * other positions untouched.  dy planes [n][(h+1)/2][(w+1)/2][cdy], wgt_t [cdx][cdy], dx / mask planes [n][h][w][cdx] (mask may
 * be NULL; it is the one already applied to dx, so masking the sum equals masking the new term).  cdx % 8 == 0.  Replaces
 * b200r_conv2d_dgrad_nhwc on the small map + b200r_dilate2_nhwc + the residual operand of the main branch's last dgrad. */
int b200r_conv2d_dgrad1x1s2_acc_nhwc(const uint16_t* dy, const uint16_t* wgt_t, const uint16_t* mask, uint16_t* dx,
                                     int n, int h, int w, int cdy, int cdx, int passes, b200r_stream_t stream);
/* ReLU backward: out = (act > 0 ? dy : 0) + add; add may be NULL (a second gradient branch joining here,
 * e.g. the identity path of a residual block).  count = elements per plane, multiple of 8. */
int b200r_relu_bwd(const uint16_t* dy, const uint16_t* act, const uint16_t* add, uint16_t* out,
                   size_t count, b200r_stream_t stream);
/* y[n, 2i, 2j, :] = x[n, i, j, :], zero elsewhere; x planes [n,h,w,c] -> y planes [n,2h,2w,c] */
int b200r_dilate2_nhwc(const uint16_t* x, uint16_t* y, int n, int h, int w, int c,
                       b200r_stream_t stream);
/* MaxPool2d(3,2,1) backward (resnet_official.py:227): x = the pool's forward input planes [n,h,w,c],
 * dy planes [n,ho,wo,c] -> dx planes [n,h,w,c]; ties go to the first maximum in (ky,kx) scan order as in
 * the forward kernel and in PyTorch.  workspace: n*ho*wo*c bytes (arg-max codes), 8-byte aligned. */
int b200r_maxpool3x3s2_bwd_nhwc(const uint16_t* x, const uint16_t* dy, uint16_t* dx, void* workspace,
                                size_t ws_bytes, int n, int h, int w, int c, b200r_stream_t stream);
/* MaxPool2d(3,2,1) backward with the backward of the ReLU that produced x fused in (a window whose maximum is not positive routes
 * nothing: relu'(0) = 0 as torch has it), ONE fp16 plane out whatever the precision: the ResNet stem's gradient GEMM reads a single
 * plane (resnet_official.py:225-227).  x / dy: split planes (planes = 2) or one fp16 plane (planes = 1); dx_hi: [n, h, w, c] fp16. */
int b200r_maxpool3x3s2_relu_bwd_hi(const uint16_t* x, const uint16_t* dy, uint16_t* dx_hi, void* workspace,
                                   size_t ws_bytes, int n, int h, int w, int c, int planes, b200r_stream_t stream);
/* The same in two halves for a saved forward (the attack path): the forward pool also writes, per pooled element, the window position
 * of its first maximum (one byte; 0xF where the maximum is not positive = the producing ReLU's backward folded in), and the backward
 * routes dy through those codes without touching the 112 x 112 activation again.
 *   codes: [n, ho, wo, c] bytes, 8-byte aligned.  b200r_maxpool3x3s2_nhwc_codes takes split planes; dy of the backward: planes = 2 / 1. */
int b200r_maxpool3x3s2_nhwc_codes(const uint16_t* x, uint16_t* y, void* codes, int n, int h, int w, int c,
                                  b200r_stream_t stream);
int b200r_maxpool3x3s2_bwd_codes_hi(const void* codes, const uint16_t* dy, uint16_t* dx_hi, int n, int h, int w, int c,
                                    int planes, b200r_stream_t stream);
/* AdaptiveAvgPool2d(1) backward: dy planes [n,c] -> dx planes [n,hw,c] = dy / hw */
int b200r_global_avgpool_bwd_nhwc(const uint16_t* dy, uint16_t* dx, int n, int hw, int c,
                                  b200r_stream_t stream);
/* Transpose of b200r_stem_im2col_f32 composed with Normalize: dcols planes [n*(h/2)*(w/2), 192] (column order
 * of b200r_stem_im2col_u8) -> float32 NCHW gradient w.r.t. the [0,1] image, dx[n,c,y,x] = sum(taps) / std[c]. */
int b200r_stem_col2im_f32(const uint16_t* dcols, float* dx, int n, int h, int w, const float* std_host,
                          b200r_stream_t stream);
/* fp16 single-plane twins (gradients produced with passes = B200R_PASSES_F16).  Gradients of a classifier are
 * small (1e-4 .. 1e-10 on the golden ResNets), so the fp16 pass runs on dlogits * S (b200r_f32_to_f16's scale; the
 * host mirror uses S = 4096) and `unscale` = 1/S is applied with 1/std when the image gradient is written. */
int b200r_relu_bwd_f16(const uint16_t* dy, const uint16_t* act, const uint16_t* add, uint16_t* out,
                       size_t count, b200r_stream_t stream);
int b200r_dilate2_nhwc_f16(const uint16_t* x, uint16_t* y, int n, int h, int w, int c,
                           b200r_stream_t stream);
int b200r_maxpool3x3s2_bwd_nhwc_f16(const uint16_t* x, const uint16_t* dy, uint16_t* dx, void* workspace,
                                    size_t ws_bytes, int n, int h, int w, int c, b200r_stream_t stream);
int b200r_global_avgpool_bwd_nhwc_f16(const uint16_t* dy, uint16_t* dx, int n, int hw, int c,
                                      b200r_stream_t stream);
int b200r_stem_col2im_f32_f16(const uint16_t* dcols, float* dx, int n, int h, int w,
                              const float* std_host, float unscale, b200r_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Mobile families (MobileNetV2: prototype/prototype/model/mobilenet_v2.py:31-202; EfficientNet-B0:
 * prototype/prototype/model/efficientnet.py:289-495): the layers that are not dense contractions.
 * ------------------------------------------------------------------------------------------ */
/* depthwise k x k conv + folded BN + act on split planes NHWC.  wgt: float32 [k*k][c] (tap major) */
int b200r_dwconv_nhwc(const uint16_t* x, const float* wgt, const float* scale, const float* bias,
                      uint16_t* y, int n, int h, int w, int c, int k, int stride, int pad, int act,
                      b200r_stream_t stream);
/* 1x1 convolution with a small input width on CUDA cores, exact fp32: y = act(x W^T + bias (+ res)) for cin in {8, 16, 24, 32} -- the
 * expansion / projection layers of the mobile families at 112 x 112 and 56 x 56 (mobilenet_v2.py:52-60; efficientnet.py:312-321), where
 * a tensor-core tile is one mostly-empty k-block and the layer is pure output streaming.
 *   x: split planes [m, cin]; wgt: float32 [cout][cin] (BN scale folded in); bias: float32 [cout] (nullable);
 *   res: split planes [m, cout] (nullable); y: split planes [m, cout]; cout % 8 == 0, cout * cin * 4 <= 96 KB. */
int b200r_pointwise_smallk_nhwc(const uint16_t* x, const float* wgt, const float* bias, const uint16_t* res, uint16_t* y,
                                size_t m, int cin, int cout, int act, b200r_stream_t stream);
/* squeeze-excite scaling: y[n, p, c] = x[n, p, c] * s[n, c]  (s: split planes [n, s_stride]) */
int b200r_channel_scale(const uint16_t* x, const uint16_t* s, uint16_t* y, int n, int hw, int c,
                        int s_stride, b200r_stream_t stream);
/* k x k patches of the 3-channel input image as a GEMM operand: planes [n*ho*wo, kpad],
 * column = (ky*k + kx)*3 + c, zero padded to kpad; ToTensor+Normalize fused */
int b200r_image_im2col_u8(const uint8_t* img, uint16_t* planes, int n, int h, int w, int k, int stride,
                          int pad, int kpad, const float* mean_host, const float* std_host,
                          b200r_stream_t stream);
int b200r_image_im2col_f32(const float* img, uint16_t* planes, int n, int h, int w, int k, int stride,
                           int pad, int kpad, const float* mean_host, const float* std_host,
                           b200r_stream_t stream);

/* Input-gradient pieces of the mobile families (autograd of mobilenet_v2.py:31-77 / efficientnet.py:312-360 as the attack loops
 * see it).  The depthwise convolution's input gradient is b200r_dwconv_nhwc on flipped taps (stride 2: after b200r_dilate2_nhwc),
 * the 1x1 convolutions are b200r_conv2d_dgrad_nhwc, activations b200r_act_bwd_planes (ReLU6 from the saved output, swish / sigmoid
 * from the saved pre-activation). */
/* squeeze-excite: ds planes [n, s_stride] = sum over pixels of a * b (a, b planes [n, hw, c]); columns >= c are zeroed */
int b200r_channel_dot(const uint16_t* a, const uint16_t* b, uint16_t* ds, int n, int hw, int c, int s_stride,
                      b200r_stream_t stream);
/* out = a + b on split planes (two gradient branches joining); count = elements per plane, multiple of 8 */
int b200r_planes_add(const uint16_t* a, const uint16_t* b, uint16_t* out, size_t count, b200r_stream_t stream);
/* transpose of b200r_image_stem3x3s2_f32 composed with Normalize: dy planes [n, ho, wo, cout] -> float32 NCHW gradient w.r.t. the
 * [0,1] image, times unscale; wgt float32 [cout][27] with the BatchNorm scale folded in */
int b200r_image_stem3x3s2_bwd(const uint16_t* dy, const float* wgt, float* dx, int n, int h, int w, int cout,
                              const float* std_host, float unscale, b200r_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Token models (ViT-B/16: prototype/prototype/model/vision_transformer.py:44-349; MLP-Mixer-B/16:
 * prototype/prototype/model/vit/mlp_mixer.py:7-159), split planes [rows, c].
 * ------------------------------------------------------------------------------------------ */
/* nn.LayerNorm over the last dimension (biased variance) */
int b200r_layernorm(const uint16_t* x, uint16_t* y, const float* gamma, const float* beta, int rows,
                    int c, float eps, b200r_stream_t stream);
/* patch embedding operand: [n*(h/p)*(w/p), 3*p*p], column = c*p*p + ky*p + kx (== conv weight
 * .reshape(D, -1)), ToTensor+Normalize fused (u8) / Normalize fused (float NCHW in [0,1]) */
int b200r_patch_gather_u8(const uint8_t* img, uint16_t* y, int n, int h, int w, int patch,
                          const float* mean_host, const float* std_host, b200r_stream_t stream);
int b200r_patch_gather_f32(const float* img, uint16_t* y, int n, int h, int w, int patch,
                           const float* mean_host, const float* std_host, b200r_stream_t stream);
/* y[n, 1+np, c] = cat(cls, x[n, np, c]) + pos[1+np, c]   (cls, pos: float32) */
int b200r_assemble_tokens(const uint16_t* x, const float* cls, const float* pos, uint16_t* y, int n,
                          int num_patches, int c, b200r_stream_t stream);
/* out[n*t, h*d] = softmax(q k^T * scale) v with qkv[n*t, 3*h*d] packed "(qkv h d)" */
int b200r_attention(const uint16_t* qkv, uint16_t* out, int n, int tokens, int heads, int head_dim,
                    float scale, b200r_stream_t stream);
/* Mixer token mixing: y[b, c, t_pad] = x[b, t, c] (zero padded), and out[b,t,c] = res[b,t,c] + y[b,c,t] */
int b200r_tokens_to_channels(const uint16_t* x, uint16_t* y, int b, int t, int c, int t_pad,
                             b200r_stream_t stream);
int b200r_channels_to_tokens_add(const uint16_t* y, const uint16_t* res, uint16_t* out, int b, int t,
                                 int c, int t_pad, b200r_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Input-gradient pass of the token models (what autograd.grad(loss, x) hands the attacks:
 * RobustART/noise/utils/adv/autoattack/autopgd_base.py:371-376, Attacks/imfgsm_attack.py:77-80); split planes.
 * The Linear layers' input gradients are b200r_linear calls with the transposed weights.
 * ------------------------------------------------------------------------------------------ */
/* nn.LayerNorm backward w.r.t. its input: dx = add + rstd (dy g - mean(dy g) - xhat mean(dy g xhat)); `add` (the
 * gradient arriving over the residual connection) may be NULL */
int b200r_layernorm_bwd(const uint16_t* dy, const uint16_t* x, const float* gamma, const uint16_t* add,
                        uint16_t* dx, int rows, int c, float eps, b200r_stream_t stream);
/* out = act(pre) and dx = dy * act'(pre) on a saved pre-activation; act in {GELU_TANH, GELU_ERF, TANH};
 * count = elements per plane, a multiple of 8 */
int b200r_act_planes(const uint16_t* pre, uint16_t* out, size_t count, int act, b200r_stream_t stream);
int b200r_act_bwd_planes(const uint16_t* dy, const uint16_t* pre, uint16_t* dx, size_t count, int act,
                         b200r_stream_t stream);
/* transpose of b200r_patch_gather_f32: dcols planes [n*(h/p)*(w/p), 3*p*p] -> float32 NCHW gradient w.r.t. the
 * [0,1] image, dx[n,c,y,x] = dcols[...] / std[c]; patch a multiple of 8 */
int b200r_patch_scatter_f32(const uint16_t* dcols, float* dx, int n, int h, int w, int patch,
                            const float* std_host, b200r_stream_t stream);
/* backward of b200r_attention: dout planes [n*t, h*d] -> dqkv planes [n*t, 3*h*d] in the packing of qkv; the
 * probabilities are recomputed from qkv (nothing else is saved by the forward) */
int b200r_attention_bwd(const uint16_t* qkv, const uint16_t* dout, uint16_t* dqkv, int n, int tokens, int heads,
                        int head_dim, float scale, b200r_stream_t stream);
/* the same on the tensor cores (csrc/attention_bwd_sm100.cu: two tcgen05 launches, query tiles then key tiles) when tokens <= 256;
 * workspace: b200r_attention_bwd_workspace_bytes() bytes for the per-query softmax statistics.  Falls back to the CUDA-core kernel
 * above for longer sequences or without a workspace. */
int b200r_attention_bwd_workspace_bytes(int n, int tokens, int heads, size_t* bytes);
int b200r_attention_bwd_ws(const uint16_t* qkv, const uint16_t* dout, uint16_t* dqkv, void* workspace, size_t ws_bytes, int n,
                           int tokens, int heads, int head_dim, float scale, b200r_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Model handles: build a classifier from the reference's state_dict tensors and run it without Python.
 * Replaces model_entry() + nn.Module.forward for the ResNet family (prototype/prototype/model/resnet_official.py:40-140,
 * 221-239,330-346), ViT-B/16 (vision_transformer.py:44-349) and MLP-Mixer-B/16 (vit/mlp_mixer.py:7-159), and autograd.grad(loss, x) of the attack loops (autopgd_base.py:371-376; foolbox value_and_grad behind
 * adv/attack.py:20-33).  The handle owns its device weights and an activation arena; one handle per (process, device);
 * calls on one handle must not overlap.  Every layer is a launch of the entry points above on `stream`; nothing synchronises.
 * ------------------------------------------------------------------------------------------ */
typedef struct b200r_model b200r_model;
enum b200r_arch { B200R_ARCH_RESNET18 = 0, B200R_ARCH_RESNET34 = 1, B200R_ARCH_RESNET50 = 2, B200R_ARCH_RESNET101 = 3,
                  /* ViT-B/16 (vision_transformer.py:44-349; state_dict keys embedding.*, cls_token, pos_embedding,
                   * transformer.encoders.encoder_<i>.{norm1,norm2,attention.to_qkv,attention.to_out,feedforward.mlp1,feedforward.mlp2}.*,
                   * transformer.encoder_norm.*, [pre_logits.*,] head.*): split precision only */
                  B200R_ARCH_VIT_B16 = 4,
                  /* MLP-Mixer-B/16 (vit/mlp_mixer.py:7-159; keys patch_embed.proj.*, blocks.<i>.{norm1,norm2,token_mix.fc1,token_mix.fc2,
                   * channel_mix.fc1,channel_mix.fc2}.*, norm.*, head.*): split precision only */
                  B200R_ARCH_MIXER_B16 = 5,
                  /* MobileNetV2 x1.0 (mobilenet_v2.py:80-202; keys features.*, classifier.1.*) and EfficientNet-B0 (efficientnet.py:91-125,
                   * 289-495; keys stem.*, blocks.<i>.{in_conv,se_block,out_conv}.*, head.*, fc.*): split precision, INFERENCE handles --
                   * b200r_model_forward_u8 only (the ImageNet-C sweep is evaluation); forward_f32 / input_grad return B200R_ENOTSUP */
                  B200R_ARCH_MOBILENET_V2 = 6, B200R_ARCH_EFFICIENTNET_B0 = 7 };
/* one state_dict entry: the reference's key ("layer1.0.conv1.weight", "bn1.running_var", "fc.bias", ...; optional "module." /
 * "base_model." prefixes are stripped, benchmark_eval_adv.py:162-168), HOST float32 data in PyTorch layout, element count */
typedef struct b200r_weight { const char* name; const float* data; int64_t numel; } b200r_weight;
/* passes: 3 = split planes (fp16 hi + lo, three MMAs per product, fp32-faithful) or B200R_PASSES_F16 (one fp16 plane, TF32-class).
 * BatchNorm is folded (eval mode), weights are re-laid out and uploaded; blocking. */
int b200r_model_create(int arch, const b200r_weight* weights, int n_weights, int passes, b200r_model** out);
int b200r_model_destroy(b200r_model* model);
int b200r_model_num_classes(const b200r_model* model);
/* size the activation arena for batches of n images of h x w (for_input_grad: also for forward_f32 + input_grad).  The forward
 * calls grow it on demand (cudaMalloc: blocking, not capturable); reserve first when capturing into a CUDA graph. */
int b200r_model_reserve(b200r_model* model, int n, int h, int w, int for_input_grad);
/* logits[n, classes] (device float32) from raw uint8 NHWC pixels (ToTensor + Normalize fused into the stem) */
int b200r_model_forward_u8(b200r_model* model, const uint8_t* images_nhwc, float* logits, int n, int h, int w,
                           b200r_stream_t stream);
/* the same from a float32 NCHW image in [0,1] (the attack loops' iterate); keeps the activations for b200r_model_input_grad */
int b200r_model_forward_f32(b200r_model* model, const float* x01_nchw, float* logits, int n, int h, int w,
                            b200r_stream_t stream);
/* dx[n,3,h,w] = d loss / d x01 from dlogits[n, classes] = d loss / d logits of the last b200r_model_forward_f32 */
int b200r_model_input_grad(b200r_model* model, const float* dlogits, float* dx, b200r_stream_t stream);

/* The evaluation's one collective: in-place sum of int64 counters over an ncclComm_t, on `stream`.  Replaces the
 * reference's per-rank result files + merge (base_dataset.py:116-133) and barrier all-reduces (linklink/__init__.py:37-41).
 * NCCL is resolved at run time (already-loaded symbol, else libnccl.so.2); the library does not link it. */
int b200r_allreduce_counts(void* nccl_comm, int64_t* dev_counts, int count, b200r_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200R_H_ */
