/* contrad_b200 C ABI  --  the drop-in boundary of the B200-native ContraD hot path.
 *
 * The reference (jh-jeong/ContraD) has no C FFI for this path: its operators are Python
 * modules that call ATen.  The product keeps those Python surfaces
 * (contrad_b200/{augment,training,third_party,models}) and routes each operator, as a
 * torch.autograd.Function, through the plain-C entry points declared here.  Each entry point
 * cites the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions (all entry points):
 *   - raw device pointers + explicit sizes; fp32, contiguous unless a leading dimension is given;
 *   - the caller allocates every output and workspace; nothing is allocated or synchronised inside;
 *   - `stream` is a cudaStream_t (the caller's current stream); kernels run on the device that is
 *     current in the calling thread; no global mutable state besides the launch counter;
 *   - return 0 on success, a cudaError_t value or CB200_ERR_* otherwise; cb200_last_error() then
 *     returns a thread-local message.  There is no CPU fallback.
 */
#ifndef CONTRAD_B200_H_
#define CONTRAD_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define CB200_ERR_ARG 1000
#define CB200_ERR_TMAP 1001

/* ---- library plumbing ------------------------------------------------------------------- */
int cb200_version(void);
const char* cb200_last_error(void);
unsigned long long cb200_launch_count(void);     /* kernels launched by this library so far */
void cb200_reset_launch_count(void);
/* Launches recorded while a CUDA graph is being captured did not run, and a replay bypasses the counter: the host
 * corrects it by a signed delta. */
void cb200_add_launch_count(long long n);
int cb200_device_arch(int device, int* major, int* minor);

/* ---- fused SimCLR augmentation ------------------------------------------------------------
 * Replaces nn.Sequential(RandomResizeCropLayer, HorizontalFlipLayer, RandomApply(ColorJitterLayer,.8),
 * RandomApply(RandomColorGrayLayer,.2)) = augment.simclr()  (augment/__init__.py:106-112;
 * augment/spatial.py:84-148; augment/color_jitter.py:44-104; augment/utils.py:27-63).
 *   x, y    [B,3,H,W] images in [0,1], NCHW
 *   params  [11,B] SoA: sx, sy, bx, by (affine theta of the crop), flip (+-1), cj_on (0/1),
 *           contrast, hue, sat, val factors, gray_on (0/1)   -- drawn on the host in the
 *           reference's numpy/torch RNG order
 *   order   0: [contrast, hsv]   1: [hsv, contrast]   (color_jitter.py:65-70, one draw per batch)
 *           -1: read per image from an extra row 11 of `params` ([12, B]; != 0 means 1) - keeps the launch free of
 *               step-dependent host scalars so it can be replayed inside a CUDA graph
 * Backward = autograd of the reference chain (HSV straight-through, color_jitter.py:97-104). */
int cb200_augment_simclr_fwd(const float* x, float* y, const float* params, int B, int H, int W,
                             int order, void* stream);
int cb200_augment_simclr_bwd(const float* x, const float* dy, float* dx, const float* params,
                             int B, int H, int W, int order, void* stream);

/* The [11|12, B] parameter block from raw draws in one launch: boxes [4,B] = the host (numpy) crop draws of
 * RandomResizeCropLayer (augment/spatial.py:119-143), u [7,B] = device U[0,1) draws for flip, apply-jitter, contrast,
 * hue, saturation, value, apply-gray, mapped like the reference (bernoulli(p) = u < p: spatial.py:86-88,
 * augment/__init__.py:100-103; uniform_(lo,hi) = lo + (hi-lo) u: color_jitter.py:44-63).  cfg11 (HOST pointer) =
 * {p_flip, p_jitter, p_gray, contrast lo,hi, hue lo,hi, saturation lo,hi, value lo,hi}; order_src = device scalar
 * written to row 11 when rows == 12 (may be NULL). */
int cb200_augment_simclr_params(const float* boxes, const float* u, const float* order_src, float* params,
                                int B, int rows, const float* cfg11, void* stream);

/* Same chain for images of any size (the entry points above keep one image per CTA in shared memory: H*W <= 4096;
 * these run from global memory and are what the 512x512 StyleGAN2 configs use).  means [B,3] (per-channel mean at the
 * contrast input, written by fwd, read by bwd) and gsums [B,3] (bwd scratch) are caller-allocated. */
int cb200_augment_simclr_large_fwd(const float* x, float* y, const float* params, float* means, int B, int H, int W,
                                   int order, void* stream);
int cb200_augment_simclr_large_bwd(const float* x, const float* dy, float* dx, const float* params, const float* means,
                                   float* gsums, int B, int H, int W, int order, void* stream);

/* Row f3: the dataset's `ToTensor` (datasets.py:10-21: uint8 HWC -> fp32 CHW / 255, here on NCHW bytes),
 * the fp32 host->device copy of train_gan.py:153-154 and the `torch.cat([images, images, gen_images])` of
 * training/gan/contrad.py:38-40 folded into the forward of the chain.  The launch produces B views; view b reads
 *     b <  n_u8_views : the uint8 image (b mod n_u8) of x_u8 [n_u8,3,H,W]      (value / 255, correctly rounded)
 *     b >= n_u8_views : the fp32 image (b - n_u8_views) of x_f32 [B - n_u8_views,3,H,W]
 * params / order as above, over all B views.  means [B,3]: caller-allocated scratch, written on the any-size path
 * (H, W other than 32x32 / 64x64) exactly like cb200_augment_simclr_large_fwd.  Gradients flow to the fp32 images
 * only: call cb200_augment_simclr_bwd / _large_bwd on them with the parameter columns (and means rows) of their views. */
int cb200_augment_simclr_mixed_fwd(const unsigned char* x_u8, int n_u8, int n_u8_views, const float* x_f32, float* y,
                                   const float* params, float* means, int B, int H, int W, int order, void* stream);

/* Tail of `simclr_hq` / `simclr_hq_cutout` (augment/__init__.py:52-78,115-133; augment/spatial.py:151-181):
 * gaussian_blur: y[b] = on[b] ? blur(x[b]) : x[b]; `taps` = the k normalised 1-D Gaussian weights (device memory; the
 *                reference's dense k x k outer-product kernel with 'reflect' padding, applied separably); `tmp` scratch
 *                of x's size; adjoint = 1 applies the transpose (backward).  P = planes per image (3).
 * cutout:        params [3, B] = {on, h centre, w centre}; zeroes the clipped (length x length) square; the backward
 *                pass is the same call on the gradient. */
int cb200_gaussian_blur(const float* x, float* tmp, float* y, const float* taps, const float* on, int B, int P, int H,
                        int W, int k, int adjoint, void* stream);
int cb200_cutout(const float* x, float* y, const float* params, int B, int P, int H, int W, int length, void* stream);

/* Row f4: light augmentations of the CR / bCR baselines (`--aug hfrt`, `gaussian`).
 * shift_flip:  HorizontalFlipRandomCrop / RandomCrop (augment/spatial.py:14-67) = grid_sample(mode='nearest',
 *              padding_mode, align_corners=False) on theta = [[sign, 0, bias_x], [0, 1, bias_y]].
 *              params [3, B] = {sign (+-1), bias_x, bias_y} (bias = integer shift / (width / 2), as the reference draws
 *              it); padding_mode 0 'zeros', 1 'border', 2 'reflection'; x, y [B,P,H,W].  bwd: dx = transposed gather of
 *              dy (dx is cleared inside the call).
 * noise_clamp: Gaussian (augment/__init__.py:40-49): y = clamp(x + noise * sigma, 0, 1) on n elements, `noise` drawn by
 *              the caller (torch.randn_like keeps the reference's random stream); bwd: dx = dy where the sum is inside
 *              [0, 1]. */
int cb200_shift_flip_fwd(const float* x, float* y, const float* params, int B, int P, int H, int W, int padding_mode,
                         void* stream);
int cb200_shift_flip_bwd(const float* dy, float* dx, const float* params, int B, int P, int H, int W, int padding_mode,
                         void* stream);
int cb200_noise_clamp_fwd(const float* x, const float* noise, float* y, float sigma, long long n, void* stream);
int cb200_noise_clamp_bwd(const float* x, const float* noise, const float* dy, float* dx, float sigma, long long n,
                          void* stream);

/* DiffAugment (third_party/diffaug.py:8-76; augment/__init__.py:136-145, `--aug diffaug` = policy 'color,cutout'):
 * x -> 2x-1 -> [color: brightness, saturation, contrast] -> [translation: integer shift, zero fill] -> [cutout: half-size
 * square zeroed] -> 0.5x+0.5, each stage per sample.  flags = 1 (color) | 2 (translation) | 4 (cutout), applied in that
 * order.  params [7,B] = {r_brightness, r_saturation, r_contrast (raw U[0,1) draws), shift along H, shift along W, cutout
 * offset along H, along W (integers stored as floats)}; x, y, dy, dx [B,3,H,W]; sums / gsums [B] caller-allocated scratch. */
int cb200_diffaug_fwd(const float* x, float* y, const float* params, float* sums, int B, int H, int W, int flags,
                      void* stream);
int cb200_diffaug_bwd(const float* dy, float* dx, const float* params, float* gsums, int B, int H, int W, int flags,
                      void* stream);

/* ---- tcgen05 tensor-core GEMM / implicit-GEMM convolutions (TF32 in, FP32 accumulate) --------
 * Replace F.linear / nn.Conv2d / nn.ConvTranspose2d behind models/gan/sndcgan.py:24-38,91-109 and
 * models/gan/base.py:14-35,92-101 (cuBLAS / cuDNN in the reference).  Activations are NHWC.
 *
 * gemm_nt:   out[M,N] = lrelu_slope(A[M,K] * Bw[N,K]^T + bias)      (slope 1 = no activation);
 *            with dact != NULL: out = (A * Bw^T + bias) * lrelu'(dact[M,N])  (row strides lda/ldb/ldo)
 * conv fwd:  y[B,Ho,Wo,Cout] = lrelu_slope(conv(x[B,H,W,Cin]) + bias); wmat = [Cout, ks*ks*Cin],
 *            column (kh*ks+kw)*Cin+ci = W[co,ci,kh,kw]; (ks,stride) in {(3,1),(4,2)}, pad 1
 * conv dgrad: dx[B,H,W,Cin] = conv^T(dy[B,Ho,Wo,Cout]) (* lrelu'(act_in) when act_in != NULL,
 *            + bias_out then lrelu_slope otherwise); wmat_t layouts are produced by
 *            cb200_sn_pack_weights.  Also serves ConvTranspose2d forward in G_SNDCGAN.
 * round_out: round outputs to TF32 (nearest) because they feed another tensor-core GEMM.
 * colsum   : optional [N] / [Cin] buffer (zeroed inside) that receives the column sums of the STORED outputs.  In the
 *            backward pass the output of a data-gradient GEMM is dL/d(pre-activation) of the previous layer, whose
 *            column sum is that layer's bias gradient (autograd of `nn.Conv2d` / `nn.Linear` bias): fusing it into
 *            the epilogue saves one full read of the gradient tensor per layer. */
int cb200_gemm_nt_tf32(const float* a, long long lda, const float* bw, long long ldb, const float* bias,
                       const float* dact, float* out, long long ldo, int M, int N, int K, float slope,
                       int round_out, float* colsum, void* stream);
int cb200_conv2d_nhwc_fwd(const float* x, const float* wmat, const float* bias, float* y, int B, int H,
                          int W, int Cin, int Cout, int ks, int stride, float slope, int round_out,
                          void* stream);
int cb200_conv2d_nhwc_dgrad(const float* dy, const float* wmat_t, const float* act_in,
                            const float* bias_out, float* dx, int B, int H, int W, int Cin, int Cout,
                            int ks, int stride, float slope, int round_out, float* colsum, void* stream);

/* Split-K for small tile lists (the deep layers and the heads at a small per-GPU batch: DDP b512 over 8 GPUs leaves 64
 * images per rank, train_gan.py:247): the three calls above split the reduction over up to 32 CTAs per output tile when
 * the tile list cannot fill the GPU; partials are parked in a caller-provided scratch and the last CTA to arrive at a
 * tile adds them in split order and runs the fused epilogue (deterministic, one launch).  cb200_tapgemm_workspace hands
 * that scratch to the NEXT of those calls made by this host thread (one-shot, thread-local): `ws` = `bytes` bytes of
 * device scratch, `counters` = `n_counters` int32 that are zero and are left zero (one persistent buffer per device and
 * stream).  Without it the calls run unsplit. */
int cb200_tapgemm_workspace(void* ws, long long bytes, int* counters, int n_counters);

/* Error-compensated TF32 ("3xTF32") operands for the strict-parity mode: x = hi + lo, hi = rn_tf32(x),
 * lo = rn_tf32(x - hi).  (a_hi + a_lo)(b_hi + b_lo) ~ a_hi b_hi + a_lo b_hi + a_hi b_lo is an ordinary GEMM / convolution
 * over a concatenated reduction axis, so the unchanged tcgen05 kernels above evaluate it on operands written as
 * mode 0: out[rows, 3C] = [hi | lo | hi]; mode 1: out[rows, 2C] = [hi | hi]; mode 2: out[2, rows, C] = hi ; lo.
 * Where the reference computes in fp32 (cuBLAS SGEMM for nn.Linear, models/gan/base.py:14-35) or TF32 (cuDNN default
 * for nn.Conv2d) the product's default is single-pass TF32; this mode is what holds the generator's gradient norm to
 * 1e-3 of the fp32 CPU reference at initialisation (tools/tf32_sensitivity.py). */
int cb200_split_tf32(const float* x, float* out, long long rows, int C, int mode, void* stream);

/* wgrad: dw_hat[Cout, ks*ks*Cin] (forward-pack layout) = sum over pixels dy (x) shifted x; both operands are
 * MN-major tcgen05 tiles, split-K over pixels with fp32 atomics (buffer is zeroed inside).
 * gemm_tn_wgrad: dw[N,K] = dy[M,N]^T * x[M,K] (linear layers).  Replace autograd's cuDNN/cuBLAS wgrad. */
int cb200_conv2d_nhwc_wgrad(const float* x, const float* dy, float* dw_hat, int B, int H, int W, int Cin,
                            int Cout, int ks, int stride, void* stream);
int cb200_gemm_tn_wgrad(const float* dy, long long ldy, const float* x, long long ldx, float* dw,
                        long long ldw, int M, int N, int K, void* stream);

/* ---- spectral norm + weight packing ---------------------------------------------------------
 * Replaces torch.nn.utils.spectral_norm's pre-forward hook (call sites models/gan/sndcgan.py:111-118;
 * arithmetic torch/nn/utils/spectral_norm.py:92-114): one power iteration in train mode (u, v updated
 * in place), sigma = u^T W v kept on the device as sigma[2] = {sigma, 1/sigma}.
 * sn_pack_weights writes W/sigma in the GEMM layouts the tensor-core kernels consume
 *   fwd   [Cout][KH][KW][Cin] (row stride ld_fwd);
 *   dgrad mode 1: [Cin][KH][KW][Cout]; mode 2: [ph][pw][Cin][jh][jw][Cout] (4x4 stride 2);
 *         mode 3: [KH][KW][Cin] rows x ldt columns at column offset col0 (transposed linear weight).
 * sn_weight_bwd maps dW_hat (forward-pack layout) back to dW (OIHW) including the sigma term:
 *   dW = (dW_hat - <dW_hat, W_hat> u v^T) / sigma        (autograd of W / (u^T W v), u, v constant). */
int cb200_sn_power_iter(const float* w, float* u, float* v, float* sigma, float* t_scratch,
                        float* s_scratch, int Cout, int F, float eps, int training, void* stream);
int cb200_sn_pack_weights(const float* w, const float* sigma, float* fwd, long long ld_fwd, float* dgrad,
                          int dgrad_mode, long long ldt, int col0, int Cout, int Cin, int KH, int KW,
                          int round_out, void* stream);
int cb200_sn_weight_bwd(const float* dw_hat_packed, long long ld_fwd, const float* w, const float* u,
                        const float* v, const float* sigma, float* acc_scratch, float* dw, int accumulate,
                        int Cout, int Cin, int KH, int KW, void* stream);

/* Batched forms: all spectrally-normalised layers of D in one launch per phase (descriptor arrays live in
 * HOST memory; n <= 16).  power_iter_batched expects every layer's `t` scratch zeroed by the caller,
 * weight_bwd_batched every job's `acc` scalar zeroed. */
struct cb200_sn_layer {
    const float* w; float* u; float* v; float* sigma; float* t; float* s;
    int cout, f;
};
struct cb200_sn_pack_job {
    const float* w; const float* sigma; float* fwd; float* dgrad; long long ld_fwd, ldt;
    int cout, cin, kh, kw, dgrad_mode, col0, round_out;
};
struct cb200_sn_bwd_job {
    const float* dw_hat_packed; const float* w; const float* u; const float* v; const float* sigma; float* acc; float* dw;
    long long ld_fwd; int cout, cin, kh, kw;
};
int cb200_sn_power_iter_batched(const struct cb200_sn_layer* layers, int n, float eps, int training, void* stream);
int cb200_sn_pack_batched(const struct cb200_sn_pack_job* jobs, int n, void* stream);
int cb200_sn_weight_bwd_batched(const struct cb200_sn_bwd_job* jobs, int n, void* stream);

/* ---- first discriminator layer: Conv2d(3->64,3,1,1) + bias + LeakyReLU with x*2-1 folded in ---
 * (models/gan/sndcgan.py:91-93,122-124).  NCHW image in, NHWC activation out (SIMT, HBM-bound).
 * wgrad accumulates into dw_hat[64,27] (OIHW order) and db[64]; dgrad_finish extracts the 3 real
 * channels of the 32-channel padded tensor-core data gradient, applies the factor 2, NHWC->NCHW. */
int cb200_conv_first_fwd(const float* x, const float* w, const float* sigma, const float* bias, float* y,
                         int B, int H, int W, float slope, int round_out, float in_scale, float in_shift,
                         void* stream);
int cb200_conv_first_wgrad(const float* x, const float* dy, float* dw_hat, float* db, int B, int H, int W,
                           float in_scale, float in_shift, void* stream);
int cb200_conv_first_dgrad_finish(const float* dpad, float* dx, int B, int H, int W, int cpad, void* stream);

/* ---- contrastive + GAN losses (fp32 SIMT, flash-style: no R x R matrix is materialised) --------
 * rownorm      = F.normalize(x, dim=1, eps)                       (training/gan/contrad.py:43,48)
 * contrastive  mode 0 = nt_xent(out1,out2) on z=[out1;out2] (training/criterion.py:24-45)
 *              mode 1 = supcon_fake(out1,out2,others) on z=[out1;out2;others] (training/gan/contrad.py:8-32)
 *              z rows are L2-normalised, width 128; diagonal sentinel -5e4 after the 1/temperature scale.
 *              fwd writes row log-sum-exps (needed by bwd) and loss[0]; bwd writes dz = gscale[0]*dL/dz.
 * gan_d_loss   kind 0 nonsat, 1 hinge, 2 wgan, 3 lsgan (contrad.py:52-64): out3 = {L_dis, mean d_real,
 *              mean d_gen}, g_real/g_gen = dL_dis/dd.   gan_g_loss (contrad.py:75-80): out1 = {L_gen}.
 * colsum       out[n] = sum_m x[m,n]  (bias gradients). */
int cb200_rownorm_fwd(const float* x, long long ldx, float* y, float* inv_norm, int rows, int d, float eps,
                      void* stream);
int cb200_rownorm_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, long long lddx,
                      int rows, int d, int round_out, void* stream);
int cb200_contrastive_fwd(const float* z, int N, int d, int mode, float temperature, float* lse,
                          float* scratch /* 48 floats per loss row */, float* loss, void* stream);
int cb200_contrastive_bwd(const float* z, int N, int d, int mode, float temperature, const float* lse,
                          const float* gscale, float* dz, void* stream);

/* Tensor-core formulation of the same two losses (north_star: "the NT-Xent pairwise-cosine similarity matrix uses
 * tcgen05 MMA fed by TMA; the temperature-softmax / CE are warp-shuffle reductions"): the caller forms S = Z_A Z^T with
 * cb200_gemm_nt_tf32 on error-compensated operands (cb200_split_tf32: the logits are S / tau with tau = 0.1, so
 * single-pass TF32 would be amplified 10x) and these two kernels do the row reductions of training/criterion.py:35-45 /
 * training/gan/contrad.py:13-32 on it: S [Ra, lds] = dot products of the loss rows (global row indices row0 .. row0+Ra-1)
 * against all R rows, mode 0 NT-Xent (R = 2N), mode 1 supcon-fake (R = 3N, row0 = 2N).  fwd: lse [Ra] and rowloss [Ra]
 * (loss = sum).  bwd: G [Ra, ldg] = gscale[0] * dLoss/dS (columns >= R and the diagonal are zero); the embedding
 * gradient is then two more GEMMs, dZ_A += G Z and dZ += G^T Z_A. */
int cb200_sim_rows_fwd(const float* S, long long lds, int Ra, int R, int N, int mode, int row0, float temperature,
                       float* lse, float* rowloss, void* stream);
int cb200_sim_rows_bwd(const float* S, long long lds, int Ra, int R, int N, int mode, int row0, float temperature,
                       const float* lse, const float* gscale, float* G, long long ldg, int cols, void* stream);
int cb200_gan_d_loss(const float* d_real, const float* d_gen, long long stride, int N, int kind, float* out3,
                     float* g_real, float* g_gen, void* stream);
int cb200_gan_g_loss(const float* d_gen, long long stride, int N, int kind, float* out1, float* g_gen,
                     void* stream);
int cb200_colsum(const float* x, long long ld, int M, int N, float* out, void* stream);
/* out = dy * lrelu'(act) from the saved output activation (nn.LeakyReLU backward, sndcgan.py:92-108). */
int cb200_lrelu_bwd(const float* dy, const float* act, float* out, long long n, float slope, int round_out,
                    void* stream);

/* ---- generator-side kernels (G_SNDCGAN, models/gan/sndcgan.py:24-48) ------------------------------
 * Train-mode BatchNorm2d(+ReLU) on NHWC activations x[M,C] (M = batch*H*W), split so the host can
 * all-reduce the partial sums (SyncBatchNorm under DDP, train_gan.py:268):
 *   bn_stats      sums[2,C] = {sum x, sum x^2}
 *   bn_finalize   stats[2,C] = {mean, rstd}; running_mean/var updated (momentum, unbiased variance)
 *   bn_apply_relu y = relu(gamma*(x-mean)*rstd + beta); remap_s > 0: x is [M, C] with feature index
 *                 c*S+s ((c,h,w) flattening of norm_init, sndcgan.py:42-45) and y is NHWC [M, S, C/S]
 *   bn_bwd_reduce sums[2,C] = {sum dz, sum dz*xhat}, dz = dy*1[y>0]
 *   bn_bwd_apply  dx = gamma*rstd*(dz - sum_dz/count - xhat*sum_dz_xhat/count)
 * g_final_fwd: out[B,3,H,W] = 0.5*tanh(pre[B,H,W,cpad][..., :3] + bias) + 0.5 (nn.Tanh + `0.5*y+0.5`);
 * g_final_bwd: its backward from the saved output (+ bias gradient).  round_tf32: y = rna_tf32(x). */
int cb200_bn_stats(const float* x, int M, int C, float* sums, void* stream);
int cb200_bn_finalize(const float* sums, float count, int C, float eps, float momentum, float* stats,
                      float* running_mean, float* running_var, void* stream);
int cb200_bn_apply_relu(const float* x, const float* stats, const float* gamma, const float* beta, float* y,
                        int M, int C, int remap_s, int round_out, void* stream);
int cb200_bn_bwd_reduce(const float* dy, const float* y, const float* x, const float* stats, int M, int C,
                        int remap_s, float* sums, void* stream);
int cb200_bn_bwd_apply(const float* dy, const float* y, const float* x, const float* stats, const float* gamma,
                       const float* sums, float count, float* dx, int M, int C, int remap_s, int round_out,
                       void* stream);
int cb200_g_final_fwd(const float* pre, const float* bias, float* out, int B, int H, int W, int cpad,
                      void* stream);
int cb200_g_final_bwd(const float* dout, const float* out, float* dpre, float* dbias, int B, int H, int W,
                      void* stream);
int cb200_round_tf32(const float* x, float* y, long long n, void* stream);

/* ---- fused multi-tensor Adam (torch.optim.Adam of train_gan.py:273-274: no weight decay / amsgrad) ----
 * One launch per 48 tensors; `step` is the 1-based step count used for the bias corrections. */
struct cb200_adam_tensor {
    float* p; const float* g; float* m; float* v; long long numel;
};
int cb200_adam_step(const struct cb200_adam_tensor* tensors, int n, float lr, float beta1, float beta2, float eps,
                    int step, void* stream);
/* Same update; {lr, 1 - beta1^t, sqrt(1 - beta2^t)} are read from 3 floats in DEVICE memory (CUDA-graph replay). */
int cb200_adam_step_dev(const struct cb200_adam_tensor* tensors, int n, const float* hyper, float beta1, float beta2,
                        float eps, void* stream);

/* ---- StyleGAN2 side of the path (models/gan/stylegan2/*) ---------------------------
 * All activations NHWC fp32 unless strides are given.  The dense contractions of ResidualDiscriminatorP and
 * Generator run on the tensor-core entry points above (3x3 stride-1 convolutions: cb200_conv2d_nhwc_*; every
 * 1x1 convolution, linear layer, the blurred 3x3 stride-2 convolution and the stride-2 transposed convolution:
 * cb200_gemm_nt_tf32 / cb200_gemm_tn_wgrad on the patch matrices written by cb200_patch_s2_*).
 *
 * upfirdn2d      replaces op/upfirdn2d.py:145-200 + upfirdn2d_kernel.cu (`upfirdn2d(input, kernel, up, down, pad)`):
 *                out = decimate_down(FIR(pad(zero_stuff_up(x)))); strides are {n, c, h, w} in elements, so the same
 *                kernel serves the reference's NCHW op and the NHWC activations; negative pads crop; flip = 1 uses the
 *                kernel as given (the backward pass, upfirdn2d.py:113), flip = 0 flips it like the reference forward;
 *                c_fast = 1 makes the channel the fastest thread index (NHWC outputs).
 * patch_s2       u[B,Ho,Wo,9,C] <-> x[B,2Ho+1,2Wo+1,C]: gather = im2col of `F.conv2d(stride=2, padding=0)` with a 3x3
 *                kernel (layers.py:174-198 after Blur), scatter = its transpose = `F.conv_transpose2d(stride=2)`
 *                (generator.py:66-72).
 * bias_act       mode 0: y = lrelu_slope(x + bias[c]) * gain (+ res)  (op/fused_act.py:86-94, `(out+skip)/sqrt2` of
 *                discriminator.py:70-76 folded in); mode 1: y = x * ((ref + bias[c]) > 0 ? gain : gain*slope), the
 *                backward AND the backward-of-backward of mode 0 (op/fused_act.py:18-52).
 * modulate       y[b,p,c] = x[b,p,c] * s[b,c] * alpha (generator.py:55-56 moved from the weights to the activations;
 *                x_batch_stride 0 broadcasts ConstantInput); mul_reduce: out[b,c] = sum_p a[b,p,c] * w[b,p,c].
 * mod_epilogue   y = lrelu(x * demod[b,c] + noise[b,p] * noise_weight[0] + bias[c]) * gain
 *                (demodulation generator.py:58-60, NoiseInjection :85-94, FusedLeakyReLU); noise_grad: its weight grad.
 * stddev_*       _minibatch_stddev_layer (discriminator.py:22-33): fwd -> std[B/G]; concat appends it as channel C of a
 *                Cp-channel tensor (zero padded so that Cp % 32 == 0); split / bwd / bwd_bwd are the backward passes
 *                (bwd_bwd: the R1 double backward).
 * rgb_to_nhwc    y[b,h,w,c<3] = x[b,c,h,w]*scale + shift, other channels 0 (`input * 2. - 1.` discriminator.py:229);
 *                nhwc_to_rgb: out[b,c,h,w] = src[b,h,w,c]*scale (+ res), the ToRGB skip sum (generator.py:133-143).
 * pixelnorm      layers.py:15-20.  row_sqsum / row_scale: `grad.pow(2).reshape(B,-1).sum(1)` of r1_loss
 *                (train_stylegan2.py:106-113) and its backward.  axpby: out = alpha*a + beta*b + gamma.
 * ema_lerp       utils.py:130-143 `accumulate`: dst = decay*dst + (1-decay)*src for a table of tensors (host memory). */
int cb200_upfirdn2d(const float* x, const long long* x_strides, float* y, const long long* y_strides, const float* fir,
                    int N, int C, int Hi, int Wi, int Ho, int Wo, int up, int down, int pad_x0, int pad_y0, int kh, int kw,
                    int flip, float gain, int c_fast, int round_out, void* stream);
int cb200_patch_s2_gather(const float* x, float* u, int B, int Ho, int Wo, int C, int round_out, void* stream);
int cb200_patch_s2_scatter(const float* u, float* x, int B, int Ho, int Wo, int C, int round_out, void* stream);
int cb200_bias_act(const float* x, const float* bias, const float* ref, const float* res, float* y, long long n, int C,
                   int mode, float slope, float gain, int round_out, void* stream);
int cb200_modulate(const float* x, long long x_batch_stride, const float* s, float* y, int B, long long P, int C, float alpha,
                   int round_out, void* stream);
int cb200_mul_reduce(const float* a, const float* w, long long w_batch_stride, float* out, int B, long long P, int C,
                     void* stream);
int cb200_mod_epilogue(const float* x, const float* demod, const float* noise, const float* noise_weight, const float* bias,
                       float* y, int B, long long P, int C, float slope, float gain, int round_out, void* stream);
int cb200_noise_grad(const float* g, const float* noise, float* out1, long long rows, int C, void* stream);
int cb200_stddev_fwd(const float* x, float* std, int B, long long F, void* stream);
int cb200_stddev_bwd(const float* dstd, const float* x, float* dx, int B, long long F, void* stream);
int cb200_stddev_bwd_bwd(const float* gg, const float* dstd, const float* x, float* d_dstd, float* d_x, int B, long long F,
                         void* stream);
int cb200_stddev_concat(const float* x, const float* std, float* y, int B, long long P, int C, int Cp, int round_out,
                        void* stream);
int cb200_stddev_split(const float* dy, float* dx, float* dstd, int B, long long P, int C, int Cp, void* stream);
int cb200_rgb_to_nhwc(const float* x, float* y, int B, int H, int W, int cpad, float scale, float shift, int round_out,
                      void* stream);
int cb200_nhwc_to_rgb(const float* src, const float* res, float* out, int B, int H, int W, int cpad, float scale,
                      void* stream);
int cb200_pixelnorm(const float* x, float* y, int rows, int d, int round_out, void* stream);
int cb200_row_sqsum(const float* x, float* out, int B, long long n, void* stream);
int cb200_row_scale(const float* x, const float* s, float* y, int B, long long n, float alpha, void* stream);
int cb200_axpby(const float* a, const float* b, float* out, long long n, float alpha, float beta, float gamma, int round_out,
                void* stream);
/* y[rows, Cp] = [x[rows, C], 0]: zero-extension of the channel dimension (operands of the weight-gradient kernels). */
int cb200_pad_channels(const float* x, float* y, long long rows, int C, int Cp, void* stream);
struct cb200_ema_tensor {
    float* dst; const float* src; long long numel;
};
int cb200_ema_lerp(const struct cb200_ema_tensor* tensors, int n, float decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CONTRAD_B200_H_ */
