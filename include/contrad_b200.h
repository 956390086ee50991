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
 * Backward = autograd of the reference chain (HSV straight-through, color_jitter.py:97-104). */
int cb200_augment_simclr_fwd(const float* x, float* y, const float* params, int B, int H, int W,
                             int order, void* stream);
int cb200_augment_simclr_bwd(const float* x, const float* dy, float* dx, const float* params,
                             int B, int H, int W, int order, void* stream);

/* ---- tcgen05 tensor-core GEMM / implicit-GEMM convolutions (TF32 in, FP32 accumulate) --------
 * Replace F.linear / nn.Conv2d / nn.ConvTranspose2d behind models/gan/sndcgan.py:24-38,91-109 and
 * models/gan/base.py:14-35,92-101 (cuBLAS / cuDNN in the reference).  Activations are NHWC.
 *
 * gemm_nt:   out[M,N] = lrelu_slope(A[M,K] * Bw[N,K]^T + bias)      (slope 1 = no activation)
 * conv fwd:  y[B,Ho,Wo,Cout] = lrelu_slope(conv(x[B,H,W,Cin]) + bias); wmat = [Cout, ks*ks*Cin],
 *            column (kh*ks+kw)*Cin+ci = W[co,ci,kh,kw]; (ks,stride) in {(3,1),(4,2)}, pad 1
 * conv dgrad: dx[B,H,W,Cin] = conv^T(dy[B,Ho,Wo,Cout]) (* lrelu'(act_in) when act_in != NULL,
 *            + bias_out then lrelu_slope otherwise); wmat_t layouts are produced by
 *            cb200_sn_pack_weights.  Also serves ConvTranspose2d forward in G_SNDCGAN.
 * round_out: round outputs to TF32 (nearest) because they feed another tensor-core GEMM. */
int cb200_gemm_nt_tf32(const float* a, long long lda, const float* bw, const float* bias, float* out,
                       long long ldo, int M, int N, int K, float slope, int round_out, void* stream);
int cb200_conv2d_nhwc_fwd(const float* x, const float* wmat, const float* bias, float* y, int B, int H,
                          int W, int Cin, int Cout, int ks, int stride, float slope, int round_out,
                          void* stream);
int cb200_conv2d_nhwc_dgrad(const float* dy, const float* wmat_t, const float* act_in,
                            const float* bias_out, float* dx, int B, int H, int W, int Cin, int Cout,
                            int ks, int stride, float slope, int round_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CONTRAD_B200_H_ */
