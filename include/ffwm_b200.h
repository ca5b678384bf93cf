/*
 * ffwm_b200 — C ABI of the B200-native FFWM flow-warping hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces
 * one function a reference pybind module exports; the Python shims in
 * ffwm_b200/dropin/ re-export them under the reference's module names
 * (resample2d_cuda, block_extractor_cuda, local_attn_reshape_cuda) with the
 * reference's signatures, so models/external_function.py runs unmodified.
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions (identical for every call):
 *   - tensors are 4-D (N,C,H,W), described by ffwm_tensor4: raw DEVICE
 *     pointer, sizes and ELEMENT strides (the reference kernels take long4
 *     size/stride pairs and honour arbitrary strides; so do these);
 *   - dtype is FFWM_F32 or FFWM_F64 (AT_DISPATCH_FLOATING_TYPES);
 *   - the caller owns and allocates everything, the callee writes in place
 *     and never allocates, frees or synchronises;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL is
 *     the legacy default stream) of the CURRENT device;
 *   - return 0 on success, FFWM_ERR_* (<0) for a rejected argument, or a
 *     positive cudaError_t from the launch; ffwm_last_error() describes the
 *     last failure of the calling thread.  (The reference returns 1 and
 *     never checks; the Python shims translate 0 -> 1 / raise RuntimeError.)
 *
 * Output initialisation: the reference's callers zero-fill every output and
 * gradient buffer and the reference kernels atomicAdd into them.  Here only
 * the true scatter targets ACCUMULATE and therefore must be zero-filled by
 * the caller, exactly as the reference requires:
 *       grad_input1 (resample2d), grad_source (block_extractor),
 *       grad_images (grid_warp).
 * Every other output is OVERWRITTEN (each element is produced exactly once,
 * deterministically), which is indistinguishable from accumulate-into-zero.
 */
#ifndef FFWM_B200_H_
#define FFWM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFWM_ABI_VERSION 2

enum { FFWM_F32 = 0, FFWM_F64 = 1 };

enum {
    FFWM_OK = 0,
    FFWM_ERR_NULL = -1,      /* null tensor / data pointer with non-zero numel   */
    FFWM_ERR_SHAPE = -2,     /* inconsistent sizes                               */
    FFWM_ERR_ARG = -3,       /* bad kernel_size / dilation / dtype               */
    FFWM_ERR_TOO_LARGE = -4  /* a single (H,W) plane spans more than 2^31 elems  */
};

typedef struct ffwm_tensor4 {
    void* data;          /* device pointer */
    int64_t size[4];     /* N, C, H, W */
    int64_t stride[4];   /* in elements */
} ffwm_tensor4;

int ffwm_abi_version(void);
const char* ffwm_last_error(void);
/* Number of kernels this process has enqueued through the library (diagnostic counter used by
 * bench.py's `gpu_launches`; the reference has no equivalent). */
unsigned long long ffwm_kernel_launches(void);

/* resample2d_cuda.forward   (cuda/resample2d_package/resample2d_cuda.cc:6-15,
 *                            resample2d_kernel.cu:20-95,335-375)
 * input1 (B,C,Hi,Wi); input2 (B,3,H,W) = (dx, dy, sigma); output (B,C,H,W). */
int ffwm_resample2d_forward(const ffwm_tensor4* input1, const ffwm_tensor4* input2,
                            const ffwm_tensor4* output,
                            int kernel_size, int dilation, int dtype, void* stream);

/* resample2d_cuda.backward  (resample2d_cuda.cc:17-28, resample2d_kernel.cu:98-330,377-454)
 * grad_input1 (B,C,Hi,Wi) accumulates (zero-fill it); grad_input2 (B,3,H,W) is overwritten.
 * One fused pass over grad_output instead of the reference's two kernels.    */
int ffwm_resample2d_backward(const ffwm_tensor4* input1, const ffwm_tensor4* input2,
                             const ffwm_tensor4* grad_output,
                             const ffwm_tensor4* grad_input1, const ffwm_tensor4* grad_input2,
                             int kernel_size, int dilation, int dtype, void* stream);

/* block_extractor_cuda.forward  (cuda/block_extractor/block_extractor_cuda.cc:5-12,
 *                                block_extractor_kernel.cu:20-85,172-217)
 * source (B,C,Hs,Ws); flow (B,2,Hf,Wf); output (B,C,k*Hf,k*Wf).              */
int ffwm_block_extractor_forward(const ffwm_tensor4* source, const ffwm_tensor4* flow,
                                 const ffwm_tensor4* output,
                                 int kernel_size, int dtype, void* stream);

/* block_extractor_cuda.backward (block_extractor_cuda.cc:14-27, block_extractor_kernel.cu:89-170,222-278)
 * grad_source accumulates (zero-fill it); grad_flow is overwritten (no atomics).  */
int ffwm_block_extractor_backward(const ffwm_tensor4* source, const ffwm_tensor4* flow,
                                  const ffwm_tensor4* grad_output,
                                  const ffwm_tensor4* grad_source, const ffwm_tensor4* grad_flow,
                                  int kernel_size, int dtype, void* stream);

/* local_attn_reshape_cuda.forward  (cuda/local_attn_reshape/local_attn_reshape_cuda.cc:5-11,
 *                                   local_attn_reshape_kernel.cu:20-61,110-148)
 * inputs (B,k*k,H,W) -> output (B,1,k*H,k*W).                                 */
int ffwm_local_attn_reshape_forward(const ffwm_tensor4* inputs, const ffwm_tensor4* output,
                                    int kernel_size, int dtype, void* stream);

/* local_attn_reshape_cuda.backward (local_attn_reshape_cuda.cc:13-23, local_attn_reshape_kernel.cu:65-108,153-195)
 * grad_output (B,Cg,k*H,k*W) -> grad_inputs (B,k*k,H,W), overwritten with the
 * sum over Cg (Cg is 1 for every caller in the reference).                    */
int ffwm_local_attn_reshape_backward(const ffwm_tensor4* grad_output, const ffwm_tensor4* grad_inputs,
                                     int kernel_size, int dtype, void* stream);

/* WarpNet.forward (models/base_networks.py:168-173) and
 * PerceptualCorrectness.bilinear_warp (models/losses.py:392-396):
 *   F.grid_sample(images, flow.permute(0,2,3,1)), bilinear, zeros padding,
 *   align_corners=False.  flow stays in the network's (B,2,H,W) layout:
 *   channel 0 = x, channel 1 = y, both in [-1,1]; no permuted copy is made.
 * images (B,C,Hi,Wi); flow (B,2,H,W); output (B,C,H,W).                       */
int ffwm_grid_warp_forward(const ffwm_tensor4* images, const ffwm_tensor4* flow,
                           const ffwm_tensor4* output, int dtype, void* stream);

/* grad_images accumulates (zero-fill it); grad_flow (B,2,H,W) is overwritten.
 * grad_images->data or grad_flow->data may be NULL to skip that gradient.     */
int ffwm_grid_warp_backward(const ffwm_tensor4* images, const ffwm_tensor4* flow,
                            const ffwm_tensor4* grad_output,
                            const ffwm_tensor4* grad_images, const ffwm_tensor4* grad_flow,
                            int dtype, void* stream);

/* ---- runtime options (A/B switches and test hooks; no reference counterpart) -------------------------------
 * Each option NAME ("DISABLE_TILED", "FORCE_TILED", "CONV_MATH", ...: the list is in csrc/common.cuh) is read once
 * from the environment variable FFWM_<NAME> when the library is first used; launch paths read a cached int.
 * ffwm_set_option changes it afterwards (process-wide), ffwm_get_option returns it (-1: unknown name). */
int ffwm_set_option(const char* name, int value);
int ffwm_get_option(const char* name);

/* ---- 3x3 / stride 1 / pad 1 convolution on the tcgen05 tensor cores (csrc/conv3x3_tc.cu; fp32 in/out) -----------------
 * Replaces the cuDNN call behind nn.Conv2d(Cin, Cout, 3, 1, 1).forward (models/base_networks.py:218-222,235-246: the
 * generator's ResidualBlock / ConvBlock convolutions, and every other 3x3 stride-1 layer of the path) for maps of width
 * 128, 64, 32 or 16, and — with weights packed with dgrad=1 — its data gradient.  Weights are packed once per update.
 *   nt   = output channels per CTA (the MMA N) the image is laid out for: 64, or 128 (W = 128 only);
 *   math = operand split: FFWM_MATH_TF32X3 (a = hi + lo in tf32, three MMAs: library-grade fp32 accuracy, used for
 *          forward passes) or FFWM_MATH_BF16X3 (two bf16 parts, three MMAs at twice the rate: used for gradients, which
 *          are linear in the operand error).  pack and forward must agree on nt and math. */
enum { FFWM_MATH_TF32X3 = 0, FFWM_MATH_BF16X3 = 1 };
int64_t ffwm_conv3x3_packed_floats(int cout, int cin, int nt, int math);
/* weight (Cout,Cin,3,3) fp32 (any strides) -> packed; dgrad=1 packs the weights of the data-gradient convolution
 * (channels swapped, taps flipped; needs ffwm_conv3x3_packed_floats(cin, cout, nt, math) floats). */
int ffwm_conv3x3_pack_weights(const ffwm_tensor4* weight, int dgrad, float* packed, int64_t packed_floats, int nt, int math, void* stream);
/* out (B,Cout,H,W) = conv2d(x (B,Cin,H,W), weight, bias, stride 1, padding 1), W in {128,64,32,16}; bias may be NULL. */
int ffwm_conv3x3_forward(const ffwm_tensor4* x, const float* packed, const float* bias, const ffwm_tensor4* out, int nt, int math, void* stream);

/* ---- general dense convolution on the tcgen05 tensor cores (csrc/conv_gen_tc.cu; fp32 in/out; math as above) ----------
 * Replaces the cuDNN calls behind nn.Conv2d / nn.ConvTranspose2d for every other shape of the path: kernels up to 7x7,
 * stride 1 or 2, any padding and map size (models/base_networks.py:30-57 conv / deconv / predict_flow, :208-246, :274-312
 * generator stem, strided encoders, 1x1 residual inputs, :354-437 discriminator; lightcnn/light_cnn.py:13-26 5x5 / 1x1
 * MFM layers; models/losses.py:398-519 VGG19 blocks 4-5), and the data gradient of each (aten::convolution_backward's
 * grad_input).  groups = 1, dilation = 1, zero padding. */
int64_t ffwm_conv_packed_bytes(int n_out, int n_in, int kh, int kw, int math);
/* in_major = 0: weight[n_out][n_in][kh][kw] (Conv2d forward, ConvTranspose2d data gradient);
 * in_major = 1: weight[n_in][n_out][kh][kw] (ConvTranspose2d forward, Conv2d data gradient).  Any strides. */
int ffwm_conv_pack_weights(const ffwm_tensor4* weight, int in_major, int stride, int pad, int transposed, int math, void* packed,
                           int64_t packed_bytes, void* stream);
/* transposed = 0: out = conv2d(x, W, bias, stride, pad);  transposed = 1: out = conv_transpose2d(x, W, bias, stride, pad)
 * with out's H, W in [(Hi-1)*stride - 2*pad + kh, +stride) (output_padding; as a data gradient: the forward input size).
 * bias (Cout floats) may be NULL.  Layers too small to fill the GPU split K; their partial sums meet as fp32 REDs. */
int ffwm_conv_forward(const ffwm_tensor4* x, const void* packed, const float* bias, const ffwm_tensor4* out, int kh, int kw,
                      int stride, int pad, int transposed, int math, void* stream);

/* Weight gradient of EVERY convolution of the path, the 3x3 stride-1 layers included (csrc/conv_gen_wgrad_tc.cu;
 * aten::convolution_backward's grad_weight; 3xBF16 split, <= 1.5e-5 of max|dW| where cuDNN's fp32 engines measure 3-7e-5):
 *   grad_weight[a][b][ky][kx] = sum_{n,y,x} small[n,a,y,x] * large[n,b,y*stride-pad+ky,x*stride-pad+kx]   (OVERWRITTEN)
 * nn.Conv2d: small = grad_out, large = input; nn.ConvTranspose2d: small = input, large = grad_out; (a, b) are the first
 * two dimensions of the weight either way.  Deterministic (split-K partials in `workspace`, summed in a fixed order). */
int64_t ffwm_conv_wgrad_workspace_bytes(int n, int ca, int cb, int hs, int ws, int hl, int wl, int kh, int kw, int stride, int pad);
int ffwm_conv_wgrad(const ffwm_tensor4* small, const ffwm_tensor4* large, const ffwm_tensor4* grad_weight, int stride, int pad,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* ---- affine regularisation of FlowNet pre-training as one kernel per direction (csrc/affine_reg.cu) -------
 * Replaces the five-pass chain of models/losses.py:211-219 (conv2d with the fixed (kz^2,1,kz,kz) kernel ->
 * LocalAttnReshape -> BlockExtractor(flow = kz//2) -> multiply -> avg_pool2d):
 *   out[b,0,y,x] = w^T Q w / kz^2,  w = the kz x kz window of grid (B,1,H,W) at (y,x),  out (B,1,H-kz+1,W-kz+1);
 * Q (kz^2 x kz^2, row-major, device memory, dtype of the call) is the reference's K^T K; kz in {3,5,7}.
 * backward: grad_grid (B,1,H,W) ACCUMULATES (zero-fill it) the gradient of sum(grad_out * out). */
int ffwm_affine_reg_forward(const ffwm_tensor4* grid, const void* q, const ffwm_tensor4* out, int kz, int dtype, void* stream);
int ffwm_affine_reg_backward(const ffwm_tensor4* grid, const void* q, const ffwm_tensor4* grad_out, const ffwm_tensor4* grad_grid,
                             int kz, int dtype, void* stream);

/* ---- training-mode BatchNorm2d (+ LeakyReLU, + residual add) as two streaming kernels per direction (csrc/batch_norm.cu) ----
 * Replaces nn.BatchNorm2d in training mode and the nn.LeakyReLU that follows it in every conv unit of the reference
 * (models/base_networks.py:15-45, :179-206, :397-410), and `activ(blocks(x) + input(x))` of ResidualBlock (:208-233):
 *   y = lrelu_slope( (x - mean_c) * gamma_c / sqrt(var_c + eps) + beta_c [+ residual] ),  mean / biased var over (N, H, W);
 *   running_mean/var (may be NULL) get torch's momentum update (unbiased variance); save_mean / save_invstd (C floats) feed backward.
 * act_slope = 1: no activation.  x, residual, y, grad_*: contiguous (N, C, H*W) fp32.  Deterministic.
 * backward: grad_x (overwritten), grad_gamma / grad_beta (C floats, overwritten, may be NULL); when the forward added a
 * residual, pass y_out (the forward's output: its sign is the activation's) and grad_residual (overwritten with
 * grad_out * lrelu'(y), the gradient of the residual operand).
 * workspace: ffwm_batch_norm_workspace_bytes(N, C, H*W) bytes of device memory, no initialisation needed. */
int64_t ffwm_batch_norm_workspace_bytes(int n, int c, int64_t hw);
int ffwm_batch_norm_forward(const float* x, const float* residual, const float* gamma, const float* beta, float* running_mean,
                            float* running_var, float momentum, float eps, float act_slope, float* y, float* save_mean,
                            float* save_invstd, int n, int c, int64_t hw, void* workspace, int64_t workspace_bytes, void* stream);
int ffwm_batch_norm_backward(const float* x, const float* grad_out, const float* y_out, const float* gamma, const float* beta,
                             const float* save_mean, const float* save_invstd, float act_slope, float* grad_x, float* grad_residual,
                             float* grad_gamma, float* grad_beta, int n, int c, int64_t hw, void* workspace, int64_t workspace_bytes,
                             void* stream);

/* ---- direct fp32 convolution for degenerate channel counts (csrc/conv_few.cu) -----------------------------------------------
 * out = conv2d(x, W', bias, stride, pad) where W' has at most 4 input channels (LightCNN's 5x5 stem, the 7x7 / 3x3 stems on RGB:
 * lightcnn/light_cnn.py:96-100, models/base_networks.py:59-75,230,397-399, models/losses.py:430) or at most 4 output channels at
 * stride 1 (flow heads, reconstructions, the stems' data gradients: models/base_networks.py:45-57,241); FFWM_ERR_ARG otherwise.
 * W' is `weight` read through (in_major, flip): in_major = 0: W'[o][i] = weight[o][i]; 1: W'[o][i] = weight[i][o]; flip: taps
 * reversed — the data gradient of a stride-1 convolution is ffwm_conv_few(grad_out, weight, 1, 1, NULL, grad_in, 1, k-1-pad).
 * Exact fp32 FFMA accumulation; any strides; kernels up to 7x7; stride 1 or 2 (few inputs only); bias may be NULL. */
int ffwm_conv_few(const ffwm_tensor4* x, const ffwm_tensor4* weight, int in_major, int flip, const float* bias, const ffwm_tensor4* out,
                  int stride, int pad, void* stream);

/* ---- 2x2 / stride-2 max pooling (csrc/pool.cu): nn.MaxPool2d(2, 2[, ceil_mode=True]) / F.max_pool2d(x, 2) of LightCNN and VGG19
 * (lightcnn/light_cnn.py:38-42,96-124, models/losses.py:430-470) without ATen's int64 index map: x (planes, h, w) -> out (planes,
 * ho, wo), contiguous fp32, ho in {floor(h/2), ceil(h/2)}; backward recomputes the argmax from x (first maximum wins, NaN wins:
 * ATen's rule) and OVERWRITES grad_x. */
int ffwm_max_pool2x2_forward(const float* x, float* out, int64_t planes, int h, int w, int ho, int wo, void* stream);
int ffwm_max_pool2x2_backward(const float* x, const float* grad_out, float* grad_x, int64_t planes, int h, int w, int ho, int wo,
                              void* stream);

/* ---- spectral norm of all layers of a network at once (csrc/spectral_norm.cu) --------------------------------------------
 * Replaces torch.nn.utils.spectral_norm's per-layer pre-forward hooks (models/base_networks.py:204-246, :397-410; one power
 * iteration, dim 0): v = normalize(W^T u), u = normalize(W v), sigma = u . (W v), weight = W / sigma — three launches for
 * all layers forward, two backward.  `table` (device memory): `layers` rows of 8 int64 {weight_orig pointer (h x w row-major
 * fp32), u pointer (h), v pointer (w), h, w, offset of the layer in the flat element buffers, in the flat t / v buffers (sum of
 * w), in the flat s / u buffers (sum of h)}, then three arrays of layers + 1 int64 block prefix sums with ceil(w / 32),
 * ceil(h / 8), ceil(h * w / 4096) blocks per layer (blocks1..3 = their totals).
 * forward: out (sum of h*w) = W / sigma per layer; t, v_saved (sum of w), s, u_saved (sum of h) and sigma (layers) are
 * scratch / saved for backward; update = 1 runs the power iteration and overwrites u and v, 0 uses them as stored (eval mode).
 * backward: grad_w (flat, like out) = gradient of sum(grad_out * W / sigma) with u, v constants; partial: blocks3 doubles. */
int ffwm_spectral_norm_forward(const void* table, int layers, int update, float eps, float* out, float* t, float* s, float* u_saved,
                               float* v_saved, float* sigma, int blocks1, int blocks2, int blocks3, void* stream);
int ffwm_spectral_norm_backward(const void* table, int layers, const float* grad_out, const float* u_saved, const float* v_saved,
                                const float* sigma, float* grad_w, void* partial, int blocks3, void* stream);

/* out[c] = sum over (n, hw) of x (N, C, H*W contiguous fp32): the bias gradient of a convolution (grad_out.sum((0,2,3)) in
 * aten::convolution_backward), two kernels, deterministic.  workspace: ffwm_batch_norm_workspace_bytes(N, C, H*W) bytes. */
int ffwm_channel_sum(const float* x, float* out, int n, int c, int64_t hw, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- correlation column-max of PerceptualCorrectness (csrc/corr_max.cu) ------------------------------------
 * Replaces models/losses.py:347-353 (per-pixel cosine normalisation, bmm to a [b, N2, N2] product — 1.07 GB per sample at
 * relu1_1 — and max over the source axis) with a normalising prepass + one tcgen05 GEMM whose epilogue keeps a running
 * column max (3xBF16 split, fp32 accumulation):
 *   cmax[b, j] = max_i < source[b,:,i] / (|source[b,:,i]| + eps) , target[b,:,j] / (|target[b,:,j]| + eps) >
 * source, target (B, C, H, W) fp32, contiguous planes, C a multiple of 64 up to 256; cmax (B, H*W) floats.
 * workspace: ffwm_corr_max_workspace_bytes(B, C, H*W) bytes of device memory (0 = unsupported shape). */
int64_t ffwm_corr_max_workspace_bytes(int b, int c, int n);
int ffwm_corr_max(const ffwm_tensor4* source, const ffwm_tensor4* target, float eps, float* cmax, void* workspace,
                  int64_t workspace_bytes, void* stream);

/* ---- input pipeline, tensor half (csrc/ingest.cu) -----------------------------------------------------------
 * Replaces the per-sample host work of data/face_dataset.py:66-80 (left-right flip `img[:, ::-1, :]`, HWC -> CHW transpose,
 * astype(float32), div(255)) for a whole batch: the batch crosses PCIe as uint8 and is converted on the device.
 * dst (B,C,H,W) float32 = transpose(src (B,H,W,C) uint8, flipped left-right where flip[b] != 0) / 255; flip may be NULL;
 * C in {1,3}; dense tensors.  Bit-identical to the reference's numpy / torch arithmetic. */
int ffwm_ingest_u8(const void* src, const void* flip, float* dst, int b, int h, int w, int c, void* stream);

/* ---- LightCNN max-feature-map activation (lightcnn/light_cnn.py:13-26: `torch.max(out[0], out[1])` over the two
 * channel halves of the preceding conv / linear output) and its gradient, one streaming kernel each; ATen's
 * semantics incl. NaN propagation and tie splitting.  x (n, 2*chw) and out (n, chw) contiguous fp32; chw = C*H*W.
 * Default on (FFWM_FUSED_MFM=0 for the A/B); B200 parity: tests/test_mfm_gpu.py. */
int ffwm_mfm_forward(const float* x, float* out, int64_t n, int64_t chw, void* stream);
int ffwm_mfm_backward(const float* x, const float* grad_out, float* grad_x, int64_t n, int64_t chw, void* stream);

/* ---- Guided filter (models/external_function.py:164-195 BoxFilter, :239-277 GuidedFilter.forward) and its gradient
 * with respect to x, four kernels per direction; box sums are truncated window sums taken directly (separable).
 * x, y, q, grad_q, grad_x: (planes = B*C, H, W) contiguous fp32, H and W > 2r+1 (the reference's assert).
 * save: 5*planes*H*W floats written by forward, read by backward; scratch: 5 (forward) / 6 (backward) * planes*H*W.
 * Default on (FFWM_FUSED_GF=0 for the A/B); B200 parity against autograd of the reference formula: tests/test_guided_filter_gpu.py. */
int ffwm_guided_filter_forward(const float* x, const float* y, float* q, float* save, float* scratch,
                               int64_t planes, int h, int w, int r, float eps, void* stream);
int ffwm_guided_filter_backward(const float* x, const float* y, const float* grad_q, const float* save,
                                float* grad_x, float* scratch, int64_t planes, int h, int w, int r, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FFWM_B200_H_ */
