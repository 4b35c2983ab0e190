/*
 * histogan_b200.h -- C ABI of libhistogan_b200.so (sm_100a CUDA kernels for the
 * HistoGAN training hot path).
 *
 * The reference (mahmoudnafifi/HistoGAN) is pure Python/PyTorch: it has no FFI
 * of its own, so this header DEFINES the boundary.  Every entry point cites the
 * reference code it replaces (paths relative to the reference repo root); the
 * Python classes in histogan_b200/ (same names / constructor kwargs / forward
 * signatures as the reference classes) are thin callers of these functions via
 * ctypes.  See INTEGRATION.md for the reference-side binding.
 *
 * Conventions
 *  - all pointers are DEVICE pointers unless the name ends in _host;
 *    tensors are float32, owned by the caller, never retained;
 *  - every call is asynchronous on `stream` (a cudaStream_t / CUstream);
 *  - return value: 0 on success, otherwise a negative HG_E* code or a positive
 *    cudaError_t; hg_last_error() returns a thread-local message;
 *  - `ws` is a caller-provided scratch buffer of at least the size reported by
 *    the matching *_workspace_bytes() call (16-byte aligned);
 *  - no hidden allocations, no host synchronisation.
 */
#ifndef HISTOGAN_B200_H_
#define HISTOGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HG_ABI_VERSION 4

/* error codes (negative; positive values are cudaError_t) */
#define HG_EINVAL   (-1)   /* bad argument (message in hg_last_error)          */
#define HG_ENOSUP   (-2)   /* configuration not supported by the CUDA path     */
#define HG_EWS      (-3)   /* workspace too small                              */
#define HG_EARCH    (-4)   /* device is not sm_100                             */

typedef void* hg_stream_t;             /* cudaStream_t */

int         hg_abi_version(void);
const char* hg_last_error(void);
/* number of CUDA kernels this library has launched so far in this process
 * (bench.py reports the per-step delta as "gpu_launches").                  */
uint64_t    hg_launch_count(void);
/* 0 when device `dev` can run the kernels (compute capability 10.x). */
int         hg_device_check(int dev);

/* ------------------------------------------------------------------------ *
 * RGB-uv histogram block
 *   replaces histogram_classes/RGBuvHistBlock.py:75-228 (forward) and the
 *   autograd backward PyTorch derives from it.
 * ------------------------------------------------------------------------ */

enum { HG_RESIZE_INTERPOLATION = 0, HG_RESIZE_SAMPLING = 1 };
enum { HG_METHOD_THRESHOLDING = 0, HG_METHOD_RBF = 1, HG_METHOD_INVERSE_QUADRATIC = 2 };
/* colour projection: RGB-uv log-chroma, 3 (or green-only 1) channels
 * (histogram_classes/RGBuvHistBlock.py:112-115,150-153,190-193) or rg-chromaticity,
 * 1 channel: u = R/(R+G+B+eps), v = G/(R+G+B+eps) (histogram_classes/rgChromaHistBlock.py:111-112),
 * or Lab, 1 channel: u = a, v = b, weight L (histogram_classes/LabHistBlock.py:104-111; the input
 * is already Lab scaled to [0,1]) */
enum { HG_PROJ_RGB_UV = 0, HG_PROJ_RG_CHROMA = 1, HG_PROJ_LAB = 2 };

typedef struct hg_hist_params {
  /* input image batch x: (B, C>=3, H, W) float32, element strides sb/sc/sh/sw */
  int32_t B, C, H, W;
  int64_t sb, sc, sh, sw;
  /* constructor arguments of RGBuvHistBlock (RGBuvHistBlock.py:29-73) */
  int32_t h;                /* bins per axis (default 64)                     */
  int32_t insz;             /* resize when H > insz or W > insz (default 150) */
  int32_t resizing;         /* HG_RESIZE_*                                     */
  int32_t method;           /* HG_METHOD_*                                     */
  double  sigma;            /* kernel width (default 0.02)                    */
  double  lo, hi;           /* sorted hist_boundary (default -3, 3)           */
  int32_t intensity_scale;  /* bool                                            */
  int32_t green_only;       /* bool: output has 1 channel (the G histogram)   */
  int32_t projection;       /* HG_PROJ_* (rg-chroma / Lab: 1 output channel)   */
} hg_hist_params;

/* number of pixels per image that enter the histogram after the optional
 * resize (RGBuvHistBlock.py:77-95): H*W, insz*insz or h*h.  <0 on error. */
int64_t hg_hist_num_pixels(const hg_hist_params* p);

size_t  hg_hist_fwd_workspace_bytes(const hg_hist_params* p);
size_t  hg_hist_bwd_workspace_bytes(const hg_hist_params* p);

/* Forward.  Outputs (nc = green_only ? 1 : 3):
 *   hist      (B, nc, h, h)  normalised histogram  == RGBuvHistBlock.forward(x)
 *   hist_sum  (B)            per-image sum of the un-normalised histogram
 *                            (saved for backward; RGBuvHistBlock.py:225-226)   */
int hg_hist_fwd(const float* x, const hg_hist_params* p,
                float* hist, float* hist_sum,
                void* ws, size_t ws_bytes, hg_stream_t stream);

/* Backward: grad_x (same geometry / strides as x, fully written, channels >= 3
 * get zeros) = d<grad_hist, hist>/dx, i.e. what autograd produces for the
 * reference forward given the upstream gradient grad_hist (B, nc, h, h).
 * method == thresholding: the bin masks are comparison results (constants for
 * autograd, RGBuvHistBlock.py:126-131), so the gradient flows only through the
 * intensity weight Iy and the normalisation; with intensity_scale == 0 the
 * reference's output does not require grad and grad_x is zero-filled.           */
int hg_hist_bwd(const float* x, const hg_hist_params* p,
                const float* hist, const float* hist_sum,
                const float* grad_hist, float* grad_x,
                void* ws, size_t ws_bytes, hg_stream_t stream);

/* Test/debug hooks used by the parity tests to separate kernel arithmetic from
 * libm differences: the pre-processed pixels (clamp + resize + first 3
 * channels; RGBuvHistBlock.py:76-99) as (B,3,N) and the float32 natural log
 * the kernels use.                                                            */
int hg_hist_preprocess(const float* x, const hg_hist_params* p, float* pixels,
                       hg_stream_t stream);
int hg_debug_logf(const float* in, float* out, int64_t n, hg_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Hellinger histogram loss
 *   replaces histoGAN/histoGAN.py:957-960 and ReHistoGAN/rehistoGAN.py:1011-1014
 *     loss = alpha * (1/sqrt(2)) * sqrt(sum((sqrt(T) - sqrt(H))^2)) / B
 * ------------------------------------------------------------------------ */

/* loss, q: device scalars (q = the sum under the square root, saved for bwd). */
size_t hg_hellinger_workspace_bytes(void);
int hg_hellinger_fwd(const float* target, const float* hist, int64_t numel, int32_t B,
                     float alpha, float* loss, float* q, void* ws, size_t ws_bytes,
                     hg_stream_t stream);
/* grad_hist / grad_target (either may be NULL) = grad_loss[0] * dloss/d(.)    */
int hg_hellinger_bwd(const float* target, const float* hist, int64_t numel, int32_t B,
                     float alpha, const float* q, const float* grad_loss,
                     float* grad_hist, float* grad_target, hg_stream_t stream);

/* ------------------------------------------------------------------------ *
 * NHWC convolution on tcgen05 tensor cores (TF32 operands, fp32 accumulate)
 *   the dense contraction of Conv2DMod.forward (histoGAN/histoGAN.py:420-440),
 *   RGBBlock's 1x1 conv (:375,382), and the DiscriminatorBlock convs (:505-526);
 *   with mode-1 packed weights also their input gradients.
 *   x: (B,H,W,Cin) float32 NHWC, already TF32-rounded by its producer; y: (B,OH,OW,Cout);
 *   Cin and Cout multiples of 4.  w_packed: [round_up(Cout,32)][KH][KW][round_up(Cin,32)]
 *   from hg_pack_conv_weight (zero padded; tensors with fewer than 32 channels are completed
 *   by TMA's out-of-bounds zero fill).
 * ------------------------------------------------------------------------ */
#define HG_CONV_LRELU       1   /* LeakyReLU(slope) after scale/bias/noise       */
#define HG_CONV_ROUND_TF32  2   /* round the stored result to TF32 (RN-even)     */

typedef struct hg_conv_params {
  int32_t B, H, W, Cin;
  int32_t Cout, KH, KW, stride, pad;
  int32_t OH, OW;                       /* output extent: normally (H + 2 pad - KH)/stride + 1;
                                         * a larger value (<= H + 2 pad) is allowed -- windows
                                         * that reach past the input read zeros                */
} hg_conv_params;

/* Fused epilogue, applied in this order to the accumulator of output (b,oh,ow,co):
 *   v *= scale[b][co]                                  demodulation   (:427-429)
 *   v += bias[co]                                      nn.Conv2d bias (:506-518)
 *   v += noise[b][ow][oh] * noise_w[co] + noise_b[co]  to_noise Linear + the
 *                                                      spatial transpose (:465-467)
 *   v  = LeakyReLU(v)                                  if HG_CONV_LRELU (:471,476)
 *   v += residual[b][oh][ow][co]                       DiscriminatorBlock (:524)
 * Any pointer may be NULL to skip its step.
 * out_*_stride (in floats, all three 0 = dense NHWC): where output pixel (b,oh,ow) goes,
 *   y + b*out_img_stride + oh*out_row_stride + ow*out_pix_stride -- lets a convolution
 *   write one parity class of a twice-as-large tensor (the stride-2 input gradient is four
 *   such convolutions over dy, ops.py); `residual` is not supported together with it.    */
typedef struct hg_conv_epilogue {
  const float* scale;
  const float* bias;
  const float* noise;      /* (B, noise_size, noise_size) */
  const float* noise_w;
  const float* noise_b;
  const float* residual;
  int32_t noise_size;
  int32_t flags;
  float   lrelu_slope;
  int32_t reserved_;
  int64_t out_img_stride, out_row_stride, out_pix_stride;
} hg_conv_epilogue;

int hg_conv2d_fwd(const float* x, const float* w_packed, float* y, const hg_conv_params* p,
                  const hg_conv_epilogue* ep, hg_stream_t stream);

/* (Cout,Cin,KH,KW) parameter -> packed K-major TF32 weight, N and K extents zero-padded to
 * multiples of 32.
 * mode 0: forward  [Cout_p][KH][KW][Cin_p];
 * mode 1: dgrad    [Cin_p][KH][KW][Cout_p], taps flipped (conv of dy with it = dx);
 * mode | HG_PACK_FROM_OHWI: `w` is stored channels_last, i.e. memory [Cout][KH][KW][Cin] (the
 * storage this package's modules give their conv weights: forward packing is then a rounding
 * copy and the weight gradient of hg_conv2d_wgrad already has the parameter's layout);
 * otherwise `w` is plain contiguous OIHW (the reference's state_dict storage).          */
#define HG_PACK_FROM_OHWI 2
int hg_pack_conv_weight(const float* w, float* w_packed, int32_t Cout, int32_t Cin,
                        int32_t KH, int32_t KW, int32_t mode, hg_stream_t stream);

/* Weight gradient: dw_packed [Cout][KH][KW][round_up(Cin,32)] (fully written) =
 *   sum over pixels of dy (B,OH,OW,Cout) x shifted x (B,H,W,Cin); both TF32-rounded
 *   by their producers.  hg_unpack_conv_wgrad converts to the OIHW layout of the
 *   nn.Parameter (accumulate != 0: dw_oihw += ...).                              */
int hg_conv2d_wgrad(const float* dy, const float* x, float* dw_packed, const hg_conv_params* p,
                    hg_stream_t stream);
int hg_unpack_conv_wgrad(const float* dw_packed, float* dw_oihw, int32_t Cout, int32_t Cin,
                         int32_t KH, int32_t KW, int32_t accumulate, hg_stream_t stream);

/* Producer-side helpers of the modulated convolution (NHWC, C % 4 == 0):
 *   hg_modulate_round: out = tf32_round?(x * mod[b][c])   -- the style modulation
 *     w2 * (w1 + 1) of Conv2DMod (histoGAN.py:423-425) moved onto the activations;
 *     mod may be NULL (round only), do_round = 0 gives the plain product (backward).
 *   hg_channel_dot:    out[b][c] = sum_hw a * g            -- d/d(mod) of the above. */
int hg_modulate_round(const float* x, const float* mod, float* out, int32_t B, int32_t HW,
                      int32_t C, int32_t do_round, hg_stream_t stream);
int hg_channel_dot(const float* a, const float* g, float* out, int32_t B, int32_t HW, int32_t C,
                   hg_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Fused element-wise / reduction halves of GeneratorBlock and RGBBlock (NHWC)
 * ------------------------------------------------------------------------ */

/* adjoint of the conv epilogue y = lrelu(d[b,c]*z + noise[b,ow,oh]*nw[c] + nb[c])
 * (histoGAN/histoGAN.py:427-429,465-471): from dy and y
 *   dz (B,H,W,C) = tf32_round(dpre * d)   (input of dgrad / wgrad)
 *   gd (B,C)     = d loss / d d ;  gnw (C), gnb (C) = grads of the to_noise Linear
 * d / noise may be NULL (then gd / gnw,gnb are ignored).                           */
int hg_modconv_epilogue_bwd(const float* dy, const float* y, const float* d, const float* noise,
                            const float* noise_w, const float* noise_b, float* dz, float* gd,
                            float* gnw, float* gnb, int32_t B, int32_t H, int32_t W, int32_t C,
                            int32_t noise_size, float slope, hg_stream_t stream);

/* adjoint of xm = x * mod[b,c] (the style modulation, :423-425): dxm_inout *= mod in place
 * (-> dx) and gmod (B,C) = sum_hw dxm * x.                                         */
int hg_modulate_bwd(float* dxm_inout, const float* x, const float* mod, float* gmod, int32_t B,
                    int32_t HW, int32_t C, hg_stream_t stream);

/* the same two with the generator's 2x bilinear up-sampling (nn.Upsample(scale_factor=2,
 * mode='bilinear', align_corners=False), :446-447,462-463) folded in: x is the LOW-resolution
 * (B,H,W,C) activation, xm / dxm the (B,2H,2W,C) conv input / its gradient; the up-sampled
 * activation is never materialised.                                                 */
int hg_upsample_modulate_round(const float* x, const float* mod, float* xm, int32_t B, int32_t H,
                               int32_t W, int32_t C, hg_stream_t stream);
int hg_upsample_modulate_bwd(const float* dxm, const float* x, const float* mod, float* dx,
                             float* gmod, int32_t B, int32_t H, int32_t W, int32_t C,
                             hg_stream_t stream);

/* RGBBlock (:380-386): rgb (B,3,HW) planar = sum_c x[b,p,c] * wmod[b,o,c] (+ prev), with
 * wmod (B,3,C) = conv.weight * (style + 1); backward: dx (B,HW,C) and gw (B,3,C).  */
int hg_torgb_fwd(const float* x, const float* wmod, const float* prev, float* rgb, int32_t B,
                 int32_t HW, int32_t C, hg_stream_t stream);
int hg_torgb_bwd(const float* drgb, const float* x, const float* wmod, float* dx, float* gw,
                 int32_t B, int32_t HW, int32_t C, int32_t accumulate_dx, hg_stream_t stream);

/* adjoint of y = LeakyReLU(conv + bias) (DiscriminatorBlock, :507-518): dpre = TF32-rounded
 * dy * (y > 0 ? 1 : slope) (y NULL: no activation) and gb (C) = sum over batch and pixels
 * of the unrounded dpre (gb may be NULL).                                          */
int hg_bias_act_bwd(const float* dy, const float* y, float* dpre, float* gb, int32_t B,
                    int32_t HW, int32_t C, float slope, hg_stream_t stream);

/* ------------------------------------------------------------------------ *
 * The generator's "style path" (csrc/style.cu): every small dense op between the latents
 * and the modulated convolutions as GROUPED skinny GEMMs (batch <= 32 on the lanes of a warp),
 * one launch for all layers of a generator pass.  HOST arrays of `count` (<= 24) DEVICE
 * pointers / extents.  x_g (B,K_g), W_g (J_g,K_g) row-major, y_g (B,J_g); K_g % 4 == 0.
 *   fwd:  y_g = f(x_g' W_g^T + bias_g),  x' = x or x^2 (HG_LIN_SQUARE_INPUT),
 *         f = [rsqrt(. + eps)] -> [LeakyReLU(slope)] -> [+ 1]   in that order, as flagged:
 *         to_style Linear + 1  (GeneratorBlock.to_style1/2, RGBBlock.to_style, histoGAN.py:372,451,455
 *         and the `y + 1` of Conv2DMod :423-425):            HG_LIN_ADD_ONE
 *         demodulation d = rsqrt(mod^2 Wsq^T + 1e-8) (:427-429): HG_LIN_SQUARE_INPUT | HG_LIN_RSQRT_EPS
 *   bwd (B <= 32): gw_g = gy_g^T x_g', gb_g = sum_b gy_g, gx_g = gy_g W_g [* 2 x_g: HG_LIN_POST_2X]
 *         [added to gx_g: HG_LIN_ACCUMULATE]; any of the three output tables / entries may be NULL.
 *         Deterministic (no atomics).
 * ------------------------------------------------------------------------ */
enum { HG_LIN_ADD_ONE = 1, HG_LIN_LRELU = 2, HG_LIN_SQUARE_INPUT = 4, HG_LIN_RSQRT_EPS = 8,
       HG_LIN_POST_2X = 16, HG_LIN_ACCUMULATE = 32 };
int hg_grouped_linear_fwd(int32_t count, const float* const* x, const float* const* w,
                          const float* const* bias, float* const* y, const int32_t* J,
                          const int32_t* K, int32_t B, int32_t flags, float slope, float eps,
                          hg_stream_t stream);
int hg_grouped_linear_bwd(int32_t count, const float* const* x, const float* const* w,
                          const float* const* gy, float* const* gw, float* const* gb,
                          float* const* gx, const int32_t* J, const int32_t* K, int32_t B,
                          int32_t flags, hg_stream_t stream);
/* Wsq (Cout,Cin) = sum over the T = KH*KW taps of w^2, w stored channels_last [Cout][T][Cin]
 * (the batch-independent part of the demodulation, histoGAN.py:428).                       */
int hg_weight_sqsum(const float* w, float* wsq, int32_t Cout, int32_t T, int32_t Cin,
                    hg_stream_t stream);
/* adjoint of d = rsqrt(mod^2 Wsq^T + eps) given gd (B,Cout):  gmod_accum (B,Cin) += d d/d mod,
 * dw_accum [Cout][T][Cin] += 2 w * (sum_b t mod^2), t = -1/2 gd d^3.  Either accumulator may be
 * NULL.  t_ws: B*Cout floats of scratch.  B <= 32.                                          */
int hg_demod_bwd(const float* gd, const float* d, const float* mod, const float* wsq,
                 const float* w, float* gmod_accum, float* dw_accum, float* t_ws, int32_t B,
                 int32_t Cout, int32_t T, int32_t Cin, hg_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Convolutions whose INPUT is the 3/4-channel image (DiscriminatorBlock 0: conv_res 1x1, net[0] 3x3;
 * histoGAN/histoGAN.py:507-511,520-523) on the CUDA cores, exact fp32, reading the planar image with
 * its own strides (csrc/conv_small.cu).  k in {1,3}, stride 1, pad k/2, Cin <= 4, Cout % 4 == 0, <= 64.
 * w: contiguous OIHW.  NHWC tensors (y, dy, residual) carry Cp >= Cout channels (Cp % 4 == 0; y's
 * padding channels are written as zeros).  sb,sc,sh,sw: element strides of the planar tensor.
 *   fwd   y = [round]( [lrelu](conv(x, w) + bias) [+ residual] )     flags: HG_CONV_LRELU | HG_CONV_ROUND_TF32
 *   dgrad dx (planar, fully written) = conv^T(dy, w)
 *   wgrad dw (contiguous OIHW) = sum over batch and pixels; deterministic (per-CTA partials in ws,
 *         hg_conv_small_wgrad_workspace_bytes, added in index order)
 * Cin == 3, Cout == Cp == 16 (network_capacity 16) runs specialised kernels whose filter lives in constant
 * memory: the calls of one device must then be issued on ONE stream (each call rewrites the constant bank
 * in stream order); HG_SMALL_GENERIC=1 selects the generic kernels.
 * ------------------------------------------------------------------------ */
int hg_conv_small_fwd(const float* x, const float* w, const float* bias, const float* residual, float* y,
                      int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t Cp, int32_t k,
                      int64_t sb, int64_t sc, int64_t sh, int64_t sw, int32_t flags, float slope,
                      hg_stream_t stream);
int hg_conv_small_dgrad(const float* dy, const float* w, float* dx, int32_t B, int32_t Cin, int32_t H,
                        int32_t W, int32_t Cout, int32_t Cp, int32_t k, int64_t sb, int64_t sc, int64_t sh,
                        int64_t sw, hg_stream_t stream);
size_t hg_conv_small_wgrad_workspace_bytes(int32_t Cin, int32_t Cout, int32_t k);
int hg_conv_small_wgrad(const float* dy, const float* x, float* dw, void* ws, size_t ws_bytes, int32_t B,
                        int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t Cp, int32_t k, int64_t sb,
                        int64_t sc, int64_t sh, int64_t sw, hg_stream_t stream);

/* Discriminator input staging (histoGAN/histoGAN.py:613-617): x (B,C,H,W) float32 with element
 * strides sb,sc,sh,sw -> out NHWC (B,H,W,Cp), channels zero-padded to Cp (multiple of 4),
 * TF32-rounded (round to nearest even), in one pass.                                   */
int hg_pad_round_nhwc(const float* x, float* out, int32_t B, int32_t C, int32_t H, int32_t W,
                      int32_t Cp, int64_t sb, int64_t sc, int64_t sh, int64_t sw, hg_stream_t stream);

/* 2x bilinear up-sampling (align_corners = False) of the planar RGB skip tensor of RGBBlock
 * (histoGAN/histoGAN.py:377-378,388-389).  backward == 0: x (planes,H,W) -> y (planes,2H,2W);
 * backward != 0: the adjoint, x = dy (planes,2H,2W) -> y = dx (planes,H,W).  H, W always name
 * the LOW-resolution extent.                                                         */
int hg_upsample2x_planar(const float* x, float* y, int32_t planes, int32_t H, int32_t W,
                         int32_t backward, hg_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Fused multi-tensor DiffGrad step (torch_optimizer.DiffGrad as used at
 * histoGAN/histoGAN.py:670-671,932,989).  HOST arrays of `count` DEVICE pointers;
 * step_size = lr * sqrt(1 - beta2^t) / (1 - beta1^t) is computed by the caller.
 * packed (ABI v4; table may be NULL, entries may be NULL): where given, the kernel also
 * writes the TF32-rounded copy of the UPDATED parameter there (numel floats, same linear
 * order) -- the forward tensor-core operand of a channels_last conv weight
 * (hg_pack_conv_weight mode 0 | HG_PACK_FROM_OHWI with Cout, Cin multiples of 32).
 * ------------------------------------------------------------------------ */
int hg_diffgrad_step(int32_t count, float* const* p, const float* const* g, float* const* m,
                     float* const* v, float* const* prev, float* const* packed,
                     const int64_t* numel, float beta1, float beta2, float eps, float step_size,
                     float weight_decay, hg_stream_t stream);

/* Fused multi-tensor exponential moving average (HistoGAN.EMA, histoGAN/histoGAN.py:62-69,
 * 698-707): ma_i = beta * ma_i + (1 - beta) * cur_i for `count` tensors (HOST arrays of DEVICE
 * pointers), one pass.                                                                     */
int hg_ema_update(int32_t count, float* const* ma, const float* const* cur, const int64_t* numel,
                  float beta, hg_stream_t stream);

/* ------------------------------------------------------------------------ *
 * ReHistoGAN recolouring step (ReHistoGAN/rehistoGAN.py), SURVEY 8f-1.
 * ------------------------------------------------------------------------ */

/* nn.InstanceNorm2d(affine=False) + LeakyReLU of EncoderBlock (:489-495) on an NHWC
 * activation x (B,HW,C), C % 4 == 0: y = lrelu((x - mean_bc) * rsqrt(var_bc + eps)), biased
 * variance; `stats` (B*C*2 doubles: sum, sum of squares) is written by fwd and read by bwd.
 * round_tf32: store y / dx rounded to TF32 (they feed a tensor-core convolution).
 * bwd: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * lrelu'(xhat); `ws` is
 * B*C*2 doubles of scratch.                                                             */
int hg_instnorm_lrelu_fwd(const float* x, float* y, double* stats, int32_t B, int32_t HW, int32_t C,
                          float eps, float slope, int32_t round_tf32, hg_stream_t stream);
int hg_instnorm_lrelu_bwd(const float* dy, const float* x, const double* stats, float* dx, double* ws,
                          int32_t B, int32_t HW, int32_t C, float eps, float slope,
                          int32_t round_tf32, hg_stream_t stream);

/* reconstruction_loss('2nd gradient').compute_loss (:293-299,321-324): a, b planar (B,3,H,W);
 * loss[0] = mean over (B,H,W) of | lap(sum_c a_c) - lap(sum_c b_c) |, 5-point laplacian with
 * zero padding; `sign` (B*H*W int8) keeps sign(lap a - lap b) for the backward, which writes
 * d loss / d b = -(gout[0] / (B*H*W)) * lap(sign) into all three planes of db.           */
size_t hg_laplacian_l1_workspace_bytes(void);
int hg_laplacian_l1_fwd(const float* a, const float* b, float* loss, int8_t* sign, void* ws,
                        size_t ws_bytes, int32_t B, int32_t H, int32_t W, hg_stream_t stream);
int hg_laplacian_l1_bwd(const int8_t* sign, const float* gout, float* db, int32_t B, int32_t H,
                        int32_t W, hg_stream_t stream);

/* gaussian_op (:228-232; the filter of get_gaussian_kernel :207-225 repeated over the planes):
 * y[pl] = x[pl] (*) kernel (K x K, K <= 15), zero padding `pad` on every side, output
 * (H + 2 pad - K + 1) x (W + 2 pad - K + 1).  pad = 0 is the reference's valid convolution;
 * pad = K - 1 with flip = 1 is its adjoint (the backward).                               */
int hg_depthwise_conv(const float* x, const float* kernel, float* y, int32_t planes, int32_t H,
                      int32_t W, int32_t K, int32_t pad, int32_t flip, hg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HISTOGAN_B200_H_ */
