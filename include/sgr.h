/*
 * sgr.h — C ABI of libsgr.so, the sm_100a StyleGAN2 synthesis path.
 *
 * Drop-in boundary.  The reference repo's native ABI for this path is two pybind11
 * functions JIT-built by torch.utils.cpp_extension.load:
 *     fused_bias_act(input, bias, refer, act, grad, alpha, scale)        libs/gan/StyleGAN2/op/fused_bias_act.cpp:14-24
 *     upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0..y1)   libs/gan/StyleGAN2/op/upfirdn2d.cpp:16-26
 * plus F.conv2d / F.conv_transpose2d(groups=batch) called from ModulatedConv2d.forward
 * (libs/gan/StyleGAN2/model.py:232-273).  The entry points below replace all three.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates); nothing is allocated here;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*) and never synchronises;
 *   - return value: 0 on success, non-zero on error (sgr_last_error() gives a thread-local message);
 *   - all tensors are fp32 unless stated; activations crossing this boundary are NCHW contiguous like the
 *     reference's; the bf16 hi/lo "C8" layout is internal: [plane(hi,lo)][B][C/8][H][W][8] bf16.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns an error.
 */
#ifndef SGR_H_
#define SGR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGR_MAX_STYLED 24
#define SGR_MAX_RGB 12
#define SGR_STYLE_DIM 512

/* Operand formats of the split-precision tensor-core GEMM (3 MMAs per product either way):
 *   SGR_FMT_BF16  bf16 hi/lo planes: 16 mantissa bits, fp32 range (~2e-5 relative)            — gradient GEMMs
 *   SGR_FMT_FP16  fp16 hi/lo planes: 22 mantissa bits (~5e-7 relative); activations are stored x 2^-4 and weights
 *                 x 2^8 inside the library and saturate at +-65504 (|activation * style| up to 1e6)  — forward pass */
#define SGR_FMT_BF16 0
#define SGR_FMT_FP16 1

const char* sgr_version(void);
const char* sgr_last_error(void);
/* number of kernels of this library launched by the calling thread since the last reset (bench bookkeeping) */
long long sgr_launch_count(void);
void sgr_reset_launch_count(void);

/* Per-launch timing of the modconv kernel (bench.py's roofline): while enabled, every sgr_modconv_forward records a
 * CUDA event pair on its stream around each of its launches.  sgr_profile_collect synchronises on the recorded events, writes up
 * to `cap` durations (milliseconds, launch order) and clears the list; returns the number of launches recorded. */
void sgr_profile_enable(int on);
int sgr_profile_collect(float* ms, int cap);
/* same, also returning the kind of each timed launch: 0 = tensor-core convolution GEMM, 1 = FIR/epilogue pass of an
 * up == 2 layer (HBM-bound) */
int sgr_profile_collect_tagged(float* ms, int* tags, int cap);

/* ---------------------------------------------------------------------------------------------------------
 * upfirdn2d: zero-insert upsample x`up`, pad (pad0 before / pad1 after, negative crops), true 2-D convolution
 * with taps[kh][kw], decimate x`down`.  Replaces op/upfirdn2d.cpp:16-26 + op/upfirdn2d_kernel.cu:52-272
 * (same factors on both axes, as every call site in model.py uses).  x: [planes,in_h,in_w], y: [planes,out_h,out_w]
 * with out = (in*up + pad0 + pad1 - k + down) / down  (op/upfirdn2d.py:104-105).
 */
int sgr_upfirdn2d(const float* x, float* y, const float* taps, int planes, int in_h, int in_w, int up, int down,
                  int pad0, int pad1, int kh, int kw, void* stream);

/* fused bias + leaky-relu * scale.  Replaces op/fused_bias_act.cpp:14-24 / fused_bias_act_kernel.cu:18-99.
 * grad == 0: y = lrelu(x + bias[c], slope) * scale          (act*10+grad == 30; bias may be NULL)
 * grad == 1: y = (ref > 0 ? x : slope * x) * scale           (== 31; `ref` is the saved forward OUTPUT)
 * x: [outer, channels, inner] contiguous. */
int sgr_fused_bias_act(const float* x, const float* bias, const float* ref, float* y, long long outer, int channels,
                       long long inner, int grad, float slope, float scale, void* stream);

/* Output stage: frames [B,3,h,w] fp32 in [-1,1] -> uint8 [B,out_h,out_w,3] (HWC, RGB order) with the
 * reference's arithmetic  uint8((clamp(x,-1,1) + 1) / (2 + 1e-5) * 255)  (libs/utilities/image_utils.py:97-111, np.uint8 at
 * libs/utilities/utils_inference.py:16); out_h/out_w < h/w additionally applies the block-mean AdaptiveAvgPool2d of
 * generate_image (libs/utilities/generic.py:146-148; h % out_h == 0, w % out_w == 0).  4x less device->host traffic. */
int sgr_frames_to_uint8(const float* frames, unsigned char* out, int batch, int h, int w, int out_h, int out_w, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Weight packing for the tcgen05 implicit-GEMM convolution.
 * weight: [cout, cin, k, k] fp32 (the reference parameter ModulatedConv2d.weight[0], model.py:216-218), k = 3 or 1.
 * up == 1: the stride-2 transposed conv + 4x4 FIR of model.py:246-257 is folded into four 3x3 phase kernels
 *          (W (*) fir, 6x6); fir = the layer's blur.kernel buffer [4,4].  36 Cin Cout MACs / input pixel.
 * up == 2: scatter form at the minimal 9 Cin Cout MACs / input pixel: the 9 taps are packed by (row,column) shift of
 *          the input tile and output parity; the FIR runs afterwards on the parity planes (sgr_modconv_forward does
 *          both).  fir is not needed for packing.  Fixed column tile: 256 (cout >= 64) or 4*cout.
 * transpose != 0 packs the adjoint (data-gradient) operator instead: GEMM columns = cin, K = cout (x4 for up).
 * packed: bf16 hi/lo slabs in shared-memory image order, sgr_packed_weight_bytes() bytes.
 * wsq:    [cout, cin] fp32 = sum_k (weight*scale)^2 for the demodulation mini-GEMM (may be NULL).
 */
size_t sgr_packed_weight_bytes(int cout, int cin, int ksize, int up, int transpose);
/* bytes of the fp32 parity-plane scratch an up == 2 convolution needs: [B][4][cout/4][h_in+1][w_in+1][4] */
size_t sgr_up_scratch_bytes(int batch, int cout, int h_in, int w_in);
/* GEMM column tile (32/64/128/256) the library would pick for a layer with n_total GEMM columns (cout, x4 for up
 * layers; cin for the adjoint) on an h_in x w_in grid at this batch: fills the 148 SMs on the small layers.  The packed
 * weight layout depends on it, so pack and convolve with the same value (0 = the default min(n_total, 256)). */
int sgr_choose_column_tile(int batch, int h_in, int w_in, int n_total);
int sgr_pack_modconv_weight(const float* weight, const float* fir, int cout, int cin, int ksize, int up,
                            int transpose, int format, int column_tile, void* packed, float* wsq, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Layout helpers (module-level calls and tests; the fused network path never needs them).
 * NCHW fp32 (optionally * scale[b,c]) -> C8 bf16 hi/lo planes; s2d != 0 additionally folds 2x2 pixel phases
 * into channels (channel = phase*C + c at half resolution), the layout the up-layer adjoint consumes. */
int sgr_nchw_to_c8(const float* x, const float* scale, void* out_c8, int batch, int channels, int h, int w, int s2d,
                   int format, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * One modulated convolution (ModulatedConv2d.forward + NoiseInjection + FusedLeakyReLU, model.py:232-287,331-337)
 * on tensor cores.  Input is C8 hi/lo planes ALREADY multiplied by the style s[b,cin] (the modulation of
 * model.py:235-236 moved from the weights to the activations).
 *   z   = conv(x_c8, W) (3x3 pad 1 | 1x1 | polyphase up)            fp32 accumulate of bf16x3 products
 *   t   = z * demod[b,o] + noise_weight[0] * noise[y,x] + bias[o]    (each term optional: NULL pointer skips it)
 *   t   = act ? max(t, 0.2 t) : t
 *   out_f32[b,o,y,x]  = t * act_gain                                  (optional, NCHW fp32)
 *   out_c8            = split_bf16(t * (s2 ? s2[b,o] : act_gain))     (optional, next layer's input)
 *   rgb_partial[n_tile][b,c,y,x] = sum_{o in column tile} t * rgb_coef[b,c,o]   (optional fused ToRGB; one slot per
 *                                    column tile, plain stores: summing the slots in order is deterministic)
 * up == 2 runs two kernels: the tensor-core GEMM leaves the four parity planes of conv_transpose2d in t_scratch, then
 * an HBM-bound kernel applies the 4x4 FIR (Blur, model.py:72-88) and the same fused epilogue.
 */
typedef struct sgr_conv_args {
  int batch, cin, cout, h_in, w_in;
  int ksize;              /* 3 or 1 */
  int up;                 /* 0: same resolution; 1: output is 2x, polyphase weights; 2: output is 2x, scatter weights + FIR pass */
  int act;                /* apply leaky-relu 0.2 */
  float act_gain;         /* sqrt(2) for StyledConv, 1 for raw conv */
  int operand_format;     /* SGR_FMT_* of x_c8 and w_packed */
  int column_tile;        /* the value w_packed was packed with (0 = default) */
  int out_format;         /* SGR_FMT_* written to out_c8 */
  const void* x_c8;       /* [2][B][cin/8][h_in][w_in][8] bf16 */
  const void* w_packed;
  const float* demod;     /* [B,cout] or NULL */
  const float* bias;      /* [cout] or NULL */
  const float* noise;     /* [h_out,w_out] (or [B,h_out,w_out] with noise_batch_stride = h_out*w_out) or NULL */
  long long noise_batch_stride; /* elements between samples' noise maps; 0 = shared (registered buffer) */
  const float* noise_weight; /* [1] device scalar (required when noise != NULL) */
  const float* s2;        /* [B,cout] or NULL */
  void* out_c8;           /* or NULL */
  float* out_f32;         /* or NULL */
  const float* rgb_coef;  /* [B,3,cout] or NULL */
  float* rgb_partial;     /* [cout/column_tile][B,3,h_out,w_out], fully overwritten */
  float* t_scratch;       /* up == 2: sgr_up_scratch_bytes() of scratch */
  const float* fir;       /* up == 2: blur.kernel [4,4]; must be an outer product (rank 1), as make_kernel builds it */
  void* splitk_scratch;   /* optional: >= splitk_scratch_bytes of scratch enabling split-K on layers with fewer output
                             tiles than SMs (4x4 .. 16x16 grids, small batches); NULL = never split */
  size_t splitk_scratch_bytes;
  int single_pass;        /* 0: split-precision products (3 MMAs, fp32 parity); 1: hi x hi only = plain bf16/fp16
                             tensor-core precision with fp32 accumulation (BASELINE config 5), 1 MMA per product */
} sgr_conv_args;
int sgr_modconv_forward(const sgr_conv_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Weight gradient of one modulated 3x3 convolution (the cuDNN wgrad behind ATen autograd of model.py:254,263,269,
 * needed when the generator itself is trained: optimize_g, libs/optimization.py:25-72) as ONE tensor-core GEMM over all
 * pixels of all samples, bf16 hi/lo operands (3 MMAs per product), deterministic:
 *   up == 0:  gw[o,i,ky,kx] = sum_{b,y,x} gz[b,o,y,x] * x[b,i,y+ky-1,x+kx-1]
 *   up == 2:  gw[o,i,ky,kx] = sum_{b,y,x} G[b,o,2y+ky,2x+kx] * x[b,i,y,x],  G = gradient of the (2h+1)x(2w+1) output of
 *             conv_transpose2d(stride 2), passed as its four parity planes stacked as channels:
 *             gz_c8 channel (2*pu+pv)*cout + o at (I,J) = G[o, 2I+pu, 2J+pv] on the (h_in+1) x (w_in+1) grid (0 outside G).
 * x_c8 is the layer input ALREADY multiplied by the style (what the forward GEMM consumed); both operands are C8 planes
 * in SGR_FMT_BF16.  gw is the gradient with respect to W_bar = scale * weight[0] before the demodulation term. */
typedef struct sgr_wgrad_args {
  int batch, cin, cout, h_in, w_in;
  int up;                 /* 0 or 2 */
  const void* x_c8;       /* [2][B][cin/8][h_in][w_in][8] bf16 */
  const void* gz_c8;      /* up == 0: [2][B][cout/8][h_in][w_in][8]; up == 2: [2][B][4*cout/8][h_in+1][w_in+1][8] */
  float* gw;              /* [cout,cin,3,3], fully overwritten */
  void* scratch;          /* >= sgr_wgrad_scratch_bytes(cout, cin): per-slice fp32 partial sums */
  size_t scratch_bytes;
} sgr_wgrad_args;
size_t sgr_wgrad_scratch_bytes(int cout, int cin);
int sgr_modconv_wgrad(const sgr_wgrad_args* args, void* stream);

/* styles s[b,i] = latent_row[b,:] . mod_weight[i,:] / sqrt(512) + mod_bias[i]   (EqualLinear, model.py:148-157,235)
 * demod d[b,o]  = rsqrt(sum_i s[b,i]^2 wsq[o,i] + 1e-8)                          (model.py:238-240 in the form of §9.1) */
int sgr_style_affine(const float* latent, int latent_stride, int batch, const float* mod_weight,
                     const float* mod_bias, int cin, float* s_out, void* stream);
int sgr_demod(const float* s, const float* wsq, int batch, int cin, int cout, float* d_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Whole synthesis network (Generator.forward after the mapping/truncation glue, model.py:519-534).
 */
typedef struct sgr_styled_layer {
  int cin, cout;
  int up;                    /* 0, or the packing mode of w_packed for an upsampling layer: 1 polyphase, 2 scatter */
  int latent_row;
  int column_tile;           /* of w_packed (0 = default) */
  int column_tile_t;         /* of w_packed_t */
  const void* w_packed;      /* forward operator, sgr_pack_modconv_weight(transpose=0) */
  const void* w_packed_t;    /* adjoint operator (transpose=1); only sgr_synthesis_backward reads it, may be NULL otherwise */
  const float* wsq;          /* [cout,cin] */
  const float* mod_weight;   /* [cin,512] */
  const float* mod_bias;     /* [cin] */
  const float* noise;        /* [res,res], or [B,res,res] when noise_batch_stride != 0 */
  long long noise_batch_stride;
  const float* noise_weight; /* [1] */
  const float* act_bias;     /* [cout] */
  const float* fir;          /* blur.kernel [4,4] of an upsampling layer (read when up == 2) */
} sgr_styled_layer;

typedef struct sgr_rgb_layer {
  int cin, latent_row;
  const float* weight;       /* [3,cin] (ToRGB.conv.weight[0,:,:,0,0]) */
  const float* mod_weight;   /* [cin,512] */
  const float* mod_bias;     /* [cin] */
  const float* bias;         /* [3] */
  const float* fir;          /* upsample.kernel [4,4] (NULL for to_rgb1) */
  const float* fir_flipped;  /* flip(fir) in both axes, for the adjoint; only sgr_synthesis_backward reads it */
} sgr_rgb_layer;

typedef struct sgr_synthesis {
  int size;                  /* output resolution, power of two >= 8 */
  int n_styled;              /* conv1 + convs.* = 2*log2(size) - 3 */
  int n_rgb;                 /* to_rgb1 + to_rgbs.* = log2(size) - 1 */
  int n_latent;
  int format;                /* SGR_FMT_* of the forward pass (w_packed of every layer must be packed with it) */
  int single_pass;           /* see sgr_conv_args.single_pass; applies to every layer of sgr_synthesis_forward */
  const float* const_input;  /* [512,4,4] */
  sgr_styled_layer styled[SGR_MAX_STYLED];
  sgr_rgb_layer rgb[SGR_MAX_RGB];
} sgr_synthesis;

size_t sgr_synthesis_workspace_bytes(const sgr_synthesis* net, int batch);
/* latent: [B,n_latent,512] (already truncated / shifted); image: [B,3,size,size] fp32 NCHW.
 * feats: NULL, or n_styled device pointers (entries may be NULL) receiving each StyledConv output [B,cout,res,res]. */
int sgr_synthesis_forward(const sgr_synthesis* net, const float* latent, int batch, float* image, void* workspace,
                          size_t workspace_bytes, float* const* feats, void* stream);

/* Same, with the output stage of the callers folded into the last ToRGB tail: `frames_u8`
 * [B,u8_h,u8_w,3] receives uint8 HWC frames computed as libs/utilities/generic.py:146-148 (AdaptiveAvgPool2d when
 * u8_h < size; u8_h, u8_w must divide size), libs/utilities/image_utils.py:97-111 (tensor_to_image: clamp to [-1,1],
 * (v+1)/(2+1e-5)*255) and np.uint8 (libs/utilities/utils_inference.py:16) produce them.  `image` may then be NULL: the fp32
 * frame is never written. */
typedef struct sgr_forward_extras {
  unsigned char* frames_u8;
  int u8_h, u8_w;
} sgr_forward_extras;
int sgr_synthesis_forward_ex(const sgr_synthesis* net, const float* latent, int batch, float* image, void* workspace,
                             size_t workspace_bytes, float* const* feats, const sgr_forward_extras* extras, void* stream);

/* Gradient of sum(image * grad_image) with respect to `latent` (dlatent: [B,n_latent,512], overwritten).
 * feats: the n_styled saved StyledConv outputs of the forward call (all required).  Replaces ATen autograd through
 * F.conv2d / F.conv_transpose2d and the modulation graph (model.py:232-273) for the A-matrix training step
 * (libs/trainer.py:177-189); generator weight gradients are not produced. */
/* Optional extra outputs of the backward pass: what a caller needs to assemble the gradients of the GENERATOR's own
 * parameters (optimize_g, libs/optimization.py:25-72).  Any pointer (array or entry) may be NULL.
 *   gfeats[l]     [B,cout_l,res,res]  dL/d(output of StyledConv l)                       (n_styled entries)
 *   ds_styled[l]  [B,cin_l]           dL/d(style s_l), convolution + demodulation terms  (n_styled entries)
 *   ds_rgb[r]     [B,cin_r]           dL/d(style of ToRGB r)                             (n_rgb entries)
 *   g_input       [B,cin_0,4,4]       dL/d(const_input * s_0), the modulated input of conv1 */
typedef struct sgr_backward_extras {
  float* const* gfeats;
  float* const* ds_styled;
  float* const* ds_rgb;
  float* g_input;
  const struct sgr_param_grads* params;  /* NULL, or where to write the gradient of every generator parameter (below) */
  void* wgrad_scratch;      /* required with params: sgr_synthesis_wgrad_scratch_bytes() bytes, 256-byte aligned */
  size_t wgrad_scratch_bytes;
} sgr_backward_extras;

/* Gradients of the generator's own parameters (train() mode: optimize_g, libs/optimization.py:25-72),
 * produced inside sgr_synthesis_backward_ex: the weight gradients by the tcgen05 GEMM of sgr_modconv_wgrad on the
 * operands the backward pass already holds (demodulation term and EqualLR scale fused into its reduction pass), the
 * per-channel ones by small reduction kernels.  Every g_* pointer is an output, fully overwritten, shaped like the
 * reference parameter (model.py:216-224,280,348; op/fused_act.py:77); `weight` is an input.  Requires up != 1 layers. */
typedef struct sgr_styled_param_grads {
  const float* weight;      /* in: conv.weight[0] [cout,cin,3,3] fp32 */
  float* g_weight;          /* [cout,cin,3,3] */
  float* g_mod_weight;      /* [cin,512] */
  float* g_mod_bias;        /* [cin] */
  float* g_noise_weight;    /* [1] */
  float* g_act_bias;        /* [cout] */
} sgr_styled_param_grads;
typedef struct sgr_rgb_param_grads {
  float* g_weight;          /* [3,cin] */
  float* g_mod_weight;      /* [cin,512] */
  float* g_mod_bias;        /* [cin] */
  float* g_bias;            /* [3] */
} sgr_rgb_param_grads;
typedef struct sgr_param_grads {
  sgr_styled_param_grads styled[SGR_MAX_STYLED];
  sgr_rgb_param_grads rgb[SGR_MAX_RGB];
  float* g_const_input;     /* [cin_0,4,4] */
} sgr_param_grads;
size_t sgr_synthesis_wgrad_scratch_bytes(const sgr_synthesis* net, int batch);

size_t sgr_synthesis_backward_workspace_bytes(const sgr_synthesis* net, int batch);
int sgr_synthesis_backward(const sgr_synthesis* net, const float* latent, int batch, const float* const* feats,
                           const float* grad_image, float* dlatent, void* workspace, size_t workspace_bytes,
                           void* stream);
int sgr_synthesis_backward_ex(const sgr_synthesis* net, const float* latent, int batch, const float* const* feats,
                              const float* grad_image, float* dlatent, void* workspace, size_t workspace_bytes,
                              const sgr_backward_extras* extras, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGR_H_ */
