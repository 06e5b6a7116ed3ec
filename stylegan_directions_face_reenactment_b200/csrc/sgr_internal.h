// Internal declarations shared by the .cu files of libsgr.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/sgr.h"

namespace sgr {

constexpr int kTileM = 128;      // GEMM rows (pixels) per CTA tile == TMEM lanes
constexpr int kBlockK = 32;      // input channels per pipeline stage (4 chunks of 8)
constexpr int kABytes = 2 * kTileM * kBlockK * 2;   // hi+lo planes of one A stage = 16 KiB

// Static scales of the fp16 operand format (SGR_FMT_FP16): activations are stored x 2^-4 (range 1e6, absolute
// resolution 5e-7), weights x 2^8 (keeps the lo halves of 1/sqrt(fan_in)-sized weights out of the fp16 subnormals).
constexpr float kActScaleFP16 = 0.0625f;
constexpr float kWScaleFP16 = 256.f;
inline float act_scale(int fmt) { return fmt == SGR_FMT_FP16 ? kActScaleFP16 : 1.f; }
inline float w_scale(int fmt) { return fmt == SGR_FMT_FP16 ? kWScaleFP16 : 1.f; }

// GEMM column tile for a layer with n_total columns.
inline int pick_nt(int n_total) { return n_total >= 256 ? 256 : n_total; }
// Scatter up-conv: 4 parity blocks of NT/4 output channels per column tile.
inline int up2_nt(int cout) { return cout >= 64 ? 256 : 4 * cout; }
int choose_nt(int batch, int h, int w, int n_total);
void tile_box(int h, int w, int* bw, int* bh, int* bb);
void tile_box_search(int batch, int h, int w, int* bw, int* bh, int* bb);

void set_error(const char* fmt, ...);
bool pdl_enabled();          // programmatic dependent launch: small-batch forward calls, or SGR_PDL=1 (sgr_api.cu)
void pdl_small_batch(bool on);

// Launch with programmatic stream serialization: the kernel may start while its predecessor drains (sgr_ptx.cuh pdl_wait()).
// ONLY for kernels that call pdl_wait() before their first dependent global access.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
void count_launch();
bool check_launch(const char* what);

struct ConvKernelParams {
  int B, H, W;                  // pixel space of the implicit GEMM (input resolution)
  int debug;                    // experiment switches (SGR_DEBUG env)
  int mode;                     // 0 plain / 1x1, 1 polyphase up-conv, 2 scatter up-conv (raw parity planes -> t_out)
  int bw, bh, bb;               // tile box (rows = bw*bh*bb <= 128)
  int rows;
  int tiles_x, tiles_y, tiles_b;
  int m_tiles, n_tiles;
  int kchunks;                  // cin / 32
  int ntaps;                    // 9 or 1
  signed char tap_dy[9], tap_dx[9];   // pixel shift of tap t
  int tap_chunk[9];             // first 8-channel chunk of tap t's operand (parity planes stacked as channels)
  int cout;                     // real output channels (power of two)
  int up;
  int Hout, Wout;
  int act;
  float act_gain;
  int fmt;                      // operand format of x_c8 / wpacked (SGR_FMT_*)
  int single;                   // 1: hi x hi products only (plain bf16/fp16 tensor-core precision, 1 MMA instead of 3);
                                //    only the hi plane of x_c8 / hi half of the weight slabs is loaded
  int single_out;               // 1: out_c8 consumers are single-pass too: the lo plane is not written
  float acc_scale;              // undoes the operand scales on the accumulator
  int out_fmt;                  // format of out_c8
  float out_scale;              // activation scale of out_fmt
  const __nv_bfloat16* wpacked;
  const float* demod;
  const float* bias;
  const float* noise;
  const float* noise_w;
  long long noise_bstride;       // 0: one noise map shared by the batch
  const float* s2;
  __nv_bfloat16* out_c8;
  float* out_f32;
  const float* rgb_coef;
  float* rgb_part;              // [n_tiles][B,3,H,W] partial ToRGB sums, one slot per column tile
  float* t_out;                 // mode 2: [B][4][cout/4][H][W][4] fp32 parity planes (H, W = grid dims above)
  // split-K (small pixel grids / small batches: fewer tiles than SMs): the K range of a tile is cut into `ksplit` slices
  // run by different CTAs; raw fp32 partials go to kpart[ks][n_tile][m_tile * sub + s][128][NT] and
  // splitk_finish_kernel adds them in slice order (deterministic) and applies the fused epilogue
  int ksplit;
  int halo_mt;                  // 0: modconv_kernel row -> pixel mapping; 1, 2: modconv_halo_kernel sub-tiles per tile
  int nt;
  float* kpart;
  float acc_base, acc_mmas;     // acc_scale = acc_base * (1 + 1.16e-8 * acc_mmas / ksplit)
  int tma_store;                // scatter up-conv: the epilogue stages 32 columns in shared memory and stores them with one TMA box
  // scatter up-conv, wrapped-halo tiles (modconv_scatter_sm100.cu): ONE dense TMA box of bw x (bh + 1) pixels per channel block
  // whose first row / column are the halo; the four shifts are descriptor offsets into it.  bw - 1 valid columns, bh valid rows.
  int halo;
  int box_rows;                 // rows of that TMA box (bw * box_rows = 144 or 140 entries: fixed chunk strides in the kernel)
};

// FIR pass of the preceding scatter up-conv folded into a halo convolution's producer warps (fir_producer.cuh)
struct FusedFirParams {
  const float* t;               // [B][4 (oe,ee,eo,oo)][C/4][Hin+1][Win+1][4] fp32 parity planes
  const float* fir;             // [4][4] blur.kernel
  int C, Hin, Win;              // the up layer's output channels (= this conv's cin) and INPUT resolution
  float plane_scale[4];
  const float* demod;           // [B,C] of the up layer
  const float* bias;            // [C] or NULL
  const float* noise;           // [2Hin,2Win] (+ batch stride) or NULL
  long long noise_bstride;
  const float* noise_w;
  const float* s2;              // [B,C]: sqrt2 * this conv's style
  int act;
  float act_gain, out_scale;
};

// modconv_sm100.cu
int launch_modconv(const ConvKernelParams& p, const CUtensorMap& tmap, int nt, cudaStream_t stream);
// host: build the 5-D tensor map over C8 activation planes
int make_act_tensor_map(CUtensorMap* map, const void* base, int batch, int channels, int h, int w, int bw, int bh,
                        int bb, int planes = 2, int chunk_box = kBlockK / 8, bool wide = false);
int conv_fill_params(const sgr_conv_args* a, ConvKernelParams* p, int* nt);
// 5-D map over fp32 parity planes [B][4][C/4][Hp][Wp][4]: box = (cols x 4 floats, rows, groups of 4 channels, planes, 1 sample);
// the defaults are the window the fused FIR producers load, the scatter GEMM stores (cols x rows, 8 groups, 1 plane) boxes
int make_plane_tensor_map(CUtensorMap* map, const float* base, int batch, int channels, int hp, int wp, int cols,
                          int rows = 12, int groups = 4, int planes = 4);
int num_sms();
// modconv_scatter_sm100.cu: scatter-form upsampling convolution (parity planes -> p.t_out)
int launch_upconv_scatter(const ConvKernelParams& p, const CUtensorMap& tmap, int nt, cudaStream_t stream);
// modconv_halo_sm100.cu: resident-halo variant for plain 3x3 layers of at least 16x16 pixels
bool halo_eligible(const sgr_conv_args* a);
// split-K: slices for a layer with `tiles` output tiles and `k_units` K units (0/1 = off); finish pass
constexpr size_t kSplitKScratchBytes = 160u * 128u * 1024u;      // >= 148 CTA tiles x 128 rows x 256 columns x 4 B
int choose_ksplit(const sgr_conv_args* a, int tiles, int k_units, int min_units_per_slice, size_t tile_bytes);
void set_ksplit(ConvKernelParams* p, int ksplit);
int splitk_finish_launch(const ConvKernelParams& p, cudaStream_t stream);
int launch_modconv_halo(const sgr_conv_args* a, ConvKernelParams p, cudaStream_t stream, const FusedFirParams* fused = nullptr);
bool halo_fusable(const sgr_conv_args* a);

// wgrad_sm100.cu: weight-gradient GEMM (K = pixels, MN-major operands)
size_t wgrad_scratch_bytes(int cout, int cin);
struct WgradFinish {      // optional fused tail: demodulation term + the EqualLR scale -> gradient of conv.weight itself
  const float* weight;    // [cout,cin,3,3] fp32 parameter (NULL: raw convolution term)
  const float* q;         // [B,cout]
  const float* demod;     // [B,cout]
  const float* style;     // [B,cin]
  int batch;
  float scale;            // 1/sqrt(cin*9)
};
int wgrad_launch(const sgr_wgrad_args* a, const WgradFinish* finish, cudaStream_t stream);

// prep_kernels.cu
struct StyleJob {
  const float* mod_weight;   // [cin, 512]
  const float* mod_bias;     // [cin]
  float* out;                // [B, cin]
  int cin;
  int latent_row;
};
struct StyleJobs {
  StyleJob job[SGR_MAX_STYLED + SGR_MAX_RGB];
  int n;
};
struct TableJob {
  const float* s;        // [B, cin]   this layer's style
  const float* wsq;      // [cout, cin]
  float* demod;          // [B, cout]
  const float* s_next;   // [B, cout]  next conv's style or NULL
  float* s2;             // [B, cout]  sqrt2 * s_next
  const float* s_rgb;    // [B, cout]  following ToRGB's style or NULL
  const float* w_rgb;    // [3, cout]
  float* rgb_coef;       // [B, 3, cout]
  int cin, cout;
};
struct TableJobs {
  TableJob job[SGR_MAX_STYLED];
  int n;
};
int pack_weight_launch(const float* w, const float* fir, int cout, int cin, int ks, int up, int transpose, int fmt,
                       int nt, void* packed, float* wsq, cudaStream_t st);
int style_jobs_launch(const StyleJobs& jobs, const float* latent, int latent_stride, int batch, cudaStream_t st);
int table_jobs_launch(const TableJobs& jobs, int batch, cudaStream_t st);
int demod_launch(const float* s, const float* wsq, int batch, int cin, int cout, float* d, cudaStream_t st);
int nchw_to_c8_launch(const float* x, const float* scale, void* out, int batch, int C, int H, int W, int s2d, int fmt,
                      cudaStream_t st);
int const_input_launch(const float* cinput, const float* s, int batch, int C, int fmt, void* out, cudaStream_t st);
// upfirdn2d_sm100.cu
int upfirdn2d_launch(const float* x, float* y, const float* taps, int planes, int in_h, int in_w, int up, int down,
                     int pad0, int pad1, int kh, int kw, cudaStream_t st);
int bias_act_launch(const float* x, const float* bias, const float* ref, float* y, long long outer, int channels,
                    long long inner, int grad, float slope, float scale, cudaStream_t st);
int torgb_tail_launch(const float* rgb_acc, int slots, const float* bias, const float* skip_in, const float* fir,
                      float* out, int batch, int H, int W, cudaStream_t st);
int torgb_tail_u8_launch(const float* rgb_acc, int slots, const float* bias, const float* skip_in, const float* fir,
                         unsigned char* out, int batch, int H, int W, int out_h, int out_w, cudaStream_t st);
int frames_to_uint8_launch(const float* x, unsigned char* y, int batch, int H, int W, int out_h, int out_w, cudaStream_t st);
// up_finish_sm100.cu: FIR + fused epilogue over the parity planes of a scatter up-conv
// planes / nslices: the plane tensor when it is not a->t_scratch (split-K scatter GEMM: nslices copies, added on load)
int up_finish_launch(const sgr_conv_args* a, float acc_scale, float comp_per_tap, cudaStream_t st, const float* planes = nullptr,
                     int nslices = 1);
bool acc_comp_enabled();

}  // namespace sgr
