// extern "C" surface of libsgr.so (include/sgr.h) and the whole-network orchestration.
#include <algorithm>

#include <stdarg.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include "sgr_internal.h"

namespace sgr {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch() { ++g_launches; }
// Programmatic dependent launch of the forward chain (sgr_internal.h launch_pdl, sgr_ptx.cuh pdl_wait).
// Large batches: OFF.  Measured on one B200 (B = 32, 256^2, tools/gpu/call25.sh, same box): 3.04 ms per step without, 3.17 ms
// with it (sustained 3.27 vs 3.43 ms) — the persistent GEMM CTAs hold a whole SM's shared memory and TMEM, so a dependent CTA
// can only become resident when a primary CTA has already exited, and the early-resident CTAs then sit in
// griddepcontrol.wait holding that SM; nothing of the ~3 us prologue is recovered.
// Small batches (<= 4 frames per call, how run_inference.py drives the generator): ON.  Most grids there are smaller than
// the machine, the dependents set themselves up (barrier init, TMEM allocation, descriptor prefetch) on idle SMs while the
// predecessor drains: batch 1 0.429 -> 0.413 ms per frame, batch 2 0.519 -> 0.492 ms (captured graph, same box).
// SGR_PDL=1 / 0 forces it on / off for every batch (results are identical either way).
static const int g_pdl_mode = [] {
  const char* e = getenv("SGR_PDL");
  return !e ? 2 : (e[0] == '1' ? 1 : 0);
}();
static thread_local bool g_pdl_small_batch = false;
bool pdl_enabled() { return g_pdl_mode == 1 || (g_pdl_mode == 2 && g_pdl_small_batch); }
void pdl_small_batch(bool on) { g_pdl_small_batch = on; }
bool check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return false;
  }
  return true;
}

// ---- optional per-launch event timing of the modconv kernel
static bool g_prof_on = false;
static const int kProfCap = 8192;
static cudaEvent_t g_prof_ev[kProfCap][2];
static int g_prof_tag[kProfCap];
static int g_prof_made = 0, g_prof_n = 0;

static int g_prof_layer = 0;          // styled-layer index of the launches being timed (sgr_synthesis_forward sets it)
static bool prof_begin(cudaStream_t st, int tag = 0) {
  if (!g_prof_on || g_prof_n >= kProfCap) return false;
  tag += 16 * g_prof_layer;           // tag & 15: 0 GEMM, 1 FIR pass; tag >> 4: layer
  while (g_prof_made <= g_prof_n) {
    cudaEventCreate(&g_prof_ev[g_prof_made][0]);
    cudaEventCreate(&g_prof_ev[g_prof_made][1]);
    ++g_prof_made;
  }
  g_prof_tag[g_prof_n] = tag;
  cudaEventRecord(g_prof_ev[g_prof_n][0], st);
  return true;
}
static void prof_end(cudaStream_t st) {
  cudaEventRecord(g_prof_ev[g_prof_n][1], st);
  ++g_prof_n;
}

// Side stream for the ToRGB tails: rgb_r = slot sum + bias + 2x FIR(skip_{r-1}) only feeds the next tail and, at the end, the
// image — never a convolution — so the seven small, latency-bound launches (6..9 us each, 2 % of a B = 32 step) leave the
// critical path: forked after the convolution that produced the partial sums, joined once before the call returns (the
// pattern is capturable into a CUDA graph).  One stream + event set per device, created on first use.  SGR_TAIL_STREAM=0: off.
struct TailStream {
  cudaStream_t side = nullptr;
  cudaEvent_t fork[SGR_MAX_RGB] = {};
  cudaEvent_t join = nullptr;
  cudaEvent_t prep_fork = nullptr, prep_join = nullptr;      // styles / tables of the later layers (see sgr_synthesis_forward)
  bool ok = false;
};
static TailStream* tail_stream() {
  static const bool off = [] { const char* e = getenv("SGR_TAIL_STREAM"); return e && e[0] == '0'; }();
  if (off) return nullptr;
  static TailStream per_dev[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  TailStream& t = per_dev[dev];
  if (!t.ok) {
    if (cudaStreamCreateWithFlags(&t.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (int i = 0; i < SGR_MAX_RGB; ++i)
      if (cudaEventCreateWithFlags(&t.fork[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&t.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&t.prep_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&t.prep_join, cudaEventDisableTiming) != cudaSuccess)
      return nullptr;
    t.ok = true;
  }
  return &t;
}

static bool have_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("no CUDA device available: libsgr has no CPU fallback");
    return false;
  }
  return true;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Workspace carve-up for the whole network (all offsets 256-byte aligned).
struct SynthPlan {
  size_t style_off[SGR_MAX_STYLED], rgbstyle_off[SGR_MAX_RGB];
  size_t demod_off[SGR_MAX_STYLED], s2_off[SGR_MAX_STYLED], coef_off[SGR_MAX_STYLED];
  size_t rgbacc_off[SGR_MAX_RGB];
  int rgb_slots[SGR_MAX_RGB];
  size_t t_off;
  size_t splitk_off;
  size_t skip_off[2];
  size_t act_off[2];
  size_t total;
};

static int plan_synthesis(const sgr_synthesis* net, int batch, SynthPlan* pl) {
  if (!net || batch <= 0) {
    set_error("synthesis: null network or batch <= 0");
    return 1;
  }
  if (net->n_styled < 1 || net->n_styled > SGR_MAX_STYLED || net->n_rgb < 1 || net->n_rgb > SGR_MAX_RGB ||
      net->n_styled != 2 * net->n_rgb - 1) {
    set_error("synthesis: inconsistent layer counts n_styled=%d n_rgb=%d", net->n_styled, net->n_rgb);
    return 1;
  }
  size_t off = 0;
  const size_t B = static_cast<size_t>(batch);
  for (int l = 0; l < net->n_styled; ++l) {
    const sgr_styled_layer& L = net->styled[l];
    pl->style_off[l] = off; off = align_up(off + B * L.cin * 4, 256);
    pl->demod_off[l] = off; off = align_up(off + B * L.cout * 4, 256);
    pl->s2_off[l] = off; off = align_up(off + B * L.cout * 4, 256);
    pl->coef_off[l] = off; off = align_up(off + B * 3 * L.cout * 4, 256);
  }
  for (int r = 0; r < net->n_rgb; ++r) {
    pl->rgbstyle_off[r] = off; off = align_up(off + B * net->rgb[r].cin * 4, 256);
  }
  size_t max_act = 0;
  for (int r = 0; r < net->n_rgb; ++r) {
    // one partial-sum slot per column tile of the conv that feeds this ToRGB (layer 0, 2, 4, ...)
    const sgr_styled_layer& L = net->styled[r == 0 ? 0 : 2 * r];
    const int nt = L.column_tile > 0 ? L.column_tile : pick_nt(L.cout);
    pl->rgb_slots[r] = (L.cout + nt - 1) / nt;
    const size_t res = static_cast<size_t>(4) << r;
    pl->rgbacc_off[r] = off; off = align_up(off + pl->rgb_slots[r] * B * 3 * res * res * 4, 256);
  }
  // parity planes of the scatter up-convs: [B][4][cout/8][H+1][W+1][8] fp32
  size_t max_t = 0;
  for (int l = 0; l < net->n_styled; ++l)
    if (net->styled[l].up == 2) {
      const size_t res_in = static_cast<size_t>(4) << ((l + 1) / 2 - 1);
      const size_t e = B * 4 * net->styled[l].cout * (res_in + 1) * (res_in + 1) * 4;
      if (e > max_t) max_t = e;
    }
  pl->t_off = off; off = align_up(off + max_t, 256);
  pl->splitk_off = off; off = align_up(off + kSplitKScratchBytes, 256);
  const size_t S = static_cast<size_t>(net->size);
  for (int i = 0; i < 2; ++i) {
    pl->skip_off[i] = off; off = align_up(off + B * 3 * S * S * 4, 256);
  }
  // activation ping-pong: hi+lo bf16 = 4 bytes per element; the input of layer 0 is [B,cin,4,4]
  max_act = B * net->styled[0].cin * 16;
  for (int l = 0; l + 1 < net->n_styled; ++l) {      // the last layer's activation never leaves the chip
    const size_t res = static_cast<size_t>(4) << ((l + 1) / 2);
    const size_t e = B * net->styled[l].cout * res * res;
    if (e > max_act) max_act = e;
  }
  for (int i = 0; i < 2; ++i) {
    pl->act_off[i] = off; off = align_up(off + max_act * 4, 256);
  }
  pl->total = off;
  return 0;
}

}  // namespace sgr

using namespace sgr;

extern "C" {

const char* sgr_version(void) { return "sgr 0.1 (sm_100a: tcgen05/TMEM/TMA modconv, bf16x3)"; }
const char* sgr_last_error(void) { return g_err; }
long long sgr_launch_count(void) { return g_launches; }
void sgr_reset_launch_count(void) { g_launches = 0; }

int sgr_upfirdn2d(const float* x, float* y, const float* taps, int planes, int in_h, int in_w, int up, int down,
                  int pad0, int pad1, int kh, int kw, void* stream) {
  if (!have_device()) return 1;
  if (!x || !y || !taps) {
    set_error("upfirdn2d: null pointer");
    return 1;
  }
  return upfirdn2d_launch(x, y, taps, planes, in_h, in_w, up, down, pad0, pad1, kh, kw,
                          static_cast<cudaStream_t>(stream));
}

int sgr_fused_bias_act(const float* x, const float* bias, const float* ref, float* y, long long outer, int channels,
                       long long inner, int grad, float slope, float scale, void* stream) {
  if (!have_device()) return 1;
  if (!x || !y) {
    set_error("fused_bias_act: null pointer");
    return 1;
  }
  return bias_act_launch(x, bias, ref, y, outer, channels, inner, grad, slope, scale, static_cast<cudaStream_t>(stream));
}

int sgr_frames_to_uint8(const float* frames, unsigned char* out, int batch, int h, int w, int out_h, int out_w, void* stream) {
  if (!have_device()) return 1;
  if (!frames || !out) {
    set_error("frames_to_uint8: null pointer");
    return 1;
  }
  return frames_to_uint8_launch(frames, out, batch, h, w, out_h, out_w, static_cast<cudaStream_t>(stream));
}

int sgr_choose_column_tile(int batch, int h_in, int w_in, int n_total) {
  if (batch <= 0 || h_in <= 0 || w_in <= 0 || n_total < 32) return 0;
  return choose_nt(batch, h_in, w_in, n_total);
}

size_t sgr_packed_weight_bytes(int cout, int cin, int ksize, int up, int transpose) {
  if (up == 2) return static_cast<size_t>(cout) * cin * 9 * 2 * 2;   // scatter / its gather adjoint: the 9 real taps
  const size_t n_total = transpose ? cin : static_cast<size_t>(cout) * (up ? 4 : 1);
  const size_t k_total = transpose ? static_cast<size_t>(cout) * (up ? 4 : 1) : cin;
  return n_total * k_total * ksize * ksize * 2 /*planes*/ * 2 /*bf16*/;
}

size_t sgr_up_scratch_bytes(int batch, int cout, int h_in, int w_in) {
  return static_cast<size_t>(batch) * 4 * cout * (h_in + 1) * (w_in + 1) * 4;
}

int sgr_pack_modconv_weight(const float* weight, const float* fir, int cout, int cin, int ksize, int up, int transpose,
                            int format, int column_tile, void* packed, float* wsq, void* stream) {
  if (!have_device()) return 1;
  const int n_total = transpose ? cin : cout * (up ? 4 : 1);
  const int k_total = transpose ? cout * (up == 1 ? 4 : 1) : cin;
  if (up == 2 && !transpose) {
    if (!weight || !packed || ksize != 3 || cin % kBlockK != 0 || cout < 32 || (cout & (cout - 1)) != 0 ||
        (format != SGR_FMT_BF16 && format != SGR_FMT_FP16) || (column_tile != 0 && column_tile != up2_nt(cout))) {
      set_error("pack_modconv_weight: unsupported scatter up-conv cout=%d cin=%d k=%d tile=%d", cout, cin, ksize,
                column_tile);
      return 1;
    }
    return pack_weight_launch(weight, fir, cout, cin, ksize, 2, 0, format, column_tile, packed, wsq,
                              static_cast<cudaStream_t>(stream));
  }
  if (!weight || !packed || (up == 1 && !fir) || (ksize != 3 && ksize != 1) || (up && ksize != 3) ||
      k_total % kBlockK != 0 || n_total < 32 || (n_total & (n_total - 1)) != 0 ||
      (format != SGR_FMT_BF16 && format != SGR_FMT_FP16) ||
      (column_tile != 0 && (column_tile > n_total || n_total % column_tile != 0 ||
                            (column_tile != 32 && column_tile != 64 && column_tile != 128 && column_tile != 256)))) {
    set_error("pack_modconv_weight: unsupported cout=%d cin=%d k=%d up=%d transpose=%d", cout, cin, ksize, up,
              transpose);
    return 1;
  }
  return pack_weight_launch(weight, fir, cout, cin, ksize, up, transpose, format, column_tile, packed, wsq,
                            static_cast<cudaStream_t>(stream));
}

int sgr_nchw_to_c8(const float* x, const float* scale, void* out_c8, int batch, int channels, int h, int w, int s2d,
                   int format, void* stream) {
  if (!have_device()) return 1;
  if (!x || !out_c8 || channels % 8 != 0 || (s2d && ((h | w) & 1))) {
    set_error("nchw_to_c8: bad arguments (C=%d H=%d W=%d)", channels, h, w);
    return 1;
  }
  return nchw_to_c8_launch(x, scale, out_c8, batch, channels, h, w, s2d, format, static_cast<cudaStream_t>(stream));
}

// plane scales of the FIR pass of an up == 2 layer: undo the operand scales, compensate the accumulate truncation
static void up_plane_scales(const sgr_conv_args* args, float* base, float* comp) {
  *base = 1.f / (act_scale(args->operand_format) * w_scale(args->operand_format));
  *comp = acc_comp_enabled() ? 1.16e-8f * (args->single_pass ? 1.f : 3.f) * static_cast<float>(args->cin / 16) : 0.f;
}

// The FIR pass of up-layer `up` expressed as the producer of the following halo convolution (fir_producer.cuh).
static void fill_fused_fir(const sgr_conv_args* up, FusedFirParams* f) {
  memset(f, 0, sizeof(*f));
  float base, comp;
  up_plane_scales(up, &base, &comp);
  const int taps[4] = {2, 4, 2, 1};                  // MMA chains accumulated per plane: oe, ee, eo, oo (up_finish_launch)
  for (int i = 0; i < 4; ++i) f->plane_scale[i] = base * (1.f + comp * taps[i]);
  f->t = up->t_scratch; f->fir = up->fir;
  f->C = up->cout; f->Hin = up->h_in; f->Win = up->w_in;
  f->demod = up->demod; f->bias = up->bias;
  f->noise = up->noise; f->noise_bstride = up->noise_batch_stride; f->noise_w = up->noise_weight;
  f->s2 = up->s2; f->act = up->act; f->act_gain = up->act_gain; f->out_scale = act_scale(up->out_format);
}

// Split-K of the scatter up-conv (modconv_scatter_sm100.cu): with fewer tiles than half the SMs (batch 1-4, the 5^2 .. 33^2
// grids: 8 .. 36 CTAs walking all channel blocks one after the other, ~25 us each at batch 1) the channel blocks are cut into
// up to 8 slices; every slice writes its own copy of the raw parity planes into the split-K scratch and the FIR pass
// (up_finish_kernel) adds the copies on load, in slice order.  Not combined with the fused FIR producers, which read the
// planes by TMA.  SGR_UP_SPLITK=0 disables it, any other number caps the slice count.
static int scatter_ksplit(const sgr_conv_args* a, const ConvKernelParams& p) {
  static const int cap = [] {
    const char* e = getenv("SGR_UP_SPLITK");
    const int v = e ? atoi(e) : 8;
    return v < 1 ? 1 : (v > 8 ? 8 : v);          // up_finish_kernel adds at most 8 copies
  }();
  if (cap < 2 || !a->splitk_scratch || p.debug) return 1;
  const int sms = num_sms();
  const int tiles = p.m_tiles * p.n_tiles;
  if (sms <= 0 || tiles <= 0 || 2 * tiles > sms) return 1;
  int s = std::min(cap, sms / tiles);
  s = std::min(s, p.kchunks / 2);                    // at least two 32-channel blocks per slice
  const size_t plane_bytes = static_cast<size_t>(a->batch) * 4 * a->cout * (a->h_in + 1) * (a->w_in + 1) * 4;
  while (s > 1 && s * plane_bytes > a->splitk_scratch_bytes) --s;
  while (s & (s - 1)) --s;                           // 1, 2, 4 or 8 copies (up_finish_kernel instantiations)
  return s < 2 ? 1 : s;
}

// defer_fir: (up == 2) run the scatter GEMM only; the consumer applies the FIR pass (fused_src = that layer's arguments)
static int modconv_forward_impl(const sgr_conv_args* args, const sgr_conv_args* fused_src, bool defer_fir, void* stream) {
  if (!have_device()) return 1;
  ConvKernelParams p;
  int nt = 0;
  if (conv_fill_params(args, &p, &nt)) return 1;
  int rc;
  if (halo_eligible(args)) {           // wide 3x3 layers: resident halo tile (modconv_halo_sm100.cu)
    FusedFirParams f;
    if (fused_src) fill_fused_fir(fused_src, &f);
    const bool prof = prof_begin(static_cast<cudaStream_t>(stream));
    rc = launch_modconv_halo(args, p, static_cast<cudaStream_t>(stream), fused_src ? &f : nullptr);
    if (prof) prof_end(static_cast<cudaStream_t>(stream));
    return rc;
  }
  if (fused_src) {
    set_error("modconv_forward: fused FIR input needs a halo-eligible layer");
    return 1;
  }
  CUtensorMap tmap;
  if (args->up == 3) {       // gather adjoint: the operand is the 4-plane tensor on the (h+1) x (w+1) grid
    if (make_act_tensor_map(&tmap, args->x_c8, args->batch, 4 * args->cin, args->h_in + 1, args->w_in + 1, p.bw, p.bh, p.bb,
                            p.single ? 1 : 2))
      return 1;
  } else if (make_act_tensor_map(&tmap, args->x_c8, args->batch, args->cin, args->h_in, args->w_in, p.bw, p.halo ? p.box_rows : p.bh,
                                 p.bb, p.single ? 1 : 2, kBlockK / 8, p.halo == 2)) {
    return 1;
  }
  int up_slices = 1;
  if (args->up != 2) {
    set_ksplit(&p, choose_ksplit(args, p.m_tiles * p.n_tiles, p.ntaps * p.kchunks, 8, static_cast<size_t>(kTileM) * nt * 4));
  } else if (!defer_fir) {
    up_slices = scatter_ksplit(args, p);
    if (up_slices > 1) {                 // (p.acc_scale is not used by the scatter kernel: up_finish applies the plane scales)
      p.ksplit = up_slices;
      p.t_out = static_cast<float*>(args->splitk_scratch);
    }
  }
  const bool prof = prof_begin(static_cast<cudaStream_t>(stream));
  rc = args->up == 2 ? launch_upconv_scatter(p, tmap, nt, static_cast<cudaStream_t>(stream))
                     : launch_modconv(p, tmap, nt, static_cast<cudaStream_t>(stream));
  if (rc == 0 && args->up != 2 && p.ksplit > 1) rc = splitk_finish_launch(p, static_cast<cudaStream_t>(stream));
  if (prof) prof_end(static_cast<cudaStream_t>(stream));
  if (rc == 0 && args->up == 2 && !defer_fir) {
    float base, comp;
    up_plane_scales(args, &base, &comp);
    comp /= static_cast<float>(up_slices);           // each slice's accumulation chain is 1 / up_slices as long
    const bool prof2 = prof_begin(static_cast<cudaStream_t>(stream), 1);
    rc = up_finish_launch(args, base, comp, static_cast<cudaStream_t>(stream), up_slices > 1 ? p.t_out : nullptr, up_slices);
    if (prof2) prof_end(static_cast<cudaStream_t>(stream));
  }
  return rc;
}

int sgr_modconv_forward(const sgr_conv_args* args, void* stream) { return modconv_forward_impl(args, nullptr, false, stream); }

void sgr_profile_enable(int on) { g_prof_on = on != 0; }

int sgr_profile_collect_tagged(float* ms, int* tags, int cap) {
  const int n = g_prof_n;
  for (int i = 0; i < n; ++i) {
    float t = 0.f;
    cudaEventSynchronize(g_prof_ev[i][1]);
    cudaEventElapsedTime(&t, g_prof_ev[i][0], g_prof_ev[i][1]);
    if (ms && i < cap) ms[i] = t;
    if (tags && i < cap) tags[i] = g_prof_tag[i];
  }
  g_prof_n = 0;
  return n;
}

int sgr_profile_collect(float* ms, int cap) { return sgr_profile_collect_tagged(ms, nullptr, cap); }

int sgr_style_affine(const float* latent, int latent_stride, int batch, const float* mod_weight, const float* mod_bias,
                     int cin, float* s_out, void* stream) {
  if (!have_device()) return 1;
  StyleJobs jobs;
  jobs.n = 1;
  jobs.job[0] = StyleJob{mod_weight, mod_bias, s_out, cin, 0};
  return style_jobs_launch(jobs, latent, latent_stride, batch, static_cast<cudaStream_t>(stream));
}

int sgr_demod(const float* s, const float* wsq, int batch, int cin, int cout, float* d_out, void* stream) {
  if (!have_device()) return 1;
  return demod_launch(s, wsq, batch, cin, cout, d_out, static_cast<cudaStream_t>(stream));
}

size_t sgr_synthesis_workspace_bytes(const sgr_synthesis* net, int batch) {
  SynthPlan pl;
  if (plan_synthesis(net, batch, &pl)) return 0;
  return pl.total;
}

int sgr_synthesis_forward(const sgr_synthesis* net, const float* latent, int batch, float* image, void* workspace,
                          size_t workspace_bytes, float* const* feats, void* stream) {
  return sgr_synthesis_forward_ex(net, latent, batch, image, workspace, workspace_bytes, feats, nullptr, stream);
}

int sgr_synthesis_forward_ex(const sgr_synthesis* net, const float* latent, int batch, float* image, void* workspace,
                             size_t workspace_bytes, float* const* feats, const sgr_forward_extras* extras, void* stream) {
  if (!have_device()) return 1;
  SynthPlan pl;
  if (plan_synthesis(net, batch, &pl)) return 1;
  struct PdlScope {                                    // programmatic dependent launch for small-batch calls (pdl_enabled())
    explicit PdlScope(bool on) { pdl_small_batch(on); }
    ~PdlScope() { pdl_small_batch(false); }
  } pdl_scope(batch <= 4);
  unsigned char* frames_u8 = extras ? extras->frames_u8 : nullptr;
  if (frames_u8 && (extras->u8_h <= 0 || extras->u8_w <= 0 || net->size % extras->u8_h != 0 || net->size % extras->u8_w != 0)) {
    set_error("synthesis_forward: uint8 frame size %dx%d must divide the network size %d", extras->u8_h, extras->u8_w, net->size);
    return 1;
  }
  if (!latent || (!image && !frames_u8) || !workspace || workspace_bytes < pl.total) {
    set_error("synthesis_forward: workspace too small (%zu < %zu) or null pointer", workspace_bytes, pl.total);
    return 1;
  }
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) {
    set_error("synthesis_forward: workspace must be 256-byte aligned");
    return 1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  const int latent_stride = net->n_latent * SGR_STYLE_DIM;

  // 1 + 2. style vectors, demodulation and epilogue tables (next layer's style, fused ToRGB coefficients).  The first kSplit
  // layers (4^2, 8^2: latency-bound GEMMs that leave the SMs mostly idle) only need their own share; the rest of the
  // modulation GEMVs / table GEMMs (~80 % of the work, ~35 us at B = 32) runs on the side stream underneath them and is joined
  // before layer kSplit.
  TailStream* tails = g_prof_on ? nullptr : tail_stream();      // per-launch event timing keeps everything on one stream
  constexpr int kSplit = 4;
  const bool split_prep = tails != nullptr && net->n_styled > kSplit + 1;
  const int n_first = split_prep ? kSplit : net->n_styled;      // tables of layers [0, n_first) come first
  StyleJobs sj[2];
  sj[0].n = sj[1].n = 0;
  for (int l = 0; l < net->n_styled; ++l) {
    const sgr_styled_layer& L = net->styled[l];
    StyleJobs& d = (l <= n_first) ? sj[0] : sj[1];                // table l needs style l and style l + 1
    d.job[d.n++] = StyleJob{L.mod_weight, L.mod_bias, F(pl.style_off[l]), L.cin, L.latent_row};
  }
  for (int r = 0; r < net->n_rgb; ++r) {
    const sgr_rgb_layer& R = net->rgb[r];
    StyleJobs& d = (2 * r <= n_first) ? sj[0] : sj[1];            // ToRGB r hangs off styled layer 2r (layer 0 for r = 0)
    d.job[d.n++] = StyleJob{R.mod_weight, R.mod_bias, F(pl.rgbstyle_off[r]), R.cin, R.latent_row};
  }
  TableJobs tj[2];
  tj[0].n = tj[1].n = 0;
  for (int l = 0; l < net->n_styled; ++l) {
    const sgr_styled_layer& L = net->styled[l];
    TableJobs& d = l < n_first ? tj[0] : tj[1];
    TableJob& j = d.job[d.n++];
    j.s = F(pl.style_off[l]);
    j.wsq = L.wsq;
    j.demod = F(pl.demod_off[l]);
    j.cin = L.cin;
    j.cout = L.cout;
    const bool has_next = l + 1 < net->n_styled;
    j.s_next = has_next ? F(pl.style_off[l + 1]) : nullptr;
    j.s2 = F(pl.s2_off[l]);
    const bool has_rgb = !L.up;                 // conv1 and every odd convs.* feed a ToRGB
    const int r = (l + 1) / 2;
    j.s_rgb = has_rgb ? F(pl.rgbstyle_off[r]) : nullptr;
    j.w_rgb = has_rgb ? net->rgb[r].weight : nullptr;
    j.rgb_coef = F(pl.coef_off[l]);
    if (has_next && net->styled[l + 1].cin != L.cout) {
      set_error("synthesis_forward: layer %d cout %d != layer %d cin %d", l, L.cout, l + 1, net->styled[l + 1].cin);
      return 1;
    }
    if (has_rgb && net->rgb[r].cin != L.cout) {
      set_error("synthesis_forward: rgb %d cin mismatch", r);
      return 1;
    }
  }
  if (style_jobs_launch(sj[0], latent, latent_stride, batch, st)) return 1;
  bool prep_pending = false;
  if (split_prep && sj[1].n + tj[1].n > 0) {
    if (cudaEventRecord(tails->prep_fork, st) != cudaSuccess || cudaStreamWaitEvent(tails->side, tails->prep_fork, 0) != cudaSuccess) {
      set_error("synthesis_forward: forking the side stream failed");
      return 1;
    }
    if (sj[1].n > 0 && style_jobs_launch(sj[1], latent, latent_stride, batch, tails->side)) return 1;
    if (tj[1].n > 0 && table_jobs_launch(tj[1], batch, tails->side)) return 1;
    if (cudaEventRecord(tails->prep_join, tails->side) != cudaSuccess) {
      set_error("synthesis_forward: side stream event failed");
      return 1;
    }
    prep_pending = true;
  }
  if (table_jobs_launch(tj[0], batch, st)) return 1;

  // 3. modulated constant input
  int cur = 0;
  if (const_input_launch(net->const_input, F(pl.style_off[0]), batch, net->styled[0].cin, net->format,
                         ws + pl.act_off[cur], st))
    return 1;

  // 4. the layer chain
  bool forked = false;
  int res = 4;
  int skip_cur = 0;
  const float* prev_skip = nullptr;
  sgr_conv_args pending_up;                 // an up layer whose FIR pass the next convolution applies (fused producer)
  bool have_pending = false;
  for (int l = 0; l < net->n_styled; ++l) {
    const sgr_styled_layer& L = net->styled[l];
    const bool last = l + 1 == net->n_styled;
    sgr_conv_args a;
    memset(&a, 0, sizeof(a));
    a.batch = batch;
    a.cin = L.cin;
    a.cout = L.cout;
    a.h_in = res;
    a.w_in = res;
    a.ksize = 3;
    a.up = L.up;
    a.act = 1;
    a.act_gain = 1.4142135623730951f;
    a.operand_format = net->format;
    a.single_pass = net->single_pass;
    a.splitk_scratch = ws + pl.splitk_off;
    a.splitk_scratch_bytes = kSplitKScratchBytes;
    a.column_tile = L.column_tile;
    a.out_format = net->format;
    a.x_c8 = ws + pl.act_off[cur];
    a.w_packed = L.w_packed;
    a.demod = F(pl.demod_off[l]);
    a.bias = L.act_bias;
    a.noise = L.noise;
    a.noise_batch_stride = L.noise_batch_stride;
    a.noise_weight = L.noise_weight;
    a.s2 = last ? nullptr : F(pl.s2_off[l]);
    a.out_c8 = last ? nullptr : ws + pl.act_off[1 - cur];
    a.out_f32 = feats ? feats[l] : nullptr;
    const int r = (l + 1) / 2;
    if (!L.up) {
      a.rgb_coef = F(pl.coef_off[l]);
      a.rgb_partial = F(pl.rgbacc_off[r]);
    } else if (L.up == 2) {
      if (!L.fir) {
        set_error("synthesis_forward: layer %d (scatter up-conv) lacks its blur kernel", l);
        return 1;
      }
      a.fir = L.fir;
      a.t_scratch = F(pl.t_off);
    }
    bool defer = false;
    if (L.up == 2 && !last && !a.out_f32 && net->styled[l + 1].up == 0) {
      sgr_conv_args nx;                      // the consumer, as far as halo_fusable() looks at it
      memset(&nx, 0, sizeof(nx));
      nx.batch = batch; nx.cin = L.cout; nx.cout = net->styled[l + 1].cout; nx.h_in = nx.w_in = 2 * res; nx.ksize = 3;
      nx.single_pass = net->single_pass;
      nx.column_tile = net->styled[l + 1].column_tile;
      defer = halo_fusable(&nx);
    }
    if (prep_pending && l == n_first) {                      // the tables of layers >= kSplit come from the side stream
      if (cudaStreamWaitEvent(st, tails->prep_join, 0) != cudaSuccess) {
        set_error("synthesis_forward: joining the side stream failed");
        return 1;
      }
      prep_pending = false;
      forked = true;                                         // (the side stream has work of this call: join it at the end too)
    }
    g_prof_layer = l;
    const int rc_layer = modconv_forward_impl(&a, have_pending ? &pending_up : nullptr, defer, stream);
    g_prof_layer = 0;
    if (rc_layer) return 1;
    have_pending = defer;
    if (defer) pending_up = a;
    if (L.up) res *= 2;
    cur = 1 - cur;
    if (!L.up) {
      const sgr_rgb_layer& R = net->rgb[r];
      float* dst = last ? image : F(pl.skip_off[skip_cur]);
      if (prev_skip && !R.fir) {
        set_error("synthesis_forward: rgb %d needs an upsample kernel", r);
        return 1;
      }
      cudaStream_t ts = st;
      if (tails && cudaEventRecord(tails->fork[r], st) == cudaSuccess &&
          cudaStreamWaitEvent(tails->side, tails->fork[r], 0) == cudaSuccess) {
        ts = tails->side;
        forked = true;
      }
      if (last && frames_u8 &&
          torgb_tail_u8_launch(F(pl.rgbacc_off[r]), pl.rgb_slots[r], R.bias, prev_skip, R.fir, frames_u8, batch, res, res,
                               extras->u8_h, extras->u8_w, ts))
        return 1;
      if ((!last || image) &&
          torgb_tail_launch(F(pl.rgbacc_off[r]), pl.rgb_slots[r], R.bias, prev_skip, R.fir, dst, batch, res, res, ts))
        return 1;
      prev_skip = dst;
      skip_cur = 1 - skip_cur;
    }
  }
  if (forked && (cudaEventRecord(tails->join, tails->side) != cudaSuccess ||
                 cudaStreamWaitEvent(st, tails->join, 0) != cudaSuccess)) {
    set_error("synthesis_forward: joining the ToRGB side stream failed");
    return 1;
  }
  if (res != net->size) {
    set_error("synthesis_forward: layer chain ends at %d, expected %d", res, net->size);
    return 1;
  }
  return 0;
}

}  // extern "C"
