/*
 * viai_b200.h -- C ABI of libviai_b200.so: the sm_100a kernels behind the VIAI hot path.
 *
 * The reference (Hangz-nju-cuhk/Vision-Infused-Audio-Inpainter-VIAI) is pure Python; its seam for this
 * path is the nn.Module API.  There is no FFI to mirror, so every entry point below names
 * the ATen operator call site in the reference that it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - All tensors are dense, fp32 unless stated, activation layout is NHWC ("rows x C").
 *   - The caller owns every buffer; kernels borrow raw device pointers for the stream-ordered call.
 *     Nothing is allocated inside the library; workspaces are arguments.
 *   - Every function returns 0 on success, a negative viai_status otherwise; viai_last_error() returns a
 *     thread-local human-readable message.  No function synchronises the device.
 *   - `stream` is a cudaStream_t passed as void*.  All entry points are re-entrant (no global mutable
 *     state besides the thread-local error string) and capturable into CUDA graphs.
 */
#ifndef VIAI_B200_H
#define VIAI_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* viai_stream_t;

enum viai_status { VIAI_OK = 0, VIAI_ERR_ARG = -1, VIAI_ERR_CUDA = -2, VIAI_ERR_UNSUPPORTED = -3 };
enum viai_act { VIAI_ACT_NONE = 0, VIAI_ACT_RELU = 1, VIAI_ACT_LRELU = 2, VIAI_ACT_SIGMOID = 3 };

const char* viai_last_error(void);
int viai_version(void);
/* number of kernel launches issued through this library by the calling process (for bench.py gpu_launches) */
long long viai_launch_count(void);

/* Geometry of one implicit-GEMM convolution.
 *   out[n, y, x, o] = bias[o] + sum_{r,s,i} in[n, Y, X, i] * wp[o][r][s][i]
 *   mode 0 (forward gather):     Y = y*stride_h - pad_h + r,            X likewise
 *   mode 1 (transposed gather):  Y = (y + pad_h - r) / stride_h  when divisible, else the tap is skipped
 * mode 0 is nn.Conv2d forward (networks/Inpainting_Networks.py:55-63, networks/Discriminator_Networks.py:17-33,
 * networks/Image_Embedding.py:18, networks/ResNet.py:22) and the data-gradient of a stride-1 ConvTranspose2d;
 * mode 1 is nn.ConvTranspose2d forward (networks/New_Inpainting_Networks.py:24,53-63) and the data-gradient
 * (`convolution_backward`) of nn.Conv2d. */
typedef struct {
  int32_t N, Hin, Win, Cin;   /* gathered tensor */
  int32_t Hout, Wout, Cout;   /* produced tensor */
  int32_t R, S;               /* filter taps */
  int32_t stride_h, stride_w, pad_h, pad_w;
  int32_t mode;
} viai_conv_geom;

/* dst[o][r'][s'][i] = src[o*so + i*si + r*sr + s*ss];  r' = flip ? R-1-r : r, s' likewise.
 * Re-lays a parameter stored in the reference's state_dict layout ((Cout,Cin,kh,kw) for Conv2d,
 * (Cin,Cout,kh,kw) for ConvTranspose2d) into the [O][R][S][I] operand the kernels read. */
int viai_pack_weight(const float* src, float* dst, int O, int I, int R, int S, int64_t so, int64_t si,
                     int64_t sr, int64_t ss, int flip, viai_stream_t stream);

/* CUDA-core fp32 implicit-GEMM convolution (exact-fp32 path; also the on-device validator of the
 * tensor-core path).  bias may be NULL. */
int viai_conv2d_simt(const viai_conv_geom* g, const float* in, const float* wp, const float* bias, float* out,
                     viai_stream_t stream);

/* ---- tensor-core (tcgen05, kind::tf32, fp32 accumulate) implicit-GEMM convolution -------------------------------
 * Same geometry and operand meaning as viai_conv2d_simt; replaces the same ATen call sites.  The weight operand is
 * pre-packed by viai_pack_weight_tc (arguments as viai_pack_weight; values rounded to tf32) into
 * viai_tc_packed_size(O, I, R, S) floats.  If stat_sum/stat_sumsq are non-NULL the epilogue also produces the
 * per-(group, channel) sum and sum of squares of the OUTPUT (double[groups*Cout], zeroed by the call; groups = 1:
 * BatchNorm2d batch statistics, groups = N: InstanceNorm2d), i.e. viai_channel_stats fused into the convolution.
 * viai_conv2d_tc_supported() says whether a geometry is handled (Cin, Cout multiples of 4 and >= 16, <= 16 taps,
 * strides 1 or 2); callers use viai_conv2d_simt otherwise.
 * flags: VIAI_TC_X3 (4) = error-compensated 3-term product  x*w ~ hi(x)*hi(w) + hi(x)*lo(w) + lo(x)*hi(w)  with
 * hi = tf32(.), lo = tf32(. - hi): three tensor-core MMAs per tile instead of one, product error ~2^-19 instead of
 * ~2^-10 (the precision needed for the north star's 1e-3 end-to-end bound; see DESIGN.md "Precision").  It needs
 * weights packed with split = 1.
 * VIAI_TC_BF16X3 (8) = the same 3-term scheme on bf16 pairs (hi = bf16(.), lo = bf16(. - hi); kind::f16 MMAs, K = 16, at
 * twice the tf32 rate; product error ~2^-17).  The kernel splits each fp32 activation slab in shared memory in place into
 * [32 x hi | 32 x lo] per 128-byte pixel row; weights are packed with split = 2 (same size as split = 0).
 * VIAI_TC_F16 (32, together with VIAI_TC_BF16X3) = the pairs are fp16 instead of bf16: activations are scaled by 2^3 and weights
 * (packed with split = 3) by 2^10 before the split, x = hi + lo keeps 22 significand bits (product error ~2^-21, the accuracy of
 * VIAI_TC_X3 at the bf16 MMA rate) and the epilogue multiplies the accumulator by 2^-13 (exact).  |x| >= 8188 or |w| >= 64
 * saturates (the result stays finite) and is counted: viai_tc_f16_overflow(reset, &count) returns the number of saturating threads
 * since the last reset (synchronises with the device; the host side checks it where it already reads the loss).
 * Other bits select alternative shared-memory layouts used as cross-checks. */
#define VIAI_TC_X3 4
#define VIAI_TC_BF16X3 8
#define VIAI_TC_F16 32
int viai_tc_bn(int Cout);
int viai_tc_f16_overflow(int reset, unsigned int* count);
int64_t viai_tc_packed_size(int O, int I, int R, int S, int split);
int viai_pack_weight_tc(const float* src, float* dst, int O, int I, int R, int S, int64_t so, int64_t si, int64_t sr,
                        int64_t ss, int flip, int split, viai_stream_t stream);
/* Every operand re-layout a training-step segment needs in ONE launch (per 24 tensors): kind 0 = viai_pack_weight (fp32
 * [O][R][S][I]), 1 / 2 / 3 / 4 = viai_pack_weight_tc with split 0 / 1 / 2 / 3; dst sized by O*I*R*S floats resp.
 * viai_tc_packed_size.  descs is a HOST array, copied into the kernel's parameters. */
typedef struct viai_pack_desc {
  const float* src; void* dst;
  int32_t O, I, R, S;
  int64_t so, si, sr, ss;     /* element strides of src along O, I, kh, kw */
  int32_t flip, kind;
} viai_pack_desc;
int viai_pack_weights_batched(const viai_pack_desc* descs, int n, viai_stream_t stream);
int viai_conv2d_tc_supported(const viai_conv_geom* g);
int viai_conv2d_tc(const viai_conv_geom* g, const float* in, const float* wp_tc, const float* bias, float* out,
                   double* stat_sum, double* stat_sumsq, int stat_groups, int flags, viai_stream_t stream);

/* Data gradient fused with the first pass of the producing layer's norm backward.  Same convolution as viai_conv2d_tc without
 * bias; `out` is dz, the gradient w.r.t. the post-activation tensor z = act(norm(y)) that was this convolution's forward
 * input.  The epilogue also accumulates, per channel (BatchNorm batch statistics, i.e. one group),
 *     s1 = sum g,  s2 = sum g * xhat,   g = dz * act'(xhat*gamma + beta),  xhat = (y - mean) * invstd
 * which is exactly viai_norm_act_bwd_reduce(dz, y, ...) -- the separate two-stream pass over dz and y disappears
 * (autograd of native_batch_norm_backward + leaky_relu/relu backward, networks/Inpainting_Networks.py:72-76 etc.).
 * y has the shape of `out`; mean/invstd/gamma/beta: float[Cout] or NULL (as in viai_norm_act_fwd).  s1/s2: double[Cout],
 * zeroed by the call. */
typedef struct {
  const float* y;
  const float* mean;
  const float* invstd;
  const float* gamma;
  const float* beta;
  int32_t act;     /* viai_act */
  float slope;
} viai_norm_bwd_ctx;
int viai_conv2d_tc_bwd_reduce(const viai_conv_geom* g, const float* in, const float* wp_tc, float* out,
                              const viai_norm_bwd_ctx* nb, double* s1, double* s2, int flags, viai_stream_t stream);

/* Tensor-core weight gradient: same meaning as viai_conv2d_wgrad_simt (below).  `workspace` is a caller-owned scratch of
 * viai_wgrad_tc_workspace(g) floats (every CTA of the split-K grid stores its per-tap partial sums in its own slice; a
 * second kernel adds the slices and scatters into the gradient layout -- no atomics). */
int viai_conv2d_wgrad_tc_supported(const viai_conv_geom* g);
int64_t viai_wgrad_tc_workspace(const viai_conv_geom* g);
int viai_conv2d_wgrad_tc(const viai_conv_geom* g, const float* U, const float* G, float* dw, int64_t sa, int64_t sb,
                         int64_t sr, int64_t ss, int accumulate, float* workspace, viai_stream_t stream);

/* Weight gradient.  dw[a*sa + b*sb + r*sr + s*ss] (+)= sum_{n,y,x} U[n,y,x,a] * G[n, y*stride-pad+r, x*stride-pad+s, b]
 * U is (N,Hout,Wout,Cout=A), G is (N,Hin,Win,Cin=B) in the geometry struct (mode ignored).
 * For nn.Conv2d: U = dOut, G = input.  For stride-1 nn.ConvTranspose2d: U = input, G = dOut.
 * The destination is zeroed first unless accumulate != 0.  (`convolution_backward` weight branch.) */
int viai_conv2d_wgrad_simt(const viai_conv_geom* g, const float* U, const float* G, float* dw, int64_t sa,
                           int64_t sb, int64_t sr, int64_t ss, int accumulate, viai_stream_t stream);

/* Convolutions with one input or one output channel (the first / last layers of the generator and the discriminator):
 * HBM-bound CUDA-core kernels with coalesced 128-bit accesses.  Operands as viai_conv2d_simt / viai_conv2d_wgrad_simt.
 * viai_conv2d_thin_supported: 1 (Cout == 1), 2 (Cin == 1) or 0;  viai_conv2d_wgrad_thin_supported: A == 1 or B == 1. */
int viai_conv2d_thin_supported(const viai_conv_geom* g);
int viai_conv2d_thin(const viai_conv_geom* g, const float* in, const float* wp, const float* bias, float* out,
                     viai_stream_t stream);
/* Cin == 1 convolution followed by BatchNorm (MelEncoder.conv1 -> bn1, networks/Inpainting_Networks.py:54-55,72-73;
 * MelDiscriminator.conv1 -> bn1, networks/Discriminator_Networks.py:14-15,39): the same kernel also leaves the per-channel
 * sum / sum of squares of `out` (one statistics group; zeroed here first) in stat_sum / stat_sumsq (Cout doubles each) --
 * viai_channel_stats's result without its read pass over the wide tensor.  3x3 and 1x4 filters, Cout <= 64
 * (viai_conv2d_thin_stats_supported; VIAI_CIN1_STATS=0 disables it). */
int viai_conv2d_thin_stats_supported(const viai_conv_geom* g);
int viai_conv2d_thin_stats(const viai_conv_geom* g, const float* in, const float* wp, const float* bias, float* out,
                           double* stat_sum, double* stat_sumsq, viai_stream_t stream);
int viai_conv2d_wgrad_thin_supported(const viai_conv_geom* g);
int64_t viai_wgrad_thin_workspace(const viai_conv_geom* g);   /* floats of caller-owned scratch (per-CTA partial sums) */
int viai_conv2d_wgrad_thin(const viai_conv_geom* g, const float* U, const float* G, float* dw, int64_t sa, int64_t sb,
                           int64_t sr, int64_t ss, int accumulate, float* workspace, viai_stream_t stream);

/* Sweep direction of the normalisation passes below (viai_channel_stats, viai_norm_*_fwd, viai_norm_act_bwd_*).  On tensors of
 * at least `mb` MiB every pass walks its tensors starting where the previous launch of this library ended (back to front after a
 * front-to-back kernel and vice versa), so that it begins on lines still resident in L2; smaller tensors always walk front to
 * back.  Results do not depend on it.  mb >= 0 sets the threshold (0: every tensor), mb < 0 only queries; returns the previous
 * value.  Initial value: VIAI_NORM_WALK_MB or 24; VIAI_NORM_WALK=0 disables the alternation altogether. */
int viai_norm_walk_mb(int mb);

/* Per-(group,channel) sum and sum of squares over rows of an NHWC tensor: groups = 1 is BatchNorm2d's
 * batch statistics, groups = N is InstanceNorm2d's (rows_per_group = H*W).  sum/sumsq are double[groups*C],
 * zeroed by the call.  sumsq may be NULL (plain channel sum: the bias gradient). */
int viai_channel_stats(const float* y, int64_t rows_per_group, int groups, int C, double* sum, double* sumsq,
                       viai_stream_t stream);

/* mean/invstd (float[groups*C]) from the sums; if running_mean != NULL also performs BatchNorm's running
 * update (momentum, unbiased variance) and increments *num_batches_tracked (int64, may be NULL).
 * Replaces native_batch_norm / instance_norm statistics (networks/Inpainting_Networks.py:56-64 etc.). */
int viai_norm_finalize(const double* sum, const double* sumsq, int64_t rows_per_group, int groups, int C, float eps,
                       float* mean, float* invstd, float* running_mean, float* running_var, float momentum,
                       int64_t* num_batches_tracked, viai_stream_t stream);

/* out = act(((y - mean) * invstd) * gamma + beta).  mean/invstd NULL => no normalisation; gamma/beta NULL => 1/0.
 * stat index = (row / rows_per_group) * C + c  (groups = 1: per channel).  In-place allowed. */
int viai_norm_act_fwd(const float* y, int64_t rows_per_group, int groups, int C, const float* mean,
                      const float* invstd, const float* gamma, const float* beta, int act, float slope, float* out,
                      viai_stream_t stream);

/* Backward of viai_norm_act_fwd in two passes.  g = dz * act'(.);  s1 = sum g, s2 = sum g*xhat per (group,channel).
 * s1/s2: double[groups*C], zeroed by the call. */
int viai_norm_act_bwd_reduce(const float* dz, const float* y, int64_t rows_per_group, int groups, int C,
                             const float* mean, const float* invstd, const float* gamma, const float* beta, int act,
                             float slope, double* s1, double* s2, viai_stream_t stream);
/* dy = gamma*invstd*(g - s1/cnt - xhat*s2/cnt)   (training-mode statistics);  with mean == NULL: dy = g.
 * dgamma/dbeta (float[C], may be NULL) receive sum_groups s2 / s1. */
int viai_norm_act_bwd_apply(const float* dz, const float* y, int64_t rows_per_group, int groups, int C,
                            const float* mean, const float* invstd, const float* gamma, const float* beta, int act,
                            float slope, const double* s1, const double* s2, float* dy, float* dgamma, float* dbeta,
                            viai_stream_t stream);
/* The same with the two viai_fold_groups launches folded into the kernel (its first block writes dgamma / dbeta; accumulate != 0
 * adds into them, which is how they land in a gradient bucket).  Needs the statistics path (mean != NULL). */
int viai_norm_act_bwd_apply_fold(const float* dz, const float* y, int64_t rows_per_group, int groups, int C,
                                 const float* mean, const float* invstd, const float* gamma, const float* beta, int act,
                                 float slope, const double* s1, const double* s2, float* dy, float* dgamma, float* dbeta,
                                 int accumulate, viai_stream_t stream);
/* viai_norm_finalize + viai_norm_act_fwd in ONE launch: every thread derives its channels' mean / invstd from the double sums
 * (same arithmetic), the first block also publishes them (mean / invstd: float[groups*C], for the backward pass) and updates the
 * running buffers / num_batches_tracked (groups == 1 only; may be NULL). */
int viai_norm_finalize_act_fwd(const float* y, int64_t rows_per_group, int groups, int C, const double* sum, const double* sumsq,
                               float eps, const float* gamma, const float* beta, int act, float slope, float* out, float* mean,
                               float* invstd, float* running_mean, float* running_var, float momentum,
                               int64_t* num_batches_tracked, viai_stream_t stream);
/* out[c] (+)= sum_g sums[g*C+c]  -- double -> float fold used for bias gradients. */
int viai_fold_groups(const double* sums, int groups, int C, float* out, int accumulate, viai_stream_t stream);

/* F.interpolate(mode='bilinear', align_corners=True) (networks/New_Inpainting_Networks.py:78,83) written into the
 * channel slice [coff, coff+C) of an NHWC tensor with Ctot channels (this is how torch.cat :81 is made copy-free). */
int viai_bilinear_fwd(const float* in, int N, int Hin, int Win, int C, float* out, int Hout, int Wout, int Ctot,
                      int coff, viai_stream_t stream);
int viai_bilinear_bwd(const float* dout, int N, int Hin, int Win, int C, float* din, int Hout, int Wout, int Ctot,
                      int coff, viai_stream_t stream);
/* dst[row][doff + c] = src[row][soff + c], c < C   (torch.cat / its backward slice) */
int viai_copy_channels(const float* src, int64_t rows, int Csrc, int soff, float* dst, int Cdst, int doff, int C,
                       viai_stream_t stream);
/* nn.AvgPool2d((kh,1)) on NHWC (networks/Inpainting_Networks.py:65,77) */
int viai_avgpool_h_fwd(const float* in, int N, int H, int W, int C, int kh, float* out, viai_stream_t stream);
int viai_avgpool_h_bwd(const float* dout, int N, int H, int W, int C, int kh, float* din, viai_stream_t stream);
/* nn.MaxPool2d(3, 2, 1) (networks/Image_Embedding.py:21) forward/backward, NHWC */
int viai_maxpool3s2_fwd(const float* in, int N, int H, int W, int C, float* out, int Ho, int Wo, viai_stream_t stream);
int viai_maxpool3s2_bwd(const float* in, const float* dout, int N, int H, int W, int C, float* din, int Ho, int Wo,
                        viai_stream_t stream);
/* out = a * b  (mask application, bit exact), out = a + b, out = relu(a + b) and its backward */
int viai_mul(const float* a, const float* b, float* out, int64_t n, viai_stream_t stream);
int viai_add_act(const float* a, const float* b, float* out, int64_t n, int act, viai_stream_t stream);
int viai_add_act_bwd(const float* out, const float* dout, float* din, int64_t n, int act, viai_stream_t stream);

/* STFT -> mel front end: utils/audio.py:70-75 melspectrogram (lws framing :90-108, |STFT| -> mel basis :116-120 -> dB :130-132
 * -> minus ref_level_db :72 -> normalise/clip :139-140) as one kernel.  y: T samples; frame m covers samples
 * [m*hop - pad_left, m*hop - pad_left + fft_size) with zeros outside [0, T); window: fft_size floats; mel_basis:
 * (n_mels, fft_size/2+1) row-major; mel_span: int[n_mels][2] = [first, last+1) non-zero bin of each filter;
 * out: (n_mels, num_frames) in [0,1]; mag_out (optional): |STFT| (fft_size/2+1, num_frames). */
int viai_stft_mel(const float* y, int64_t T, int fft_size, int hop, int pad_left, int num_frames, const float* window,
                  const float* mel_basis, const int* mel_span, int n_mels, float min_level_db, float ref_level_db,
                  float* out, float* mag_out, viai_stream_t stream);

/* WaveNet vocoder synthesis: wavenet_vocoder/wavenet.py:237-364 incremental_forward (+ modules.py:162-210, conv.py:17-62,
 * mixture.py:117-153) as one persistent cooperative kernel that runs all T autoregressive steps on the device.
 * L layers (dilation 2^(l % layers_per_stack)), R residual / G gate / S skip / C conditioning channels, K taps, O = 3*nr_mix
 * outputs, B <= 4 utterances.  nC = viai_wavenet_num_ctas(...) CTAs cooperate (0: unsupported configuration).
 * packed_layers / first / head1 / head2: parameter blocks in the layout documented in csrc/wavenet_synth.cu (produced by
 * WaveNet.pack_for_synthesis); cond: (B, T, C) upsampled conditioning; uniforms: (T, B, nr_mix + 1) draws in (0, 1);
 * test_inputs (optional): (B, Ttest) teacher-forced inputs for the first Ttest steps.  ring (sum_l ((K-1)*d_l + 1) * B * R
 * floats at offsets ring_off[l]): zero-initialised state.  gbuf (B*G/2), sbuf (B*S), hbuf (B*S), xchg (B*R): zero-initialised,
 * 8-byte aligned exchange buffers of 64-bit {value, stage tag} words, i.e. TWO 32-bit words per element (the CTAs synchronise
 * through these tagged words; there is no grid barrier).
 * out: (B, T) samples in [-1, 1]; logits (optional): (B, T, O) pre-sampling outputs. */
int viai_wavenet_num_ctas(int R, int G, int S, int C, int K, int O, int B);
int viai_wavenet_synth(int L, int layers_per_stack, int R, int G, int S, int C, int K, int O, int B, int T, int nC,
                       const float* packed_layers, const float* first, const float* head1, const float* head2,
                       const float* cond, const float* uniforms, const float* test_inputs, int Ttest, float log_scale_min,
                       float* ring, const int64_t* ring_off, float* gbuf, float* sbuf, float* hbuf, unsigned* xchg, float* out,
                       float* logits, viai_stream_t stream);

/* The same synthesis on ONE 16-CTA thread-block cluster: cross-CTA vectors travel through distributed shared memory instead of
 * global memory (an exchange costs a few hundred ns instead of ~1.3 us), the weights are streamed by 16 SMs.  Operands as
 * viai_wavenet_synth with the parameter blocks packed for nC = 16; B <= 2; only `ring` is caller-owned state.
 * viai_wavenet_cluster_supported returns 16 when the configuration fits, else 0 (use viai_wavenet_synth). */
int viai_wavenet_cluster_supported(int R, int G, int S, int C, int K, int O, int B);
int viai_wavenet_synth_cluster(int L, int layers_per_stack, int R, int G, int S, int C, int K, int O, int B, int T,
                               const float* packed_layers, const float* first, const float* head1, const float* head2,
                               const float* cond, const float* uniforms, const float* test_inputs, int Ttest,
                               float log_scale_min, float* ring, const int64_t* ring_off, float* out, float* logits,
                               viai_stream_t stream);

/* Folded schedule of the same synthesis (csrc/wavenet_synth2.cu): one dependent cross-CTA exchange per layer instead of two
 * (the current-time tap of layer l is folded through layer l-1's residual 1x1 when the parameters are packed).  Same contract
 * and arguments as viai_wavenet_synth except: packed_layers / last come from WaveNet.pack_for_synthesis_folded;  gbuf holds
 * 2 * B * (G/2) and xchg 3 * B * R 64-bit words (zeroed, 8-byte aligned).  viai_wavenet2_num_ctas returns 0 when the
 * configuration is unsupported (L < 3, or it does not fit) -- use viai_wavenet_synth then. */
int viai_wavenet2_num_ctas(int L, int R, int G, int S, int C, int K, int O, int B);
int viai_wavenet_synth2(int L, int layers_per_stack, int R, int G, int S, int C, int K, int O, int B, int T, int nC,
                        const float* packed_layers, const float* last, const float* first, const float* head1,
                        const float* head2, const float* cond, const float* uniforms, const float* test_inputs, int Ttest,
                        float log_scale_min, float* ring, const int64_t* ring_off, float* gbuf, float* sbuf, float* hbuf,
                        unsigned* xchg, float* out, float* logits, viai_stream_t stream);

/* Warp-specialised variant of the folded schedule (csrc/wavenet_synth3.cu): 8 warps run the dependent chain (one warp per
 * gate-row pair / residual row / skip row, no CTA-wide barrier), 8 warps evaluate the next layer's independent part, one
 * producer lane streams the weight blocks through a shared-memory ring with bulk copies.  Same arguments as
 * viai_wavenet_synth2 except for the exchange buffers, which hold NREP = viai_wavenet3_replicas() copies of every vector
 * (a CTA reads copy cta % NREP; one copy made 128 CTAs poll the same 8 L2 slices): gbuf 3 * NREP * B * (G/2), xchg
 * 3 * NREP * B * R, sbuf / hbuf NREP * B * S 64-bit words.  viai_wavenet3_num_ctas returns 0 when the configuration is
 * unsupported (use viai_wavenet_synth2 / viai_wavenet_synth then). */
int viai_wavenet3_replicas(void);
int viai_wavenet3_num_ctas(int L, int R, int G, int S, int C, int K, int O, int B);
int viai_wavenet_synth3(int L, int layers_per_stack, int R, int G, int S, int C, int K, int O, int B, int T, int nC,
                        const float* packed_layers, const float* last, const float* first, const float* head1,
                        const float* head2, const float* cond, const float* uniforms, const float* test_inputs, int Ttest,
                        float log_scale_min, float* ring, const int64_t* ring_off, float* gbuf, float* sbuf, float* hbuf,
                        unsigned* xchg, float* out, float* logits, viai_stream_t stream);

/* Debug aid: VIAI_WN3_PROF=1 selects a profiling instantiation of viai_wavenet_synth3; per-phase clocks of CTA 0 (phase list
 * in csrc/wavenet_synth3.cu); synchronises the device. */
int viai_wavenet3_profile(long long* out32);

/* Debug aid: with VIAI_WN2_PROF=1 in the environment, clocks CTA 0 spent per phase of the last viai_wavenet_synth2 launch
 * (phase list in csrc/wavenet_synth2.cu); synchronises the device. */
int viai_wavenet2_profile(long long* out16);

/* Losses (loss_functions.py:79-104 GANLoss = MSELoss / BCELoss against an expanded scalar; nn.L1Loss).
 * kind 0: mean (p-t)^2   1: BCE(p, t)   2: mean |p - q|  (q = other tensor).  acc: double[1] workspace.
 * loss_out: float[1] on device. */
int viai_loss_fwd(int kind, const float* p, const float* q, float target, int64_t n, double* acc, float* loss_out,
                  viai_stream_t stream);
/* dp = (*gout) * dloss/dp;  gout: float[1] on device (upstream scalar gradient). */
int viai_loss_bwd(int kind, const float* p, const float* q, float target, int64_t n, const float* gout, float* dp,
                  viai_stream_t stream);

/* Fused Adam over one flat buffer (torch.optim.Adam semantics, no weight decay / amsgrad).
 * step_dev: float[1] device counter, incremented first when tick != 0; lr_dev: float[1] device learning rate (so
 * schedules work under CUDA graphs); grad_scale multiplies g (1/world_size after a sum all-reduce). */
int viai_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev, double beta1,
                   double beta2, double eps, float* step_dev, int tick, float grad_scale, viai_stream_t stream);
/* out[0] = wa * a[0] + wb * b[0]   (b may be NULL) -- scalar loss arithmetic kept on the device */
int viai_lincomb2(const float* a, float wa, const float* b, float wb, float* out, viai_stream_t stream);
int viai_fill(float* p, int64_t n, float value, viai_stream_t stream);
/* out[i] = 1/sqrt(var[i] + eps)  (eval-mode BatchNorm uses running_var) */
int viai_rsqrt_eps(const float* var, int n, float eps, float* out, viai_stream_t stream);

/* ---- WaveNet teacher-forced training path: wavenet_vocoder/wavenet.py:177-235 forward,
 * wavenet_vocoder/modules.py:162-210 ResidualConv1dGLU._forward (is_incremental = False), wavenet_vocoder/mixture.py:25-105,
 * loss_functions.py:11-76.  Activations are (B, T, C) rows ("NHWC" with H*W = B*T pixels). */
/* Operand of a dilated causal nn.Conv1d (modules.py:175-178: conv + "remove future time steps") joined with the
 * local-conditioning features (modules.py:183-187), so that conv + conv1x1c is ONE 1x1 GEMM with the linearised weight of
 * conv.py:51-62:  out[b,t,k*R + r] = x[b, t - (K-1-k)*dilation, r] (0 before t = 0), out[b,t,K*R + j] = c[b,t,j], zero up to
 * Kpad.  R, Cc, Kpad multiples of 4; c may be NULL when Cc == 0.  mask (B,T,R; may be NULL) and scale fuse the dropout of
 * modules.py:173: x is read as x * mask * scale. */
int viai_shiftcat_fwd(const float* x, const float* c, const float* mask, float scale, int B, int T, int R, int Cc, int K, int dilation,
                      int Kpad, float* out, viai_stream_t stream);
/* dx (B,T,R) and dc (B,T,Cc; may be NULL) from dout (B,T,Kpad) */
int viai_shiftcat_bwd(const float* dout, const float* mask, float scale, int B, int T, int R, int Cc, int K, int dilation, int Kpad,
                      float* dx, float* dc, viai_stream_t stream);
/* DeepVoice3-style weight normalisation (wavenet_vocoder/modules.py:22-32 via nn.utils.weight_norm, dim 0):
 * w[row,:] = v[row,:] * g[row] / ||v[row,:]||; norms: float[rows] kept for the backward (dv, dg from dw). */
int viai_weight_norm_fwd(const float* v, const float* g, int rows, int cols, float* w, float* norms, viai_stream_t stream);
int viai_weight_norm_bwd(const float* v, const float* g, const float* norms, const float* dw, int rows, int cols, float* dv, float* dg,
                         viai_stream_t stream);
/* out[row,i] = tanh(y[row,i]) * sigmoid(y[row,G/2+i])  (modules.py:180,196) and its backward; G % 8 == 0 */
int viai_glu_fwd(const float* y, int64_t rows, int G, float* out, viai_stream_t stream);
int viai_glu_bwd(const float* y, const float* dout, int64_t rows, int G, float* dy, viai_stream_t stream);
/* out = alpha*a + beta*b (b may be NULL; out may alias a): (x + residual)*sqrt(0.5) modules.py:204, skip accumulation
 * wavenet.py:222-223, ExponentialMovingAverage.update loss_functions.py:70-73 */
int viai_axpby(const float* a, float alpha, const float* b, float beta, float* out, int64_t n, viai_stream_t stream);
/* discretized_mix_logistic_loss(reduce=False) (mixture.py:25-105): y_hat rows of 3*nr_mix floats (logits | means | log
 * scales), target rows in [-1,1].  nll[row] = -log_sum_exp(log_probs) (may be NULL); with dnll/dy_hat also the analytic
 * gradient dy_hat[row,:] = dnll[row] * d nll / d y_hat.  nr_mix <= 32. */
int viai_dmol_nll(const float* y_hat, const float* target, int64_t rows, int nr_mix, int num_classes, float log_scale_min, float* nll,
                  const float* dnll, float* dy_hat, viai_stream_t stream);
/* sample_from_discretized_mix_logistic (mixture.py:117-153): uniforms rows of nr_mix + 1 draws in (1e-5, 1 - 1e-5)
 * ([0:nr_mix] Gumbel-max over the logits, [nr_mix] the logistic sample); out[row] in [-1, 1] */
int viai_dmol_sample(const float* y_hat, const float* uniforms, int64_t rows, int nr_mix, float log_scale_min, float* out,
                     viai_stream_t stream);
/* mean != 0: (v*mask).sum() / mask.sum() (loss_functions.py:59); mean == 0: (v*mask).sum().  mask may be NULL (ones).
 * acc: double[2] workspace kept for the backward (sum v*m, sum m); out / gout: float[1] on device. */
int viai_masked_sum_fwd(const float* v, const float* mask, int64_t n, int mean, double* acc, float* out, viai_stream_t stream);
int viai_masked_sum_bwd(const float* mask, int64_t n, int mean, const double* acc, const float* gout, float* dv, viai_stream_t stream);
/* sequence_mask (loss_functions.py:11-21): out[b,t] = t < lengths[b] */
int viai_sequence_mask(const int64_t* lengths, int B, int T, float* out, viai_stream_t stream);

/* ---- Audio-visual synchronisation heads */
/* utils/util.py:94-96 l2_norm = F.normalize(x, p=2, dim=1): y = x / max(||x||, eps); norms: float[rows] kept for the backward */
int viai_l2norm_fwd(const float* x, int rows, int cols, float eps, float* y, float* norms, viai_stream_t stream);
int viai_l2norm_bwd(const float* y, const float* norms, const float* dy, int rows, int cols, float eps, float* dx, viai_stream_t stream);
/* loss_functions.py:106-108 l2_sim: scores[a][b] = ||f1[a] - f2[b]||_2 (also the distance matrix of utils/util.py:99-121 L2retrieval) */
int viai_pairdist_fwd(const float* f1, const float* f2, int n1, int n2, int F, float* scores, viai_stream_t stream);
int viai_pairdist_bwd(const float* f1, const float* f2, const float* scores, const float* dscores, int n1, int n2, int F, float* df1,
                      float* df2, viai_stream_t stream);
/* utils/util.py:99-121 L2retrieval on the (captions x clips) distance matrix of viai_pairdist_fwd: rank[i] = position of clip i in
 * row i sorted ascending (ties: lower index first), top1[i] = index of the closest clip; int64[n1] each.  Replaces np.argsort +
 * np.where of the reference without sorting. */
int viai_retrieval_ranks(const float* dist, int n1, int n2, long long* rank, long long* top1, viai_stream_t stream);
/* loss_functions.py:111-148 L2ContrastiveLoss on a (B,B) score matrix: loss (float[1], may be NULL) and, with gout/dscores,
 * dscores = gout * d loss / d scores */
int viai_l2_contrastive(const float* scores, int B, float margin, int max_violation, float* loss, const float* gout, float* dscores,
                        viai_stream_t stream);

/* ---- Loader: video-frame preprocessing, Data_loaders/audio_loader.py:199-246 (sample_data_new) and :262-300.
 * src: n_frames decoded uint8 frames (src_h, src_w, src_c) as cv2.imread returns them (BGR for src_c = 3, gray for 1).
 * Per frame: cv2.resize(frame, (resize_w, resize_h)) (INTER_LINEAR 8-bit fixed point, bit exact) -> BGR->RGB when swap_rb ->
 * np.fliplr when flip -> (v - 127) / 128 -> crop [crop_row, crop_row+out_h) x [crop_col, crop_col+out_w) -> written to channels
 * [c_off, c_off+src_c) of the float NHWC block out (n_frames, out_h, out_w, out_c) (flow_x / flow_y fill channels 0 / 1). */
int viai_frames_preprocess(const uint8_t* src, int n_frames, int src_h, int src_w, int src_c, int swap_rb, int resize_h, int resize_w,
                           int flip, int crop_row, int crop_col, int out_h, int out_w, int out_c, int c_off, float* out,
                           viai_stream_t stream);

/* ---- Opt-in fast paths of the ResNet stem (networks/Image_Embedding.py:18-21), off by default (VIAI_FAST_STEM=1): they were
 * written from the C3 per-kernel table (profiles/r01_c3_step_kernels.csv: the CUDA-core 7x7 weight gradient is 56 % of that step,
 * the max-pool backward 15 %) after the round's GPU budget was spent and are NOT yet validated on a B200. */
/* im2col of a mode-0 convolution, channel-major columns k = c*R*S + r*S + s, rows = output pixels, zero-padded to Kpad
 * (multiple of 4): turns the weight gradient of a few-channel convolution into a 1x1 weight gradient (viai_conv2d_wgrad_tc). */
int viai_im2col(const viai_conv_geom* g, const float* in, int Kpad, float* out, viai_stream_t stream);
/* nn.MaxPool2d(3, 2, 1) backward that also reads the forward output (rejects non-maximal elements after one load); C % 4 == 0 */
int viai_maxpool3s2_bwd_out(const float* in, const float* out, const float* dout, int N, int H, int W, int C, float* din, int Ho,
                            int Wo, viai_stream_t stream);
/* Max-pool with a saved argmax (1 byte per output element: the window position 0..8 of the FIRST maximum, row-major, ATen's
 * tie-breaking rule): the backward pass gathers dout through it and reads neither the forward input nor output.  C % 4 == 0.
 * Replaces F.max_pool2d / max_pool2d_with_indices_backward at networks/Image_Embedding.py:21. */
int viai_maxpool3s2_fwd_idx(const float* in, int N, int H, int W, int C, float* out, unsigned char* idx, int Ho, int Wo,
                            viai_stream_t stream);
int viai_maxpool3s2_bwd_idx(const unsigned char* idx, const float* dout, int N, int H, int W, int C, float* din, int Ho, int Wo,
                            viai_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VIAI_B200_H */
