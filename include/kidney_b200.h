/*
 * libkidney_b200 -- C ABI of the B200-native (sm_100a) sampling hot path of
 * jameshball/kidney-diffusion.
 *
 * The reference is pure Python (no native components, no FFI).
 * Its boundary for this path is the Python API of imagen-pytorch==1.18.5
 * (requirements.txt:37) as constructed in train_ultra_res_v_param.py:27-92 /
 * train.py:28-95 / train_uncond.py:28-93 and called at
 * sample_ultra_res.py:183-195, sample_cond.py:40-48, sample_uncond.py:49-55.
 * Each entry point below replaces the group of PyTorch-eager library calls that
 * imagen-pytorch issues for one step of that path; the "replaces" note on every
 * function names the reference-side operation (module of imagen_pytorch.py as
 * restated in oracle/imagen_oracle.py, which cites the call sites).
 *
 * Conventions
 *  - plain pointers + sizes, no torch types; every pointer is DEVICE memory
 *    owned by the caller unless stated otherwise; the library never allocates or
 *    frees device memory (one exception: the kd_peer_* mailbox) and keeps no pointer after return.
 *  - all work is enqueued on `stream` (a cudaStream_t); no implicit device
 *    synchronisation; every call is CUDA-graph capturable.
 *  - return 0 on success, negative KdStatus otherwise; message via
 *    kd_last_error() (thread-local).
 *  - activations inside the UNet are NHWC fp16 ("[B,H,W,C]"), image state at the
 *    sampler level is NCHW fp32 exactly as in the reference.
 */
#ifndef KIDNEY_B200_H_
#define KIDNEY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* kd_stream_t; /* cudaStream_t */

enum KdStatus {
  KD_OK = 0,
  KD_ERR_BAD_ARG = -1,
  KD_ERR_CUDA = -2,
  KD_ERR_LAUNCH = -3,
  KD_ERR_ARCH = -4,
  KD_ERR_UNSUPPORTED = -5
};

enum KdAct { KD_ACT_NONE = 0, KD_ACT_SILU = 1, KD_ACT_GELU = 2, KD_ACT_SIGMOID = 3 };
enum KdObjective { KD_PRED_NOISE = 0, KD_PRED_V = 1, KD_PRED_X0 = 2 };

int kd_version(void);
const char* kd_last_error(void);
/* 0 iff the current CUDA device is compute capability 10.0 (B200). No fallback exists. */
int kd_check_device(void);

/* ------------------------------------------------------------------ K1: implicit-GEMM convolution (tcgen05 + TMEM + TMA)
 * replaces: nn.Conv2d 3x3 / 1x1 in Block.project, ResnetBlock.res_conv, Downsample (pixel-unshuffle + 1x1),
 *           PixelShuffleUpsample (1x1 + SiLU + PixelShuffle), Parallel(3x3,1x1), and every nn.Linear applied to
 *           image tokens (Attention/CrossAttention to_q,to_kv,to_out, ChanFeedForward).
 * out[b,h,w,n] = act( sum_{tap,c} A[b, h+dy(tap), w+dx(tap), c] * Wt[n, tap, c] + bias[n] )
 *                + addend_scale[b,n] * addend[b,h,w,n]
 * A is the channel concatenation of two NHWC fp16 sources (xa: Ca channels, xb: Cb channels; Cb may be 0).
 */
typedef struct KdConvDesc {
  int mode;         /* 0: ksize x ksize, stride 1, zero pad ksize/2.  1: 2x2 stride-2 "pixel-unshuffle" taps (Downsample);
                       2: plain GEMM on a [M,K] row-major matrix given as xa (H=1, W=M, Ca=K) */
  int B, H, W;      /* OUTPUT batch / height / width (mode 1: the input is [B,2H,2W,C]) */
  int Ca, Cb;       /* input channels of the two sources; each a multiple of 64 */
  int Cout;         /* output channels (multiple of 8) */
  int ksize;        /* 1 or 3 for mode 0; ignored otherwise */
  int act;          /* KdAct applied before the addend */
  int out_mode;     /* 0: NHWC [B,H,W,Cout].  1: pixel-shuffle(2): weight rows ordered (dy,dx,c) -> out [B,2H,2W,Cout/4] */
  int out_f32;      /* 0: fp16 output, 1: fp32 output */
  int addend_f32;   /* dtype of addend (same layout as out) */
} KdConvDesc;

/* Test / profiling hook: 0 = automatic kernel choice (default), 1 = force the single-CTA 128x128 kernel, 2 = force the
 * CTA-pair (cta_group::2) kernels (halo variant where it applies), 4 = CTA-pair tap-loop kernel only, 8 = automatic but never
 * split-K.  Kernels 1 / 2 / 4 accumulate each output in the same k order: their results are bit-identical (split-K adds partial
 * sums, so it agrees with them to fp32 rounding only). */
int kd_set_conv_impl(int impl);

int kd_conv_gemm(const KdConvDesc* desc, const void* xa, const void* xb,
                 const void* w /* fp16 [Cout, taps*(Ca+Cb)], K ordered (tap, channel) */, const float* bias /* [Cout] or NULL */,
                 const void* addend /* or NULL */, const float* addend_scale /* [B,Cout] or NULL (=1) */, void* out,
                 kd_stream_t stream);

/* Operator fusion around kd_conv_gemm (every pointer may be NULL):
 *   stats      : OUT, fused GroupNorm statistics of the stored output, stats[row][Cout/8] = {sum, sumsq} (fp32) over 32 output
 *                pixels x 8 channels, rows = 4 per 128-pixel M-tile; geometry from kd_conv_stats_layout, consumed by kd_oct_reduce;
 *   logit_w / logit_parts : the GlobalContext to_k 1x1 conv (replaces kd_rowdot): logit_parts[g][b*H*W + n] = sum over the
 *                64 output channels of group g of out[b,n,c] * logit_w[c]; kd_gca_pool adds the Cout/64 parts in fixed order
 *                (needs out_mode 0, Cout % 64 == 0, no addend_scale);
 *   pre_coef   : IN, [B][Ca+Cb] x {A, B} fp32 from kd_gn_finalize_oct: the convolution consumes SiLU(A * x + B) instead of
 *                x -- GroupNorm + scale/shift + SiLU of Block.forward applied to the raw tensors while the halo tile sits in
 *                shared memory (replaces kd_gn_apply and its write + re-read of the activated tensor).
 * kd_conv_stats_layout: layout[0] = rows of the stats buffer, 0 if this shape's kernel cannot produce stats / logits;
 * [1] = M-tiles per batch group; [2] = batch images per tile; [3] = 1 if pre_coef is supported (3x3 halo kernel). */
typedef struct KdConvFusion {
  float* stats;
  const float* logit_w;
  float* logit_parts;
  const float* pre_coef;
  void* splitk_ws;        /* caller-owned scratch for shapes that run split-K, >= kd_conv_splitk_workspace_bytes(desc) */
  size_t splitk_ws_bytes;
} KdConvFusion;
/* Convolutions on tiny images (<= 16 x 16 output pixels per sample, K = taps * Cin in the thousands: the 8^2 / 16^2 levels of
 * the 64^2 base UNet) divide K over several CTAs; a second kernel adds the fp32 partial tiles in fixed split order and applies
 * the fused epilogue.  The split count depends on the per-sample shape only (never on B), so results stay batch-invariant.
 * Returns 0 when `desc` does not split (no workspace needed). */
size_t kd_conv_splitk_workspace_bytes(const KdConvDesc* desc);
int kd_conv_stats_layout(const KdConvDesc* desc, int* layout /* [4] */);
int kd_conv_gemm_fused(const KdConvDesc* desc, const void* xa, const void* xb, const void* w, const float* bias, const void* addend,
                       const float* addend_scale, void* out, const KdConvFusion* fusion, kd_stream_t stream);

/* ------------------------------------------------------------------ small-M linear (time / conditioning towers, GCA MLP)
 * replaces: nn.Linear on (B, features) tensors: to_time_hiddens, to_time_cond, to_time_tokens, ResnetBlock.time_mlp,
 *           GlobalContext.net, CrossAttention.to_kv on the conditioning tokens.
 * y[m, n] = post_act( sum_k pre_act(x[m,k]) * w[n,k] + bias[n] ), fp32 everywhere, M <= 65536 rows (meant for M of a few hundred at most). */
int kd_linear_small(const float* x, int M, int K, long ldx, const float* w, const float* bias, float* y, int N, long ldy,
                    int pre_act, int post_act, kd_stream_t stream);

/* replaces: LearnedSinusoidalPosEmb.forward -> out[b] = [t, sin(2 pi t w), cos(2 pi t w)] (1 + 2*half values) */
int kd_sinu_emb(const float* t, const float* weights, int B, int half, float* out, kd_stream_t stream);

/* ------------------------------------------------------------------ K3: GroupNorm (statistics / finalize / apply)
 * replaces: nn.GroupNorm + (scale+1)*x+shift + SiLU in Block.forward, including GroupNorm over the channel concat
 *           cat(x, skip * 2^-0.5) of the up path (two sources, the second pre-scaled by src_scale).
 * kd_gn_stats writes deterministic per-block partial sums  partial[b][blk][g] = {sum, sumsq} over the channels of this
 * source that fall into global group g (global channel = c_offset + c; group = global channel / group_size). */
int kd_gn_stats(const void* x /* fp16 [B,HW,C] */, int B, long HW, int C, int c_offset, int group_size, int num_groups,
                float* partial /* [B][nblk][num_groups][2] */, int nblk, kd_stream_t stream);
int kd_gn_finalize(const float* partial_a, int nblk_a, float scale_a, const float* partial_b, int nblk_b, float scale_b, int B,
                   int num_groups, double count /* elements per (b, group) */, float eps, float* mean_rstd /* [B][G][2] */,
                   kd_stream_t stream);
/* y = act( ((x*src_scale - mean) * rstd * gamma + beta) * (scale + 1) + shift ), fp16 out.
 * scale_shift: fp32 rows of 2*Ctot values laid out as time_mlp output (scale = first Ctot, shift = last Ctot), row b at
 * scale_shift + b*ss_stride (so one launch of kd_linear_small can produce every block's time_mlp at once), or NULL. */
int kd_gn_apply(const void* x, void* y, int B, long HW, int C, int c_offset, int group_size, int num_groups, float src_scale,
                const float* mean_rstd, const float* gamma, const float* beta, const float* scale_shift, long ss_stride,
                int Ctot, int act, kd_stream_t stream);

/* Octet-granular statistics (8 channels), the form the conv epilogue emits: any GroupNorm grouping (group sizes are
 * multiples of 8), including groups that straddle the two sources of a channel concat, is derived from them afterwards.
 * kd_oct_stats:  partial[b][blk][C/8] = {sum, sumsq} of x over a chunk of pixels (standalone pass, same as kd_gn_stats).
 * kd_oct_reduce: first-level sum of partial rows into out[b][NS][C/8][2], NS = kd_oct_reduce_splits(rpt, tiles, TB); physical
 *                row of logical (b, i): tile_b = b / TB, sub = b % TB, rpb = rpt / TB,
 *                row = ((tile_b * tiles + i / rpb) * rpt) + sub * rpb + i % rpb, for i < tiles * rpb  (conv: rpt = 4).
 * kd_gn_finalize_oct: mean / rstd per (b, group) from the split sums of one or two (concatenated) sources; with coef != NULL
 *                also the per-channel affine coef[b][c] = {A, B} of GroupNorm + (scale+1, shift) on the RAW sources
 *                (y = A*x + B; source scales folded into A), input of KdConvFusion.pre_coef. */
int kd_oct_stats(const void* x, int B, long HW, int C, float* partial /* [B][nblk][C/8][2] */, int nblk, kd_stream_t stream);
int kd_oct_reduce_splits(int rpt, int tiles, int TB);
int kd_oct_reduce(const float* partial, int rpt, int tiles, int TB, int B, int n_oct, float* out /* [B][NS][n_oct][2] */, kd_stream_t stream);
int kd_gn_finalize_oct(const float* sum_a, int n_oct_a, int ns_a, float scale_a, const float* sum_b, int n_oct_b, int ns_b, float scale_b,
                       int B, int num_groups, int group_size, double count, float eps, float* mean_rstd,
                       const float* gamma /* [C] */, const float* beta, const float* scale_shift /* [B][ss_stride]: scale | shift, or NULL */,
                       long ss_stride, float* coef /* [B][C][2] or NULL */, kd_stream_t stream);

/* kd_oct_reduce + kd_gn_finalize_oct in ONE launch (bit-identical results): the last block to finish an image finalizes it.
 * scratch: B * (NS_a * n_oct_a + NS_b * n_oct_b) * 2 floats (NS = kd_oct_reduce_splits of each source); counter: >= B
 * unsigned ints, zero before the first call (the kernel leaves them zero). */
int kd_gn_reduce_finalize(const float* partial_a, int rpt_a, int tiles_a, int TB_a, int n_oct_a, float scale_a, const float* partial_b,
                          int rpt_b, int tiles_b, int TB_b, int n_oct_b, float scale_b, int B, int num_groups, int group_size,
                          double count, float eps, float* scratch, unsigned int* counter, float* mean_rstd, const float* gamma,
                          const float* beta, const float* scale_shift, long ss_stride, float* coef, kd_stream_t stream);

/* ------------------------------------------------------------------ K4: GlobalContext gate
 * replaces: GlobalContext.forward (to_k 1x1 conv -> softmax over H*W -> weighted channel sum) and h * gate + residual. */
int kd_rowdot(const void* x /* fp16 [B,HW,C] */, const float* w /* [C] */, const float* bias /* [1] or NULL */, float* out /* [B,HW] */,
              int B, long HW, int C, kd_stream_t stream);
int kd_gca_pool(const void* x, const float* logits /* [n_parts][B*HW]: partial logits, summed per pixel */, int n_parts, int B, long HW,
                int C, int nblk, float* part /* [B][nblk][C] */, float* ml /* [B][nblk][2] = {max, sumexp} */, kd_stream_t stream);
int kd_gca_finalize(const float* part, const float* ml, int B, int nblk, int C, float* pooled /* [B][C] */, kd_stream_t stream);
/* kd_gca_finalize + GlobalContext.net (Conv1x1 C -> hid, SiLU, Conv1x1 hid -> C, Sigmoid) in ONE launch: a cluster of 8 CTAs per
 * image merges the pooling partials, splits both matrix-vector products by output rows and passes the hidden vector through
 * distributed shared memory.  gate[b][c] is what kd_gate_residual / the residual conv's epilogue multiplies h by. */
int kd_gca_gate(const float* part, const float* ml, int B, int nblk, int C, int hid, const float* w0 /* [hid][C] */, const float* b0,
                const float* w1 /* [C][hid] */, const float* b1, float* gate /* [B][C] */, kd_stream_t stream);
/* out = h * gate[b,c] + res   (gate NULL -> 1, res NULL -> 0); fp16 in/out.  oct_partial (optional, [B][nblk][C/8][2]):
 * fused statistics of `out` in kd_oct_stats form, nblk = kd_elementwise_blocks(HW, C). */
int kd_gate_residual(const void* h, const float* gate, const void* res, void* out, float* oct_partial, int B, long HW, int C,
                     kd_stream_t stream);
int kd_elementwise_blocks(long HW, int C);

/* ------------------------------------------------------------------ LayerNorm over channels of NHWC tokens
 * replaces: imagen-pytorch LayerNorm / ChanLayerNorm (gain only, eps 1e-5) and nn.LayerNorm (gain + bias).
 * y = (x - mean) * rsqrt(var + eps) * g (+ bias) (+ residual) */
int kd_layernorm_h16(const void* x, const float* g, const float* bias, const void* residual, void* y, long M, int C, float eps,
                      kd_stream_t stream);
int kd_layernorm_f32(const float* x, const float* g, const float* bias, float* y, long M, int C, float eps, kd_stream_t stream);

/* ------------------------------------------------------------------ K5: attention
 * replaces: Attention.forward (multi-query: one shared 64-d K/V head, null k/v and optional context k/v prepended) and
 *           CrossAttention.forward (full multi-head K/V from <= 64 conditioning tokens + null k/v). */
int kd_kv_assemble(const void* qkv /* fp16 [B,N,ld] */, long ld, int kv_col, const float* ctx_kv /* [B,Jc,128] or NULL */, int Jc,
                   const float* null_kv /* [2,64] */, void* kv_out /* fp16 [B, Jc+1+N, 128] */, int B, int N, kd_stream_t stream);
int kd_attn_mqa(const void* q /* fp16 [B,N,*] */, long ldq, const void* kv /* fp16 [B,J,128] */, void* out /* fp16 [B,N,heads*64] */,
                int B, int N, int J, int heads, float scale, kd_stream_t stream);
/* kd_attn_mqa on the 5th-gen tensor cores (tcgen05.mma with S and PV accumulators in TMEM, Q / K / V^T tiles by TMA, online softmax
 * in registers; csrc/kd_attn_tc.cu), used for N >= 256 tokens.  vt_scratch: kd_attn_vt_elems(B, J) fp16 elements of caller-owned
 * scratch (the kernel first writes V^T there: the K-major B operand of the P V product). */
int kd_attn_vt_elems(int B, int J);
int kd_attn_mqa_tc(const void* q, long ldq, const void* kv, void* vt_scratch, void* out, int B, int N, int J, int heads, float scale,
                   kd_stream_t stream);
int kd_attn_cross(const void* q, long ldq, const float* kv /* [B,Jc,2*heads*64] */, const float* null_kv /* [2,64] */,
                  void* out /* fp16 [B,N,heads*64] */, int B, int N, int Jc, int heads, float scale, kd_stream_t stream);

/* ------------------------------------------------------------------ linear attention (Unet(use_linear_attn=...), north_star (b))
 * replaces: imagen_pytorch.LinearAttention.forward after the 1x1 convolutions (which run on kd_conv_gemm):
 *   kd_dwconv3x3      : the depthwise 3x3 convolutions of to_q / to_k / to_v (groups = channels, no bias) on the concatenated
 *                       q|k|v map, NHWC fp16; w fp32 [C][3][3];
 *   kd_linattn_context: k <- softmax over positions (N pixels of columns [k_col, k_col+heads*64) of `qkv` plus J fp32 context-token
 *                       rows ctx_kv [B][J][2*heads*64] = k | v), ctx[b][h][d][e] = sum_n k[n,h,d] v[n,h,e] (fp32 [B][heads][64][64]);
 *                       pixel chunks are reduced in fixed order (chunk count = kd_linattn_blocks(N), independent of B); N = 0 = tokens
 *                       only (LinearCrossAttention);
 *   kd_linattn_apply  : out[n, h*64+e] = act(scale * sum_d softmax_d(q[n,h,:])[d] * ctx[h][d][e]), fp16 [B][N][heads*64].
 * The products are 64 x 64 per head (no tensor-core tile fits); the block is bound by one pass over q, k and v. */
int kd_dwconv3x3(const void* x, const float* w, void* y, int B, int H, int W, int C, kd_stream_t stream);
int kd_linattn_blocks(int N);
size_t kd_linattn_workspace_bytes(int B, int N, int heads);
int kd_linattn_context(const void* qkv, long ld, int k_col, int v_col, int B, int N, int heads, const float* ctx_kv, int J, float* workspace,
                       size_t ws_bytes, float* ctx, kd_stream_t stream);
int kd_linattn_apply(const void* q, long ld, int q_col, const float* ctx, void* out, int B, int N, int heads, float scale, int act,
                     kd_stream_t stream);

/* replaces: PerceiverAttention of the text-conditioning tower (Unet.attn_pool; train.py models only): fp32 multi-head
 *           attention of Nq latent queries over J keys, q [B,Nq,heads*64], kv [B,J,2*heads*64] (k | v), once per sample(). */
int kd_attn_small_f32(const float* q, const float* kv, float* out, int B, int Nq, int J, int heads, float scale, kd_stream_t stream);
/* out = a*x + b*y (fp32): residual adds of the conditioning towers and classifier-free guidance
 * (Unet.forward_with_cond_scale: null + (cond - null) * scale = scale*cond + (1-scale)*null). */
int kd_axpby(const float* x, const float* y, float a, float b, float* out, long n, kd_stream_t stream);

/* ------------------------------------------------------------------ init / final convolutions
 * replaces: CrossEmbedLayer (3 convs k=3,7,15 concatenated) via an im2col panel consumed by kd_conv_gemm mode 2,
 *           and Unet.final_conv (3x3, Cout = 3) on cat(x, lowres_cond_img). */
int kd_im2col_nchw(const float* x /* fp32 [B,C,H,W] */, int B, int C, int H, int W, int ksize, void* out /* fp16 [B*H*W, Kp] */,
                   int Kp, kd_stream_t stream);
/* CrossEmbedLayer without a panel, for <= 3 image channels per call and Cout in {64, 128}: tcgen05 implicit GEMM over
 * element-shifted halo rows kept in shared memory (see csrc/kd_init_conv.cu).  w_packed: fp16 [Cout][Kp], Kp =
 * kd_init_conv_kp(C, ksize), column (ky*C + c)*16 + kx holds the merged ksize x ksize filter tap (ky, kx) of channel c
 * (kx >= ksize: zero).  out = conv + bias + addend (addend: fp16 NHWC or NULL; more channels = chained calls). */
int kd_init_conv_kp(int C, int ksize);
int kd_init_conv(const float* x /* fp32 [B,C,H,W] */, int B, int C, int H, int W, int ksize, const void* w_packed, const float* bias,
                 const void* addend, void* out /* fp16 [B,H,W,Cout] */, int Cout, kd_stream_t stream);
/* w_split: the Ca-channel part of the filter pre-split into fp16 hi + lo for the tensor cores, kd_final_conv_pack_elems(Ca)
 * fp16 elements written once per model by kd_final_conv_pack. */
long kd_final_conv_pack_elems(int Ca);
int kd_final_conv_pack(const float* w /* fp32 [Cout][3][3][Ca+Cb] */, int Cout, int Ca, int Cb, void* w_split, kd_stream_t stream);
int kd_final_conv(const void* xa /* fp16 [B,H,W,Ca] */, int Ca, const float* xb /* fp32 NCHW [B,Cb,H,W] or NULL */, int Cb,
                  const float* w /* fp32 [Cout][3][3][Ca+Cb] */, const void* w_split, const float* bias,
                  float* out /* fp32 NCHW [B,Cout,H,W] */, int B, int H, int W, int Cout, kd_stream_t stream);

/* ------------------------------------------------------------------ K6 / K7: sampler update
 * replaces: Imagen.p_mean_variance + p_sample (x0 from eps / v, dynamic threshold = torch.quantile(|x0|, 0.95) per sample,
 *           clamp(min=1), q_posterior mean, ancestral noise), the RePaint blend / re-noise of p_sample_loop and the
 *           final clamp + paste + un-normalise.  All NCHW fp32, scalars computed by the caller in fp32. */
size_t kd_dynthresh_workspace_bytes(int B);
int kd_dynthresh(const float* x_t, const float* pred, int B, long n_per, int objective, float alpha, float sigma, long rank_lo,
                 long rank_hi, float weight, void* workspace, size_t ws_bytes, float* s_out /* [B] */, kd_stream_t stream);
int kd_ddpm_step(const float* x_t, const float* pred, const float* noise, const float* s /* [B] or NULL: static clamp */,
                 float* out, float* x0_out /* or NULL */, int B, long n_per, int objective, float alpha, float sigma,
                 float one_minus_c, float c, float alpha_next, float std, const float* renoise /* or NULL */, float rn_k1,
                 float rn_num, float rn_alpha, kd_stream_t stream);
int kd_inpaint_blend(float* img, const float* inpaint, const uint8_t* mask /* [B,HW] */, const float* noise /* or NULL */,
                     float alpha, float sigma, int B, int C, long HW, kd_stream_t stream);
int kd_finalize_image(float* img, const float* inpaint /* or NULL */, const uint8_t* mask, int B, int C, long HW,
                      kd_stream_t stream);
int kd_q_sample(const float* x0, const float* noise, float alpha, float sigma, float* out, long n, kd_stream_t stream);
/* counter-based N(0,1): Philox4x32-10 keyed by (seed, key), Box-Muller */
int kd_randn(float* out, long n, uint64_t seed, uint64_t key, kd_stream_t stream);

/* ------------------------------------------------------------------ K8: overlap-border pack for the patch-grid sampler
 * replaces: the inpaint canvas construction of generate_image_distributed (sample_ultra_res.py:149-170): writes the overlap
 *           strips of up to three finished neighbour patches into inpaint_patch [3,S,S] / inpaint_mask [S,S] (write order
 *           above -> side -> corner; the corner sets no mask bits).  Each neighbour is a strided strip view
 *           (element (c,y,x) = ptr[c*cs + y*rs + x]): above = bottom `overlap_pos` rows [3,ov,S], side = facing columns
 *           [3,S,ov], corner = facing box [3,ov,ov]; a full resident patch and a strip received over NVLink use the same call. */
int kd_border_pack(float* inpaint, uint8_t* mask, const float* above, long above_cs, long above_rs, const float* side, long side_cs,
                   long side_rs, const float* corner, long corner_cs, long corner_rs, int S, int overlap_pos, int orientation,
                   kd_stream_t stream);

/* ------------------------------------------------------------------ K9: peer mailbox of the patch-grid sampler (one process per GPU)
 * replaces: the mp.Manager dict through which generate_image_distributed hands finished patches to other workers
 *           (sample_ultra_res.py:125-131, 204-205; whole patches pickled through the CPU).  Here a producer copies exactly the
 *           overlap strip (or previous-stage patch) a dependent on ANOTHER GPU needs into that GPU's mailbox over NVLink (peer
 *           stores into CUDA-IPC mapped memory) and then raises a flag word there; the consumer's stream waits for the flag on
 *           the device.  One-sided: no collective, no send/recv matching, no host synchronisation.
 * The mailbox is the one buffer this library allocates itself (cudaMalloc: IPC export needs a whole allocation): kd_peer_alloc
 * returns zeroed memory; kd_peer_export writes the 64-byte IPC handle another process passes to kd_peer_open (same node). */
int kd_peer_alloc(size_t bytes, void** ptr);
int kd_peer_free(void* ptr);
int kd_peer_export(const void* ptr, uint8_t* handle /* [64] */);
int kd_peer_open(const uint8_t* handle /* [64] */, void** ptr);
int kd_peer_close(void* ptr);
/* dst[c][y][x] = src[c*cs + y*rs + x] (dst contiguous [C,rows,cols], normally peer memory), then *flag = value with release
 * semantics at system scope (all strip bytes are visible to a reader that observes the flag). */
int kd_strip_push(const float* src, long cs, long rs, int C, int rows, int cols, float* dst, uint32_t* flag, uint32_t value,
                  kd_stream_t stream);
/* Blocks `stream` (one spinning thread, ld.acquire.sys + nanosleep) until *flag >= value.  timeout_s > 0: gives up after that
 * many seconds and increments *status (the GPU is never left hanging; the host raises). */
int kd_flag_wait(const uint32_t* flag, uint32_t value, double timeout_s, uint32_t* status, kd_stream_t stream);

/* ------------------------------------------------------------------ N2: conditioning window of one patch
 * replaces: the per-patch torch.roll of the WHOLE zoomed image + gap fill + CenterCrop(1024) of get_cond_images
 *           (sample_ultra_res.py:358-395) by a gather of the P x P window only.  Output position o along an axis reads shifted
 *           position p = o + off (off = CenterCrop offset; outside [0, W) -> 0, the crop's zero padding); p is fill-coloured when
 *           shift > 0 ? p < shift : (shift < 0 ? p >= W + shift : true)  [shift == 0 fills everything: reference quirk kept];
 *           else zoomed[(p - shift) mod W].  channels_out 6 (version v2, :392-395) appends CenterCrop(patch_width) of the window
 *           (offset center_top) nearest-upsampled to P. */
int kd_cond_gather(const float* zoomed /* fp32 [3,W,W] */, int W, float* out /* fp32 [channels_out,P,P] */, int channels_out, int P, int off,
                   int shift_y, int shift_x, float fill, int patch_width, int center_top, kd_stream_t stream);

/* ------------------------------------------------------------------ N3: stitch
 * replaces: generate_high_res_image's canvas = F.interpolate(zoomed, bilinear) followed by row-major patch pastes where later
 *           patches overwrite earlier ones (sample_ultra_res.py:440-446; outpainting.py:236-241 with a zero canvas).
 * cell_index [n*n]: position of grid cell (i, j) in the patch list, -1 when the cell holds no patch.  A canvas pixel belongs to
 * the covering patch with the LARGEST list index (the survivor of the sequential pastes); kd_patch_paste writes only the pixels
 * its patch owns and kd_canvas_fill only pixels no patch covers (zoomed == NULL: zeros), so any number of GPUs can paste into
 * one canvas (peer memory) concurrently and the result equals the sequential loop bit for bit. */
int kd_canvas_fill(const float* zoomed /* fp32 [3,W,W] or NULL */, int W, float* canvas /* fp32 [3,Wc,Wc] */, int Wc, const int* cell_index,
                   int n, int patch_dist, int P, kd_stream_t stream);
int kd_patch_paste(const float* patch /* fp32 [3,P,P] */, float* canvas, int Wc, const int* cell_index, int n, int patch_dist, int P, int k,
                   int i, int j, kd_stream_t stream);

/* ------------------------------------------------------------------ overflow guard of the fp16 activation path
 * The reference runs fp32; this path stores UNet activations as fp16 with saturating conversions (+-65504).  Adds to *counter the
 * number of elements of the fp16 tensor x[n] whose magnitude is >= 65504 (clipped values) or that are not finite.  Debug aid:
 * Imagen.check_saturation runs it on every stored activation tensor of a sample() call. */
int kd_count_saturated(const void* x, long n, unsigned long long* counter, kd_stream_t stream);

/* ------------------------------------------------------------------ fp32 ("precise") path (kd_precise.cu)
 * The reference runs the whole UNet in fp32 (imagen_pytorch/imagen_pytorch.py, Unet.forward; ImagenTrainer(fp16=False) in
 * train_ultra_res_v_param.py:109-115).  These entry points compute the same operations as the fp16 tensor-core path on fp32 NHWC
 * activations with fp32 weights and fp32 accumulation on the CUDA cores: the "fp32 path" of BASELINE.json's north_star (parity
 * 1e-4).  Selected explicitly (Imagen.set_precision("fp32")); 20-50x slower than the tensor-core path.
 *
 * kd_conv_f32: Conv2d ksize x ksize / stride / pad over the channel concat [xa (Ca) | xb (Cb)], weights [Cout, ksize*ksize*(Ca+Cb)]
 * (tap-major, channel-minor); out = act(conv + bias) + addend * addend_scale[b, n]; out_mode 1 stores through PixelShuffle(2)
 * (conv channel n = (dy*2+dx) * Cout/4 + c).  Downsample's pixel-unshuffle + 1x1 conv is ksize 2 / stride 2 / pad 0. */
int kd_conv_f32(const float* xa, int Ca, const float* xb, int Cb, const float* w, const float* bias, const float* addend,
                const float* addend_scale /* [B, Cout] or NULL */, float* out, int B, int Hin, int Win, int Cout, int ksize, int stride,
                int pad, int act, int out_mode, kd_stream_t stream);
/* GroupNorm over the concat [xa | scale_b * xb] (Block.forward: groupnorm -> * (scale + 1) + shift -> SiLU): fp64 partial sums per
 * (sample, group, pixel chunk) -> mean / rstd -> apply.  kd_gn_chunks_f32(HW) chunks per group (a function of HW only). */
int kd_gn_chunks_f32(long HW);
int kd_gn_stats_f32(const float* xa, int Ca, const float* xb, int Cb, float scale_b, int B, long HW, int G,
                    double* partial /* [B, G, kd_gn_chunks_f32(HW), 2] */, kd_stream_t stream);
int kd_gn_finalize_f32(const double* partial, int B, long HW, int G, int group_size, float eps, float* mean_rstd /* [B, G, 2] */,
                       kd_stream_t stream);
int kd_gn_apply_f32(const float* x, float* y, int B, long HW, int C, int c_offset, int group_size, int G, float src_scale,
                    const float* mean_rstd, const float* gamma, const float* beta, const float* scale_shift /* rows [scale | shift] or NULL */,
                    long ss_stride, int ctot, int act, kd_stream_t stream);
/* GlobalContext (to_k logits, softmax over pixels, pooled = softmax @ x, h * gate + residual) */
int kd_rowdot_f32(const float* x /* [M, C] */, const float* w, const float* bias /* [1] or NULL */, float* out /* [M] */, long M, int C,
                  kd_stream_t stream);
int kd_softmax_pool_f32(const float* x /* [B, HW, C] */, const float* logits /* [B, HW] */, int B, long HW, int C, float* ml /* [B, 2] */,
                        float* part /* [B, kd_gn_chunks_f32(HW), C] */, float* pooled /* [B, C] */, kd_stream_t stream);
int kd_gate_residual_f32(const float* h, const float* gate /* [B, C] */, const float* res, float* out, int B, long HW, int C, kd_stream_t stream);
/* Softmax attention with head dim 64: q [B, N, ldq] (head h at column 64 h), key j of head h of sample b at
 * k + b * k_batch + j * ldk + h * k_head (k_head = 0: one shared K/V head, Attention; 64: per-head, CrossAttention), v likewise;
 * out [B, N, heads * 64].  Replaces Attention.forward / CrossAttention.forward after their projections. */
int kd_attn_f32(const float* q, long ldq, const float* k, long ldk, long k_batch, int k_head, const float* v, long ldv, long v_batch, int v_head,
                float* out, int B, int N, int J, int heads, float scale, kd_stream_t stream);
/* LinearAttention / LinearCrossAttention in fp32: depthwise 3x3 (zero padding, no bias; w [C,3,3]); then for every (sample, head)
 * ctx = softmax over the J key positions of k (per head-dim column) transposed times v, and out = act(scale * softmax_d(q) ctx).
 * q [B, N, ldq] (head h at column 64 h); k / v rows of length ldkv with head h at column 64 h, J positions, sample stride kv_batch. */
int kd_dwconv3x3_f32(const float* x /* NHWC */, const float* w, float* y, int B, int H, int W, int C, kd_stream_t stream);
int kd_linattn_f32(const float* q, long ldq, const float* k, const float* v, long ldkv, long kv_batch, int B, int N, int J, int heads, float scale,
                   int act, float* ctx /* [B, heads, 64, 64] workspace / result */, float* out /* [B, N, heads*64] */, kd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
