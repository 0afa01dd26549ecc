/*
 * disco_b200.h -- C ABI of libdisco_b200.so: the B200 (sm_100a) kernels behind the DISCO
 * colorization forward (`AnchorColorProb.forward`, reference models/model.py:103-199).
 *
 * The reference is pure Python/PyTorch and has NO FFI of its own, so each
 * entry point cites the reference *function* it replaces; the Python side that binds these symbols
 * with ctypes is disentangledcolorization_b200/_lib.py (shown in INTEGRATION.md).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller
 *     (the host side allocates through the torch caching allocator and passes data_ptr()).
 *   - every call is asynchronous on the `stream` argument (a cudaStream_t passed as void*).
 *   - every call returns 0 on success or a negative disco_status; disco_last_error() returns a
 *     thread-local message.  Nothing throws across the boundary.
 *   - one handle per (process, device); calls on one handle are not re-entrant (mirrors the
 *     reference's one single-threaded Python process per GPU, train_colorizer_ddp.py:22-31).
 *   - activations are NHWC, DISCO_F32 (exact path) or DISCO_BF16 (tensor-core path); the API-edge
 *     tensors (gray, affinity, logits, pred_colors ...) are NCHW fp32 exactly as the reference returns.
 */
#ifndef DISCO_B200_H_
#define DISCO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct disco_handle disco_handle;

enum disco_status {
  DISCO_OK = 0,
  DISCO_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  DISCO_ERR_CUDA = -2,      /* a CUDA runtime/driver call failed */
  DISCO_ERR_UNSUPPORTED = -3
};

enum disco_dtype { DISCO_F32 = 0, DISCO_BF16 = 1 };
enum disco_act { DISCO_ACT_NONE = 0, DISCO_ACT_RELU = 1, DISCO_ACT_LRELU = 2 };
/* heads write fp32 NCHW: softmax over 9 channels, tanh of 2 channels, or the 2 raw channels (HourGlass2.forward
 * standalone returns the pre-tanh map, models/network.py:134,144) */
enum disco_head { DISCO_HEAD_NONE = 0, DISCO_HEAD_SOFTMAX9 = 1, DISCO_HEAD_TANH2 = 2, DISCO_HEAD_RAW2 = 3 };
enum disco_conv_kind { DISCO_CONV3 = 0, DISCO_DECONV4 = 1 };

int disco_version(void);
/* sizes/offsets of the descriptor structs as compiled (for language bindings to verify their mirrors):
 * 0 sizeof(disco_conv_src), 1 sizeof(disco_conv_desc), 2 sizeof(disco_linear_desc), 3 offsetof(conv_desc, out),
 * 4 offsetof(conv_desc, bias_host) */
int disco_abi_size(int which);
const char* disco_last_error(void);
int disco_create(disco_handle** out, int device);
int disco_destroy(disco_handle* h);
/* number of kernel launches issued through this handle since creation / last reset */
int64_t disco_launch_count(disco_handle* h);
void disco_reset_launch_count(disco_handle* h);
/* accounts for kernels replayed from a CUDA graph that was captured from calls on this handle */
void disco_add_launch_count(disco_handle* h, int64_t n);

/* ---------------------------------------------------------------------------------------------
 * Fused convolution.  Replaces every nn.Conv2d / nn.ConvTranspose2d (+ bias, activation,
 * eval-mode BatchNorm, residual add, nn.Upsample, torch.cat, Softmax(1), tanh) call of
 * SpixelNet / ColorProbNet / HourGlass2 (reference models/network.py:10-47,66-101,125-313 and
 * models/model.py:196-197).
 *
 *   out[n,oy,ox,co] = post( act( bias[co] + res[n,oy,ox,co]
 *                     + sum_s sum_tap sum_ci W_s[tap][ci][co] * src_s[n, iy, ix, ci] ) )
 *   DISCO_CONV3  : iy = (oy*stride + ky - 1) >> up2_s   (zero outside the virtual (H<<up2) map)
 *   DISCO_DECONV4: iy = (oy + 1 - ky) / 2 for ky in 0..3 where (oy + 1 - ky) is even and in range
 *   post(v) = v*post_scale[co] + post_shift[co] (when non-NULL);  head: softmax over the 9
 *   channels or tanh of the 2 channels, written as fp32 NCHW.
 *
 * Weight packing (done by the host, see engine.py):
 *   dtype F32 : per source a block [taps][cin_s][cout] fp32, source s starting at src[s].w_off
 *   dtype BF16: tensor-core packing described in csrc/conv_tc.cu
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* ptr;   /* NHWC activations of this source (dtype of the descriptor; `gray_f32`=1 -> fp32) */
  int32_t H, W, C;   /* stored per-image dims */
  int32_t up2;       /* 1: source is nearest-upsampled x2 on the fly */
  int32_t is_f32;    /* 1: this source is fp32 even when the descriptor dtype is bf16 (the L channel) */
  int64_t w_off;     /* element offset of this source's weight block inside `weights` */
} disco_conv_src;

typedef struct {
  int32_t kind;      /* disco_conv_kind */
  int32_t stride;    /* 1 | 2 (DISCO_CONV3 only) */
  int32_t dtype;     /* disco_dtype of activations in/out */
  int32_t batch, Ho, Wo, Cout;
  int32_t n_src;
  disco_conv_src src[2];
  const void* weights;
  const float* bias;        /* [Cout] */
  const float* post_scale;  /* [Cout] or NULL */
  const float* post_shift;  /* [Cout] or NULL */
  const void* residual;     /* NHWC [batch,Ho,Wo,Cout] or NULL */
  int32_t act;              /* disco_act */
  float slope;              /* LeakyReLU negative slope */
  int32_t head;             /* disco_head */
  void* out;                /* NHWC (dtype) or, with a head, fp32 NCHW */
  const float* gray_weights; /* tensor-core path only: fp32 [9][Cout] weights of the fp32 1-channel source */
  /* Optional HOST copies of bias / post_scale / post_shift / gray_weights (same values as the device arrays; NULL = not
   * supplied).  When present, narrow (Cout <= 64) tensor-core layers pass them in the kernel-parameter block so the
   * epilogue reads them as constant-bank operands instead of through shared memory.  Read at every disco_conv call. */
  const float* bias_host;
  const float* post_scale_host;
  const float* post_shift_host;
  const float* gray_weights_host;
} disco_conv_desc;

int disco_conv(disco_handle* h, const disco_conv_desc* d, void* stream);

/* Tensor-core (tcgen05) routing of bf16 descriptors.  disco_conv sends a DISCO_BF16 descriptor to the tcgen05
 * implicit-GEMM kernel when disco_conv_tc_supported() is 1; `weights` must then hold the bf16 packing produced by
 * disco_conv_tc_pack_weights (host-side repack of the fp32 [tap][cin][cout] blocks; both pointers are HOST
 * pointers there, `src[].w_off` and `src[].C` of the descriptor describe the fp32 blocks).  Otherwise the
 * descriptor runs on the CUDA-core kernel with fp32 weights. */
int disco_set_tensor_core(disco_handle* h, int enable);
int disco_conv_tc_supported(disco_handle* h, const disco_conv_desc* d);
int64_t disco_conv_tc_weight_elems(const disco_conv_desc* d);
int disco_conv_tc_pack_weights(const disco_conv_desc* d, const float* w_f32_host, uint16_t* w_bf16_host);
/* Drops the cached launch plans (tensor maps, tap tables) of this handle's device.  The cache is keyed on the whole
 * descriptor, pointers included; a host executor that frees an activation workspace calls this so that plans of dead
 * buffers do not accumulate (the CLI's --no_resize mode sees a new image size per file).  Synchronises the device. */
int disco_conv_tc_cache_clear(disco_handle* h);
/* debug aid: per-role clock64 timeline of CTA 0 of the last tensor-core launch made with DISCO_TC_DEBUG=1 */
int disco_debug_timeline(long long* out_host);

/* ---------------------------------------------------------------------------------------------
 * Fused head of SpixelNet: conv0a -> conv0b -> conv1a (models/network.py:264-266,293-295; conv(no bias) + BatchNorm +
 * LeakyReLU each, BatchNorm folded) in one launch on the bf16 path.  The 16-channel intermediate of conv0a stays in
 * shared memory; only the tensors later layers read are written: out1 = conv0b output (NHWC bf16 [B,H,W,16]) and
 * a1 = conv1a output (NHWC bf16 [B,H/2,W/2,32]).
 *   gray fp32 [B,H,W];  w0a fp32 [9][16], w0b bf16 [9][16 co][16 ci], w1a bf16 [9][32 co][16 ci] (taps row-major,
 *   BatchNorm scale folded in), b0a/b0b/b1a fp32 folded biases;  slope = LeakyReLU negative slope (0.1)
 * ------------------------------------------------------------------------------------------- */
int disco_segnet_head(disco_handle* h, const float* gray, const float* w0a, const float* b0a, const uint16_t* w0b,
                      const float* b0b, const uint16_t* w1a, const float* b1a, float slope, int batch, int H, int W,
                      void* out1, void* a1, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Super-pixel pooling.  Replaces basic.poolfeat (models/basic.py:274-324) applied to
 * cat[pred_feats, input_colors] and basic.get_spixel_size (models/basic.py:327-335), i.e.
 * models/model.py:114-117,121, in one pass over the feature map.
 *   feats   : NHWC [B,H,W,C] (dtype), C == 64
 *   ab      : NCHW fp32 [B,2,H,W]
 *   affinity: NCHW fp32 [B,9,H,W]
 *   partial : workspace fp32 [B, H/16, W/16, 9, 68]
 *   tokens  : fp32 [B, S, 64] (S = H/16*W/16, row-major cells)   == feat_tokens
 *   spix_ab : fp32 NCHW [B,2,h,w]                                 == spix_colors (pooled)
 *   conf    : fp32 [B,S]  soft mass;  sizes: fp32 [B,S]  hard-assignment mass (spixel_sizes)
 * ------------------------------------------------------------------------------------------- */
int disco_poolfeat(disco_handle* h, int dtype, const void* feats, const float* ab, const float* affinity,
                   int batch, int H, int W, int C, float* partial, float* tokens, float* spix_ab,
                   float* conf, float* sizes, void* stream);

/* Replaces basic.upfeat (models/basic.py:338-376), models/model.py:195.
 *   tokens fp32 [B,S,C] -> out NHWC [B,H,W,C] (dtype), C == 64 */
int disco_upfeat(disco_handle* h, int dtype, const float* tokens, const float* affinity, int batch, int H, int W,
                 int C, void* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Token-path linear layer: Y = epi(X' W^T + b).  Replaces the nn.Linear / in_proj / out_proj /
 * LayerNorm calls of EncoderLayer (models/transformer2d.py:52-60), mid_word_prj / trg_word_prj
 * (models/model.py:134-135,187-189) and the trg_word_emb hint embedding (models/model.py:175-185).
 *   X [M,K] fp32 row-major; W [N,K]; b [N] or NULL.
 *   pos/pos_cols/S : X' = X + pos[row % S] for output columns < pos_cols (q,k of MHA), else X' = X
 *   col_scale/scale_cols : columns < scale_cols multiplied by col_scale after bias (q scaling)
 *   relu; residual R [M,N] added; ln_gamma/ln_beta: LayerNorm over N (requires N == 64), eps 1e-5
 *   hint_mask [M], labels [M], emb [314,N]: Y += hint_mask[row] * (emb[labels[row]] + emb[313])
 *   transpose_S > 0: output written as [M/S][N][S] (the reference's permute(1,2,0).view(N,C,h,w))
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* X; const float* W; const float* b;
  int32_t M, N, K;
  const float* pos; int32_t pos_cols; int32_t S;
  float col_scale; int32_t scale_cols;
  int32_t relu;
  const float* residual;
  const float* ln_gamma; const float* ln_beta;
  const float* hint_mask; const int32_t* labels; const float* emb;
  int32_t transpose_S;
  float* Y;
} disco_linear_desc;
int disco_linear(disco_handle* h, const disco_linear_desc* d, void* stream);

/* Fused tail of one EncoderLayer (models/transformer2d.py:55-59): x1 = LayerNorm1(x + att Wo^T + bo);
 * y = LayerNorm2(x1 + W2 relu(W1 x1 + b1) + b2).  att, x, y: fp32 [M,64]; Wo [64,64]; W1 [256,64]; W2 [64,256]. */
int disco_encoder_tail(disco_handle* h, const float* att, const float* x, float* y, int M, const float* wo,
                       const float* bo, const float* ln1_g, const float* ln1_b, const float* w1, const float* b1,
                       const float* w2, const float* b2, const float* ln2_g, const float* ln2_b, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused encoder stack: `n_layers` EncoderLayers (models/transformer2d.py:52-60) = one TransformerEncoder.forward
 * (models/transformer2d.py:17-28, use_dense_pos: pos added to q and k in every layer) in ONE launch.  Replaces the
 * per-layer disco_linear / disco_attention / disco_encoder_tail sequence on the bf16 path (models/model.py:133,186).
 * Tensor cores (mma.sync bf16, every operand split hi + lo -> fp32-grade products), tokens resident in registers,
 * one CTA per 128 tokens, images of more than 128 tokens span a thread-block cluster (S <= 1024).
 *   x_in, y   fp32 [B, S, 64] (may not alias);  pos fp32 [S, 64]
 *   w_packed, vec: produced by disco_encoder_stack_pack from the layers' fp32 parameters (HOST pointers there):
 *             w_packed bf16 [L][12][2][64][64], vec fp32 [L][832]
 *   kv_scratch bf16, disco_encoder_stack_scratch_elems(B, S) elements (device; contents undefined between calls)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* in_w;  const float* in_b;    /* self_attn.in_proj_weight [192,64], in_proj_bias [192] */
  const float* out_w; const float* out_b;   /* self_attn.out_proj.weight [64,64], bias [64] */
  const float* l1_w;  const float* l1_b;    /* linear1 [256,64], [256] */
  const float* l2_w;  const float* l2_b;    /* linear2 [64,256], [64] */
  const float* n1_w;  const float* n1_b;    /* norm1 [64], [64] */
  const float* n2_w;  const float* n2_b;    /* norm2 [64], [64] */
} disco_encoder_layer_weights;
int64_t disco_encoder_stack_scratch_elems(int batch, int S);
int disco_encoder_stack_pack(const disco_encoder_layer_weights* layers, int n_layers, uint16_t* w_packed_host,
                             float* vec_host);
int disco_encoder_stack(disco_handle* h, const float* x_in, const float* pos, const uint16_t* w_packed, const float* vec,
                        int n_layers, int batch, int S, uint16_t* kv_scratch, float* y, void* stream);

/* Multi-head self-attention core: softmax(q k^T) v per (image, head); q already scaled.
 * Replaces the attention inside nn.MultiheadAttention (models/transformer2d.py:36,54).
 *   qkv [B*S, 192] (q | k | v, head h = columns 8h..8h+7 of each third) -> out [B*S, 64] */
int disco_attention(disco_handle* h, const float* qkv, int batch, int S, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Anchor selection.  Replaces clusterkit.batch_kmeans_pytorch (models/clusterkit.py:31-58,
 * 99-208,253-269) + AnchorAnalysis.__call__ 'clustering' (models/anchor_gen.py:92-101).
 *   X        fp32 [B,S,64] tokens (enc_out)
 *   init_idx int32 [B,K]   host-drawn np.random.choice(S,K,replace=False), batch order
 *   draws    int32 [n_draws] host-pre-drawn torch.randint(S,(1,)) stream for empty clusters
 *   sizes    fp32 [B,S]    spixel_sizes
 *   assign   int32 [B,S]   final cluster ids (out);  hint_mask fp32 [B,S] (out)
 *   events   int32 [B+2]   (out) per-image number of draws consumed; [B] = total; [B+1] = error flag
 *   iters    int32 [B]     (out) Lloyd iterations executed
 * Draw order is the reference's (image-major, then iteration, then cluster): a speculative
 * parallel pass is followed by a sequential fix-up of the (rare) images whose draw offset moved.
 * ------------------------------------------------------------------------------------------- */
int disco_kmeans_anchor(disco_handle* h, const float* X, const int32_t* init_idx, const int32_t* draws,
                        int n_draws, const float* sizes, int batch, int S, int K, int iter_limit, float tol,
                        int32_t* assign, float* hint_mask, int32_t* events, int32_t* iters, void* stream);

/* Token labels.  mode 0: argmax over the 313 logits (== _sample_anchor_colors T=0 followed by
 * encode_ab2ind(...).max, models/anchor_gen.py:54-66 + models/model.py:161,166);
 * mode 1: nearest gamut bin of ab (== encode_ab2ind(ab).max, models/basic.py:177-194).
 *   logits fp32 [B,313,S] (the pal_logit layout) or ab fp32 NCHW [B,2,h,w];  labels int32 [B*S];
 *   colors fp32 NCHW [B,2,h,w] = q_to_ab[label]/110 (mode 0 only; NULL ok) */
int disco_token_labels(disco_handle* h, int mode, const float* src, const float* q_to_ab, int batch, int S,
                       int32_t* labels, float* colors, void* stream);

/* Diverse anchor colours (sampled_T > 0, the CLI's --diverse): for every token the T=0, T=1 and T=2 picks of
 * AnchorAnalysis._sample_anchor_colors (models/anchor_gen.py:54-90) among the 10 most probable bins.
 *   logits fp32 [B,313,S];  labels3 int32 [3][B*S] (picked bin = encode_ab2ind(...).max of the pick);
 *   colors3 fp32 [3][B,2,S] = q_to_ab[pick]/110 */
int disco_token_sample3(disco_handle* h, const float* logits, const float* q_to_ab, int batch, int S,
                        int32_t* labels3, float* colors3, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training-side losses (BASELINE config 5; first slice of SURVEY 8f N3/N4 -- the conv / transformer backward is not built).
 * disco_ce_rebalance: one token-level term of AnchorColorProbLoss.__call__ (models/loss.py:59-76): nn.CrossEntropyLoss
 * (mean, ignore_index -1) of a [B,313,S] logit map against hard labels, and d loss / d logits as autograd delivers it
 * through basic.RebalanceLoss (models/basic.py:120-134): (softmax - onehot) / n_valid * token_weights[token].
 *   logits fp32 [B,313,S]; labels int32 [B*S] (disco_token_labels mode 1 of the pooled GT colours, -1 = ignored);
 *   token_weights fp32 [B*S] (data['class_weight'] = ColorLabel.get_classweights(labels), models/basic.py:173-175);
 *   token_loss fp32 [B*S] (out);
 *   loss_out fp32 [2] = {mean CE, n_valid}; dlogits fp32 [B,313,S] or NULL
 * disco_spixel_recon_loss: SPixelLoss.__call__ (models/loss.py:17-30) after recon = upfeat(poolfeat(target, prob), prob):
 *   loss_out[0] = mean_pixels ||recon - target||_2 over channels [0, C-2), loss_out[1] = the same over the last 2 channels
 *   (the caller forms 10 * feat + 0.003 * pos / kernel_size); partial: fp32 workspace [n_partial][2]
 * ------------------------------------------------------------------------------------------- */
 /* ColorLabel.encode_ab2ind (models/basic.py:177-194): ab fp32 [B,2,S] (ab/110) -> soft code q fp32 [B,313,S]
  * (5 nearest bins, Gaussian sigma 5, normalised) */
int disco_encode_ab2ind(disco_handle* h, const float* ab, const float* q_to_ab, int batch, int S, float* q, void* stream);
int disco_ce_rebalance(disco_handle* h, const float* logits, const int32_t* labels, const float* token_weights, int batch, int S,
                       float* token_loss, float* loss_out, float* dlogits, void* stream);
int disco_spixel_recon_loss(disco_handle* h, const float* recon, const float* target, int batch, int C, int H, int W,
                            float* partial, int n_partial, float* loss_out, void* stream);

/* Lab -> sRGB uint8, the image the CLI saves (main/colorizer/inference.py:119-127 + utils/util.py:91-106:
 * L = (gray + 1) * 50, ab * 110, cv2.cvtColor(COLOR_LAB2RGB) on float32, * 255, astype(uint8)).
 *   gray fp32 [B,1,H,W], ab fp32 [B,2,H,W] (normalised, as the forward returns pred_colors);
 *   rgb uint8 [B, crop_h, crop_w, 3] = the top-left crop_h x crop_w pixels (--no_resize de-padding) */
int disco_lab2rgb_u8(disco_handle* h, const float* gray, const float* ab, int batch, int H, int W, int crop_h, int crop_w,
                     uint8_t* rgb, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Host helper (no device work): the k-means initialisation draws of clusterkit.initialize
 * (models/clusterkit.py:99-109), `np.random.choice(S, K, replace=False)` once per image in batch order, for `rows`
 * images in one call.  Bit-identical to numpy's legacy global RandomState: choice(replace=False) is
 * permutation(S)[:K], a Fisher-Yates shuffle of arange(S) from the top index down whose j = random_interval(i) is a
 * masked, rejection-sampled 32-bit MT19937 output.  The caller passes the generator state it read with
 * np.random.get_state() (key[624], pos) and writes the updated state back with np.random.set_state(), so numpy's stream
 * ends exactly where `rows` calls of np.random.choice would have left it.  Rows in [keep_lo, keep_hi) are written to
 * out[(keep_hi-keep_lo)*K] (a rank of a sharded job walks the whole global batch but keeps its slice).
 * 16 us -> ~1 us of host time per image: at 8 ranks x 64 images the python loop was on the critical path. */
int disco_host_choice_rows(uint32_t* mt_key, int32_t* mt_pos, int S, int K, int rows, int keep_lo, int keep_hi,
                           int32_t* out);

/* split_spixels of main/spixelseg/inference.py:67-75 on the neighbour-id grid of basic.init_spixel_grid
 * (models/basic.py:221-251): ids[n,y,x] = sum over the channels k with prob[n,k,y,x] == max_k prob of the id of the k-th
 * neighbour cell (edge-replicated) of the sp x sp cell that holds (y, x).  prob fp32 [B,9,H,W], ids int32 [B,1,H,W]. */
int disco_spixel_ids(disco_handle* h, const float* prob, int batch, int H, int W, int sp, int32_t* ids, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Perceptual-loss side ops (config 5; AnchorColorProbLoss._perceptual_loss, models/loss.py:45-49 and VGG19Loss,
 * models/loss.py:138-223).  The VGG19 convolutions themselves are disco_conv calls (3x3, bias, ReLU).
 *   disco_lab2rgb_norm: basic.lab2rgb (models/basic.py:431-475; normalised Lab = gray (L-50)/50 and ab/110, fp32 NCHW)
 *     -> RGB in [0,1] as fp32 NCHW (`rgb_nchw`, may be NULL) and/or VGG19Loss.normalize((rgb-mean)/std) as NHWC
 *     activations of `dtype` with `cpad` >= 3 channels, channels 3.. zero (`norm_nhwc`, may be NULL; mean3/std3 are HOST arrays).
 *   disco_rgb_norm: the normalisation / packing alone, for VGG19Loss.forward(x, y) on RGB images (B,3,H,W).
 *   disco_maxpool2: nn.MaxPool2d(2, 2) on NHWC activations (C % 8 == 0 for bf16, % 4 for fp32; H, W even).
 *   disco_l1_mean: nn.L1Loss(): out[0] = (accumulate ? out[0] : 0) + weight * mean|x - y| over n elements; `partial` is
 *     scratch of n_partial floats (the reduction order is fixed by n_partial: deterministic). */
int disco_lab2rgb_norm(disco_handle* h, const float* gray, const float* ab, int batch, int H, int W, float* rgb_nchw,
                       void* norm_nhwc, int dtype, int cpad, const float* mean3, const float* std3, void* stream);
int disco_rgb_norm(disco_handle* h, const float* rgb, int batch, int H, int W, void* norm_nhwc, int dtype, int cpad,
                   const float* mean3, const float* std3, void* stream);
int disco_maxpool2(disco_handle* h, int dtype, const void* x, int batch, int H, int W, int C, void* y, void* stream);
int disco_l1_mean(disco_handle* h, int dtype, const void* x, const void* y, long long n, float weight, int accumulate,
                  float* partial, int n_partial, float* out, void* stream);
/* AnchorColorProbLoss._laplace_gradient (models/loss.py:51-57, `with_grad=True`): out[0] = mean |Lap(target) - Lap(pred)| with
 * the 3x3 Laplacian [[1,1,1],[1,-8,1],[1,1,1]] per channel, no padding; fp32 NCHW.  grad_pred (may be NULL) receives
 * d out / d pred; sign_scratch holds batch*C*(H-2)*(W-2) floats (required with grad_pred). */
int disco_laplace_l1(disco_handle* h, const float* pred, const float* target, int batch, int C, int H, int W,
                     float* sign_scratch, float* partial, int n_partial, float* out, float* grad_pred, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Host helper (no device work): PNG encoder for the CLI's writer threads, replacing the reference's
 * `Image.fromarray(rgb).save(path, "PNG")` (utils/util.py:91-106 through main/colorizer/inference.py:131-135).
 * 8-bit RGB, H x W x 3 with `row_stride` bytes between rows.  Same decoded pixels as any PNG writer; the file bytes differ
 * (Sub filter on every row, one dynamic-Huffman deflate block of literals + distance-1 runs, no LZ77 search): ~0.5 ms per
 * 256 x 256 image against 7.7 ms (OpenCV/libpng level 1) and 18-26 ms (PIL level 6) on one host core.
 * disco_host_png_bound: capacity `out` must have.  disco_host_png_encode: file image into out[0 .. *out_len).
 * disco_host_png_write: encode + write to `path`. */
long long disco_host_png_bound(int H, int W);
int disco_host_png_encode(const uint8_t* rgb, int H, int W, long long row_stride, uint8_t* out, long long cap,
                          long long* out_len);
int disco_host_png_write(const char* path, const uint8_t* rgb, int H, int W, long long row_stride);

#ifdef __cplusplus
}
#endif
#endif /* DISCO_B200_H_ */
