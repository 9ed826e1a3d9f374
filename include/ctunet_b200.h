/* ctunet_b200.h — C ABI of the B200 (sm_100a) CTUNet hot path.
 *
 * The reference (shouwangzhe134/Hybrid-CTUNet) has no FFI/plugin interface: every operator below replaces a
 * torch library call made from the reference's Python modules (cited per entry point as file:line under
 * /root/reference).  The boundary is therefore this C ABI, bound from Python with ctypes
 * (hybrid_ctunet_b200/lib.py) underneath drop-in nn.Modules that keep the reference's constructors,
 * forward signatures and state_dict layout (hybrid_ctunet_b200/networks/*).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless stated otherwise; `stream` is a cudaStream_t passed as void*;
 *  - activations are channels-last ("token major"): a reference NCDHW tensor [B,C,X,Y,Z] lives here as
 *    [B,X,Y,Z,C] bf16; d1 is the fastest spatial extent (Z), d3 the slowest (X), d4 the batch;
 *  - all functions are asynchronous on `stream`, allocate nothing, keep no state and return 0 on success,
 *    a negative CTU_E_* code for a rejected argument, or a positive cudaError_t from the launch;
 *  - there is no CPU fallback: on a machine without an sm_100 GPU every launch returns an error.
 */
#ifndef CTUNET_B200_H
#define CTUNET_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTU_E_BADARG (-1)
#define CTU_E_UNSUPPORTED (-2)
#define CTU_E_DRIVER (-3)

#define CTU_OUT_BF16_ROWS 0 /* out[row*ldc + col] bf16 */
#define CTU_OUT_F32_ROWS 1  /* out[row*ldc + col] fp32 */
#define CTU_OUT_F32_CF 2    /* channel-first fp32: out[(batch*n_real + col)*S + s], S = d1*d2*d3 (NCDHW heads) */

#define CTU_ACT_NONE 0
#define CTU_ACT_GELU 1 /* exact erf GELU, vit.py:37 / hybrid_CTUNet.py:520 */

#define CTU_RES_NONE 0
#define CTU_RES_BF16 1
#define CTU_RES_F32 2
#define CTU_RES_GELU_BWD 3 /* `residual` = bf16 pre-activation x of a GELU: out = acc * gelu'(x)  (backward of
                              Linear -> GELU -> Linear: the GELU derivative rides on the input-gradient GEMM) */

/* One tcgen05 tensor-core contraction: out = epilogue(A (*) W^T).
 *   k1=k2=k3=1 : plain GEMM over tokens — nn.Linear (vit.py:36,39,59,62,117; hybrid_CTUNet.py:402,457,465,
 *                519,522,632-641,679), 1x1x1 Conv3d (resnet.py:96,100,197; hybrid_CTUNet.py:75; UnetOutBlock).
 *   k1=k2=k3=3 : stride-1 "same" 3x3x3 Conv3d as an implicit GEMM, the 27 taps fetched as shifted TMA boxes
 *                with hardware zero fill for the padding (resnet.py:98; hybrid_CTUNet.py:57-74).
 *   convt_cout>0: ConvTranspose3d with kernel == stride (u3,u2,u1), padding 0 (hybrid_CTUNet.py:232-240,
 *                286-294): a GEMM whose N axis is (sub-voxel, Cout) scattered to the up-sampled grid.
 * W is packed bf16 [n_pad][k_total] with k_total = k3*k2*k1*a_c ordered (tap, channel); tap = (t3*k2+t2)*k1+t1.
 */
typedef struct ctu_gemm_desc {
  const void* a;        /* bf16 [d4][d3][d2][d1][lda], first a_c channels of each row are used */
  const void* w;        /* bf16 [n_pad][k_total] */
  void* out;            /* see out_mode */
  const float* bias;    /* fp32 [n_real] or NULL */
  const void* residual; /* same row indexing as out (ldr), bf16 or fp32, or NULL */
  double* stats;        /* fp64 [d4][stats_ld][2] (sum, sum of squares per batch & column), accumulated with
                           atomics over valid rows — feeds InstanceNorm3d (resnet.py:97; hybrid_CTUNet.py:85) */
  int32_t a_c, lda;
  int32_t d1, d2, d3, d4;
  int32_t b1, b2, b3;   /* tile box, b1*b2*b3 == 128 */
  int32_t k1, k2, k3;   /* filter extents, all 1 or all 3 */
  int32_t n_pad, n_real, k_total;
  int32_t block_n;      /* 16, 32, 64, 128 or 256; divides n_pad */
  int32_t out_mode, ldc;
  int32_t act;
  int32_t res_mode, ldr;
  int32_t convt_cout, u1, u2, u3;
  int32_t stats_ld;     /* columns per batch in `stats` (>= n_real) */
  int32_t out_col0;     /* first output column inside the ldc-wide output rows (concat-by-offset) */
  int32_t a_c_live;     /* 0, or a multiple of 16 <= a_c: channels [a_c_live, a_c) of every `a` row are known to be zero
                           (ResNet layer-1 planes = 32 live in 64-channel rows, resnet.py:181-186): the 3x3x3 kernel skips
                           their K steps */
  const void* w_x3;     /* NULL, or (k = 3, block_n = 64) the CTU_PACK_X3_FROM_PACKED copy of `w`: enables the kernel that
                           computes two output x-planes per tile with N = 128 instructions (umma_conv3_halo.cu) */
} ctu_gemm_desc;

int ctu_umma_gemm(const ctu_gemm_desc* desc, void* stream);

/* Fused token FeedForward for inference (hybrid_CTUNet.py:513-526 inside Residual :434-440, C = 128 stage):
 *   out[r,:] = residual[r,:] + W2 · GELU(W1 · a[r,:] + b1) + b2      a = LayerNorm(x) (ctu_layernorm), residual = x
 * a / residual / out: bf16 rows of C channels (row strides lda / ldr / ldc elements, multiples of 8); w1: bf16 [hidden][C],
 * w2: bf16 [C][hidden] (nn.Linear layouts, K contiguous); b1 / b2 fp32.  The [M, hidden] activation never leaves the SM
 * (TMEM accumulator -> GELU -> shared-memory operand of the second tcgen05 GEMM).  C == 128, hidden % 128 == 0 and hidden <= 512, else
 * CTU_E_UNSUPPORTED (the caller then runs the two-GEMM path). */
int ctu_ffn_fused(const void* a, long long lda, const void* w1, const float* b1, const void* w2, const float* b2,
                  const void* residual, long long ldr, void* out, long long ldc, long long M, int C, int hidden,
                  void* stream);

/* InstanceNorm3d statistics (resnet.py:97,99,101; hybrid_CTUNet.py:85-87): stats[b][c] += (sum, sum of squares)
 * over the S voxels of batch item b.  x: bf16 [B][S][ldx]; stats: fp64 [B][stats_ld][2], zeroed by the caller.
 * The tensor-core kernel accumulates the same format in its epilogue; this standalone pass serves the
 * CUDA-core convolutions and sub-sampled tensors. */
int ctu_in_stats(const void* x, int ldx, int B, long long S, int C, double* stats, int stats_ld, void* stream);

/* out = act( IN(x) [+ IN(res) | + res] ), IN(v) = (v - mean) * rsqrt(var + eps) with biased variance, mean/var
 * derived from the fp64 accumulators; act != 0 selects LeakyReLU(slope) (resnet.py:110-124;
 * hybrid_CTUNet.py:95-104).  res == NULL: no residual; rstats == NULL: residual added as is; else residual is
 * instance-normalised with its own statistics (downsample / conv3+norm3 branch). */
int ctu_in_apply(const void* x, int ldx, const double* xstats, int xs_ld, const void* res, int ldr,
                 const double* rstats, int rs_ld, void* out, int ldo, int B, long long S, int C, float eps, int act,
                 float slope, void* stream);

/* nn.LayerNorm over the last dim C (vit.py:35,55,116,118; hybrid_CTUNet.py:456,518,630-631), rows of M.
 * x/out are fp32 or bf16 (flags); `add` (fp32 [add_rows][C], row % add_rows) is added after the affine —
 * the ViT position embedding `x += pos_embedding` (vit.py:133). */
int ctu_layernorm(const void* x, int x_is_f32, long long ldx, const float* gamma, const float* beta, const float* add,
                  long long add_rows, void* out, int out_is_f32, long long ldo, long long M, int C, float eps,
                  void* stream);

/* ViT patchify 'b c (h 16) (w 16) (f pf) -> b (h w f) (p1 p2 pf c)' (c = 1) fused with LayerNorm(256*pf)
 * (vit.py:115-116).  img: fp32 [B][X][Y][Z]; out: bf16 [B*tokens][256*pf]. */
int ctu_patchify_ln(const float* img, int B, int X, int Y, int Z, int pf, const float* gamma, const float* beta,
                    void* out, float eps, void* stream);

/* Per-token middle of pixelweight_attention, the binary cross-weight fusion (hybrid_CTUNet.py:658-665).
 * qkv1/qkv2: bf16 [T][3C] (= [q|k|v] from to_qkv1 / to_qkv2); out: bf16 [T][C]. */
int ctu_pwa_fuse(const void* qkv1, const void* qkv2, void* out, long long T, int C, int dim_head, void* stream);

/* out[b,x,y,z,:] = in[b, x*s3, y*s2, z*s1, :] on channels-last bf16 (input side of strided convs). */
int ctu_subsample(const void* in, int ldi, int i1, int i2, int i3, void* out, int ldo, int s1, int s2, int s3, int C,
                  int B, void* stream);

/* softmax(Q K^T / sqrt(dh) + bias) V.  mode 0: ViT attention over `windows` groups of n consecutive rows
 * (vit.py:66-78); mode 1 / 2: MultiAxisAttention over the block '(h h1)' / grid '(h1 h)' partition of a
 * [batch, X, Y, Z] token grid into w^3 windows with additive relative-position bias fp32 [heads][n][n]
 * (hybrid_CTUNet.py:481-511, 559-567).  lse (fp32 [rows][heads], or NULL) receives the base-2 log-sum-exp of
 * the scaled, biased scores of every (row, head) — what the backward pass recomputes the probabilities from. */
int ctu_attention(const void* qkv, int ld_qkv, int C, int dim_head, void* out, int ldo, const float* bias, int n,
                  int windows, int mode, int batch, int X, int Y, int Z, int w, float* lse, void* stream);

/* Conv3d with one input channel on CUDA cores: ResNet stem (resnet.py:150-155) and vit_encoder0 conv1/conv3
 * (hybrid_CTUNet.py:57-83).  x: fp32 [B][X][Y][Z]; w: fp32 [kx*ky*kz][64]; out: bf16 channels-last. */
int ctu_conv_cin1(const float* x, const float* w, void* out, int ldo, int cout, int B, int X, int Y, int Z, int kx,
                  int ky, int kz, int sx, int sy, int sz, int px, int py, int pz, void* stream);

/* Sliding-window Gaussian blend (trainer_CTUNet.py:541-549, trainer_CUNet.py:388-392). */
int ctu_blend_accumulate(const float* logits0, const float* logits1, const float* imp, float* acc0, float* acc1, int C,
                         int r3, int r2, int r1, int X, int Y, int Z, int x0, int y0, int z0, void* stream);
int ctu_blend_count(const float* imp, float* cnt, int r3, int r2, int r1, int X, int Y, int Z, int x0, int y0, int z0,
                    void* stream);
int ctu_blend_normalize(const float* acc, const float* cnt, float* out, int C, long long vox, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Backward pass (training step, trainer_CTUNet.py:87-109: loss.backward() through the modules above).
 * The reference gets these from torch autograd over cuDNN / cuBLAS; here each is an explicit kernel.
 * Input gradients of the contractions reuse ctu_umma_gemm with transposed / tap-flipped weights.
 * ------------------------------------------------------------------------------------------------------------ */

/* Weight gradient of a plain GEMM / 1x1x1 Conv3d (k=1) or a stride-1 "same" 3x3x3 Conv3d (k=3):
 *   dw[(tap*x_c + ci)][co] += sum over voxels v of x[v + tap - pad][ci] * dy[v][co]
 * (the transpose of the packed forward weight [n][k_total]).  tcgen05 contraction over the voxel axis with both
 * operands MN-major; split over voxel chunks, accumulated with fp32 reductions, so dw must be zeroed (or hold a
 * running sum) before the call.  Replaces cudnn wgrad / cuBLAS for every get_conv_layer and nn.Linear on the path. */
typedef struct ctu_wgrad_desc {
  const void* x;   /* bf16 [d4][d3][d2][d1][ldx], first x_c channels used; x_c % 64 == 0 */
  const void* dy;  /* bf16 [d4][d3][d2][d1][ldy], first n channels used */
  float* dw;       /* fp32 [k1*k2*k3*x_c][ldw], columns [0, n) accumulated */
  int32_t x_c, ldx;
  int32_t n, ldy;
  int32_t ldw;
  int32_t d1, d2, d3, d4;
  int32_t b1, b2, b3; /* tile box, b1*b2*b3 == 128 */
  int32_t k1, k2, k3; /* all 1 or all 3 */
  int32_t block_n;    /* 64, 128 or 256 */
} ctu_wgrad_desc;

int ctu_umma_wgrad(const ctu_wgrad_desc* desc, void* stream);

/* Backward of ctu_in_apply: out = act(IN(x) [+ res | + IN(res)]).  With g = dout * act'(out):
 *   dx = rstd_x * (g - mean(g) - xhat * mean(g*xhat)); dres = g (res_mode 1) or the same formula with the
 *   residual's statistics (res_mode 2).  ctu_in_bwd_stats accumulates sums[b][c] = (sum g, sum g*xhat, sum g*rhat, -)
 *   as fp64 [B][C][4] (zeroed by the caller; pass rstats == NULL unless res_mode == 2); ctu_in_bwd_apply consumes it.
 *   With act != 0 and no residual, x may be NULL: xhat is then recovered from the output, xhat = LeakyReLU^-1(out),
 *   which saves one tensor read per pass. */
int ctu_in_bwd_stats(const void* dout, int ldd, const void* out, int ldo, const void* x, int ldx, const double* xstats,
                     int xs_ld, const void* res, int ldr, const double* rstats, int rs_ld, int B, long long S, int C,
                     float eps, int act, float slope, double* sums, void* stream);
int ctu_in_bwd_apply(const void* dout, int ldd, const void* out, int ldo, const void* x, int ldx, const double* xstats,
                     int xs_ld, const void* res, int ldr, const double* rstats, int rs_ld, int res_mode, int B,
                     long long S, int C, float eps, int act, float slope, const double* sums, void* dx, int lddx,
                     void* dres, int lddr, void* stream);

/* nn.LayerNorm backward.  dy: bf16 [M][C].  dx = LN'(dy) (+ dx_in, the gradient already flowing along the residual
 * stream, fp32 or bf16) written as fp32 and/or bf16; dgamma / dbeta fp32 [C] are accumulated. */
int ctu_layernorm_bwd(const void* x, int x_is_f32, long long ldx, const float* gamma, const void* dy, long long ldd,
                      const void* dx_in, int dxin_is_f32, long long ld_in, float* dx_f32, long long ld_f, void* dx_bf16,
                      long long ld_b, float* dgamma, float* dbeta, long long M, int C, float eps, void* stream);

/* Exact-erf GELU on contiguous bf16 (n % 8 == 0) and its derivative: dx = dy * gelu'(x). */
int ctu_gelu(const void* x, void* y, long long n, void* stream);
int ctu_gelu_bwd(const void* x, const void* dy, void* dx, long long n, void* stream);

/* Backward of ctu_pwa_fuse: dqkv1 / dqkv2 bf16 [T][3C] from dout bf16 [T][C]. */
int ctu_pwa_fuse_bwd(const void* qkv1, const void* qkv2, const void* dout, void* dqkv1, void* dqkv2, long long T, int C,
                     int dim_head, void* stream);

/* out[c] += sum over the M rows of x[row][c], x bf16 or fp32 [M][ldx], N columns (bias gradients, position-embedding
 * gradient, relative-position-bias gradient over windows). */
int ctu_colsum(const void* x, int x_is_f32, long long ldx, long long M, long long N, float* out, void* stream);

/* Backward of a logits head — UnetOutBlock (1x1x1 conv + bias, hybrid_CTUNet.py:781-783,810) / DecoderLinear
 * (hybrid_CTUNet.py:671-691) — in one pass: replaces torch's conv/linear backward (input gradient, weight gradient,
 * bias gradient) for the C -> n_cls heads.  g: NCDHW fp32 logit gradient [B][ncls][S]; a: the head's bf16
 * channels-last input [B*S][lda] (C channels, C in {64, 128, 256}); w: the fp32 parameter [ncls][C];
 * da [B*S][ldda] bf16 = g W (added to its previous contents when accumulate != 0); dw fp32 [C][ldw] += a^T g;
 * db fp32 [ncls] += column sums of g.  ncls <= 16. */
int ctu_head_bwd(const float* g, const void* a, long long lda, const float* w, void* da, long long ldda,
                 int accumulate, float* dw, int ldw, float* db, int B, long long S, int C, int ncls, void* stream);

/* NCDHW fp32 [B][C][S] -> channels-last bf16 [B][S][ldd] with channels [C, cpad) zero (logit gradients). */
int ctu_cf_to_cl(const float* src, void* dst, int B, int C, long long S, int ldd, int cpad, void* stream);

/* in [B][X*u3][Y*u2][Z*u1][ldi] (C channels) -> out [B][X][Y][Z][u3*u2*u1*C], column (sub*C + c),
 * sub = (a3*u2 + a2)*u1 + a1: gradient of a kernel==stride ConvTranspose3d / pixel shuffle as a GEMM output. */
int ctu_space_to_depth(const void* in, int ldi, void* out, int B, int X, int Y, int Z, int u3, int u2, int u1, int C,
                       void* stream);

/* Backward of ctu_subsample: dfull[b, x*s3, y*s2, z*s1, :] (+)= dsub[b,x,y,z,:]; without `accumulate` every other
 * position of dfull is zeroed. */
int ctu_subsample_bwd(const void* dsub, int lds, void* dfull, int ldf, int i1, int i2, int i3, int s1, int s2, int s3,
                      int C, int B, int accumulate, void* stream);

/* im2col of a single-channel fp32 volume: out bf16 [B][Xo][Yo][Zo][kpad], column = tap (x-major), zero padded — the
 * activation operand of ctu_umma_wgrad for the C_in = 1 convolutions (resnet.py:150-155; hybrid_CTUNet.py:57-83). */
int ctu_im2col_cin1(const float* img, void* out, int B, int X, int Y, int Z, int kx, int ky, int kz, int sx, int sy,
                    int sz, int px, int py, int pz, int kpad, void* stream);

/* dst[row][0..C) += src[row][0..C), src and dst independently bf16 or fp32 rows (C % 8 == 0); dst(bf16) = src(fp32). */
int ctu_accumulate(const void* src, int src_is_f32, long long lds, void* dst, int dst_is_f32, long long ldd, long long M,
                   int C, void* stream);
int ctu_cast_f32_bf16(const float* src, long long lds, void* dst, long long ldd, long long M, int C, void* stream);

/* dgamma / dbeta (fp32 [256*pf], accumulated) of the LayerNorm fused into ctu_patchify_ln. */
int ctu_patchify_ln_bwd(const float* img, int B, int X, int Y, int Z, int pf, const void* dtok, float* dgamma,
                        float* dbeta, float eps, void* stream);

/* Attention backward.  ctu_attention_delta: delta[row][head] = sum_d dO*O.  ctu_attention_bwd: dK, dV written to
 * dqkv (bf16 [rows][3C] = dq|dk|dv), dQ ACCUMULATED into dq_f32 (fp32 [rows][C], zeroed by the caller); biasT is the
 * additive bias indexed [head][key][query]; ds_out (bf16 [windows][heads][n][n], [key][query], or NULL) receives
 * dScores for the relative-position-bias gradient. */
int ctu_attention_delta(const void* o, long long ldo, const void* dout, long long ldd, float* delta, long long rows, int C,
                        int dim_head, void* stream);
int ctu_attention_bwd(const void* qkv, int ld_qkv, int C, int dim_head, const void* dout, int ldd, const float* lse,
                      const float* delta, const float* biasT, void* dqkv, int ld_dqkv, float* dq_f32, void* ds_out, int n,
                      int windows, int mode, int batch, int X, int Y, int Z, int w, void* stream);

/* Fused Dice-CE loss (trainer_CTUNet.py:92-103 calls monai.losses.DiceCELoss(to_onehot_y, softmax, squared_pred) on
 * five heads).  logits: fp32 NCDHW [B][C][S]; target: fp32 labels [B][S] (class index stored as float, as the
 * reference's loaders yield them); C in {2,3,4,14}.
 * ctu_dice_ce_fwd accumulates sums = double [B][C][3] (sum p*y, sum p^2, sum y) followed by ONE double, the sum over
 * all voxels of -log softmax[label] (zeroed by the caller).  ctu_dice_ce_bwd writes
 *   dlogits_c = p_c * (a_c - sum_k a_k p_k) + ce_scale * (p_c - y_c),  a_k = coef[b][k][0] * y_k + coef[b][k][1] * p_k
 * with coef fp32 [B][C][2] and ce_scale (device scalar) built by the caller from the forward sums and the upstream
 * gradient. */
int ctu_dice_ce_fwd(const float* logits, const float* target, int B, int C, long long S, double* sums, void* stream);
int ctu_dice_ce_bwd(const float* logits, const float* target, int B, int C, long long S, const float* coef,
                    const float* ce_scale, float* dlogits, void* stream);

/* Ensemble of two blended logit volumes (test_CTUNet.py:236-251; test_CTUNet_final.py:547-552): p1, p2 fp32 [C][V];
 * mask = argmax((softmax(p1) + softmax(p2)) / 2), mask1 / mask2 = argmax of each head (uint8 [V], any may be NULL);
 * with labels (fp32 [V]) and counts (uint64 [3][C][3], zeroed by the caller) also the per-class Dice counts
 * (|pred == c and label == c|, |pred == c|, |label == c|) of (ensemble, head 1, head 2).  C = 14. */
int ctu_ensemble_argmax(const float* p1, const float* p2, int C, long long V, uint8_t* mask, uint8_t* mask1, uint8_t* mask2,
                        const float* labels, unsigned long long* counts, void* stream);

/* Multi-tensor weight packing and gradient unpacking: ONE launch per direction over a DEVICE-resident item table
 * (the reference keeps a single fp32 layout for cuDNN / cuBLAS; the tensor-core kernels here use packed bf16 layouts
 * and produce transposed-packed fp32 gradients, so every training step re-packs ~800 matrices and unpacks ~400).
 * ctu_pack_weights : dst = bf16 [rows][cols] (zero padded), src = fp32 parameter, kind = layout map:
 *   LIN / LIN_T   (a = N, b = K)            nn.Linear / 1x1x1 conv [N][K] and its transpose
 *   CONV3 / _T    (a = co, b = ci)          [co][27*cip] tap-major and the tap-flipped [cip][27*cop] of the dgrad
 *   CONVT / _T    (a = ci, b = co, c = k^3) ConvTranspose3d k == s as [k^3*co][ci] and its transpose
 *   PS / PS_T     (a = co, b = corg, c = k^3) pixel-shuffle Linear as the block-diagonal [k^3*co][corg*k^3]
 *   CIN1          (a = co, b = taps)        C_in = 1 convolution as [co][taps padded]
 * ctu_unpack_grads : dst = fp32 gradient in the parameter's own layout (rows = its element count), src = fp32
 *   accumulator written by ctu_umma_wgrad / ctu_colsum with row pitch cols; kinds LIN, CONV3 (c = cip), CONVT, PS,
 *   PS_BIAS (a = co, c = k^3), CIN1, VEC.
 * A thread-task moves the innermost run of the item's index map (ctu_pack_item_tasks gives the task count of an item);
 * unit0 = index of the item's first work unit (256 tasks per unit); items sorted by unit0. */
#define CTU_PACK_LIN 0
#define CTU_PACK_LIN_T 1
#define CTU_PACK_CONV3 2
#define CTU_PACK_CONV3_T 3
#define CTU_PACK_CONVT 4
#define CTU_PACK_CONVT_T 5
#define CTU_PACK_PS 6
#define CTU_PACK_PS_T 7
#define CTU_PACK_CIN1 8
#define CTU_PACK_PS_BIAS 9
#define CTU_PACK_VEC 10
/* "Paired" layouts of the 32-channel ResNet layer-1 bottlenecks (resnet.py:181-186, planes = 32): two z-neighbouring voxels
 * share one dense 64-channel row (slot s = z & 1) instead of 32 live + 32 zero-padded channels per voxel, so the tensor has
 * half the rows and no padding.  A 1x1x1 convolution becomes a block-diagonal GEMM, a 3x3x3 convolution a 3x3x3 convolution
 * over pairs: z pair-tap pz and slots (s_in, s_out) select the real tap dz = 2 (pz - 1) + s_in - s_out (zero if |dz| > 1).
 *   PAIR_LIN / _T   (a = co, b = ci)  [2 co][2 ci] block diagonal, and its transpose
 *   PAIR_CONV3 / _T (a = co, b = ci)  [2 co][27 * 2 ci], and the tap-flipped [2 ci][27 * 2 co] of the input gradient
 * ctu_unpack_grads sums the blocks that hold the same real weight (kinds PAIR_LIN, PAIR_CONV3). */
#define CTU_PACK_PAIR_LIN 11
#define CTU_PACK_PAIR_LIN_T 12
#define CTU_PACK_PAIR_CONV3 13
#define CTU_PACK_PAIR_CONV3_T 14
/* Re-laid copy of an already packed 3x3x3 weight (src = bf16 [n_pad = a][27 * a_c], a_c = b, tap-major K; run in a SECOND
 * ctu_pack_weights launch after the one that produced src): dst = bf16 [9 (y,z)-taps][a / 64 N tiles][3 x-taps in the order
 * dx = +1, 0, -1][64 rows] x a_c columns, so that the weights of two adjacent x-taps are 128 consecutive rows. */
#define CTU_PACK_X3_FROM_PACKED 15
typedef struct ctu_pack_item {
  const void* src;
  void* dst;
  int32_t kind;
  int32_t rows, cols;
  int32_t a, b, c;
  int64_t unit0;
} ctu_pack_item;
long long ctu_pack_item_tasks(int unpack, int kind, int rows, int cols, int a, int b, int c);
int ctu_pack_weights(const ctu_pack_item* items_dev, int n_items, long long total_units, void* stream);
int ctu_unpack_grads(const ctu_pack_item* items_dev, int n_items, long long total_units, void* stream);

/* InstanceNorm statistics of a PAIRED tensor (see CTU_PACK_PAIR_*): columns c and c + half of stats [B][ld][width] (fp64
 * sums written by the GEMM epilogue / ctu_in_bwd_stats) belong to the same channel; both become (v[c] + v[c+half]) * scale.
 * scale = 0.5 for consumers that keep the paired rows (they divide by the paired row count), 1.0 for consumers that read the
 * tensor as un-paired rows of `half` channels. */
int ctu_stats_fold(double* stats, int B, int ld, int half, int width, double scale, void* stream);

/* The scalar of the training loss and the coefficients of its backward pass from the sums ctu_dice_ce_fwd accumulated, for
 * up to CTU_LOSS_MAX_HEADS heads at once (trainer_CTUNet.py:92-103: loss = sum_h weight_h * DiceCE_h):
 *   loss      = sum_h weight_h * ( lambda_dice * mean_{b,c}[1 - (2 I + nr) / (P + Y + dr)] + lambda_ce * CE_h / (B S_h) )
 *   coef[h]   = per (b, c) the two factors ctu_dice_ce_bwd expects, for a unit upstream gradient, weight_h folded in
 *   ce_scale[h] = weight_h * lambda_ce / (B S_h)
 * sums_off / coef_off: element offsets of head h inside `sums` (doubles, [B][C][3] + 1 per head) and `coef` (floats). */
#define CTU_LOSS_MAX_HEADS 8
typedef struct ctu_loss_heads {
  int32_t n_heads;
  int32_t B[CTU_LOSS_MAX_HEADS], C[CTU_LOSS_MAX_HEADS];
  int64_t S[CTU_LOSS_MAX_HEADS];
  int64_t sums_off[CTU_LOSS_MAX_HEADS], coef_off[CTU_LOSS_MAX_HEADS];
  double weight[CTU_LOSS_MAX_HEADS];
  double lambda_dice, lambda_ce, smooth_nr, smooth_dr;
} ctu_loss_heads;
int ctu_dice_ce_finalize(const ctu_loss_heads* heads, const double* sums, float* loss, float* coef, float* ce_scale,
                         void* stream);

/* dst[b][xo][yo][zo] = src[b][ix[xo]][iy[yo]][iz[zo]] (fp32): the deep-supervision label volumes of trainer_CTUNet.py:93-94
 * (scipy.ndimage.zoom(order=0) == a gather with fixed index tables), one launch per target.  A negative table entry marks a
 * sample scipy places outside the volume (mode='constant', cval=0): the output there is 0. */
int ctu_gather3d(const float* src, float* dst, int B, int X, int Y, int Z, int Xo, int Yo, int Zo, const int* ix,
                 const int* iy, const int* iz, void* stream);

/* InstanceNorm sums of the pointwise single-channel convolution r[v][c] = x[v] * w[c] (vit_encoder0's conv3 + norm3 residual
 * branch, hybrid_CTUNet.py:75,88-91 with in_channels = 1) WITHOUT a pass over r: stats[b][c] = (w_c sum_v x, w_c^2 sum_v x^2).
 * x: fp32 [B][S]; w: fp32 [C]; mom: fp64 [B][2] scratch, zeroed by the caller; stats: fp64 [B][stats_ld][2] (written). */
int ctu_cin1_k1_stats(const float* x, const float* w, int B, long long S, int C, double* mom, double* stats, int stats_ld,
                      void* stream);

/* Post-processing of a predicted label volume (test_CTUNet_final.py:132-190, remove_all_but_the_largest_connected_component,
 * one class or class group per call): the voxels whose label is in the set `member` (uint8 [256], non-zero = member) are
 * split into connected components (6-connectivity = scipy.ndimage.label's default structure); every component that is not
 * (one of) the largest — and, with has_min, whose size count * volume_per_voxel is < min_valid (float64, as the reference
 * compares) — is set to 0 in `image` (uint8 [X][Y][Z], in place).  Scratch: parent, sizes int32 [X*Y*Z]; summary int32 [4] =
 * (number of components, voxels of the largest, voxels of the largest one removed, 0).  X*Y*Z < 2^31. */
int ctu_cc_filter_largest(uint8_t* image, const uint8_t* member, int X, int Y, int Z, double volume_per_voxel, int has_min,
                          double min_valid, int* parent, int* sizes, int* summary, void* stream);

/* `Invertd` of the evaluation scripts (test_CTUNet.py:162-199; test_CTUNet_final.py:470-505; the inverse chain of
 * utils/data_utils.py:103-116: CropForegroundd -> zero pad, Spacingd -> trilinear resample with border padding and
 * align_corners=False computed in float64, Orientationd -> flips / transposes) folded into one index map:
 *   m         row-major 3x4, (o0, o1, o2, 1) of the OUTPUT grid (the file's own voxel grid) -> fractional voxel index in
 *             the padded grid (the grid before CropForegroundd); coordinates clamp to [0, pad_size - 1]
 *   out_size  spatial shape of the output; pad_size: shape before CropForegroundd
 *   crop_start / pred_size  where pred[.., 0, 0, 0] sits in the padded grid and pred's spatial shape: corners outside
 *             [crop_start, crop_start + pred_size) read 0 (the pad)
 *   mode      0 = nearest (round half to even), 1 = trilinear (`nearest_interp=False`, what the scripts use). */
typedef struct ctu_invert_geom {
  double m[12];
  int32_t out_size[3];
  int32_t pad_size[3];
  int32_t crop_start[3];
  int32_t pred_size[3];
  int32_t mode;
} ctu_invert_geom;

/* out[c][o] = resample(pred[c]) : pred fp32 [C][pred_size], out fp32 [C][out_size] — the tensor Invertd returns. */
int ctu_invert_resample(const float* pred, int C, const ctu_invert_geom* geom, float* out, void* stream);

/* Invertd of both models' logits fused with ctu_ensemble_argmax (same outputs, over out_size voxels; labels / counts on the
 * output grid): the two inverted [C][out_size] fp32 volumes are never written.  C = 14. */
int ctu_invert_ensemble_argmax(const float* p1, const float* p2, int C, const ctu_invert_geom* geom, uint8_t* mask,
                               uint8_t* mask1, uint8_t* mask2, const float* labels, unsigned long long* counts, void* stream);

/* Multi-tensor AdamW: the optimizer step of main_CTUNet.py:190-193 (torch.optim.AdamW(lr, weight_decay), no amsgrad) as
 * ONE launch over a device-resident item table — one item per parameter that has a gradient; all four tensors fp32 and
 * contiguous, numel elements each; unit0 = index of the item's first work unit (1024 elements per unit), items sorted by
 * unit0.  `step` is the 1-based step count of these parameters (bias corrections are formed on the host in double). */
typedef struct ctu_adamw_item {
  void* param;
  const void* grad;
  void* exp_avg;
  void* exp_avg_sq;
  int64_t numel;
  int64_t unit0;
} ctu_adamw_item;
int ctu_adamw_step(const ctu_adamw_item* items_dev, int n_items, long long total_units, double lr, double beta1,
                   double beta2, double eps, double weight_decay, long long step, void* stream);

/* Cap on the SMs a PERSISTENT tensor-core kernel (GEMM / conv / wgrad) sizes its grid for; 0 = all SMs (default).
 * The engine lowers it while it captures the two concurrent lanes of CTUNet (ViT branch || ResNet encoder,
 * hybrid_CTUNet.py:821-838) so that one lane's GPU-filling kernel leaves SMs for the other lane's short kernels. */
void ctu_set_persistent_sm_limit(int sms);

/* Number of kernels this library has launched since load (bench.py's "gpu_launches"). */
int64_t ctu_launch_count(void);
/* 1 if the current device is sm_100 and the driver entry points needed for TMA were found. */
int ctu_device_ok(void);
const char* ctu_version(void);

#ifdef __cplusplus
}
#endif
#endif
