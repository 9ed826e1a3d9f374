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
} ctu_gemm_desc;

int ctu_umma_gemm(const ctu_gemm_desc* desc, void* stream);

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
 * (hybrid_CTUNet.py:481-511, 559-567). */
int ctu_attention(const void* qkv, int ld_qkv, int C, int dim_head, void* out, int ldo, const float* bias, int n,
                  int windows, int mode, int batch, int X, int Y, int Z, int w, void* stream);

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

/* Number of kernels this library has launched since load (bench.py's "gpu_launches"). */
int64_t ctu_launch_count(void);
/* 1 if the current device is sm_100 and the driver entry points needed for TMA were found. */
int ctu_device_ok(void);
const char* ctu_version(void);

#ifdef __cplusplus
}
#endif
#endif
