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

/* Number of kernels this library has launched since load (bench.py's "gpu_launches"). */
int64_t ctu_launch_count(void);
/* 1 if the current device is sm_100 and the driver entry points needed for TMA were found. */
int ctu_device_ok(void);
const char* ctu_version(void);

#ifdef __cplusplus
}
#endif
#endif
