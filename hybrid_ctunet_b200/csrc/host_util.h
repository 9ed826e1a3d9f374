// Host-side helpers shared by the launchers: driver entry point for TMA descriptor encoding, launch counter.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>

struct ctu_gemm_desc;
struct ctu_wgrad_desc;

namespace ctu {

typedef CUresult (*tma_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled, looked up through the runtime so the library does not link against libcuda.
tma_encode_fn tma_encoder();
void count_launch(int n = 1);

// 3x3x3 convolution with shared-memory halo reuse (umma_conv3_halo.cu); CTU_E_UNSUPPORTED when not applicable.
int conv3_halo_dispatch(const ::ctu_gemm_desc* d, cudaStream_t stream);
// its weight-gradient counterpart (umma_wgrad_halo.cu)
int wgrad_halo_dispatch(const ::ctu_wgrad_desc* d, cudaStream_t stream);

}  // namespace ctu
