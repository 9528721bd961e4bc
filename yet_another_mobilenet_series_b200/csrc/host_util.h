// Host-side helpers shared by the launchers: error reporting, device queries, driver entry points.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/yamb200.h"

namespace yamb {

int set_error(int code, const char* fmt, ...);
int max_ctas();  // SM count of the current device, <= 0 without a device

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

int gemm_launch(const yamb_gemm* a, cudaStream_t stream);

}  // namespace yamb
