// preproc_tma.cuh -- placeholder until the TMA-staged kernel lands (next commit).
#pragma once
#include "cvgs_device.cuh"
#include "cvgs_runtime.hpp"
namespace cvgs {
inline bool tma_supported(const PreprocParams&, const DevCrop*, int) { return false; }
inline int launch_tma(const PreprocParams&, const ParamCropTable*, const DevCrop*, int, int, cudaStream_t) {
    return fail(CVGS_ERR_NOT_SUPPORTED, "TMA-staged kernel not built");
}
}  // namespace cvgs
