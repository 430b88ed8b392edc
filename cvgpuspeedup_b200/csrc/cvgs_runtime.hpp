// cvgs_runtime.hpp -- error reporting and launch accounting shared by the translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

namespace cvgs {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);           // records msg (thread-local), returns code
int cuda_fail(cudaError_t e, const char* what);       // same, from a CUDA runtime error
void count_launch();                                  // one more kernel launched by this thread
int sm_count_of(int device);

}  // namespace cvgs

// NVTX ranges around the C-ABI entry points, compiled in with -DCVGS_NVTX (make NVTX=1), like the reference's test
// harness does with its PUSH_RANGE / POP_RANGE macros (reference tests/nvtx.h:18-104).  Header-only NVTX v3: no
// library to link; without a profiler attached a range costs a few nanoseconds, without the flag nothing.
#ifdef CVGS_NVTX
#include <nvtx3/nvToolsExt.h>
namespace cvgs {
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace cvgs
#define CVGS_RANGE(name) cvgs::NvtxRange cvgs_nvtx_range_(name)
#else
#define CVGS_RANGE(name) ((void)0)
#endif

// Same convention as the reference's gpuErrchk (fkl/.../core/utils/utils.h:42-60), but the C-ABI
// returns the code instead of throwing; the header shim rethrows.
#define CVGS_CUDA(call)                                           \
    do {                                                          \
        cudaError_t e_ = (call);                                  \
        if (e_ != cudaSuccess) return cvgs::cuda_fail(e_, #call); \
    } while (0)
