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

// Same convention as the reference's gpuErrchk (fkl/.../core/utils/utils.h:42-60), but the C-ABI
// returns the code instead of throwing; the header shim rethrows.
#define CVGS_CUDA(call)                                           \
    do {                                                          \
        cudaError_t e_ = (call);                                  \
        if (e_ != cudaSuccess) return cvgs::cuda_fail(e_, #call); \
    } while (0)
