// explicit instantiation of the generic row-owner pass for float
#define DNMF_INSTANTIATE_ROW
#include "launch_passes.cuh"
namespace dnmf {
template int row_pass_dispatch<float>(bool, const float*, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int, float, void*, int64_t, cudaStream_t);
}
