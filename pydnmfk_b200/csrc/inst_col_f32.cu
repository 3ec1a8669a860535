// explicit instantiation of the generic col-owner pass for float
#define DNMF_INSTANTIATE_COL
#include "launch_passes.cuh"
namespace dnmf {
template int col_pass_dispatch<float>(bool, const float*, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int, float, int, void*, int64_t, cudaStream_t);
}
