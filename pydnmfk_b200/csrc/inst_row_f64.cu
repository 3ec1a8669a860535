// explicit instantiation of the generic row-owner pass for double
#define DNMF_INSTANTIATE_ROW
#include "launch_passes.cuh"
namespace dnmf {
template int row_pass_dispatch<double>(bool, const double*, int64_t, const double*, int64_t, const double*, int64_t, double*, int64_t, int64_t, int64_t, int, double, void*, int64_t, cudaStream_t);
}
