// explicit instantiation of the generic col-owner pass for double
#define DNMF_INSTANTIATE_COL
#include "launch_passes.cuh"
namespace dnmf {
template int col_pass_dispatch<double>(bool, const double*, int64_t, const double*, int64_t, const double*, int64_t, double*, int64_t, int64_t, int64_t, int, double, int, void*, int64_t, cudaStream_t);
}
