// explicit instantiation of the factor-sized kernels for double
#define DNMF_INSTANTIATE_SMALL
#include "launch_small.cuh"
namespace dnmf {
template int gram_dispatch<double>(const double*, int64_t, int64_t, int, int, double*, double*, cudaStream_t);
template int row_update_dispatch<double>(int, double*, int64_t, const double*, int64_t, const double*, int64_t, const double*, int64_t, int, double, const double*, cudaStream_t, int, int64_t);
template int col_update_dispatch<double>(int, double*, int64_t, const double*, int64_t, const double*, int64_t, int64_t, const double*, int, int64_t, double, int, const double*, cudaStream_t, int, int64_t);
template int residual_dispatch<double>(const double*, int64_t, const double*, int64_t, const double*, int64_t, int64_t, int64_t, int, int64_t, unsigned, unsigned, double*, double*, double*, cudaStream_t);
template int hals_w_col_dispatch<double>(double*, int64_t, const double*, int64_t, const double*, int64_t, int, int, double, double*, unsigned, cudaStream_t);
}
