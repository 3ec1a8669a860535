// explicit instantiation of the factor-sized kernels for float
#define DNMF_INSTANTIATE_SMALL
#include "launch_small.cuh"
namespace dnmf {
template int gram_dispatch<float>(const float*, int64_t, int64_t, int, int, float*, float*, cudaStream_t);
template int row_update_dispatch<float>(int, float*, int64_t, const float*, int64_t, const float*, int64_t, const float*, int64_t, int, float, const double*, cudaStream_t, int, int64_t);
template int col_update_dispatch<float>(int, float*, int64_t, const float*, int64_t, const float*, int64_t, int64_t, const float*, int, int64_t, float, int, const double*, cudaStream_t, int, int64_t);
template int residual_dispatch<float>(const float*, int64_t, const float*, int64_t, const float*, int64_t, int64_t, int64_t, int, int64_t, unsigned, unsigned, double*, double*, double*, cudaStream_t);
template int hals_w_col_dispatch<float>(float*, int64_t, const float*, int64_t, const float*, int64_t, int, int, float, double*, unsigned, cudaStream_t);
}
