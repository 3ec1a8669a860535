// tcgen05 / TMA / TMEM path (placeholder until the UMMA kernels land: nothing is eligible yet).
#include "tc_api.cuh"

namespace dnmf {

bool tc_eligible(int, const void*, int64_t, int64_t, int64_t, int64_t, int) { return false; }
int64_t tc_workspace_bytes(int, int64_t, int64_t, int64_t, int) { return 0; }
int tc_ah(const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int, int, void*, int64_t, cudaStream_t) { return fail(DNMF_E_UNSUPPORTED, "tc path not built"); }
int tc_wta(const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int, int, int, void*, int64_t, cudaStream_t) { return fail(DNMF_E_UNSUPPORTED, "tc path not built"); }
int tc_kl_uht(const float*, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int, float, int, void*, int64_t, cudaStream_t) { return fail(DNMF_E_UNSUPPORTED, "tc path not built"); }
int tc_kl_wtu(const float*, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int, float, int, int, void*, int64_t, cudaStream_t) { return fail(DNMF_E_UNSUPPORTED, "tc path not built"); }

}  // namespace dnmf
