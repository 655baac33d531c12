// mttkrp_tc.cu - 3xTF32 tcgen05 MTTKRP (throughput mode).  Placeholder until the tensor-core
// kernel lands: the entry point reports ADMMQ_E_UNSUPPORTED so that callers fail loudly.
#include "common.cuh"

namespace admmq {

size_t mttkrp_tc_workspace_bytes(int, int, int, int) { return 0; }

int mttkrp_tc(const float*, int, const float*, int, const float*, int, int, float*, void*, size_t, cudaStream_t) {
  return fail(ADMMQ_E_UNSUPPORTED, "admmq_mttkrp: precision 1 (3xTF32 tcgen05) is not built yet; use precision 0");
}

}  // namespace admmq
