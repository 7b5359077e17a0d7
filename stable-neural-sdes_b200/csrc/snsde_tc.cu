// placeholder until the tcgen05 kernel lands
#include "snsde_tc.cuh"
namespace snsde {
static const char* g_reason = "tensor-core path not built";
bool tc_supported(const snsde_model_desc&, int, int) { return false; }
const char* tc_unsupported_reason() { return g_reason; }
int tc_set_weights(TcPlan&, const snsde_model_desc&, const float*, int, int, cudaStream_t) { return SNSDE_ERR_UNSUPPORTED; }
cudaError_t tc_forward(TcPlan&, const TcForwardArgs&, cudaStream_t, int*) { return cudaErrorNotSupported; }
void tc_release(TcPlan&) {}
}
