// placeholder: replaced by the tcgen05 kernels
#include "kernels.h"
namespace mapnet {
struct TcConvPlan { int dummy; };
int tc_plan_create(TcConvPlan** out, const ConvGeom&, int, const bf16*) { *out = nullptr; set_last_error("tcgen05 conv path not built yet"); return 7; }
void tc_plan_destroy(TcConvPlan* p) { delete p; }
int tc_conv_run(TcConvPlan*, const bf16*, const bf16*, const bf16*, void*, cudaStream_t) { set_last_error("tcgen05 conv path not built yet"); return 7; }
}
