// placeholder - replaced by the tcgen05 implementation
#include "gan_tc.h"
#include "mb_common.h"
namespace mb {
int tc_plan_layers(std::vector<TcLayerDesc>&, size_t* b) { *b = 0; return fail(MB_ERR_STATE, "MB_PREC_F16TC not built"); }
int tc_pack_weights(const TcLayer&, const TapConv&, const float*, char*, cudaStream_t) { return fail(MB_ERR_STATE, "MB_PREC_F16TC not built"); }
size_t tc_workspace_bytes(const std::vector<TcBufReq>&, int, int, int, int) { return 0; }
int tc_forward(const std::vector<TcOp>&, const std::vector<TcBufReq>&, const char*, const float*, const int32_t*, int, int, int, int, float*, void*, cudaStream_t, cudaEvent_t*) { return fail(MB_ERR_STATE, "MB_PREC_F16TC not built"); }
int tc_debug_layer(const TcOp&, const char*, const float*, const float*, int, int, float*, void*, size_t, cudaStream_t) { return fail(MB_ERR_STATE, "MB_PREC_F16TC not built"); }
}
