// placeholder for the WaveRNN C ABI - replaced by the real implementation
#include "mb_common.h"
using namespace mb;
struct mb_wavernn { mb_wavernn_config cfg; };
extern "C" {
int mb_wavernn_create(const mb_wavernn_config* cfg, mb_wavernn** out) { if (!cfg || !out) return fail(MB_ERR_INVALID, "null"); *out = new mb_wavernn{*cfg}; return MB_OK; }
void mb_wavernn_destroy(mb_wavernn* h) { delete h; }
size_t mb_wavernn_arena_bytes(const mb_wavernn*) { return 0; }
int mb_wavernn_set_arena(mb_wavernn*, void*, size_t) { return fail(MB_ERR_STATE, "wavernn not built"); }
int mb_wavernn_set_weight(mb_wavernn*, const char*, const float*, const int64_t*, int32_t, void*) { return fail(MB_ERR_STATE, "wavernn not built"); }
int mb_wavernn_finalize(mb_wavernn*, void*) { return fail(MB_ERR_STATE, "wavernn not built"); }
size_t mb_wavernn_workspace_bytes(const mb_wavernn*, int32_t, int32_t, int32_t) { return 0; }
int mb_wavernn_condition(mb_wavernn*, const float*, int32_t, void*, size_t, void*) { return fail(MB_ERR_STATE, "wavernn not built"); }
int mb_wavernn_generate(mb_wavernn*, const int32_t*, int32_t, int32_t, int32_t, int32_t, const float*, uint64_t, int16_t*, void*, size_t, void*) { return fail(MB_ERR_STATE, "wavernn not built"); }
int mb_wavernn_last_logits(mb_wavernn*, float*, int32_t, void*, void*) { return fail(MB_ERR_STATE, "wavernn not built"); }
}
