// capi_common.cu -- error string, version and device queries of the C ABI
#include <string.h>

#include "common.cuh"

namespace ast {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
}  // namespace ast

extern "C" const char *ast_last_error(void) { return ast::g_err; }
extern "C" int ast_abi_version(void) { return AST_ABI_VERSION; }
extern "C" int ast_tile_size(void) { return AST_TILE; }
extern "C" int ast_device_sm_count(int *sm_count)
{
    AST_REQUIRE(sm_count != nullptr, "sm_count is null");
    int dev = 0;
    AST_CUDA_TRY(cudaGetDevice(&dev));
    AST_CUDA_TRY(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    return AST_OK;
}
