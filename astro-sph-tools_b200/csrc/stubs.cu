// stubs.cu -- entry points declared in the header whose kernels are not written yet (removed as they land)
#include "common.cuh"
extern "C" int ast_grid3d_workspace_bytes(const ast_grid3d_params *, size_t *) { ast::set_error("ast_grid3d: not implemented yet"); return AST_EUNSUPPORTED; }
extern "C" int ast_grid3d(const ast_grid3d_params *, const double *, const double *, const double *, double *, void *, size_t, void *, ast_project2d_stats *) { ast::set_error("ast_grid3d: not implemented yet"); return AST_EUNSUPPORTED; }
extern "C" int ast_knn_workspace_bytes(const ast_knn_params *, size_t *) { ast::set_error("ast_knn_h: not implemented yet"); return AST_EUNSUPPORTED; }
extern "C" int ast_knn_h(const ast_knn_params *, const double *, double *, int32_t *, double *, void *, size_t, void *) { ast::set_error("ast_knn_h: not implemented yet"); return AST_EUNSUPPORTED; }
