// placeholder until the KNN kernels land (same commit series)
#include "../../include/cdml.h"
#include "ctx.cuh"
extern "C" {
int cdml_mine_semihard(cdml_ctx*, const void*, int64_t, int, const float*, int64_t, const int64_t*, int64_t, int, float, int32_t*, float*, void*) { cdml::set_error("cdml_mine_semihard: not built yet"); return -3; }
int cdml_knn_index_build(cdml_ctx*, const float*, int64_t, int, int64_t, int, void*, cdml_index**) { cdml::set_error("knn: not built yet"); return -3; }
int cdml_knn_index_destroy(cdml_index*) { return 0; }
int cdml_knn_search(cdml_ctx*, cdml_index*, const float*, int64_t, int64_t, int, float*, int64_t*, int64_t, void*) { cdml::set_error("knn: not built yet"); return -3; }
int cdml_knn_last_stats(cdml_index*, int64_t*) { return -3; }
int cdml_knn_merge(cdml_ctx*, const float*, const int64_t*, int, int64_t, int, int, float*, int64_t*, void*) { cdml::set_error("knn: not built yet"); return -3; }
}
