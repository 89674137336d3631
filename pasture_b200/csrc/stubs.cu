// stubs.cu -- entry points declared in pasture_b200.h whose kernels have not landed yet.
// They fail loudly (no silent CPU fallback).
#include "internal.h"
using namespace pb200;
extern "C" {
int pb200_knn(pb200_ctx*, const pb200_buffer_desc*, uint32_t, uint32_t*, double*) {
    return set_error(PB200_ERR_UNSUPPORTED, "pb200_knn: not implemented yet");
}
int pb200_radius_search(pb200_ctx*, const pb200_buffer_desc*, double, uint32_t, uint32_t*, uint32_t*) {
    return set_error(PB200_ERR_UNSUPPORTED, "pb200_radius_search: not implemented yet");
}
int pb200_compute_normals(pb200_ctx*, const pb200_buffer_desc*, uint32_t, double*, double*) {
    return set_error(PB200_ERR_UNSUPPORTED, "pb200_compute_normals: not implemented yet");
}
int pb200_proj_pipeline_for_crs(const char*, const char*, pb200_proj_op*, uint32_t) {
    return set_error(PB200_ERR_UNSUPPORTED, "pb200_proj_pipeline_for_crs: not implemented yet");
}
int pb200_reproject(pb200_ctx*, const pb200_buffer_desc*, const pb200_buffer_desc*, const pb200_proj_op*, uint32_t) {
    return set_error(PB200_ERR_UNSUPPORTED, "pb200_reproject: not implemented yet");
}
}
