#include "mx_internal.h"
extern "C" {
int mx_qmix_param_layout(const mx_qmix_cfg*, mx_param_entry*, int32_t, int64_t*) { return -1; }
int64_t mx_qmix_workspace_bytes(const mx_qmix_cfg*) { return -1; }
int mx_qmix_create(const mx_qmix_cfg*, float*, float*, float*, float*, void*, int64_t, mx_qmix**) { return 1; }
void mx_qmix_destroy(mx_qmix*) {}
int mx_qmix_step(mx_qmix*, const mx_batch*, void*) { return 1; }
int mx_qmix_backward_only(mx_qmix*, const mx_batch*, void*) { return 1; }
int mx_qmix_apply(mx_qmix*, void*) { return 1; }
float* mx_qmix_grad_buffer(mx_qmix*, int64_t*) { return 0; }
const float* mx_qmix_info(mx_qmix*) { return 0; }
const float* mx_qmix_priorities(mx_qmix*) { return 0; }
int mx_qmix_soft_update(mx_qmix*, void*) { return 1; }
int mx_qmix_hard_update(mx_qmix*, void*) { return 1; }
int mx_qmix_ws_lookup(const mx_qmix*, const char*, int64_t*, int64_t*) { return 1; }
int mx_graph_capture(mx_replay*, mx_qmix*, int32_t, double, uint32_t, void*, mx_graph**) { return 1; }
int mx_graph_launch(mx_graph*, void*) { return 1; }
void mx_graph_destroy(mx_graph*) {}
}
