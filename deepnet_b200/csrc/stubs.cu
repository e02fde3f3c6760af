// temporary stubs
#include "common.cuh"
using namespace dn;
extern "C" {
dn_status dn_gather(const dn_tensor *, const dn_tensor *const *, int32_t, const dn_tensor *) { return set_error(DN_ERR_UNSUPPORTED, "nyi"); }
dn_status dn_scatter(const dn_tensor *, const dn_tensor *const *, int32_t, const dn_tensor *) { return set_error(DN_ERR_UNSUPPORTED, "nyi"); }
dn_status dn_count_true(const dn_tensor *, int64_t *) { return set_error(DN_ERR_UNSUPPORTED, "nyi"); }
dn_status dn_masked_get(const dn_tensor *, const dn_tensor *, const dn_tensor *const *, int32_t) { return set_error(DN_ERR_UNSUPPORTED, "nyi"); }
dn_status dn_masked_set(const dn_tensor *, const dn_tensor *const *, int32_t, const dn_tensor *) { return set_error(DN_ERR_UNSUPPORTED, "nyi"); }
dn_status dn_true_indices(const dn_tensor *, const dn_tensor *) { return set_error(DN_ERR_UNSUPPORTED, "nyi"); }
dn_status dn_vec_vec_dot(const dn_tensor *, const dn_tensor *, const dn_tensor *) { return set_error(DN_ERR_UNSUPPORTED, "nyi"); }
dn_status dn_mat_vec_dot(const dn_tensor *, const dn_tensor *, const dn_tensor *) { return set_error(DN_ERR_UNSUPPORTED, "nyi"); }
dn_status dn_mat_mat_dot(const dn_tensor *, const dn_tensor *, const dn_tensor *) { return set_error(DN_ERR_UNSUPPORTED, "nyi"); }
dn_status dn_batched_mat_mat_dot(const dn_tensor *, const dn_tensor *, const dn_tensor *) { return set_error(DN_ERR_UNSUPPORTED, "nyi"); }
}
