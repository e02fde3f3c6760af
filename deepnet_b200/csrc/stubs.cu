// temporary stubs
#include "common.cuh"
using namespace dn;
extern "C" {
dn_status dn_vec_vec_dot(const dn_tensor *, const dn_tensor *, const dn_tensor *) { return set_error(DN_ERR_UNSUPPORTED, "nyi"); }
dn_status dn_mat_vec_dot(const dn_tensor *, const dn_tensor *, const dn_tensor *) { return set_error(DN_ERR_UNSUPPORTED, "nyi"); }
dn_status dn_mat_mat_dot(const dn_tensor *, const dn_tensor *, const dn_tensor *) { return set_error(DN_ERR_UNSUPPORTED, "nyi"); }
dn_status dn_batched_mat_mat_dot(const dn_tensor *, const dn_tensor *, const dn_tensor *) { return set_error(DN_ERR_UNSUPPORTED, "nyi"); }
}
