// ew_convert.cu — dn_convert dispatch (TensorBackend.fs:69; CudaBackend.fs:300-302).
#include "ew_ops.cuh"

using namespace dn;

namespace dn {
#define DN_DECL(TT) dn_status convert_into_##TT(EwPlan &plan, int src_dtype);
DN_DECL(f32) DN_DECL(f64) DN_DECL(i8) DN_DECL(u8) DN_DECL(i16) DN_DECL(u16) DN_DECL(i32) DN_DECL(u32)
DN_DECL(i64) DN_DECL(u64) DN_DECL(bool)
#undef DN_DECL
}  // namespace dn

extern "C" dn_status dn_convert(const dn_tensor *t, const dn_tensor *a) {
    if (!tensor_valid(t) || !tensor_valid(a)) return set_error(DN_ERR_INVALID_ARG, "Convert: bad argument");
    EwPlan plan;
    const dn_tensor *srcs[1] = {a};
    dn_status st = ew_make_plan(plan, t, srcs, 1);
    if (st != DN_OK || plan.n == 0) return st;
    switch (t->dtype) {
    case DN_F32: return convert_into_f32(plan, a->dtype);
    case DN_F64: return convert_into_f64(plan, a->dtype);
    case DN_I8: return convert_into_i8(plan, a->dtype);
    case DN_U8: return convert_into_u8(plan, a->dtype);
    case DN_I16: return convert_into_i16(plan, a->dtype);
    case DN_U16: return convert_into_u16(plan, a->dtype);
    case DN_I32: return convert_into_i32(plan, a->dtype);
    case DN_U32: return convert_into_u32(plan, a->dtype);
    case DN_I64: return convert_into_i64(plan, a->dtype);
    case DN_U64: return convert_into_u64(plan, a->dtype);
    case DN_BOOL: return convert_into_bool(plan, a->dtype);
    default: return set_error(DN_ERR_INVALID_ARG, "bad dtype");
    }
}
