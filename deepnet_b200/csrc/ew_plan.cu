// ew_plan.cu — layout canonicalisation for element-wise operators (host side, no templates).
//
// The reference instantiates one NVRTC module per (dtype, rank) and walks full N-d positions on the device
// (Tensor/Tensor/Cuda/CudaKernels.fs:22-25,128; Tensor/Tensor/Cuda/Kernels/Work.cuh:25-110). Here the rank is
// erased on the host: an element-wise operator does not care in which order positions are visited, so dims can
// be dropped, flipped, reordered and merged freely as long as the same transformation is applied to every operand.
#include "elemwise.cuh"

#include <algorithm>

namespace dn {

dn_status ew_make_plan(EwPlan &plan, const dn_tensor *t, const dn_tensor *const *srcs, int nsrc, int index_operand,
                       int index_dim) {
    if (!tensor_valid(t)) return set_error(DN_ERR_INVALID_ARG, "element-wise: invalid target descriptor");
    plan = EwPlan();
    plan.nops = nsrc + 1;
    const int nd = t->ndims;
    plan.op[0].ptr = data_ptr(t);
    plan.op[0].esize = dtype_size(t->dtype);
    for (int k = 0; k < nsrc; ++k) {
        EwOperand &o = plan.op[k + 1];
        if (k == index_operand) {
            o.ptr = nullptr;
            o.esize = 1;
            o.is_index = true;
            continue;
        }
        const dn_tensor *s = srcs[k];
        if (!tensor_valid(s)) return set_error(DN_ERR_INVALID_ARG, "element-wise: invalid source descriptor %d", k);
        if (!same_shape(t, s))
            return set_error(DN_ERR_SHAPE_MISMATCH, "element-wise: source %d does not have the target's shape", k);
        o.ptr = data_ptr(s);
        o.esize = dtype_size(s->dtype);
    }
    // gather dims (outermost-first order of the descriptor), dropping size-1 dims
    struct Dim {
        int64_t size;
        int64_t st[kEwMaxOps];
    };
    Dim dims[DN_MAX_DIMS];
    int n = 0;
    int64_t total = 1;
    for (int d = 0; d < nd; ++d) {
        total *= t->shape[d];
        if (t->shape[d] == 1) continue;
        Dim &dm = dims[n++];
        dm.size = t->shape[d];
        dm.st[0] = t->stride[d];
        for (int k = 0; k < nsrc; ++k)
            dm.st[k + 1] = (k == index_operand) ? (d == index_dim ? 1 : 0) : srcs[k]->stride[d];
    }
    plan.n = total;
    if (total == 0) return DN_OK;
    // flip dims whose target stride is negative so that all target strides are positive
    for (int i = 0; i < n; ++i) {
        if (dims[i].st[0] < 0) {
            for (int k = 0; k < plan.nops; ++k) {
                plan.op[k].ptr += (dims[i].size - 1) * dims[i].st[k] * plan.op[k].esize;
                dims[i].st[k] = -dims[i].st[k];
            }
        } else if (dims[i].st[0] == 0) {
            return set_error(DN_ERR_INVALID_ARG, "element-wise: the target must not be a broadcast (stride 0) view");
        }
    }
    // innermost-first: ascending target stride (stable, so equal strides keep descriptor order reversed)
    std::reverse(dims, dims + n);
    std::stable_sort(dims, dims + n, [](const Dim &a, const Dim &b) { return a.st[0] < b.st[0]; });
    // merge dim i+1 into dim i when contiguous in every operand
    int m = 0;
    for (int i = 0; i < n; ++i) {
        if (m > 0) {
            bool ok = true;
            for (int k = 0; k < plan.nops && ok; ++k) ok = dims[i].st[k] == dims[m - 1].st[k] * dims[m - 1].size;
            if (ok) {
                dims[m - 1].size *= dims[i].size;
                continue;
            }
        }
        dims[m++] = dims[i];
    }
    if (m == 0) {  // single element
        m = 1;
        dims[0].size = 1;
        for (int k = 0; k < plan.nops; ++k) dims[0].st[k] = 0;
    }
    plan.ndims = m;
    for (int d = 0; d < m; ++d) {
        plan.shape[d] = dims[d].size;
        for (int k = 0; k < plan.nops; ++k) plan.op[k].stride[d] = dims[d].st[k];
    }
    return DN_OK;
}

// Address of the lowest byte that operand k touches for the first work item of a row (reversed sources read
// downwards from their first element).
static uintptr_t ew_first_access(const EwOperand &o, int vec, int esize) {
    return (uintptr_t)o.ptr - (o.stride[0] == -1 ? (uintptr_t)(vec - 1) * esize : 0);
}

bool ew_can_vectorize(const EwPlan &plan, int vec, const int *esize, int64_t *tail_elems) {
    *tail_elems = 0;
    const int nd = plan.ndims;
    if (plan.shape[0] < vec) return false;
    const int64_t tail = plan.shape[0] % vec;
    if (tail != 0 && nd > 1) return false;
    for (int k = 0; k < plan.nops; ++k) {
        const EwOperand &o = plan.op[k];
        if (o.is_index) {
            if (o.stride[0] != 0 && o.stride[0] != 1) return false;
            continue;
        }
        // target: unit stride; sources: unit stride, splat (0) or reversed (-1: the pack is loaded from the
        // lowest address and reversed in registers)
        if (k == 0 ? o.stride[0] != 1 : (o.stride[0] != 0 && o.stride[0] != 1 && o.stride[0] != -1)) return false;
        if (o.stride[0] == 0) continue;  // splat: scalar loads, no alignment requirement
        int64_t bytes = (int64_t)vec * esize[k];
        int64_t align = bytes >= 16 ? 16 : bytes;
        if (ew_first_access(o, vec, esize[k]) % align != 0) return false;
        for (int d = 1; d < nd; ++d)
            if ((o.stride[d] * esize[k]) % align != 0) return false;
    }
    *tail_elems = tail;
    return true;
}

// Row peeling: the innermost dim is unit-stride (or splat / reversed) in every operand and the row pitches keep
// vector alignment, but the rows START misaligned (a.[1.., 1..] views). Finds the number of leading elements h
// (< vec) after which every operand is aligned — preferably to 32 bytes for the 256-bit path — so that the
// launcher can run columns [h, h+body) vectorised and the thin head / tail columns through the scalar kernel.
bool ew_find_peel(const EwPlan &plan, int vec, const int *esize, int64_t *head) {
    const int nd = plan.ndims;
    if (plan.shape[0] < 3 * (int64_t)vec) return false;
    for (int pass = 0; pass < 2; ++pass) {  // pass 0: 32-byte alignment for packs of >= 32 bytes; pass 1: 16 bytes
        for (int64_t h = 0; h < vec; ++h) {
            bool ok = true;
            for (int k = 0; k < plan.nops && ok; ++k) {
                const EwOperand &o = plan.op[k];
                if (o.is_index) {
                    ok = o.stride[0] == 0 || o.stride[0] == 1;
                    continue;
                }
                if (k == 0 ? o.stride[0] != 1 : (o.stride[0] != 0 && o.stride[0] != 1 && o.stride[0] != -1)) ok = false;
                if (!ok || o.stride[0] == 0) continue;
                const int64_t bytes = (int64_t)vec * esize[k];
                const int64_t align = (pass == 0 && bytes >= 32) ? 32 : (bytes >= 16 ? 16 : bytes);
                EwOperand shifted = o;
                shifted.ptr += h * o.stride[0] * esize[k];
                if (ew_first_access(shifted, vec, esize[k]) % align != 0) ok = false;
                for (int d = 1; d < nd && ok; ++d)
                    if ((o.stride[d] * esize[k]) % align != 0) ok = false;
            }
            if (ok) {
                *head = h;
                return true;
            }
        }
    }
    return false;
}

int ew_pick_tiled_dim(const EwPlan &plan) {
    const int nd = plan.ndims;
    if (nd < 2 || plan.op[0].stride[0] != 1 || plan.shape[0] < 8) return -1;
    int dS = -1;
    for (int k = 1; k < plan.nops; ++k) {
        const EwOperand &o = plan.op[k];
        if (o.is_index || o.stride[0] == 0 || o.stride[0] == 1) continue;
        for (int d = 1; d < nd; ++d)
            if (o.stride[d] == 1 && plan.shape[d] >= 8) {
                if (dS < 0) dS = d;
                break;
            }
    }
    return dS;
}

int ew_xpose_mode(const EwPlan &plan, int k, int dS, int m) {
    const EwOperand &o = plan.op[k];
    if (o.is_index) return 2;
    const int64_t v = (int64_t)m * o.esize;  // vector bytes (<= 16)
    auto aligned_elsewhere = [&](int unit_dim) {
        if (((uintptr_t)o.ptr) % v != 0) return false;
        for (int d = 0; d < plan.ndims; ++d)
            if (d != unit_dim && (o.stride[d] * o.esize) % v != 0) return false;
        return true;
    };
    if (o.stride[0] == 1 && aligned_elsewhere(0)) return 0;
    if (o.stride[dS] == 1 && aligned_elsewhere(dS)) return 1;
    return 2;
}

int ew_grid_for(int64_t work_items, int items_per_cta) {
    int64_t ctas = (work_items + items_per_cta - 1) / items_per_cta;
    // Enough CTAs to fill every SM several times over, few enough that launch overhead stays negligible;
    // the kernels grid-stride over the rest.
    const int64_t cap = (int64_t)sm_count() * 32;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    return (int)ctas;
}

}  // namespace dn
