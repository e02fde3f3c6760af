// dn_backend.hpp — header-only C++ host side above the C ABI (include/dn_tensor.h).
//
// Mirrors the reference's operator interface for this path in compiled code, the way the F# binding in
// fsharp/Tensor.B200 does (which cannot be built in this image): same names, argument order and error behaviour as
//   ITensorDevice / ITensorStorage<'T> / ITensorBackend<'T> ...... Tensor/Tensor/TensorBackend.fs:14-146
//   TensorLayout ................................................. Tensor/Tensor/TensorLayout.fs:12-461
//   the slice of Tensor<'T> that feeds the backend ............... Tensor/Tensor/Tensor.fs:1342-2798,4507-4597
// The backend is parameterised on an `Api` policy that names the C entry points, so the very same frontend code
// runs on libdeepnet_b200.so (CudaApi, device memory) and — in tests only — on the CPU oracle (tests/cpp define an
// OracleApi over the dno_* symbols). No arithmetic happens in this header.
#pragma once

#include <cstdint>
#include <cstring>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "dn_tensor.h"

namespace dnhost {

// ---- exceptions: the .NET exception the reference raises for each status -------------------------
struct NotSupportedException : std::runtime_error { using std::runtime_error::runtime_error; };
struct IndexOutOfRangeException : std::out_of_range { using std::out_of_range::out_of_range; };
struct OutOfCudaMemoryException : std::runtime_error { using std::runtime_error::runtime_error; };
struct CudaException : std::runtime_error { using std::runtime_error::runtime_error; };
struct InvalidOperationException : std::logic_error { using std::logic_error::logic_error; };
struct SingularMatrixException : std::logic_error { using std::logic_error::logic_error; };
struct ArgumentException : std::invalid_argument { using std::invalid_argument::invalid_argument; };

inline void throw_status(dn_status st, const char *msg) {
    const std::string m = msg ? msg : "";
    switch (st) {
    case DN_OK: return;
    case DN_ERR_UNSUPPORTED: throw NotSupportedException(m);
    case DN_ERR_INDEX_OUT_OF_RANGE: throw IndexOutOfRangeException(m);
    case DN_ERR_OUT_OF_MEMORY: throw OutOfCudaMemoryException(m);
    case DN_ERR_INVALID_ARG: throw ArgumentException(m);
    case DN_ERR_SHAPE_MISMATCH: throw InvalidOperationException(m);
    case DN_ERR_SINGULAR_MATRIX: throw SingularMatrixException(m);
    default: throw CudaException(m);
    }
}

// ---- element types -------------------------------------------------------------------------------------------------
template <class T> struct DTypeOf;
template <> struct DTypeOf<float> { static constexpr int value = DN_F32; };
template <> struct DTypeOf<double> { static constexpr int value = DN_F64; };
template <> struct DTypeOf<int8_t> { static constexpr int value = DN_I8; };
template <> struct DTypeOf<uint8_t> { static constexpr int value = DN_U8; };
template <> struct DTypeOf<int16_t> { static constexpr int value = DN_I16; };
template <> struct DTypeOf<uint16_t> { static constexpr int value = DN_U16; };
template <> struct DTypeOf<int32_t> { static constexpr int value = DN_I32; };
template <> struct DTypeOf<uint32_t> { static constexpr int value = DN_U32; };
template <> struct DTypeOf<int64_t> { static constexpr int value = DN_I64; };
template <> struct DTypeOf<uint64_t> { static constexpr int value = DN_U64; };
template <> struct DTypeOf<bool> { static constexpr int value = DN_BOOL; };

constexpr int64_t NotFound = INT64_MIN + 4;  // TensorRng.fs:24

// ---- TensorLayout (TensorLayout.fs) ----------------------------------------------------------------------------------
struct TensorLayout {
    std::vector<int64_t> Shape, Stride;
    int64_t Offset = 0;
    int NDims() const { return (int)Shape.size(); }
    int64_t NElems() const { return std::accumulate(Shape.begin(), Shape.end(), (int64_t)1, std::multiplies<int64_t>()); }

    static TensorLayout newC(const std::vector<int64_t> &shape) {  // TensorLayout.fs:119-120
        TensorLayout l;
        l.Shape = shape;
        l.Stride.assign(shape.size(), 1);
        int64_t acc = 1;
        for (int d = (int)shape.size() - 1; d >= 0; --d) { l.Stride[d] = acc; acc *= shape[d]; }
        return l;
    }
    TensorLayout swapDim(int a, int b) const {  // :344-349
        TensorLayout l = *this;
        std::swap(l.Shape[a], l.Shape[b]);
        std::swap(l.Stride[a], l.Stride[b]);
        return l;
    }
    TensorLayout transpose() const {  // :353-356
        if (NDims() < 2) throw ArgumentException("cannot transpose non-matrix");
        return swapDim(NDims() - 2, NDims() - 1);
    }
    TensorLayout permuteAxes(const std::vector<int> &permut) const {  // :361-365: permut[i] = new position of axis i
        TensorLayout l = *this;
        for (size_t i = 0; i < permut.size(); ++i) { l.Shape[permut[i]] = Shape[i]; l.Stride[permut[i]] = Stride[i]; }
        return l;
    }
    TensorLayout reverseAxis(int ax) const {  // :368-371
        TensorLayout l = *this;
        l.Offset += (Shape[ax] - 1) * Stride[ax];
        l.Stride[ax] = -Stride[ax];
        return l;
    }
    TensorLayout padLeft() const {  // :149-150
        TensorLayout l = *this;
        l.Shape.insert(l.Shape.begin(), 1);
        l.Stride.insert(l.Stride.begin(), 0);
        return l;
    }
    TensorLayout broadcastTo(const std::vector<int64_t> &bs) const {  // broadcastToShape, :257-271
        TensorLayout l = *this;
        if ((int)bs.size() < l.NDims()) throw InvalidOperationException("Cannot broadcast to a shape of lower rank.");
        while (l.NDims() < (int)bs.size()) l = l.padLeft();
        for (size_t d = 0; d < bs.size(); ++d) {
            if (l.Shape[d] == bs[d]) continue;
            if (l.Shape[d] != 1) throw InvalidOperationException("Cannot broadcast shapes to the same size.");
            l.Shape[d] = bs[d];
            l.Stride[d] = 0;
        }
        return l;
    }
    // slice [start, stop) along `ax` (Rng.Rng, TensorLayout.fs:392-409)
    TensorLayout slice(int ax, int64_t start, int64_t stop) const {
        if (start < 0 || stop > Shape[ax] || stop < start) throw IndexOutOfRangeException("slice out of range");
        TensorLayout l = *this;
        l.Offset += start * Stride[ax];
        l.Shape[ax] = stop - start;
        return l;
    }
    dn_tensor desc(void *base, int dtype) const {
        if (NDims() > DN_MAX_DIMS) throw NotSupportedException("tensors of rank > 8 are not supported");
        dn_tensor d;
        std::memset(&d, 0, sizeof d);
        d.base = base;
        d.offset = Offset;
        d.ndims = NDims();
        d.dtype = dtype;
        for (int i = 0; i < NDims(); ++i) { d.shape[i] = Shape[i]; d.stride[i] = Stride[i]; }
        return d;
    }
};

inline std::vector<int64_t> broadcastShape(const std::vector<int64_t> &a, const std::vector<int64_t> &b) {
    const size_t n = std::max(a.size(), b.size());
    std::vector<int64_t> out(n);
    for (size_t i = 0; i < n; ++i) {
        const int64_t x = i + a.size() >= n ? a[i + a.size() - n] : 1, y = i + b.size() >= n ? b[i + b.size() - n] : 1;
        if (x != y && x != 1 && y != 1) throw InvalidOperationException("Cannot broadcast shapes to the same size.");
        out[i] = x == 1 ? y : x;
    }
    return out;
}

// ---- Api policy for libdeepnet_b200.so ------------------------------------------------------------------------------
struct CudaApi {
    static constexpr const char *Id = "Cuda";
    static void check(dn_status st) { if (st != DN_OK) throw_status(st, dn_last_error()); }
    static void *alloc(int64_t nbytes) { void *p = nullptr; check(dn_alloc(nbytes, &p)); return p; }
    static void release(void *p) { dn_free(p); }
    static void upload(void *dst, const void *src, int64_t n) { check(dn_memcpy_h2d(dst, src, n)); check(dn_sync()); }
    static void download(void *dst, const void *src, int64_t n) { check(dn_memcpy_d2h(dst, src, n)); }
#define DN_FWD(name) template <class... A> static dn_status name(A... a) { return dn_##name(a...); }
    DN_FWD(fill_const) DN_FWD(fill_incrementing) DN_FWD(copy) DN_FWD(convert) DN_FWD(unary) DN_FWD(binary) DN_FWD(compare)
    DN_FWD(is_finite) DN_FWD(if_then_else) DN_FWD(reduce_last_axis) DN_FWD(arg_reduce_last_axis) DN_FWD(find_last_axis)
    DN_FWD(gather) DN_FWD(scatter) DN_FWD(count_true) DN_FWD(masked_get) DN_FWD(masked_set) DN_FWD(true_indices)
    DN_FWD(vec_vec_dot) DN_FWD(mat_vec_dot) DN_FWD(mat_mat_dot) DN_FWD(batched_mat_mat_dot) DN_FWD(batched_invert) DN_FWD(fused_elemwise)
#undef DN_FWD
};

// ---- storage / backend / frontend -----------------------------------------------------------------------------------
template <class T, class Api>
class TensorStorage {  // TensorCudaStorage<'T>, CudaBackend.fs:51-108
  public:
    explicit TensorStorage(int64_t nElems) : n_(nElems > 0 ? nElems : 1), ptr_(Api::alloc(n_ * (int64_t)sizeof(T))) {}
    ~TensorStorage() { Api::release(ptr_); }
    TensorStorage(const TensorStorage &) = delete;
    void *Ptr() const { return ptr_; }
    int64_t DataSize() const { return n_; }
  private:
    int64_t n_;
    void *ptr_;
};

template <class T, class Api> class Tensor;

/// ITensorBackend<'T> (TensorBackend.fs:64-146): target first, sources already broadcast to the target's shape.
template <class Api>
struct Backend {
    template <class A> static dn_tensor d(const A &t) { return t.Desc(); }
    template <class TT, class TA> static void Copy(const TT &t, const TA &a) { auto x = d(t), y = d(a); Api::check(Api::copy(&x, &y)); }
    template <class TT, class TA> static void Convert(const TT &t, const TA &a) { auto x = d(t), y = d(a); Api::check(Api::convert(&x, &y)); }
    template <class TT, class V> static void FillConst(V value, const TT &t) { auto x = d(t); Api::check(Api::fill_const(&x, &value)); }
    template <class TT, class V> static void FillIncrementing(V start, V incr, const TT &t) {
        auto x = d(t); Api::check(Api::fill_incrementing(&x, &start, &incr));
    }
    template <class TT, class TA> static void Unary(int op, const TT &t, const TA &a) { auto x = d(t), y = d(a); Api::check(Api::unary(op, &x, &y)); }
    template <class TT, class TA, class TB> static void Binary(int op, const TT &t, const TA &a, const TB &b) {
        auto x = d(t), y = d(a), z = d(b); Api::check(Api::binary(op, &x, &y, &z));
    }
    template <class TT, class TA, class TB> static void Compare(int op, const TT &t, const TA &a, const TB &b) {
        auto x = d(t), y = d(a), z = d(b); Api::check(Api::compare(op, &x, &y, &z));
    }
    template <class TT, class TA> static void IsFinite(const TT &t, const TA &a) { auto x = d(t), y = d(a); Api::check(Api::is_finite(&x, &y)); }
    template <class TT, class TC, class TA, class TB> static void IfThenElse(const TT &t, const TC &c, const TA &a, const TB &b) {
        auto x = d(t), w = d(c), y = d(a), z = d(b); Api::check(Api::if_then_else(&x, &w, &y, &z));
    }
    template <class TT, class TA> static void ReduceLastAxis(int op, const TT &t, const TA &a) { auto x = d(t), y = d(a); Api::check(Api::reduce_last_axis(op, &x, &y)); }
    template <class TT, class TA> static void ArgReduceLastAxis(int op, const TT &t, const TA &a) {
        auto x = d(t), y = d(a); Api::check(Api::arg_reduce_last_axis(op, &x, &y));
    }
    template <class TT, class TA, class V> static void FindLastAxis(V value, const TT &t, const TA &a) {
        auto x = d(t), y = d(a); Api::check(Api::find_last_axis(&value, &x, &y));
    }
    template <class TT, class TA> static void TrueIndices(const TT &t, const TA &a) { auto x = d(t), y = d(a); Api::check(Api::true_indices(&x, &y)); }
    template <class TA> static int64_t CountTrue(const TA &a) { auto y = d(a); int64_t n = 0; Api::check(Api::count_true(&y, &n)); return n; }
    template <class TT, class TA, class TB> static void MatMatDot(const TT &t, const TA &a, const TB &b) {
        auto x = d(t), y = d(a), z = d(b); Api::check(Api::mat_mat_dot(&x, &y, &z));
    }
    template <class TT, class TA, class TB> static void BatchedMatMatDot(const TT &t, const TA &a, const TB &b) {
        auto x = d(t), y = d(a), z = d(b); Api::check(Api::batched_mat_mat_dot(&x, &y, &z));
    }
    // FusedElemwise (extension, include/dn_tensor.h dn_fused_elemwise)
    template <class TT> static void FusedElemwise(const TT &t, const std::vector<const dn_tensor *> &srcs,
                                                  const std::vector<dn_fused_instr> &prog) {
        auto x = d(t);
        Api::check(Api::fused_elemwise(&x, srcs.data(), (int32_t)srcs.size(), prog.data(), (int32_t)prog.size()));
    }
    // BatchedInvert (TensorBackend.fs:142)
    template <class TT, class TA> static void BatchedInvert(const TT &t, const TA &a) {
        auto x = d(t), y = d(a); Api::check(Api::batched_invert(&x, &y));
    }
    template <class TT, class TA, class TB> static void MatVecDot(const TT &t, const TA &a, const TB &b) {
        auto x = d(t), y = d(a), z = d(b); Api::check(Api::mat_vec_dot(&x, &y, &z));
    }
    template <class TT, class TA, class TB> static void VecVecDot(const TT &t, const TA &a, const TB &b) {
        auto x = d(t), y = d(a), z = d(b); Api::check(Api::vec_vec_dot(&x, &y, &z));
    }
    // index / mask lists: nullptr entry = None / NoMask
    template <class TT, class TA> static void Gather(const TT &t, const std::vector<const dn_tensor *> &idxs, const TA &a) {
        auto x = d(t), y = d(a); Api::check(Api::gather(&x, idxs.data(), (int32_t)idxs.size(), &y));
    }
    template <class TT, class TA> static void Scatter(const TT &t, const std::vector<const dn_tensor *> &idxs, const TA &a) {
        auto x = d(t), y = d(a); Api::check(Api::scatter(&x, idxs.data(), (int32_t)idxs.size(), &y));
    }
    template <class TT, class TA> static void MaskedGet(const TT &t, const TA &a, const std::vector<const dn_tensor *> &masks) {
        auto x = d(t), y = d(a); Api::check(Api::masked_get(&x, &y, masks.data(), (int32_t)masks.size()));
    }
    template <class TT, class TA> static void MaskedSet(const TT &t, const std::vector<const dn_tensor *> &masks, const TA &a) {
        auto x = d(t), y = d(a); Api::check(Api::masked_set(&x, masks.data(), (int32_t)masks.size(), &y));
    }
};

/// Tensor<'T> (Tensor.fs:50-54): a layout over a shared storage.
template <class T, class Api>
class Tensor {
  public:
    using B = Backend<Api>;
    Tensor() = default;
    Tensor(TensorLayout layout, std::shared_ptr<TensorStorage<T, Api>> storage) : layout_(std::move(layout)), storage_(std::move(storage)) {}
    explicit Tensor(const std::vector<int64_t> &shape)  // Tensor<'T>(shape, dev), row-major (Tensor.fs:317-325)
        : layout_(TensorLayout::newC(shape)), storage_(std::make_shared<TensorStorage<T, Api>>(layout_.NElems())) {}

    static Tensor ofVector(const std::vector<T> &data, const std::vector<int64_t> &shape) {
        Tensor t(shape);
        if (!data.empty()) upload_raw(t, data.data(), (int64_t)data.size());
        return t;
    }
    // bool vectors are bit-packed in C++: take bytes instead
    static Tensor ofBytes(const std::vector<uint8_t> &data, const std::vector<int64_t> &shape) {
        static_assert(sizeof(T) == 1, "ofBytes is for 1-byte element types");
        Tensor t(shape);
        if (!data.empty()) Api::upload(t.storage_->Ptr(), data.data(), (int64_t)data.size());
        return t;
    }
    std::vector<T> toVector() const {  // logical row-major contents (copies through a contiguous tensor if needed)
        Tensor c = isC() ? *this : Copy();
        std::vector<typename std::conditional<std::is_same<T, bool>::value, uint8_t, T>::type> raw((size_t)NElems());
        if (NElems() > 0)
            Api::download(raw.data(), static_cast<char *>(c.storage_->Ptr()) + c.layout_.Offset * (int64_t)sizeof(T), NElems() * (int64_t)sizeof(T));
        return std::vector<T>(raw.begin(), raw.end());
    }

    const TensorLayout &Layout() const { return layout_; }
    const std::vector<int64_t> &Shape() const { return layout_.Shape; }
    int NDims() const { return layout_.NDims(); }
    int64_t NElems() const { return layout_.NElems(); }
    dn_tensor Desc() const { return layout_.desc(storage_->Ptr(), DTypeOf<T>::value); }
    Tensor Relayout(TensorLayout l) const { return Tensor(std::move(l), storage_); }
    bool isC() const {
        const TensorLayout c = TensorLayout::newC(layout_.Shape);
        for (int d = 0; d < NDims(); ++d)
            if (layout_.Shape[d] > 1 && layout_.Stride[d] != c.Stride[d]) return false;
        return true;
    }

    // views
    Tensor T_() const { return Relayout(layout_.transpose()); }
    Tensor permuteAxes(const std::vector<int> &p) const { return Relayout(layout_.permuteAxes(p)); }
    Tensor reverseAxis(int ax) const { return Relayout(layout_.reverseAxis(ax)); }
    Tensor broadcastTo(const std::vector<int64_t> &s) const { return Relayout(layout_.broadcastTo(s)); }
    Tensor slice(int ax, int64_t start, int64_t stop) const { return Relayout(layout_.slice(ax, start, stop)); }

    // copy / fill
    Tensor Copy() const { Tensor t(layout_.Shape); B::Copy(t, *this); return t; }
    void CopyFrom(const Tensor &src) { B::Copy(*this, src.broadcastTo(Shape())); }
    void FillConst(T v) { if constexpr (std::is_same<T, bool>::value) B::FillConst((uint8_t)v, *this); else B::FillConst(v, *this); }
    void FillIncrementing(T start, T incr) { B::FillIncrementing(start, incr, *this); }
    template <class U> Tensor<U, Api> convert() const { Tensor<U, Api> t(Shape()); B::Convert(t, *this); return t; }

    // element-wise: Fill* variants (target first) and allocating operators (PrepareElemwise, Tensor.fs:4583-4597)
    void FillUnary(int op, const Tensor &a) { B::Unary(op, *this, a.broadcastTo(Shape())); }
    void FillBinary(int op, const Tensor &a, const Tensor &b) { B::Binary(op, *this, a.broadcastTo(Shape()), b.broadcastTo(Shape())); }
    void FillAdd(const Tensor &a, const Tensor &b) { FillBinary(DN_ADD, a, b); }
    void FillMultiply(const Tensor &a, const Tensor &b) { FillBinary(DN_MULTIPLY, a, b); }
    Tensor unary(int op) const { Tensor t(Shape()); B::Unary(op, t, *this); return t; }
    Tensor binary(int op, const Tensor &o) const {
        const auto s = broadcastShape(Shape(), o.Shape());
        Tensor t(s);
        B::Binary(op, t, broadcastTo(s), o.broadcastTo(s));
        return t;
    }
    Tensor<bool, Api> compare(int op, const Tensor &o) const {
        const auto s = broadcastShape(Shape(), o.Shape());
        Tensor<bool, Api> t(s);
        B::Compare(op, t, broadcastTo(s), o.broadcastTo(s));
        return t;
    }
    Tensor operator+(const Tensor &o) const { return binary(DN_ADD, o); }
    Tensor operator-(const Tensor &o) const { return binary(DN_SUBTRACT, o); }
    Tensor operator*(const Tensor &o) const { return binary(DN_MULTIPLY, o); }
    Tensor operator/(const Tensor &o) const { return binary(DN_DIVIDE, o); }
    Tensor operator%(const Tensor &o) const { return binary(DN_MODULO, o); }
    Tensor operator-() const { return unary(DN_UNARY_MINUS); }
    Tensor<bool, Api> isFinite() const { Tensor<bool, Api> t(Shape()); B::IsFinite(t, *this); return t; }
    static Tensor ifThenElse(const Tensor<bool, Api> &c, const Tensor &a, const Tensor &b) {
        const auto s = broadcastShape(broadcastShape(c.Shape(), a.Shape()), b.Shape());
        Tensor t(s);
        B::IfThenElse(t, c.broadcastTo(s), a.broadcastTo(s), b.broadcastTo(s));
        return t;
    }

    // reductions: the reduced axis is permuted to last as a view (PrepareAxisReduceSources, Tensor.fs:4544-4567)
    Tensor axisToLast(int ax) const {
        std::vector<int> perm(NDims());
        for (int d = 0; d < NDims(); ++d) perm[d] = d < ax ? d : (d == ax ? NDims() - 1 : d - 1);
        return permuteAxes(perm);
    }
    std::vector<int64_t> shapeWithout(int ax) const { auto s = Shape(); s.erase(s.begin() + ax); return s; }
    Tensor reduceAxis(int op, int ax) const { Tensor t(shapeWithout(ax)); B::ReduceLastAxis(op, t, axisToLast(ax)); return t; }
    Tensor sumAxis(int ax) const { return reduceAxis(DN_SUM, ax); }
    Tensor productAxis(int ax) const { return reduceAxis(DN_PRODUCT, ax); }
    Tensor minAxis(int ax) const { return reduceAxis(DN_MIN, ax); }
    Tensor maxAxis(int ax) const { return reduceAxis(DN_MAX, ax); }
    Tensor allAxis(int ax) const { return reduceAxis(DN_ALL, ax); }
    Tensor anyAxis(int ax) const { return reduceAxis(DN_ANY, ax); }
    Tensor<int64_t, Api> countTrueAxis(int ax) const { Tensor<int64_t, Api> t(shapeWithout(ax)); B::ReduceLastAxis(DN_COUNT_TRUE, t, axisToLast(ax)); return t; }
    Tensor<int64_t, Api> argMaxAxis(int ax) const { Tensor<int64_t, Api> t(shapeWithout(ax)); B::ArgReduceLastAxis(DN_ARG_MAX, t, axisToLast(ax)); return t; }
    Tensor<int64_t, Api> argMinAxis(int ax) const { Tensor<int64_t, Api> t(shapeWithout(ax)); B::ArgReduceLastAxis(DN_ARG_MIN, t, axisToLast(ax)); return t; }
    Tensor<int64_t, Api> findAxis(T v, int ax) const { Tensor<int64_t, Api> t(shapeWithout(ax)); B::FindLastAxis(v, t, axisToLast(ax)); return t; }
    int64_t countTrue() const { return B::CountTrue(*this); }
    Tensor<int64_t, Api> trueIdx() const {  // Tensor.fs:2259-2263
        Tensor<int64_t, Api> t({countTrue(), (int64_t)NDims()});
        B::TrueIndices(t, *this);
        return t;
    }

    // indexing (Tensor.fs:2090-2198, 3011-3069); nullptr = None / NoMask
    static Tensor gather(const std::vector<const Tensor<int64_t, Api> *> &indices, const Tensor &src) {
        std::vector<int64_t> shape;
        for (auto *i : indices) if (i) shape = shape.empty() ? i->Shape() : broadcastShape(shape, i->Shape());
        Tensor t(shape);
        std::vector<dn_tensor> descs(indices.size());
        std::vector<const dn_tensor *> ptrs(indices.size(), nullptr);
        for (size_t k = 0; k < indices.size(); ++k)
            if (indices[k]) { descs[k] = indices[k]->broadcastTo(shape).Desc(); ptrs[k] = &descs[k]; }
        B::Gather(t, ptrs, src);
        return t;
    }
    static Tensor scatter(const std::vector<const Tensor<int64_t, Api> *> &indices, const std::vector<int64_t> &trgtShape, const Tensor &src) {
        Tensor t(trgtShape);
        std::vector<dn_tensor> descs(indices.size());
        std::vector<const dn_tensor *> ptrs(indices.size(), nullptr);
        for (size_t k = 0; k < indices.size(); ++k)
            if (indices[k]) { descs[k] = indices[k]->broadcastTo(src.Shape()).Desc(); ptrs[k] = &descs[k]; }
        B::Scatter(t, ptrs, src);
        return t;
    }
    Tensor M(const Tensor<bool, Api> &mask) const {  // one mask covering the whole tensor: a.M(m)
        const Tensor<bool, Api> fm = mask.isC() ? mask : mask.Copy();
        const Tensor fs = isC() ? *this : Copy();
        const int64_t n = fm.countTrue();
        Tensor t({n});
        const Tensor<bool, Api> flat_m = fm.Relayout(TensorLayout{{fm.NElems()}, {1}, fm.Layout().Offset});
        const Tensor flat_s = fs.Relayout(TensorLayout{{fs.NElems()}, {1}, fs.Layout().Offset});
        const dn_tensor md = flat_m.Desc();
        B::MaskedGet(t, flat_s, {&md});
        return t;
    }

    // dot (Tensor.fs:2714-2798)
    Tensor dot(const Tensor &b) const {
        if (NDims() == 1 && b.NDims() == 1) { Tensor t(std::vector<int64_t>{}); B::VecVecDot(t, *this, b); return t; }
        if (NDims() == 2 && b.NDims() == 1) { Tensor t({Shape()[0]}); B::MatVecDot(t, *this, b); return t; }
        if (NDims() == 2 && b.NDims() == 2) { Tensor t({Shape()[0], b.Shape()[1]}); B::MatMatDot(t, *this, b); return t; }
        if (NDims() == b.NDims() && NDims() > 2) {
            auto s = Shape();
            s.back() = b.Shape().back();
            Tensor t(s);
            B::BatchedMatMatDot(t, *this, b);
            return t;
        }
        throw ArgumentException("Cannot compute dot product between tensors of these shapes.");
    }

    // invert (Tensor.fs:2836-2839)
    Tensor invert() const {
        if (NDims() < 2) throw ArgumentException("Need at least a matrix to invert.");
        Tensor t(Shape());
        B::BatchedInvert(t, *this);
        return t;
    }

  private:
    static void upload_raw(Tensor &t, const T *data, int64_t n) { Api::upload(t.storage_->Ptr(), data, n * (int64_t)sizeof(T)); }
    TensorLayout layout_;
    std::shared_ptr<TensorStorage<T, Api>> storage_;
};

}  // namespace dnhost
