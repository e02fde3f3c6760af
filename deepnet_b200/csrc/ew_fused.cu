// ew_fused.cu — dn_fused_elemwise: a straight-line element-wise program evaluated in ONE pass over the operands
// (SURVEY.md §8f-3). The reference has no such backend member: it fuses only through the Symbolic layer's
// NVRTC-generated "elements" kernels; its Tensor API runs `a*b + sin a` as three kernels and 512 MiB of traffic
// instead of 192 MiB (Tensor.Benchmark/Benchmark.fs).
//
// No run-time compilation: the program is interpreted, but per WORK ITEM (16 bytes), not per element. The six
// virtual registers are 16-byte slots in shared memory, private to the thread, so operand fetch and write-back are
// one 128-bit shared-memory access each and decoding costs a handful of instructions per element and program
// instruction — below the ~66 instructions per element an HBM-bound f32 kernel with three streams can afford. Every instruction rounds to the element type like the single operator it stands for; this
// translation unit is compiled with -fmad=false so that no multiply-add is contracted and the result equals the
// unfused call sequence bit for bit.
#include "ew_ops.cuh"

#include <cstdlib>
#include <string>

using namespace dn;

namespace {

// One instruction per 32-bit word: bits 0-1 kind, 2-7 op, 8-11 dst, 12-15 a, 16-19 b (one load per decode).
__host__ __device__ constexpr uint32_t fused_pack(int kind, int op, int dst, int a, int b) {
    return (uint32_t)kind | ((uint32_t)op << 2) | ((uint32_t)dst << 8) | ((uint32_t)a << 12) | ((uint32_t)b << 16);
}

template <class T, int NS>
struct FusedSig;
template <class T> struct FusedSig<T, 1> { using type = EwSig<T, T>; };
template <class T> struct FusedSig<T, 2> { using type = EwSig<T, T, T>; };
template <class T> struct FusedSig<T, 3> { using type = EwSig<T, T, T, T>; };

template <class T, int NS>
struct FusedF : FusedSig<T, NS>::type {
    static constexpr bool VectorEval = true;
    static constexpr bool Tiled = false;  // transposed operands take the strided kernel (VEC = 1)
    // 16-byte work items = one register-file slot; consecutive lanes still write consecutive 16-byte pieces,
    // i.e. whole sectors per warp-level store
    static constexpr int Vec = 16 / (int)sizeof(T);
    static constexpr bool MultiEval = true;   // decode once per thread iteration for both items in flight
    static constexpr int MaxInFlight = 2;
    static constexpr int MinBlocks = 4;       // 48 KiB of register-file slots per CTA (6 registers x 2 items)
    int32_t n;
    uint32_t ins[DN_FUSED_MAX_INSTRS];
    T imm[DN_FUSED_MAX_INSTRS];

    template <int OP, int VEC>
    __device__ __forceinline__ static void unary_all(T (&z)[VEC], const T (&x)[VEC]) {
        UnaryF<T, OP> f;
#pragma unroll
        for (int e = 0; e < VEC; ++e) z[e] = f(x[e]);
    }
    template <int OP, int VEC>
    __device__ __forceinline__ static void binary_all(T (&z)[VEC], const T (&x)[VEC], const T (&y)[VEC]) {
        BinaryF<T, OP> f;
#pragma unroll
        for (int e = 0; e < VEC; ++e) z[e] = f(x[e], y[e]);
    }

    // The virtual register file lives in shared memory, one 16-byte slot per (register, thread): operand fetch and
    // write-back are ONE 128-bit shared-memory access each whatever the register number (keeping the file in
    // hardware registers makes the compiler if-convert the register switches into ~25 predicated moves per
    // element and instruction — measured 36 % of the HBM rate). Slots are thread-private: no barriers.
    static constexpr int kSlotElems = 16 / (int)sizeof(T);
    struct alignas(16) Slot { T v[kSlotElems]; };

    template <int VEC>
    __device__ __forceinline__ void eval(T (&out)[VEC], const T (&s0)[VEC], const T (&s1)[VEC], const T (&s2)[VEC]) const {
        static_assert(VEC <= kSlotElems, "work items are at most 16 bytes");
        __shared__ Slot rf[DN_FUSED_REGS][kEwThreads];
        const int tid = threadIdx.x;
        auto put = [&](int reg, const T (&v)[VEC]) {
            Slot sl;
#pragma unroll
            for (int e = 0; e < kSlotElems; ++e) sl.v[e] = e < VEC ? v[e] : T(0);
            rf[reg][tid] = sl;
        };
        auto get = [&](int reg, T (&v)[VEC]) {
            const Slot sl = rf[reg][tid];
#pragma unroll
            for (int e = 0; e < VEC; ++e) v[e] = sl.v[e];
        };
        put(0, s0);
        if (NS > 1) put(1, s1);
        if (NS > 2) put(2, s2);
        T z[VEC];
#pragma unroll 1
        for (int k = 0; k < n; ++k) {
            const uint32_t w = ins[k];
            const int kind = w & 3, op = (w >> 2) & 63, dst = (w >> 8) & 15, ra = (w >> 12) & 15, rb = (w >> 16) & 15;
            T x[VEC], y[VEC];
            if (kind == DN_FUSED_CONST) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) z[e] = imm[k];
            } else {
                get(ra, x);
                if (kind == DN_FUSED_UNARY) {
                    switch (op) {
#define DN_U(OP) case OP: unary_all<OP, VEC>(z, x); break;
                        DN_U(DN_UNARY_MINUS) DN_U(DN_ABS) DN_U(DN_SGN) DN_U(DN_LOG) DN_U(DN_LOG10) DN_U(DN_EXP) DN_U(DN_SIN)
                        DN_U(DN_COS) DN_U(DN_TAN) DN_U(DN_ASIN) DN_U(DN_ACOS) DN_U(DN_ATAN) DN_U(DN_SINH) DN_U(DN_COSH)
                        DN_U(DN_TANH) DN_U(DN_SQRT) DN_U(DN_CEILING) DN_U(DN_FLOOR) DN_U(DN_ROUND) DN_U(DN_TRUNCATE)
#undef DN_U
                    default:  // UnaryPlus
#pragma unroll
                        for (int e = 0; e < VEC; ++e) z[e] = x[e];
                    }
                } else {
                    get(rb, y);
                    switch (op) {
#define DN_B(OP) case OP: binary_all<OP, VEC>(z, x, y); break;
                        DN_B(DN_SUBTRACT) DN_B(DN_MULTIPLY) DN_B(DN_DIVIDE) DN_B(DN_MODULO) DN_B(DN_POWER)
                        DN_B(DN_MAX_ELEMWISE) DN_B(DN_MIN_ELEMWISE)
#undef DN_B
                    default: binary_all<DN_ADD, VEC>(z, x, y);
                    }
                }
            }
            put(dst, z);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) out[e] = z[e];  // the last instruction's value
    }

    // All U work items of the thread at once: one decode per program instruction, U slots per virtual register.
    template <int VEC, int U, class PA, class PB, class PC>
    __device__ __forceinline__ void eval_multi(Pack<T, VEC> (&out)[U], const PA (&sa)[U], const PB (&sb)[U], const PC (&sc)[U]) const {
        static_assert(VEC <= kSlotElems && U <= 2, "register-file geometry");
        __shared__ Slot rf[DN_FUSED_REGS][2][kEwThreads];
        const int tid = threadIdx.x;
        auto put = [&](int reg, int j, const T (&v)[VEC]) {
            Slot sl;
#pragma unroll
            for (int e = 0; e < kSlotElems; ++e) sl.v[e] = e < VEC ? v[e] : T(0);
            rf[reg][j][tid] = sl;
        };
        auto get = [&](int reg, int j, T (&v)[VEC]) {
            const Slot sl = rf[reg][j][tid];
#pragma unroll
            for (int e = 0; e < VEC; ++e) v[e] = sl.v[e];
        };
#pragma unroll
        for (int j = 0; j < U; ++j) {
            put(0, j, sa[j].p.v);
            if constexpr (NS > 1) put(1, j, sb[j].p.v);
            if constexpr (NS > 2) put(2, j, sc[j].p.v);
        }
        T z[U][VEC];
#pragma unroll 1
        for (int k = 0; k < n; ++k) {
            const uint32_t w = ins[k];
            const int kind = w & 3, op = (w >> 2) & 63, dst = (w >> 8) & 15, ra = (w >> 12) & 15, rb = (w >> 16) & 15;
            T x[U][VEC], y[U][VEC];
            if (kind == DN_FUSED_CONST) {
#pragma unroll
                for (int j = 0; j < U; ++j)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) z[j][e] = imm[k];
            } else {
#pragma unroll
                for (int j = 0; j < U; ++j) get(ra, j, x[j]);
                if (kind == DN_FUSED_UNARY) {
                    switch (op) {
#define DN_U(OP) case OP: _Pragma("unroll") for (int j = 0; j < U; ++j) unary_all<OP, VEC>(z[j], x[j]); break;
                        DN_U(DN_UNARY_MINUS) DN_U(DN_ABS) DN_U(DN_SGN) DN_U(DN_LOG) DN_U(DN_LOG10) DN_U(DN_EXP) DN_U(DN_SIN)
                        DN_U(DN_COS) DN_U(DN_TAN) DN_U(DN_ASIN) DN_U(DN_ACOS) DN_U(DN_ATAN) DN_U(DN_SINH) DN_U(DN_COSH)
                        DN_U(DN_TANH) DN_U(DN_SQRT) DN_U(DN_CEILING) DN_U(DN_FLOOR) DN_U(DN_ROUND) DN_U(DN_TRUNCATE)
#undef DN_U
                    default:
#pragma unroll
                        for (int j = 0; j < U; ++j)
#pragma unroll
                            for (int e = 0; e < VEC; ++e) z[j][e] = x[j][e];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < U; ++j) get(rb, j, y[j]);
                    switch (op) {
#define DN_B(OP) case OP: _Pragma("unroll") for (int j = 0; j < U; ++j) binary_all<OP, VEC>(z[j], x[j], y[j]); break;
                        DN_B(DN_SUBTRACT) DN_B(DN_MULTIPLY) DN_B(DN_DIVIDE) DN_B(DN_MODULO) DN_B(DN_POWER)
                        DN_B(DN_MAX_ELEMWISE) DN_B(DN_MIN_ELEMWISE)
#undef DN_B
                    default:
#pragma unroll
                        for (int j = 0; j < U; ++j) binary_all<DN_ADD, VEC>(z[j], x[j], y[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < U; ++j) put(dst, j, z[j]);
        }
#pragma unroll
        for (int j = 0; j < U; ++j)
#pragma unroll
            for (int e = 0; e < VEC; ++e) out[j].v[e] = z[j][e];
    }

    // element-at-a-time form (not used by the kernels this functor is instantiated for, required by the interface)
    __device__ __forceinline__ T operator()(T a) const { T o[1], x[1] = {a}; eval<1>(o, x, x, x); return o[0]; }
    __device__ __forceinline__ T operator()(T a, T b) const { T o[1], x[1] = {a}, y[1] = {b}; eval<1>(o, x, y, y); return o[0]; }
    __device__ __forceinline__ T operator()(T a, T b, T c) const {
        T o[1], x[1] = {a}, y[1] = {b}, w[1] = {c};
        eval<1>(o, x, y, w);
        return o[0];
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Precompiled programs. The interpreter above is issue-bound (43 thread-instructions per element, 63-81 % of the HBM
// rate — profiles/r01d_fused_tanh_grad); the expressions that actually occur (the C1 benchmark expression, every
// element-wise chain of the MLP training step) are ALSO compiled as ordinary element-wise functors — expression
// templates over the same UnaryF / BinaryF functors, in this -fmad=false translation unit, so each node rounds exactly
// like the operator it stands for and the result is the same bits — and run through the same planner and the same
// 256-bit vector / register-transpose kernels as any single operator. A program is recognised by the canonical
// string of its expression tree (registers renamed away, constants by position); anything else is interpreted.
// ---------------------------------------------------------------------------------------------------------------
template <int K> struct XS {  // source K
    template <class T> __device__ __forceinline__ static T ev(const T (&s)[3], const T *) { return s[K]; }
};
template <int I> struct XC {  // I-th constant of the program
    template <class T> __device__ __forceinline__ static T ev(const T (&)[3], const T *c) { return c[I]; }
};
template <int OP, class A> struct XU {
    template <class T> __device__ __forceinline__ static T ev(const T (&s)[3], const T *c) {
        return UnaryF<T, OP>()(A::template ev<T>(s, c));
    }
};
template <int OP, class A, class B> struct XB {
    template <class T> __device__ __forceinline__ static T ev(const T (&s)[3], const T *c) {
        return BinaryF<T, OP>()(A::template ev<T>(s, c), B::template ev<T>(s, c));
    }
};

template <class T, int NS, class E>
struct SpecF : FusedSig<T, NS>::type {
    T c[4];
    __device__ __forceinline__ T operator()(T a) const { const T s[3] = {a, a, a}; return E::template ev<T>(s, c); }
    __device__ __forceinline__ T operator()(T a, T b) const { const T s[3] = {a, b, b}; return E::template ev<T>(s, c); }
    __device__ __forceinline__ T operator()(T a, T b, T d) const { const T s[3] = {a, b, d}; return E::template ev<T>(s, c); }
};

template <class T, int NS, class E>
dn_status run_spec(EwPlan &plan, const double *consts, int nconst) {
    SpecF<T, NS, E> f;
    for (int i = 0; i < 4; ++i) f.c[i] = i < nconst ? (T)consts[i] : T(0);
    return ew_run(plan, f);
}

// Canonical form: u<op>(x), b<op>(x,y), s<k>, k<i> (i = position of the CONST instruction among the CONSTs).
std::string canonical(const dn_fused_instr *prog, int n, int nsrc, double *consts, int *nconst) {
    std::string reg[DN_FUSED_REGS];
    for (int k = 0; k < nsrc; ++k) reg[k] = "s" + std::to_string(k);
    *nconst = 0;
    std::string last;
    for (int k = 0; k < n; ++k) {
        const dn_fused_instr &in = prog[k];
        std::string v;
        if (in.kind == DN_FUSED_CONST) {
            v = "k" + std::to_string(*nconst);
            if (*nconst < 4) consts[*nconst] = in.imm;
            ++*nconst;
        } else if (in.kind == DN_FUSED_UNARY) {
            v = "u" + std::to_string(in.op) + "(" + reg[in.a] + ")";
        } else {
            v = "b" + std::to_string(in.op) + "(" + reg[in.a] + "," + reg[in.b] + ")";
        }
        if (v.size() > 256) return std::string();  // nothing that long is precompiled
        reg[in.dst] = v;
        last = v;
    }
    return last;
}

using S0 = XS<0>;
using S1 = XS<1>;
using S2 = XS<2>;
using K0 = XC<0>;
// a*b + sin(a)                      (C1, Tensor.Benchmark-shaped)
using E_c1 = XB<DN_ADD, XB<DN_MULTIPLY, S0, S1>, XU<DN_SIN, S0>>;
// (z + b).tanh()                    (MLP hidden layer: bias + activation)
using E_bias_tanh = XU<DN_TANH, XB<DN_ADD, S0, S1>>;
// (z - c).exp()                     (softmax numerator)
using E_sub_exp = XU<DN_EXP, XB<DN_SUBTRACT, S0, S1>>;
// -(t * log p)                      (cross-entropy terms)
using E_xent = XU<DN_UNARY_MINUS, XB<DN_MULTIPLY, S0, XU<DN_LOG, S1>>>;
// (p - t) / k                       (softmax + cross-entropy gradient)
using E_diff_div = XB<DN_DIVIDE, XB<DN_SUBTRACT, S0, S1>, K0>;
// d * (k - h*h)                     (tanh gradient)
using E_tanh_grad = XB<DN_MULTIPLY, S0, XB<DN_SUBTRACT, K0, XB<DN_MULTIPLY, S1, S1>>>;
// w - g*k                           (SGD update)
using E_sgd = XB<DN_SUBTRACT, S0, XB<DN_MULTIPLY, S1, K0>>;
// (a - b) * c / (|c| + k)           (three-source example of the sweep)
using E_three = XB<DN_DIVIDE, XB<DN_MULTIPLY, XB<DN_SUBTRACT, S0, S1>, S2>, XB<DN_ADD, XU<DN_ABS, S2>, K0>>;

template <class T>
bool run_precompiled(const std::string &key, int nsrc, EwPlan &plan, const double *consts, int nconst, dn_status *st) {
    auto u = [](int op, const std::string &x) { return "u" + std::to_string(op) + "(" + x + ")"; };
    auto b = [](int op, const std::string &x, const std::string &y) { return "b" + std::to_string(op) + "(" + x + "," + y + ")"; };
    const std::string s0 = "s0", s1 = "s1", s2 = "s2", k0 = "k0";
    if (nsrc == 2) {
        if (key == b(DN_ADD, b(DN_MULTIPLY, s0, s1), u(DN_SIN, s0))) { *st = run_spec<T, 2, E_c1>(plan, consts, nconst); return true; }
        if (key == u(DN_TANH, b(DN_ADD, s0, s1))) { *st = run_spec<T, 2, E_bias_tanh>(plan, consts, nconst); return true; }
        if (key == u(DN_EXP, b(DN_SUBTRACT, s0, s1))) { *st = run_spec<T, 2, E_sub_exp>(plan, consts, nconst); return true; }
        if (key == u(DN_UNARY_MINUS, b(DN_MULTIPLY, s0, u(DN_LOG, s1)))) { *st = run_spec<T, 2, E_xent>(plan, consts, nconst); return true; }
        if (key == b(DN_DIVIDE, b(DN_SUBTRACT, s0, s1), k0)) { *st = run_spec<T, 2, E_diff_div>(plan, consts, nconst); return true; }
        if (key == b(DN_MULTIPLY, s0, b(DN_SUBTRACT, k0, b(DN_MULTIPLY, s1, s1)))) { *st = run_spec<T, 2, E_tanh_grad>(plan, consts, nconst); return true; }
        if (key == b(DN_SUBTRACT, s0, b(DN_MULTIPLY, s1, k0))) { *st = run_spec<T, 2, E_sgd>(plan, consts, nconst); return true; }
    } else if (nsrc == 3) {
        if (key == b(DN_DIVIDE, b(DN_MULTIPLY, b(DN_SUBTRACT, s0, s1), s2), b(DN_ADD, u(DN_ABS, s2), k0))) {
            *st = run_spec<T, 3, E_three>(plan, consts, nconst);
            return true;
        }
    }
    return false;
}

template <class T, int NS>
dn_status run_fused(EwPlan &plan, const dn_fused_instr *prog, int n) {
    FusedF<T, NS> f;
    f.n = n;
    for (int k = 0; k < DN_FUSED_MAX_INSTRS; ++k) {
        f.ins[k] = 0;
        f.imm[k] = T(0);
        if (k < n) {
            const bool reads_b = prog[k].kind == DN_FUSED_BINARY, reads_a = prog[k].kind != DN_FUSED_CONST;
            f.ins[k] = fused_pack(prog[k].kind, prog[k].op, prog[k].dst, reads_a ? prog[k].a : 0, reads_b ? prog[k].b : 0);
            f.imm[k] = (T)prog[k].imm;
        }
    }
    return ew_run(plan, f);
}

}  // namespace

extern "C" dn_status dn_fused_elemwise(const dn_tensor *t, const dn_tensor *const *srcs, int32_t nsrc, const dn_fused_instr *prog,
                                       int32_t ninstr) {
    if (!tensor_valid(t) || !srcs || !prog) return set_error(DN_ERR_INVALID_ARG, "FusedElemwise: bad argument");
    if (nsrc < 1 || nsrc > DN_FUSED_MAX_SRCS) return set_error(DN_ERR_INVALID_ARG, "FusedElemwise: 1 to %d sources are supported", DN_FUSED_MAX_SRCS);
    if (ninstr < 1 || ninstr > DN_FUSED_MAX_INSTRS)
        return set_error(DN_ERR_INVALID_ARG, "FusedElemwise: 1 to %d instructions are supported", DN_FUSED_MAX_INSTRS);
    if (t->dtype != DN_F32 && t->dtype != DN_F64) return set_error(DN_ERR_UNSUPPORTED, "FusedElemwise is only supported for single and double");
    for (int k = 0; k < nsrc; ++k)
        if (!tensor_valid(srcs[k]) || srcs[k]->dtype != t->dtype) return set_error(DN_ERR_INVALID_ARG, "FusedElemwise: operand types differ");
    // a register must be written (a source, or the destination of an earlier instruction) before it is read
    uint32_t written = (1u << nsrc) - 1;
    for (int k = 0; k < ninstr; ++k) {
        const dn_fused_instr &in = prog[k];
        if (in.dst < 0 || in.dst >= DN_FUSED_REGS) return set_error(DN_ERR_INVALID_ARG, "FusedElemwise: instruction %d: bad destination register", k);
        if (in.kind == DN_FUSED_UNARY) {
            if (in.op < 0 || in.op > DN_TRUNCATE) return set_error(DN_ERR_UNSUPPORTED, "FusedElemwise: instruction %d: unary op %d", k, in.op);
        } else if (in.kind == DN_FUSED_BINARY) {
            if (in.op < 0 || in.op > DN_MIN_ELEMWISE) return set_error(DN_ERR_UNSUPPORTED, "FusedElemwise: instruction %d: binary op %d", k, in.op);
        } else if (in.kind != DN_FUSED_CONST) {
            return set_error(DN_ERR_INVALID_ARG, "FusedElemwise: instruction %d: bad kind", k);
        }
        const int nread = in.kind == DN_FUSED_CONST ? 0 : (in.kind == DN_FUSED_UNARY ? 1 : 2);
        const int regs[2] = {in.a, in.b};
        for (int q = 0; q < nread; ++q)
            if (regs[q] < 0 || regs[q] >= DN_FUSED_REGS || !((written >> regs[q]) & 1u))
                return set_error(DN_ERR_INVALID_ARG, "FusedElemwise: instruction %d reads register %d before it is written", k, regs[q]);
        written |= 1u << in.dst;
    }
    EwPlan plan;
    dn_status st = ew_make_plan(plan, t, srcs, nsrc);
    if (st != DN_OK) return st;
    // a precompiled program? (DN_FUSED_INTERPRET=1, a test hook, forces the interpreter so that both stay covered)
    static const bool interpret_only = [] { const char *e = getenv("DN_FUSED_INTERPRET"); return e && e[0] == '1'; }();
    if (!interpret_only) {
        double consts[4];
        int nconst = 0;
        const std::string key = canonical(prog, ninstr, nsrc, consts, &nconst);
        if (!key.empty() && nconst <= 4) {
            const bool hit = t->dtype == DN_F32 ? run_precompiled<float>(key, nsrc, plan, consts, nconst, &st)
                                                : run_precompiled<double>(key, nsrc, plan, consts, nconst, &st);
            if (hit) return st;
        }
    }
    if (t->dtype == DN_F32) {
        if (nsrc == 1) return run_fused<float, 1>(plan, prog, ninstr);
        if (nsrc == 2) return run_fused<float, 2>(plan, prog, ninstr);
        return run_fused<float, 3>(plan, prog, ninstr);
    }
    if (nsrc == 1) return run_fused<double, 1>(plan, prog, ninstr);
    if (nsrc == 2) return run_fused<double, 2>(plan, prog, ninstr);
    return run_fused<double, 3>(plan, prog, ninstr);
}
