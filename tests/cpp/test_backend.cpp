// C++ host-side mirror (include/dn_backend.hpp) exercised the way the reference's CudaTests.fs exercises the F#
// frontend: the same templated test body runs on the CUDA backend (CudaApi -> libdeepnet_b200.so) and on the CPU
// oracle (OracleApi -> oracle/_build/libdn_oracle.so) and the results are compared. Also checks the documented
// known answers (Tensor.fs doc examples) on whichever devices are enabled.
//   usage: test_backend [oracle|both]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <string>

#include "dn_backend.hpp"

extern "C" {
const char *dno_last_error(void);
#define DNO(name, ...) dn_status dno_##name(__VA_ARGS__);
DNO(fill_const, const dn_tensor *, const void *)
DNO(fill_incrementing, const dn_tensor *, const void *, const void *)
DNO(copy, const dn_tensor *, const dn_tensor *)
DNO(convert, const dn_tensor *, const dn_tensor *)
DNO(unary, int32_t, const dn_tensor *, const dn_tensor *)
DNO(binary, int32_t, const dn_tensor *, const dn_tensor *, const dn_tensor *)
DNO(compare, int32_t, const dn_tensor *, const dn_tensor *, const dn_tensor *)
DNO(is_finite, const dn_tensor *, const dn_tensor *)
DNO(if_then_else, const dn_tensor *, const dn_tensor *, const dn_tensor *, const dn_tensor *)
DNO(reduce_last_axis, int32_t, const dn_tensor *, const dn_tensor *)
DNO(arg_reduce_last_axis, int32_t, const dn_tensor *, const dn_tensor *)
DNO(find_last_axis, const void *, const dn_tensor *, const dn_tensor *)
DNO(gather, const dn_tensor *, const dn_tensor *const *, int32_t, const dn_tensor *)
DNO(scatter, const dn_tensor *, const dn_tensor *const *, int32_t, const dn_tensor *)
DNO(count_true, const dn_tensor *, int64_t *)
DNO(masked_get, const dn_tensor *, const dn_tensor *, const dn_tensor *const *, int32_t)
DNO(masked_set, const dn_tensor *, const dn_tensor *const *, int32_t, const dn_tensor *)
DNO(true_indices, const dn_tensor *, const dn_tensor *)
DNO(vec_vec_dot, const dn_tensor *, const dn_tensor *, const dn_tensor *)
DNO(mat_vec_dot, const dn_tensor *, const dn_tensor *, const dn_tensor *)
DNO(mat_mat_dot, const dn_tensor *, const dn_tensor *, const dn_tensor *)
DNO(batched_mat_mat_dot, const dn_tensor *, const dn_tensor *, const dn_tensor *)
DNO(batched_invert, const dn_tensor *, const dn_tensor *)
DNO(fused_elemwise, const dn_tensor *, const dn_tensor *const *, int32_t, const dn_fused_instr *, int32_t)
#undef DNO
}

// HostTensor (oracle) policy: plain host memory, dno_* entry points.
struct OracleApi {
    static constexpr const char *Id = "Host";
    static void check(dn_status st) { if (st != DN_OK) dnhost::throw_status(st, dno_last_error()); }
    static void *alloc(int64_t nbytes) { return std::calloc(1, (size_t)nbytes); }
    static void release(void *p) { std::free(p); }
    static void upload(void *dst, const void *src, int64_t n) { std::memcpy(dst, src, (size_t)n); }
    static void download(void *dst, const void *src, int64_t n) { std::memcpy(dst, src, (size_t)n); }
#define DN_FWD(name) template <class... A> static dn_status name(A... a) { return dno_##name(a...); }
    DN_FWD(fill_const) DN_FWD(fill_incrementing) DN_FWD(copy) DN_FWD(convert) DN_FWD(unary) DN_FWD(binary) DN_FWD(compare)
    DN_FWD(is_finite) DN_FWD(if_then_else) DN_FWD(reduce_last_axis) DN_FWD(arg_reduce_last_axis) DN_FWD(find_last_axis)
    DN_FWD(gather) DN_FWD(scatter) DN_FWD(count_true) DN_FWD(masked_get) DN_FWD(masked_set) DN_FWD(true_indices)
    DN_FWD(vec_vec_dot) DN_FWD(mat_vec_dot) DN_FWD(mat_mat_dot) DN_FWD(batched_mat_mat_dot) DN_FWD(batched_invert) DN_FWD(fused_elemwise)
#undef DN_FWD
};

using Results = std::map<std::string, std::vector<double>>;
static int g_fail = 0;

template <class V> static std::vector<double> dbl(const V &v) { return std::vector<double>(v.begin(), v.end()); }

static void expect(bool ok, const std::string &what) {
    if (!ok) { std::printf("FAIL %s\n", what.c_str()); ++g_fail; }
}
static bool close(const std::vector<double> &a, const std::vector<double> &b, double rtol, double atol) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i) {
        if (std::isnan(a[i]) && std::isnan(b[i])) continue;
        if (std::fabs(a[i] - b[i]) > atol + rtol * std::fabs(b[i])) return false;
    }
    return true;
}

template <class Api>
Results run_suite() {
    using namespace dnhost;
    using TF = Tensor<float, Api>;
    using TD = Tensor<double, Api>;
    using TI = Tensor<int64_t, Api>;
    using TB = Tensor<bool, Api>;
    Results r;
    const std::string id = Api::Id;

    // --- documented known answers (Tensor.fs doc examples) ---
    TD a = TD::ofVector({1, 2, 3, 4, 5, 6, 7, 8}, {2, 4});
    expect(a.sumAxis(1).toVector() == std::vector<double>({10, 26}), id + " sumAxis (Tensor.fs:2277-2281)");
    expect(a.productAxis(1).toVector() == std::vector<double>({24, 1680}), id + " productAxis (Tensor.fs:2321-2325)");
    expect(a.argMaxAxis(1).toVector() == std::vector<int64_t>({3, 3}), id + " argMaxAxis (Tensor.fs:2484-2488)");
    expect(a.argMinAxis(1).toVector() == std::vector<int64_t>({0, 0}), id + " argMinAxis (Tensor.fs:2458-2462)");
    TD f = TD::ofVector({1, 2, 3, 4, 5, 6, 7, 3}, {2, 4});
    expect(f.findAxis(3.0, 1).toVector() == std::vector<int64_t>({2, 3}), id + " findAxis (Tensor.fs:2548-2552)");
    TD m5 = TD::ofVector({5, 6, 7}, {3}), m2 = TD::ofVector({2, 3, 4}, {3});
    expect((m5 % m2).toVector() == std::vector<double>({1, 0, 3}), id + " % (Tensor.fs:1522-1526)");
    expect(m5.binary(DN_POWER, m2).toVector() == std::vector<double>({25, 216, 2401}), id + " ** (Tensor.fs:1565-1569)");
    TD rr = TD::ofVector({-3.0, -2.7, 2.7, 3.0}, {4});
    expect(rr.unary(DN_ROUND).toVector() == std::vector<double>({-3, -3, 3, 3}), id + " round (Tensor.fs:1257-1260)");
    TB tb = TB::ofBytes({1, 0, 1, 0, 0, 1, 1, 0}, {2, 4});
    expect(tb.trueIdx().toVector() == std::vector<int64_t>({0, 0, 0, 2, 1, 1, 1, 2}), id + " trueIdx (Tensor.fs:2249-2256)");
    expect(tb.countTrueAxis(1).toVector() == std::vector<int64_t>({2, 2}), id + " countTrueAxis (Tensor.fs:2213-2217)");
    TD src = TD::ofVector({0.0, 0.1, 0.2, 0.3, 1.0, 1.1, 1.2, 1.3, 2.0, 2.1, 2.2, 2.3}, {3, 4});
    TI i0 = TI::ofVector({1, 2, 0, 0}, {4}), i1 = TI::ofVector({3, 1, 0, 3}, {4});
    expect(TD::gather({&i0, &i1}, src).toVector() == std::vector<double>({1.3, 2.1, 0.0, 0.3}), id + " gather (Tensor.fs:2105-2112)");
    TI j1 = TI::ofVector({3, 1, 0}, {3});
    expect(TD::gather({nullptr, &j1}, src).toVector() == std::vector<double>({0.3, 1.1, 2.0}), id + " gather None (Tensor.fs:2114-2116)");
    TI s0 = TI::ofVector({0, 0, 0, 0, 2, 2, 2, 2, 1, 1, 1, 1}, {3, 4}), s1 = TI::ofVector({3, 3, 3, 3, 0, 1, 2, 3, 0, 1, 2, 3}, {3, 4});
    expect(close(dbl(TD::scatter({&s0, &s1}, {4, 4}, src).toVector()),
                 {0, 0, 0, 0.6, 2.0, 2.1, 2.2, 2.3, 1.0, 1.1, 1.2, 1.3, 0, 0, 0, 0}, 1e-12, 1e-12), id + " scatter (Tensor.fs:2168-2185)");
    TD ma = TD::ofVector({1, 2, 3, 4, 5, 6}, {2, 3});
    TB mm = TB::ofBytes({1, 1, 0, 0, 0, 1}, {2, 3});
    expect(ma.M(mm).toVector() == std::vector<double>({1, 2, 6}), id + " a.M(m) (Tensor.fs:3120-3127)");
    TD h = TD::ofVector({0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14}, {5, 3});
    TD ii = TD::ofVector({1.1, 0.1, 0.1, 0.1, 1.1, 0.1, 0.1, 0.1, 1.1}, {3, 3});
    expect(close(dbl(h.dot(ii).toVector()), {0.3, 1.3, 2.3, 4.2, 5.2, 6.2, 8.1, 9.1, 10.1, 12, 13, 14, 15.9, 16.9, 17.9}, 1e-12, 1e-12),
           id + " h .* i (Guide-Operations.md:127-147)");
    TD inv_in = TD::ofVector({1.0, 2.0, 3.0, 4.0}, {2, 2});
    expect(close(dbl(inv_in.invert().toVector()), {-2.0, 1.0, 1.5, -0.5}, 1e-12, 1e-12), id + " invert (Tensor.fs:2821-2825)");
    bool threw = false;
    try { TD::ofVector({1, 0, 0, 1, 2, 0, 1, 0, 0}, {3, 3}).invert(); } catch (const dnhost::SingularMatrixException &) { threw = true; }
    expect(threw, id + " invert singular matrix raises (BaseTests.fs:205-211)");

    // --- seeded workload on views, returned for host-vs-CUDA comparison ---
    std::mt19937_64 gen(99);
    std::uniform_real_distribution<float> uni(-50.f, 50.f);
    const int64_t R = 67, C = 129;
    std::vector<float> xa(R * C), xb(R * C), xt(C * R), row(C);
    for (auto &v : xa) v = uni(gen);
    for (auto &v : xb) v = uni(gen);
    for (auto &v : xt) v = uni(gen);
    for (auto &v : row) v = uni(gen);
    TF A = TF::ofVector(xa, {R, C}), Bm = TF::ofVector(xb, {R, C}), At = TF::ofVector(xt, {C, R}), Row = TF::ofVector(row, {1, C});
    r["add"] = dbl((A + Bm).toVector());
    r["addT"] = dbl((At.T_() + Bm).toVector());
    r["mulrow"] = dbl((A * Row).toVector());
    r["sliced_rev"] = dbl((A.slice(0, 1, 60).slice(1, 3, 100).reverseAxis(1) - Bm.slice(0, 1, 60).slice(1, 3, 100)).toVector());
    r["sin"] = dbl(A.unary(DN_SIN).toVector());
    r["less"] = dbl(A.compare(DN_LESS, Bm).toVector());
    r["select"] = dbl(TF::ifThenElse(A.compare(DN_GREATER, Bm), A, Bm).toVector());
    r["sum1"] = dbl(A.sumAxis(1).toVector());
    r["sum0"] = dbl(A.sumAxis(0).toVector());
    r["max1"] = dbl(At.T_().maxAxis(1).toVector());
    r["argmax0"] = dbl(A.argMaxAxis(0).toVector());
    r["argmin1"] = dbl(At.T_().argMinAxis(1).toVector());
    r["convert"] = dbl(A.template convert<int32_t>().toVector());
    r["masked"] = dbl(A.M(A.compare(DN_GREATER, Bm)).toVector());
    r["trueidx"] = dbl(A.compare(DN_GREATER, Bm).trueIdx().toVector());
    r["dot"] = dbl(A.dot(At).toVector());  // [67,129] . [129,67]
    {
        std::vector<float> sq(3 * 40 * 40);
        for (size_t i = 0; i < sq.size(); ++i) sq[i] = uni(gen) / 50.f + ((i % (40 * 40)) % 41 == 0 ? 21.f : 0.f);
        r["invert"] = dbl(TF::ofVector(sq, {3, 40, 40}).invert().toVector());
    }
    {   // fused a*b + sin(a) == the three-call sequence
        TF fused(A.Shape());
        const dn_tensor da = A.Desc(), db = Bm.Desc();
        Backend<Api>::FusedElemwise(fused, {&da, &db}, {{DN_FUSED_BINARY, DN_MULTIPLY, 2, 0, 1, 0.0}, {DN_FUSED_UNARY, DN_SIN, 3, 0, 0, 0.0},
                                                {DN_FUSED_BINARY, DN_ADD, 2, 2, 3, 0.0}});
        r["fused"] = dbl(fused.toVector());
        r["unfused"] = dbl((A * Bm + A.unary(DN_SIN)).toVector());
        expect(r["fused"] == r["unfused"], id + " fused a*b+sin(a) is bit-identical to the three-call sequence");
    }
    A.FillMultiply(A, Bm);                  // in place (Guide-Operations.md:112-117)
    r["inplace"] = dbl(A.toVector());
    return r;
}

int main(int argc, char **argv) {
    const std::string mode = argc > 1 ? argv[1] : "both";
    Results host = run_suite<OracleApi>();
    if (mode == "both") {
        dnhost::CudaApi::check(dn_init(0));
        Results cuda = run_suite<dnhost::CudaApi>();
        for (auto &kv : host) {
            const std::string &k = kv.first;
            double rtol = 0, atol = 0;
            if (k == "sin") rtol = 1e-5;
            if (k == "sum1" || k == "sum0") { rtol = 1e-3; atol = 0.05; }
            if (k == "dot") { rtol = 1e-2; atol = 25.0; }  // TF32 tensor cores, |a|.|b| ~ 8e4 per element
            if (k == "invert") { rtol = 1e-4; atol = 1e-6; }
            if (k == "fused" || k == "unfused") { rtol = 1e-5; atol = 1e-4; }
            expect(close(cuda[k], kv.second, rtol, atol), "cuda vs host: " + k);
        }
    }
    if (g_fail) { std::printf("%d FAILED\n", g_fail); return 1; }
    std::printf("ALL OK (%s)\n", mode.c_str());
    return 0;
}
