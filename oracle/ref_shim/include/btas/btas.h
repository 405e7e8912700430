// btas/btas.h -- TEST INFRASTRUCTURE ONLY (oracle/ref_shim).
//
// API-compatible stand-in for the one BTAS header the reference includes (general.h:21), so that the
// reference's OWN source files (/root/reference/src/{integrals,mp2,cc,hf,general}.cpp, main.cpp) can be
// compiled unmodified into oracle/_ref/.  BTAS (github.com/ValeevGroup/btas, unpinned master,
// CMakeLists.txt:9-13) is not in this image; only the surface the reference uses is provided
// (SURVEY.md section 2.3):
//   btas::Tensor<double>: default ctor, extent ctor (value-initialised storage, i.e. zeros -- BTAS'
//     varray value-constructs its elements; the reference relies on it), copy/assign,
//     extent(k), fill, operator()(indices...), operator+=/-=, data(), size();
//   btas::contract(alpha, A, {labels}, B, {labels}, beta, C, {labels}):
//     C[labels_C] = beta * C[labels_C] + alpha * sum_{labels in A and B, not in C} A[labels_A] * B[labels_B]
//     (BTAS' documented meaning; evaluated like BTAS does, as permute -> GEMM -> permute).
// One deliberate extension: an EMPTY tensor behaves as zeros of whatever shape is asked of it
// (reads give 0, += / contract with it are no-ops).  The reference's UHF path hands an empty
// moes.F_ia to contract/+=/operator() (cc.cpp:62-67 drops it) which is an assertion
// failure / UB with real BTAS; zeros is the only defined reading (F is diagonal, cc.cpp:25-34).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <initializer_list>
#include <stdexcept>
#include <vector>
#include <math.h>
#include <stdlib.h>

namespace btas {

template<typename T>
class Tensor {
public:
    typedef T value_type;
    Tensor() {}
    template<typename... I>
    explicit Tensor(I... e) : ext_{static_cast<long>(e)...} {
        size_t n = 1;
        for (long x : ext_) n *= static_cast<size_t>(x);
        v_.assign(n, T());
    }
    size_t rank() const { return ext_.size(); }
    long extent(size_t k) const { return k < ext_.size() ? ext_[k] : 0; }
    const std::vector<long> &extents() const { return ext_; }
    size_t size() const { return v_.size(); }
    bool empty() const { return v_.empty(); }
    T *data() { return v_.data(); }
    const T *data() const { return v_.data(); }
    void fill(const T &x) { std::fill(v_.begin(), v_.end(), x); }

    template<typename... I>
    T &operator()(I... idx) {
        return v_[offset(idx...)];
    }
    template<typename... I>
    const T &operator()(I... idx) const {
        static const T zero = T();
        if (v_.empty()) return zero;// Q4: empty tensor reads as zeros
        return v_[offset(idx...)];
    }
    Tensor &operator+=(const Tensor &o) {
        if (o.v_.empty()) return *this;// Q4
        if (o.v_.size() != v_.size()) throw std::invalid_argument("btas shim: += size mismatch");
        for (size_t k = 0; k < v_.size(); ++k) v_[k] += o.v_[k];
        return *this;
    }
    Tensor &operator-=(const Tensor &o) {
        if (o.v_.empty()) return *this;
        if (o.v_.size() != v_.size()) throw std::invalid_argument("btas shim: -= size mismatch");
        for (size_t k = 0; k < v_.size(); ++k) v_[k] -= o.v_[k];
        return *this;
    }

private:
    template<typename... I>
    size_t offset(I... idx) const {
        const long ix[] = {static_cast<long>(idx)...};
        if (sizeof...(I) != ext_.size()) throw std::out_of_range("btas shim: wrong number of indices");
        size_t o = 0;
        for (size_t k = 0; k < sizeof...(I); ++k) {
            if (ix[k] < 0 || ix[k] >= ext_[k]) throw std::out_of_range("btas shim: index out of range");
            o = o * static_cast<size_t>(ext_[k]) + static_cast<size_t>(ix[k]);
        }
        return o;
    }
    std::vector<long> ext_;
    std::vector<T> v_;
};

namespace detail {
// out[m][k] (row-major, m over `rows` labels, k over `cols` labels) gathered from T with labels `lab`
inline void gather(const Tensor<double> &T, const std::vector<int> &lab, const std::vector<int> &rows,
                   const std::vector<int> &cols, std::vector<double> &out, long &nr, long &nc) {
    const size_t r = lab.size();
    std::vector<long> stride(r, 1);
    for (size_t k = r; k-- > 1;) stride[k - 1] = stride[k] * T.extent(k);
    auto pos = [&](int l) { return static_cast<size_t>(std::find(lab.begin(), lab.end(), l) - lab.begin()); };
    std::vector<long> re, rs, ce, cs;
    nr = 1, nc = 1;
    for (int l : rows) re.push_back(T.extent(pos(l))), rs.push_back(stride[pos(l)]), nr *= re.back();
    for (int l : cols) ce.push_back(T.extent(pos(l))), cs.push_back(stride[pos(l)]), nc *= ce.back();
    std::vector<long> roff(static_cast<size_t>(nr)), coff(static_cast<size_t>(nc));
    auto offsets = [](const std::vector<long> &e, const std::vector<long> &s, std::vector<long> &o) {
        const long n = static_cast<long>(o.size());
        for (long x = 0; x < n; ++x) {
            long rem = x, off = 0;
            for (size_t k = e.size(); k-- > 0;) off += (rem % e[k]) * s[k], rem /= e[k];
            o[static_cast<size_t>(x)] = off;
        }
    };
    offsets(re, rs, roff), offsets(ce, cs, coff);
    out.resize(static_cast<size_t>(nr * nc));
    const double *src = T.data();
#pragma omp parallel for schedule(static)
    for (long i = 0; i < nr; ++i)
        for (long j = 0; j < nc; ++j) out[static_cast<size_t>(i * nc + j)] = src[roff[static_cast<size_t>(i)] + coff[static_cast<size_t>(j)]];
}
}// namespace detail

// label-vector form (what the initializer-list front end below forwards to; also exported for the shim's own test)
inline void contract_labels(double alpha, const Tensor<double> &A, const std::vector<int> &la, const Tensor<double> &B,
                            const std::vector<int> &lb, double beta, Tensor<double> &C, const std::vector<int> &lc) {
    if (lc.size() != C.rank()) throw std::invalid_argument("btas shim: contract: C rank/labels mismatch");
    if (A.empty() || B.empty()) {// Q4: an empty operand is zeros
        if (static_cast<double>(beta) != 1.0)
            for (size_t k = 0; k < C.size(); ++k) C.data()[k] *= static_cast<double>(beta);
        return;
    }
    if (la.size() != A.rank() || lb.size() != B.rank()) throw std::invalid_argument("btas shim: contract: rank/labels mismatch");
    auto in = [](const std::vector<int> &v, int l) { return std::find(v.begin(), v.end(), l) != v.end(); };
    std::vector<int> con, fa, fb;
    for (int l : la) {
        if (in(lb, l)) {
            if (in(lc, l)) throw std::invalid_argument("btas shim: contract: Hadamard (batch) labels are not supported");
            con.push_back(l);
        } else {
            if (!in(lc, l)) throw std::invalid_argument("btas shim: contract: label of A missing from B and C");
            fa.push_back(l);
        }
    }
    for (int l : lb)
        if (!in(la, l)) {
            if (!in(lc, l)) throw std::invalid_argument("btas shim: contract: label of B missing from A and C");
            fb.push_back(l);
        }
    if (fa.size() + fb.size() != lc.size()) throw std::invalid_argument("btas shim: contract: C labels do not match the free labels");
    for (int l : con) {
        const size_t pa = static_cast<size_t>(std::find(la.begin(), la.end(), l) - la.begin());
        const size_t pb = static_cast<size_t>(std::find(lb.begin(), lb.end(), l) - lb.begin());
        if (A.extent(pa) != B.extent(pb)) throw std::invalid_argument("btas shim: contract: contracted extents differ");
    }
    std::vector<double> Ap, Bp;
    long M, K, K2, N;
    detail::gather(A, la, fa, con, Ap, M, K);
    detail::gather(B, lb, con, fb, Bp, K2, N);
    std::vector<double> Cp(static_cast<size_t>(M * N), 0.0);
    // Cp = Ap . Bp   (row-major; i-k-j order so the inner loop vectorises)
#pragma omp parallel for schedule(static)
    for (long i = 0; i < M; ++i) {
        double *c = &Cp[static_cast<size_t>(i * N)];
        for (long k = 0; k < K; ++k) {
            const double a = Ap[static_cast<size_t>(i * K + k)];
            const double *b = &Bp[static_cast<size_t>(k * N)];
            for (long j = 0; j < N; ++j) c[j] += a * b[j];
        }
    }
    // scatter: C[labels_C] = beta*C + alpha*Cp[(fa),(fb)]
    const size_t rc = lc.size();
    std::vector<long> cstride(rc, 1);
    for (size_t k = rc; k-- > 1;) cstride[k - 1] = cstride[k] * C.extent(k);
    auto cpos = [&](int l) { return static_cast<size_t>(std::find(lc.begin(), lc.end(), l) - lc.begin()); };
    auto offsets = [&](const std::vector<int> &labs, long n, std::vector<long> &o, const Tensor<double> &src,
                       const std::vector<int> &slab) {
        std::vector<long> e, s;
        for (int l : labs) {
            const size_t ps = static_cast<size_t>(std::find(slab.begin(), slab.end(), l) - slab.begin());
            if (src.extent(ps) != C.extent(cpos(l))) throw std::invalid_argument("btas shim: contract: C extent mismatch");
            e.push_back(src.extent(ps)), s.push_back(cstride[cpos(l)]);
        }
        o.resize(static_cast<size_t>(n));
        for (long x = 0; x < n; ++x) {
            long rem = x, off = 0;
            for (size_t k = e.size(); k-- > 0;) off += (rem % e[k]) * s[k], rem /= e[k];
            o[static_cast<size_t>(x)] = off;
        }
    };
    std::vector<long> roff, coff;
    offsets(fa, M, roff, A, la), offsets(fb, N, coff, B, lb);
    double *c = C.data();
    const double al = static_cast<double>(alpha), be = static_cast<double>(beta);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < M; ++i)
        for (long j = 0; j < N; ++j) {
            double &dst = c[roff[static_cast<size_t>(i)] + coff[static_cast<size_t>(j)]];
            dst = be * dst + al * Cp[static_cast<size_t>(i * N + j)];
        }
}

template<typename S, typename UA, typename UB, typename UC>
void contract(const S &alpha, const Tensor<double> &A, std::initializer_list<UA> aA, const Tensor<double> &B,
              std::initializer_list<UB> aB, const S &beta, Tensor<double> &C, std::initializer_list<UC> aC) {
    std::vector<int> la, lb, lc;
    for (auto x : aA) la.push_back(static_cast<int>(x));
    for (auto x : aB) lb.push_back(static_cast<int>(x));
    for (auto x : aC) lc.push_back(static_cast<int>(x));
    contract_labels(static_cast<double>(alpha), A, la, B, lb, static_cast<double>(beta), C, lc);
}

}// namespace btas
