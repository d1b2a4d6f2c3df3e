// Stand-in for the tiny subset of Boost.uBLAS that the reference encoder uses.
// TEST INFRASTRUCTURE ONLY (oracle/): Boost is not installed in this image, and the reference
// only uses ublas::matrix as a dense row-major array (no hot-path arithmetic lives in Boost,
// see SURVEY.md section 8c).  This header is written from scratch; it is not Boost code.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <utility>
#include <vector>

namespace boost { namespace numeric { namespace ublas {

template <class T> class matrix;

// half-open index interval [lo, hi)
class range {
public:
    range() : lo_(0), hi_(0) {}
    range(std::size_t lo, std::size_t hi) : lo_(lo), hi_(hi) {}
    std::size_t start() const { return lo_; }
    std::size_t size() const { return hi_ - lo_; }
private:
    std::size_t lo_, hi_;
};

template <class T>
class zero_matrix {
public:
    typedef T value_type;
    zero_matrix(std::size_t r, std::size_t c) : r_(r), c_(c) {}
    std::size_t size1() const { return r_; }
    std::size_t size2() const { return c_; }
    T operator()(std::size_t, std::size_t) const { return T(); }
private:
    std::size_t r_, c_;
};

template <class T>
class matrix {
public:
    typedef T value_type;
    typedef std::vector<T> array_type;

    matrix() : r_(0), c_(0) {}
    matrix(std::size_t r, std::size_t c) : r_(r), c_(c), d_(r * c) {}
    matrix(const matrix&) = default;
    matrix(matrix&&) = default;
    matrix& operator=(const matrix&) = default;
    matrix& operator=(matrix&&) = default;

    // converting construction / assignment from anything that looks like a 2-D expression
    template <class E>
    matrix(const E& e, decltype(e.size1())* = nullptr) { take(e); }
    template <class E>
    auto operator=(const E& e) -> decltype(e.size1(), *this) { take(e); return *this; }

    std::size_t size1() const { return r_; }
    std::size_t size2() const { return c_; }
    T& operator()(std::size_t i, std::size_t j) { return d_[i * c_ + j]; }
    const T& operator()(std::size_t i, std::size_t j) const { return d_[i * c_ + j]; }
    array_type& data() { return d_; }
    const array_type& data() const { return d_; }

    void resize(std::size_t r, std::size_t c, bool preserve = true) {
        array_type nd(r * c);
        if (preserve) {
            const std::size_t rr = std::min(r, r_), cc = std::min(c, c_);
            for (std::size_t i = 0; i < rr; ++i)
                for (std::size_t j = 0; j < cc; ++j)
                    nd[i * c + j] = std::move(d_[i * c_ + j]);
        }
        d_.swap(nd);
        r_ = r; c_ = c;
    }
    void clear() { std::fill(d_.begin(), d_.end(), T()); }

    template <class S>
    matrix& operator*=(const S& s) { for (auto& v : d_) v *= s; return *this; }

private:
    template <class E>
    void take(const E& e) {
        const std::size_t r = e.size1(), c = e.size2();
        array_type nd(r * c);
        for (std::size_t i = 0; i < r; ++i)
            for (std::size_t j = 0; j < c; ++j)
                nd[i * c + j] = static_cast<T>(e(i, j));
        d_.swap(nd);
        r_ = r; c_ = c;
    }
    std::size_t r_, c_;
    array_type d_;
};

template <class T>
void swap(matrix<T>& a, matrix<T>& b) { std::swap(a, b); }

// transpose / product, only reached from the non-default dctMat and its test helper
template <class E>
auto trans(const E& e) -> matrix<typename E::value_type> {
    matrix<typename E::value_type> t(e.size2(), e.size1());
    for (std::size_t i = 0; i < e.size1(); ++i)
        for (std::size_t j = 0; j < e.size2(); ++j)
            t(j, i) = e(i, j);
    return t;
}

template <class A, class B>
auto prod(const A& a, const B& b) -> matrix<typename A::value_type> {
    assert(a.size2() == b.size1());
    matrix<typename A::value_type> p(a.size1(), b.size2());
    for (std::size_t i = 0; i < a.size1(); ++i)
        for (std::size_t j = 0; j < b.size2(); ++j) {
            typename A::value_type s = typename A::value_type();
            for (std::size_t k = 0; k < a.size2(); ++k)
                s += a(i, k) * b(k, j);
            p(i, j) = s;
        }
    return p;
}

template <class E>
E& noalias(E& e) { return e; }

}}} // namespace boost::numeric::ublas
