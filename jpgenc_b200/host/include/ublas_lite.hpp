// Dense row-major matrix + range views: the slice of Boost.uBLAS that jpgEnc's public signatures mention
// (include/Image.hpp:14,104-114; include/Dct.hpp:47,238,264; include/Coding.hpp:7-11).  Used only when Boost itself
// is not installed, so that code written against the reference headers still compiles against this mirror.
#pragma once
#include <algorithm>
#include <cassert>
#include <cstddef>
#include <type_traits>
#include <vector>

namespace jpgenc { namespace ublas {

struct range {
    std::size_t first, last;
    range(std::size_t a = 0, std::size_t b = 0) : first(a), last(b) {}
    std::size_t start() const { return first; }
    std::size_t size() const { return last - first; }
};

template <class T>
struct zero_matrix {
    using value_type = T;
    std::size_t rows, cols;
    zero_matrix(std::size_t r, std::size_t c) : rows(r), cols(c) {}
    std::size_t size1() const { return rows; }
    std::size_t size2() const { return cols; }
    T operator()(std::size_t, std::size_t) const { return T{}; }
};

template <class E>
using is_expr = decltype(std::declval<const E&>().size1(), std::declval<const E&>()(0, 0), 0);

template <class T>
class matrix {
public:
    using value_type = T;
    using array_type = std::vector<T>;
    matrix() = default;
    matrix(std::size_t rows, std::size_t cols) : rows_(rows), cols_(cols), cells_(rows * cols) {}
    template <class E, is_expr<E> = 0>
    matrix(const E& e) { assign_from(e); }
    template <class E, is_expr<E> = 0>
    matrix& operator=(const E& e) { assign_from(e); return *this; }

    std::size_t size1() const { return rows_; }
    std::size_t size2() const { return cols_; }
    T& operator()(std::size_t r, std::size_t c) { return cells_[r * cols_ + c]; }
    const T& operator()(std::size_t r, std::size_t c) const { return cells_[r * cols_ + c]; }
    array_type& data() { return cells_; }
    const array_type& data() const { return cells_; }
    void clear() { std::fill(cells_.begin(), cells_.end(), T{}); }
    void resize(std::size_t rows, std::size_t cols, bool preserve = true) {
        array_type grown(rows * cols);
        if (preserve)
            for (std::size_t r = 0; r < std::min(rows, rows_); ++r)
                std::move(cells_.begin() + r * cols_, cells_.begin() + r * cols_ + std::min(cols, cols_), grown.begin() + r * cols);
        cells_.swap(grown);
        rows_ = rows;
        cols_ = cols;
    }
    template <class S>
    matrix& operator*=(S s) { for (T& v : cells_) v *= s; return *this; }

private:
    template <class E>
    void assign_from(const E& e) {
        array_type fresh(e.size1() * e.size2());
        for (std::size_t r = 0; r < e.size1(); ++r)
            for (std::size_t c = 0; c < e.size2(); ++c) fresh[r * e.size2() + c] = static_cast<T>(e(r, c));
        rows_ = e.size1();
        cols_ = e.size2();
        cells_.swap(fresh);
    }
    std::size_t rows_ = 0, cols_ = 0;
    array_type cells_;
};

template <class M>
class matrix_range {
public:
    using value_type = typename M::value_type;
    matrix_range(M& m, range rows, range cols) : m_(&m), r0_(rows.start()), c0_(cols.start()), nr_(rows.size()), nc_(cols.size()) {}
    matrix_range(const M& m, range rows, range cols) : matrix_range(const_cast<M&>(m), rows, cols) {}
    std::size_t size1() const { return nr_; }
    std::size_t size2() const { return nc_; }
    value_type& operator()(std::size_t r, std::size_t c) { return (*m_)(r0_ + r, c0_ + c); }
    const value_type& operator()(std::size_t r, std::size_t c) const { return (*m_)(r0_ + r, c0_ + c); }
    template <class E, is_expr<E> = 0>
    matrix_range& assign(const E& e) {
        for (std::size_t r = 0; r < nr_; ++r)
            for (std::size_t c = 0; c < nc_; ++c) (*this)(r, c) = static_cast<value_type>(e(r, c));
        return *this;
    }
    template <class E, is_expr<E> = 0>
    matrix_range& operator=(const E& e) { return assign(e); }
    template <class S>
    matrix_range& operator*=(S s) {
        for (std::size_t r = 0; r < nr_; ++r)
            for (std::size_t c = 0; c < nc_; ++c) (*this)(r, c) *= s;
        return *this;
    }

private:
    M* m_;
    std::size_t r0_, c0_, nr_, nc_;
};

template <class M>
matrix_range<M> subrange(M& m, std::size_t r0, std::size_t r1, std::size_t c0, std::size_t c1) { return {m, range(r0, r1), range(c0, c1)}; }
template <class M>
const matrix_range<M> subrange(const M& m, std::size_t r0, std::size_t r1, std::size_t c0, std::size_t c1) { return {m, range(r0, r1), range(c0, c1)}; }

template <class E>
matrix<typename E::value_type> trans(const E& e) {
    matrix<typename E::value_type> t(e.size2(), e.size1());
    for (std::size_t r = 0; r < e.size1(); ++r)
        for (std::size_t c = 0; c < e.size2(); ++c) t(c, r) = e(r, c);
    return t;
}
template <class A, class B>
matrix<typename A::value_type> prod(const A& a, const B& b) {
    matrix<typename A::value_type> out(a.size1(), b.size2());
    for (std::size_t r = 0; r < a.size1(); ++r)
        for (std::size_t c = 0; c < b.size2(); ++c) {
            typename A::value_type acc{};
            for (std::size_t k = 0; k < a.size2(); ++k) acc += a(r, k) * b(k, c);
            out(r, c) = acc;
        }
    return out;
}
template <class E>
E& noalias(E& e) { return e; }

}}  // namespace jpgenc::ublas
