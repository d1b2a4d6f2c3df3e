// Stand-in for ublas::matrix_range / subrange (see matrix.hpp in this directory). Test infrastructure only.
#pragma once
#include "matrix.hpp"

namespace boost { namespace numeric { namespace ublas {

template <class M>
class matrix_range {
public:
    typedef typename M::value_type value_type;
    matrix_range(M& m, const range& r1, const range& r2)
        : m_(&m), i0_(r1.start()), j0_(r2.start()), r_(r1.size()), c_(r2.size()) {}
    // the reference builds const ranges over members it only reads
    matrix_range(const M& m, const range& r1, const range& r2)
        : m_(const_cast<M*>(&m)), i0_(r1.start()), j0_(r2.start()), r_(r1.size()), c_(r2.size()) {}

    std::size_t size1() const { return r_; }
    std::size_t size2() const { return c_; }
    value_type& operator()(std::size_t i, std::size_t j) { return (*m_)(i0_ + i, j0_ + j); }
    const value_type& operator()(std::size_t i, std::size_t j) const { return (*m_)(i0_ + i, j0_ + j); }

    template <class E>
    matrix_range& assign(const E& e) {
        for (std::size_t i = 0; i < r_; ++i)
            for (std::size_t j = 0; j < c_; ++j)
                (*m_)(i0_ + i, j0_ + j) = static_cast<value_type>(e(i, j));
        return *this;
    }
    template <class E>
    auto operator=(const E& e) -> decltype(e.size1(), *this) { return assign(e); }
    matrix_range& operator=(const matrix_range& o) { return assign(o); }
    matrix_range(const matrix_range&) = default;

    template <class S>
    matrix_range& operator*=(const S& s) {
        for (std::size_t i = 0; i < r_; ++i)
            for (std::size_t j = 0; j < c_; ++j)
                (*m_)(i0_ + i, j0_ + j) *= s;
        return *this;
    }

private:
    M* m_;
    std::size_t i0_, j0_, r_, c_;
};

template <class M>
matrix_range<M> subrange(M& m, std::size_t r0, std::size_t r1, std::size_t c0, std::size_t c1) {
    return matrix_range<M>(m, range(r0, r1), range(c0, c1));
}
template <class M>
const matrix_range<M> subrange(const M& m, std::size_t r0, std::size_t r1, std::size_t c0, std::size_t c1) {
    return matrix_range<M>(m, range(r0, r1), range(c0, c1));
}

}}} // namespace boost::numeric::ublas
