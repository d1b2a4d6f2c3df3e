// Picks the matrix types the mirror headers expose: real Boost.uBLAS when it is installed (what the reference uses),
// otherwise the bundled ublas_lite under the same names, so that caller code written against the reference compiles.
#pragma once
#if defined(JPGENC_USE_BOOST) || (defined(__has_include) && __has_include(<boost/numeric/ublas/matrix.hpp>))
#include <boost/numeric/ublas/matrix.hpp>
#include <boost/numeric/ublas/matrix_proxy.hpp>
#else
#include "ublas_lite.hpp"
namespace boost { namespace numeric { namespace ublas { using namespace ::jpgenc::ublas; } } }
#endif
using boost::numeric::ublas::matrix;
using boost::numeric::ublas::matrix_range;
using boost::numeric::ublas::range;
