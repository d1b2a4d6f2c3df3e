// Stand-in for boost::math::constants::{pi, root_two} (the two constants Dct.hpp uses).
// The literals are the correctly rounded doubles, which is what Boost returns for T = double.
#pragma once
namespace boost { namespace math { namespace constants {
template <class T> inline T pi() { return static_cast<T>(3.141592653589793238462643383279502884L); }
template <class T> inline T root_two() { return static_cast<T>(1.414213562373095048801688724209698078L); }
}}}
