// forwards the Boost.uBLAS names the reference's tests include to the host mirror's bundled matrix types
#pragma once
#include "ublas_lite.hpp"
namespace boost { namespace numeric { namespace ublas { using namespace ::jpgenc::ublas; } } }
