// Stand-in: the reference includes ublas/io.hpp but never streams a matrix on the encode path.
#pragma once
#include "matrix.hpp"
