// Stand-in: main.cpp includes boost/dynamic_bitset.hpp and never uses it.
#pragma once
