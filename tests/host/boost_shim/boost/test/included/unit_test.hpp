// see ../unit_test.hpp; the "included" flavour also supplies main()
#pragma once
#include "../unit_test.hpp"
int main() { return ::jpgenc_boost_test::run_all(); }
