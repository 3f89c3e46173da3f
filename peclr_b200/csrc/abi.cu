#include "../../include/peclr_b200.h"
extern "C" int peclr_abi_version(void) { return 2; }  // 2: fp64 BatchNorm sums, space-to-depth stem, accumulator-set scratch
