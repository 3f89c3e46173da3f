#include "../../include/peclr_b200.h"
extern "C" int peclr_abi_version(void) { return 1; }
