#include "../../include/peclr_b200.h"
// 3: reproducible reductions (fp64 accumulators, ordered workspace slabs), workspace arguments of wgrad / sgemm
extern "C" int peclr_abi_version(void) { return 3; }
