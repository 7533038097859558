// msq_api.cu -- library-level entry points of libmsq_b200.so
#include "msq_api_internal.h"

extern "C" int msq_version(void) { return 100; }   // 0.1.0 -> round 1
