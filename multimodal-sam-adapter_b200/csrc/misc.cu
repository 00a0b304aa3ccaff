#include "common.cuh"
MMSAM_API int mmsam_arch(void) { return 100; }
