// native_tu.cu -- translation unit of a hand-written scene kernel: the pixel-loop kernel with
// SBX_APP_HEADER set to native/app_<x>_native.h by the Makefile.
#include "sbx/sbx_kernel.cuh"
