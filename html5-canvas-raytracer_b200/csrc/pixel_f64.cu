// pixel_f64.cu — float64 VALIDATION kernel.  Build this TU with -fmad=false: JS rounds
// after every operator, so no multiply-add may be contracted.
#include "rtb_internal.h"
#include "pixel_kernel.cuh"
#define REAL double
#define LAUNCH_NAME launch_pixel_f64
#include "pixel_launch.inl"
