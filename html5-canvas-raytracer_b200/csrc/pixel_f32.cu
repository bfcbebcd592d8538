// pixel_f32.cu — the per-pixel kernel in fp32 (small scenes; cross-check of the wavefront path).
#include "rtb_internal.h"
#include "pixel_kernel.cuh"
#define REAL float
#define LAUNCH_NAME launch_pixel_f32
#include "pixel_launch.inl"
