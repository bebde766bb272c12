// fused refinement kernels for texel format U8P (one translation unit per format: parallel builds)
#include "pcl_refine.cuh"
#include <string.h>

PCL_RF_INSTANTIATE(PCL_FMT_U8P)
