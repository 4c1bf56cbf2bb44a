// Single translation unit of libfoley_b200.so: the kernels live in headers (rowwise.cuh, gemm.cuh,
// attention.cuh) and the device debug words must exist exactly once, so the three sources build as one.
#include "api.cu"
#include "engine.cu"
#include "dac.cu"
#include "encoders.cu"
