/* tl/nvl reduction kernels instantiated for float */
#include "nvl_reduce_impl.cuh"
NVL_INSTANTIATE_REDUCE(float, f32)
