/* tl/nvl reduction kernels instantiated for int8_t */
#include "nvl_reduce_impl.cuh"
NVL_INSTANTIATE_REDUCE(int8_t, i8)
