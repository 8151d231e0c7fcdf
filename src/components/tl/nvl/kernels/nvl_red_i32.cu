/* tl/nvl reduction kernels instantiated for int32_t */
#include "nvl_reduce_impl.cuh"
NVL_INSTANTIATE_REDUCE(int32_t, i32)
