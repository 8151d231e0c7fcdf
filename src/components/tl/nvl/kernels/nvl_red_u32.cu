/* tl/nvl reduction kernels instantiated for uint32_t */
#include "nvl_reduce_impl.cuh"
NVL_INSTANTIATE_REDUCE(uint32_t, u32)
