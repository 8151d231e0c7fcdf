/* tl/nvl reduction kernels instantiated for uint16_t */
#include "nvl_reduce_impl.cuh"
NVL_INSTANTIATE_REDUCE(uint16_t, u16)
