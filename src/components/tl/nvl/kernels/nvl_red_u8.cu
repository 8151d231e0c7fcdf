/* tl/nvl reduction kernels instantiated for uint8_t */
#include "nvl_reduce_impl.cuh"
NVL_INSTANTIATE_REDUCE(uint8_t, u8)
