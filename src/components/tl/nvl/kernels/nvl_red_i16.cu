/* tl/nvl reduction kernels instantiated for int16_t */
#include "nvl_reduce_impl.cuh"
NVL_INSTANTIATE_REDUCE(int16_t, i16)
