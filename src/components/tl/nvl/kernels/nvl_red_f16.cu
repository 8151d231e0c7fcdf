/* tl/nvl reduction kernels instantiated for __half */
#include "nvl_reduce_impl.cuh"
NVL_INSTANTIATE_REDUCE(__half, f16)
