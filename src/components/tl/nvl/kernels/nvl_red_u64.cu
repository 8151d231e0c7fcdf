/* tl/nvl reduction kernels instantiated for uint64_t */
#include "nvl_reduce_impl.cuh"
NVL_INSTANTIATE_REDUCE(uint64_t, u64)
