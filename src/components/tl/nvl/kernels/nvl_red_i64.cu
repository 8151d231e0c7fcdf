/* tl/nvl reduction kernels instantiated for int64_t */
#include "nvl_reduce_impl.cuh"
NVL_INSTANTIATE_REDUCE(int64_t, i64)
