/* tl/nvl reduction kernels instantiated for double */
#include "nvl_reduce_impl.cuh"
NVL_INSTANTIATE_REDUCE(double, f64)
