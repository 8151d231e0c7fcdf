/* tl/nvl reduction kernels instantiated for __nv_bfloat16 */
#include "nvl_reduce_impl.cuh"
NVL_INSTANTIATE_REDUCE(__nv_bfloat16, bf16)
