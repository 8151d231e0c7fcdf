/* Build shim for the reference arm: stands where src/components/tl/cuda/tl_cuda.h is for the UNMODIFIED reference
 * kernels (`#include "../tl_cuda.h"`).  The real header needs the whole UCC/UCS component tree; the kernels only use
 * the NVLS limits (values as in the reference's tl_cuda.h:34-37) and ucc_tl_cuda_nvls_control_t, which comes from the
 * reference's own tl_cuda_nvls.h (copied next to this file by build.sh, unmodified). */
#ifndef REF_ARM_TL_CUDA_H
#define REF_ARM_TL_CUDA_H
#include <sys/types.h>
#include <stdint.h>
#include "ucc/api/ucc.h"
#define HAVE_NVLS 1
#define UCC_TL_CUDA_MAX_NVLS_PEERS (72 * 8)
#define UCC_TL_CUDA_MAX_NVLS_SM_COUNT 32
#define UCC_TL_CUDA_MAX_NVLS_THREADS 1024
#include "tl_cuda_nvls.h"
#endif
