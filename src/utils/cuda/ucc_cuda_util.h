/* Small CUDA helpers shared by the CUDA plugin modules (runtime API error
 * mapping, driver entry points resolved through the runtime so no module
 * links libcuda directly). */
#ifndef UCC_CUDA_UTIL_H_
#define UCC_CUDA_UTIL_H_
#include <cuda_runtime_api.h>
#include <cuda.h>
#include "utils/ucc_log.h"
#define UCC_CUDA_MAX_DEVICES 32

static inline ucc_status_t ucc_cuda_error_to_status(cudaError_t e)
{
    switch (e) {
    case cudaSuccess: return UCC_OK;
    case cudaErrorNotReady: return UCC_INPROGRESS;
    case cudaErrorInvalidValue: return UCC_ERR_INVALID_PARAM;
    case cudaErrorMemoryAllocation: return UCC_ERR_NO_MEMORY;
    default: return UCC_ERR_NO_MESSAGE;
    }
}
#define CUDA_CHECK(_call)                                                                              \
    do { cudaError_t _e = (_call); if (ucc_unlikely(_e != cudaSuccess)) { (void)cudaGetLastError();       \
        ucc_error("%s failed: %d (%s)", #_call, (int)_e, cudaGetErrorString(_e)); return ucc_cuda_error_to_status(_e); } } while (0)
#define CUDA_CHECK_GOTO(_call, _label, _st)                                                            \
    do { cudaError_t _e = (_call); if (ucc_unlikely(_e != cudaSuccess)) { (void)cudaGetLastError();       \
        ucc_error("%s failed: %d (%s)", #_call, (int)_e, cudaGetErrorString(_e)); _st = ucc_cuda_error_to_status(_e); goto _label; } } while (0)

#ifdef __cplusplus
extern "C" {
#endif
/* driver API table resolved lazily with cudaGetDriverEntryPoint */
typedef struct ucc_cu_api {
    CUresult (*cuMemGetAddressRange)(CUdeviceptr *, size_t *, CUdeviceptr);
    CUresult (*cuMemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long);
    CUresult (*cuMemRelease)(CUmemGenericAllocationHandle);
    CUresult (*cuMemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long);
    CUresult (*cuMemAddressFree)(CUdeviceptr, size_t);
    CUresult (*cuMemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
    CUresult (*cuMemUnmap)(CUdeviceptr, size_t);
    CUresult (*cuMemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t);
    CUresult (*cuMemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags);
    CUresult (*cuMemExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
    CUresult (*cuMemImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType);
    CUresult (*cuMulticastCreate)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *);
    CUresult (*cuMulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice);
    CUresult (*cuMulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
    CUresult (*cuMulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t);
    CUresult (*cuMulticastGetGranularity)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags);
    CUresult (*cuDeviceGet)(CUdevice *, int);
    CUresult (*cuDeviceGetAttribute)(int *, CUdevice_attribute, CUdevice);
    CUresult (*cuGetErrorString)(CUresult, const char **);
    CUresult (*cuStreamWriteValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    CUresult (*cuCtxGetDevice)(CUdevice *);
    CUresult (*cuStreamWaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    int loaded;
} ucc_cu_api_t;
extern ucc_cu_api_t ucc_cu;
ucc_status_t ucc_cu_api_load(void);
ucc_status_t ucc_cuda_get_address_range(const void *ptr, void **base, size_t *len);
const char  *ucc_cu_err_str(CUresult r);
#define CU_CHECK(_call)                                                                                  \
    do { CUresult _r = (_call); if (ucc_unlikely(_r != CUDA_SUCCESS)) {                                   \
        ucc_error("%s failed: %d (%s)", #_call, (int)_r, ucc_cu_err_str(_r)); return UCC_ERR_NO_MESSAGE; } } while (0)
#ifdef __cplusplus
}
#endif
#endif
