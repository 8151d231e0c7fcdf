#include "ucc_cuda_util.h"
#include <pthread.h>

ucc_cu_api_t ucc_cu;
static pthread_mutex_t cu_lock = PTHREAD_MUTEX_INITIALIZER;

#define LOAD(_name)                                                                                           \
    do { void *fn = NULL; enum cudaDriverEntryPointQueryResult qr;                                            \
        if (cudaGetDriverEntryPoint(#_name, &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) {             \
            (void)cudaGetLastError(); ucc_debug("driver entry point %s is not available", #_name); fn = NULL; } \
        *(void **)&ucc_cu._name = fn; } while (0)

ucc_status_t ucc_cu_api_load(void)
{
    pthread_mutex_lock(&cu_lock);
    if (!ucc_cu.loaded) {
        LOAD(cuMemGetAddressRange); LOAD(cuMemCreate); LOAD(cuMemRelease); LOAD(cuMemAddressReserve); LOAD(cuMemAddressFree);
        LOAD(cuMemMap); LOAD(cuMemUnmap); LOAD(cuMemSetAccess); LOAD(cuMemGetAllocationGranularity);
        LOAD(cuMemExportToShareableHandle); LOAD(cuMemImportFromShareableHandle); LOAD(cuMulticastCreate); LOAD(cuMulticastAddDevice);
        LOAD(cuMulticastBindMem); LOAD(cuMulticastUnbind); LOAD(cuMulticastGetGranularity); LOAD(cuDeviceGet); LOAD(cuDeviceGetAttribute);
        LOAD(cuGetErrorString); LOAD(cuStreamWriteValue32); LOAD(cuCtxGetDevice); LOAD(cuStreamWaitValue32);
        ucc_cu.loaded = 1;
    }
    pthread_mutex_unlock(&cu_lock);
    return ucc_cu.cuMemGetAddressRange ? UCC_OK : UCC_ERR_NOT_SUPPORTED;
}
const char *ucc_cu_err_str(CUresult r)
{ const char *s = "unknown"; if (ucc_cu.cuGetErrorString) ucc_cu.cuGetErrorString(r, &s); return s ? s : "unknown"; }
ucc_status_t ucc_cuda_get_address_range(const void *ptr, void **base, size_t *len)
{
    CUdeviceptr b; size_t l;
    if (ucc_cu_api_load() != UCC_OK) return UCC_ERR_NOT_SUPPORTED;
    if (ucc_cu.cuMemGetAddressRange(&b, &l, (CUdeviceptr)(uintptr_t)ptr) != CUDA_SUCCESS) return UCC_ERR_NOT_SUPPORTED;
    *base = (void *)(uintptr_t)b; *len = l;
    return UCC_OK;
}
