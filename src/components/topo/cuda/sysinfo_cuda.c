/* sysinfo/cuda: NVML scan of the host's GPUs and their NVLink connectivity
 * (role of reference components/topo/cuda/ucc_sysinfo_cuda.c).  NVML is dlopen'ed; on an
 * NVSwitch system every GPU pair is connected with all links of the GPU through the switch. */
#include "components/topo/ucc_sysinfo.h"
#include "utils/ucc_string.h"
#include "utils/ucc_log.h"
#include <cuda_runtime_api.h>
#include <nvml.h>
#include <dlfcn.h>
#include <stdlib.h>

static struct {
    void *h;
    nvmlReturn_t (*Init)(void);
    nvmlReturn_t (*Shutdown)(void);
    nvmlReturn_t (*GetCount)(unsigned *);
    nvmlReturn_t (*GetHandle)(unsigned, nvmlDevice_t *);
    nvmlReturn_t (*GetPci)(nvmlDevice_t, nvmlPciInfo_t *);
    nvmlReturn_t (*GetUUID)(nvmlDevice_t, char *, unsigned);
    nvmlReturn_t (*LinkState)(nvmlDevice_t, unsigned, nvmlEnableState_t *);
    nvmlReturn_t (*LinkRemotePci)(nvmlDevice_t, unsigned, nvmlPciInfo_t *);
    nvmlReturn_t (*LinkRemoteType)(nvmlDevice_t, unsigned, nvmlIntNvLinkDeviceType_t *);
    nvmlReturn_t (*FabricInfo)(nvmlDevice_t, nvmlGpuFabricInfo_t *);
} nv;

static ucc_status_t sysinfo_cuda_init(void)
{
    if (nv.h) return UCC_OK;
    nv.h = dlopen("libnvidia-ml.so.1", RTLD_LAZY);
    if (!nv.h) { ucc_debug("NVML is not available: %s", dlerror()); return UCC_ERR_NO_RESOURCE; }
#define SYM(_f, _n) *(void **)&nv._f = dlsym(nv.h, _n)
    SYM(Init, "nvmlInit_v2"); SYM(Shutdown, "nvmlShutdown"); SYM(GetCount, "nvmlDeviceGetCount_v2"); SYM(GetHandle, "nvmlDeviceGetHandleByIndex_v2");
    SYM(GetPci, "nvmlDeviceGetPciInfo_v3"); SYM(GetUUID, "nvmlDeviceGetUUID"); SYM(LinkState, "nvmlDeviceGetNvLinkState");
    SYM(LinkRemotePci, "nvmlDeviceGetNvLinkRemotePciInfo_v2"); SYM(LinkRemoteType, "nvmlDeviceGetNvLinkRemoteDeviceType"); SYM(FabricInfo, "nvmlDeviceGetGpuFabricInfo");
    if (!nv.Init || !nv.GetCount || !nv.GetHandle || !nv.GetPci || nv.Init() != NVML_SUCCESS) { dlclose(nv.h); nv.h = NULL; return UCC_ERR_NO_RESOURCE; }
    return UCC_OK;
}

static ucc_status_t sysinfo_cuda_get_host_info(ucc_host_info_t *info)
{
    unsigned n = 0;
    nvmlPciInfo_t pci[UCC_MAX_HOST_GPUS];
    int cur = -1, ncuda = 0;
    if (!nv.h || nv.GetCount(&n) != NVML_SUCCESS) return UCC_ERR_NO_RESOURCE;
    if (n > UCC_MAX_HOST_GPUS) n = UCC_MAX_HOST_GPUS;
    info->n_gpus = (uint8_t)n;
    for (unsigned i = 0; i < n; i++) {
        nvmlDevice_t d; char uuid[96] = ""; ucc_gpu_info_t *g = &info->gpus[i];
        memset(g, 0, sizeof(*g)); memset(&pci[i], 0, sizeof(pci[i]));
        if (nv.GetHandle(i, &d) != NVML_SUCCESS) continue;
        if (nv.GetPci(d, &pci[i]) == NVML_SUCCESS) { g->pci.domain = (uint16_t)pci[i].domain; g->pci.bus = (uint8_t)pci[i].bus; g->pci.device = (uint8_t)pci[i].device; g->pci.function = 0; }
        if (nv.GetUUID && nv.GetUUID(d, uuid, sizeof(uuid)) == NVML_SUCCESS) g->uuid_hash = ucc_str_hash_djb2(uuid);
        if (nv.FabricInfo) {
            nvmlGpuFabricInfo_t fi; memset(&fi, 0, sizeof(fi));
            if (nv.FabricInfo(d, &fi) == NVML_SUCCESS && fi.state == NVML_GPU_FABRIC_STATE_COMPLETED && fi.status == NVML_SUCCESS) {
                g->caps |= UCC_GPU_CAP_FABRIC; g->fabric_clique_id = fi.cliqueId;
                g->fabric_partition = ucc_str_hash_djb2((const char *)fi.clusterUuid) ^ ((uint64_t)fi.clusterUuid[0] << 32 | fi.clusterUuid[15]);
            }
        }
    }
    /* NVLink walk: a link whose far end is another GPU adds one edge, a switch connects the GPU to every other switch-attached GPU */
    for (unsigned i = 0; i < n && nv.LinkState; i++) {
        nvmlDevice_t d; unsigned links = 0, sw_links = 0;
        if (nv.GetHandle(i, &d) != NVML_SUCCESS) continue;
        for (unsigned l = 0; l < NVML_NVLINK_MAX_LINKS; l++) {
            nvmlEnableState_t on = NVML_FEATURE_DISABLED; nvmlPciInfo_t rp; nvmlIntNvLinkDeviceType_t rt = NVML_NVLINK_DEVICE_TYPE_UNKNOWN;
            if (nv.LinkState(d, l, &on) != NVML_SUCCESS || on != NVML_FEATURE_ENABLED) continue;
            links++;
            if (nv.LinkRemoteType && nv.LinkRemoteType(d, l, &rt) == NVML_SUCCESS && rt == NVML_NVLINK_DEVICE_TYPE_SWITCH) { sw_links++; continue; }
            if (nv.LinkRemotePci && nv.LinkRemotePci(d, l, &rp) == NVML_SUCCESS) {
                int found = 0;
                for (unsigned j = 0; j < n; j++) if (j != i && pci[j].domain == rp.domain && pci[j].bus == rp.bus && pci[j].device == rp.device) { info->nvlink_matrix[i][j]++; found = 1; }
                if (!found) sw_links++; /* far end is not a GPU we can see: a switch (or a GPU hidden by the container) */
            }
        }
        info->gpus[i].n_nvlinks = links;
        if (links) info->gpus[i].caps |= UCC_GPU_CAP_NVLINK;
        if (sw_links) { info->gpus[i].caps |= UCC_GPU_CAP_NVSWITCH; info->gpus[i].n_nvlinks = links; }
    }
    for (unsigned i = 0; i < n; i++) for (unsigned j = 0; j < n; j++)
        if (i != j && (info->gpus[i].caps & UCC_GPU_CAP_NVSWITCH) && (info->gpus[j].caps & UCC_GPU_CAP_NVSWITCH) && !info->nvlink_matrix[i][j])
            info->nvlink_matrix[i][j] = (uint8_t)ucc_min(info->gpus[i].n_nvlinks, info->gpus[j].n_nvlinks);
    /* which GPU does this process drive: match the current CUDA device to an NVML index by PCI bus id */
    info->visible_gpus = 0; info->current_gpu = -1;
    if (cudaGetDeviceCount(&ncuda) == cudaSuccess && ncuda > 0) {
        for (int c = 0; c < ncuda; c++) {
            char bus[32] = ""; unsigned dom, b, dv;
            if (cudaDeviceGetPCIBusId(bus, sizeof(bus), c) != cudaSuccess || sscanf(bus, "%x:%x:%x", &dom, &b, &dv) != 3) continue;
            for (unsigned i = 0; i < n; i++) if (pci[i].domain == dom && pci[i].bus == b && pci[i].device == dv) {
                info->visible_gpus |= (uint16_t)(1u << i);
                if (cudaGetDevice(&cur) == cudaSuccess && cur == c) { info->current_gpu = (int8_t)i; info->gpus[i].sm_count = 0; }
            }
        }
    }
    (void)cudaGetLastError();
    return UCC_OK;
}
static ucc_status_t sysinfo_cuda_finalize(void) { return UCC_OK; }

ucc_sysinfo_base_t ucc_sysinfo_cuda = {
    .super = {.name = "cuda", .score = 1}, .init = sysinfo_cuda_init, .get_host_info = sysinfo_cuda_get_host_info, .finalize = sysinfo_cuda_finalize};
