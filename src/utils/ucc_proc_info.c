#include "ucc_proc_info.h"
#include "ucc_sys.h"
#include "ucc_log.h"
#include <sched.h>
#include <stdio.h>
#include <unistd.h>

ucc_proc_info_t ucc_local_proc;

static int read_int_file(const char *path, int *v)
{
    FILE *f = fopen(path, "r"); int ok;
    if (!f) return 0;
    ok = fscanf(f, "%d", v) == 1; fclose(f); return ok;
}

/* socket / numa of the first cpu in this process' affinity mask, only if the
 * whole mask lives on one socket / numa (otherwise "unbound" = INVALID) */
ucc_status_t ucc_local_proc_info_init(void)
{
    cpu_set_t set; char path[128];
    int sock = -1, numa = -1, multi_sock = 0, multi_numa = 0;
    ucc_local_proc.host_hash = ucc_sys_host_hash();
    ucc_local_proc.pid       = getpid();
    ucc_local_proc.host_id   = 0;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        for (int c = 0; c < CPU_SETSIZE; c++) {
            int s, n = -1;
            if (!CPU_ISSET(c, &set)) continue;
            snprintf(path, sizeof(path), "/sys/devices/system/cpu/cpu%d/topology/physical_package_id", c);
            if (read_int_file(path, &s)) { if (sock < 0) sock = s; else if (sock != s) multi_sock = 1; }
            for (int k = 0; k < 64; k++) {
                snprintf(path, sizeof(path), "/sys/devices/system/cpu/cpu%d/node%d", c, k);
                if (access(path, F_OK) == 0) { n = k; break; }
            }
            if (n >= 0) { if (numa < 0) numa = n; else if (numa != n) multi_numa = 1; }
        }
    }
    ucc_local_proc.socket_id = (sock >= 0 && !multi_sock) ? (ucc_socket_id_t)sock : UCC_SOCKET_ID_INVALID;
    ucc_local_proc.numa_id   = (numa >= 0 && !multi_numa) ? (ucc_numa_id_t)numa : UCC_NUMA_ID_INVALID;
    return UCC_OK;
}
