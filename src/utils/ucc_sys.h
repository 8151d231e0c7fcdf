/* OS helpers: shared memory segments (SysV + POSIX), library path, host hash, fd passing. */
#ifndef UCC_SYS_H_
#define UCC_SYS_H_
#include "ucc_compiler_def.h"
ucc_status_t ucc_sysv_alloc(size_t *size, void **addr, int *shmid);
ucc_status_t ucc_sysv_attach(int shmid, void **addr);
ucc_status_t ucc_sysv_free(void *addr);
ucc_status_t ucc_shm_create(const char *name, size_t size, void **addr);   /* POSIX shm, O_EXCL */
ucc_status_t ucc_shm_attach(const char *name, size_t size, void **addr);
ucc_status_t ucc_shm_detach(void *addr, size_t size);
ucc_status_t ucc_shm_unlink(const char *name);
int          ucc_shm_reap_stale(const char *prefix); /* unlink /dev/shm/<prefix><pid>... whose pid is dead; returns how many */
const char  *ucc_sys_get_lib_path(void);   /* absolute path of the .so containing this function */
const char  *ucc_sys_dirname_of_lib(void);
uint64_t     ucc_sys_host_hash(void);
size_t       ucc_get_page_size(void);
/* fd passing over abstract unix sockets (VMM / multicast handle exchange) */
int          ucc_sys_fd_server_open(const char *name);                 /* returns listening socket or -1 */
int          ucc_sys_fd_server_serve_once(int lsock, int fd_to_send);  /* non-blocking: 1 served, 0 none, -1 err */
int          ucc_sys_fd_recv(const char *name, int timeout_ms);        /* returns received fd or -1 */
int          ucc_sys_pidfd_getfd(int pid, int remote_fd);              /* -1 if unsupported */
#endif
