#include "ucc_sys.h"
#include "ucc_log.h"
#include "ucc_string.h"
#include "ucc_time.h"
#include <dirent.h>
#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <libgen.h>
#include <poll.h>
#include <pthread.h>
#include <signal.h>
#include <sys/file.h>
#include <sys/ipc.h>
#include <sys/mman.h>
#include <sys/shm.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <sys/un.h>
#include <unistd.h>

size_t ucc_get_page_size(void) { static long p = 0; if (!p) p = sysconf(_SC_PAGESIZE); return (size_t)p; }

ucc_status_t ucc_sysv_alloc(size_t *size, void **addr, int *shmid)
{
    size_t ps = ucc_get_page_size(), sz = (*size + ps - 1) / ps * ps;
    int    id = shmget(IPC_PRIVATE, sz, IPC_CREAT | IPC_EXCL | 0600);
    void  *p;
    if (id < 0) { ucc_error("shmget(%zu) failed: %s", sz, strerror(errno)); return UCC_ERR_NO_RESOURCE; }
    p = shmat(id, NULL, 0);
    /* mark for deletion now: segment disappears when the last process detaches */
    shmctl(id, IPC_RMID, NULL);
    if (p == (void *)-1) { ucc_error("shmat failed: %s", strerror(errno)); return UCC_ERR_NO_RESOURCE; }
    *size = sz; *addr = p; *shmid = id;
    return UCC_OK;
}
ucc_status_t ucc_sysv_attach(int shmid, void **addr)
{
    void *p = shmat(shmid, NULL, 0);
    if (p == (void *)-1) return UCC_ERR_NO_RESOURCE;
    *addr = p; return UCC_OK;
}
ucc_status_t ucc_sysv_free(void *addr) { return shmdt(addr) == 0 ? UCC_OK : UCC_ERR_INVALID_PARAM; }

/* The creator of a named segment keeps its descriptor open with a shared flock until it unlinks the name: the kernel drops the lock
 * when the process dies, however it dies, and the lock is visible across pid namespaces - the liveness proof ucc_shm_reap_stale needs */
static struct { char name[64]; int fd; } shm_owned[256];
static pthread_mutex_t shm_owned_lock = PTHREAD_MUTEX_INITIALIZER;
static void shm_owned_add(const char *name, int fd)
{
    pthread_mutex_lock(&shm_owned_lock);
    for (size_t i = 0; i < sizeof(shm_owned) / sizeof(shm_owned[0]); i++)
        if (!shm_owned[i].name[0]) { snprintf(shm_owned[i].name, sizeof(shm_owned[i].name), "%s", name); shm_owned[i].fd = fd; fd = -1; break; }
    pthread_mutex_unlock(&shm_owned_lock);
    if (fd >= 0) close(fd);   /* table full: the segment simply is not protected by a lock (its pid still is checked) */
}
static void shm_owned_drop(const char *name)
{
    pthread_mutex_lock(&shm_owned_lock);
    for (size_t i = 0; i < sizeof(shm_owned) / sizeof(shm_owned[0]); i++)
        if (shm_owned[i].name[0] && !strcmp(shm_owned[i].name, name)) { close(shm_owned[i].fd); shm_owned[i].name[0] = 0; break; }
    pthread_mutex_unlock(&shm_owned_lock);
}

ucc_status_t ucc_shm_create(const char *name, size_t size, void **addr)
{
    int   fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    void *p;
    if (fd < 0) { ucc_debug("shm_open(%s, create) failed: %s", name, strerror(errno)); return UCC_ERR_NO_RESOURCE; }
    if (ftruncate(fd, (off_t)size) != 0) { close(fd); shm_unlink(name); return UCC_ERR_NO_MEMORY; }
    p = mmap(NULL, size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (p == MAP_FAILED) { close(fd); shm_unlink(name); return UCC_ERR_NO_MEMORY; }
    if (flock(fd, LOCK_SH | LOCK_NB) == 0) shm_owned_add(name, fd); else close(fd);
    *addr = p; return UCC_OK;
}
ucc_status_t ucc_shm_attach(const char *name, size_t size, void **addr)
{
    int   fd = shm_open(name, O_RDWR, 0600);
    void *p;
    struct stat st;
    if (fd < 0) return UCC_ERR_NOT_FOUND;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < size) { close(fd); return UCC_ERR_NOT_FOUND; /* creator has not sized it yet */ }
    p = mmap(NULL, size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return UCC_ERR_NO_MEMORY;
    *addr = p; return UCC_OK;
}
ucc_status_t ucc_shm_detach(void *addr, size_t size) { return munmap(addr, size) == 0 ? UCC_OK : UCC_ERR_INVALID_PARAM; }
ucc_status_t ucc_shm_unlink(const char *name) { shm_owned_drop(name); return shm_unlink(name) == 0 ? UCC_OK : UCC_ERR_NOT_FOUND; }

/* Named POSIX segments survive a process that dies without its destructors (SIGKILL, a crash, a job scheduler's timeout).  Every
 * segment name of this library starts with `prefix` followed by the owner's pid: remove the ones whose owner no longer exists.
 * (The reference uses SysV segments marked IPC_RMID right after creation; POSIX names have to stay visible until the peers have
 * attached, which may be at any later team creation, hence this reaper.)  Returns the number of segments removed. */
int ucc_shm_reap_stale(const char *prefix)
{
    const char *dir = "/dev/shm";
    size_t plen = strlen(prefix);
    struct dirent *de;
    DIR *d = opendir(dir);
    int n = 0;
    if (!d) return 0;
    while ((de = readdir(d)) != NULL) {
        char name[300]; char *e; long pid;
        if (strncmp(de->d_name, prefix, plen)) continue;
        pid = strtol(de->d_name + plen, &e, 10);
        if (e == de->d_name + plen || pid <= 0 || (*e != '.' && *e != '\0')) continue;
        if (kill((pid_t)pid, 0) == 0 || errno != ESRCH) continue;        /* alive (or not ours to judge) */
        snprintf(name, sizeof(name), "/%s", de->d_name);
        {   /* a pid of another pid namespace looks dead from here: the owner's flock is the second proof */
            int fd = shm_open(name, O_RDWR, 0600);
            if (fd < 0) continue;
            if (flock(fd, LOCK_EX | LOCK_NB) == 0 && shm_unlink(name) == 0) n++;
            close(fd);
        }
    }
    closedir(d);
    if (n) ucc_debug("removed %d shared-memory segments of dead processes (%s*)", n, prefix);
    return n;
}

const char *ucc_sys_get_lib_path(void)
{
    static char path[4096 + 1] = ""; /* realpath() needs PATH_MAX */
    Dl_info     info;
    if (!path[0] && dladdr((void *)ucc_sys_get_lib_path, &info) && info.dli_fname) {
        if (!realpath(info.dli_fname, path)) snprintf(path, sizeof(path), "%s", info.dli_fname);
    }
    return path;
}
const char *ucc_sys_dirname_of_lib(void)
{
    static char dir[4096 + 1] = "";
    if (!dir[0]) { char tmp[4096 + 1]; snprintf(tmp, sizeof(tmp), "%s", ucc_sys_get_lib_path()); snprintf(dir, sizeof(dir), "%s", dirname(tmp)); }
    return dir;
}
uint64_t ucc_sys_host_hash(void)
{
    char buf[256] = ""; uint64_t h;
    FILE *f = fopen("/proc/sys/kernel/random/boot_id", "r");
    size_t n = strlen(ucc_get_host_name());
    snprintf(buf, sizeof(buf), "%s:", ucc_get_host_name());
    if (f) { if (!fgets(buf + n + 1, (int)(sizeof(buf) - n - 1), f)) buf[n + 1] = 0; fclose(f); }
    h = ucc_str_hash_djb2(buf);
    return h ? h : 1;
}

static void abstract_addr(const char *name, struct sockaddr_un *sa, socklen_t *len)
{
    memset(sa, 0, sizeof(*sa));
    sa->sun_family = AF_UNIX;
    snprintf(sa->sun_path + 1, sizeof(sa->sun_path) - 1, "%s", name); /* leading NUL = abstract namespace */
    *len = (socklen_t)(offsetof(struct sockaddr_un, sun_path) + 1 + strlen(name));
}
int ucc_sys_fd_server_open(const char *name)
{
    struct sockaddr_un sa; socklen_t len;
    int s = socket(AF_UNIX, SOCK_STREAM | SOCK_NONBLOCK | SOCK_CLOEXEC, 0);
    if (s < 0) return -1;
    abstract_addr(name, &sa, &len);
    if (bind(s, (struct sockaddr *)&sa, len) != 0 || listen(s, 64) != 0) { close(s); return -1; }
    return s;
}
int ucc_sys_fd_server_serve_once(int lsock, int fd_to_send)
{
    int c = accept4(lsock, NULL, NULL, SOCK_CLOEXEC);
    struct msghdr msg; struct iovec iov; char data = 'F';
    union { struct cmsghdr h; char buf[CMSG_SPACE(sizeof(int))]; } u;
    struct cmsghdr *cm;
    if (c < 0) return (errno == EAGAIN || errno == EWOULDBLOCK) ? 0 : -1;
    memset(&msg, 0, sizeof(msg)); memset(&u, 0, sizeof(u));
    iov.iov_base = &data; iov.iov_len = 1; msg.msg_iov = &iov; msg.msg_iovlen = 1;
    msg.msg_control = u.buf; msg.msg_controllen = sizeof(u.buf);
    cm = CMSG_FIRSTHDR(&msg); cm->cmsg_level = SOL_SOCKET; cm->cmsg_type = SCM_RIGHTS; cm->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(cm), &fd_to_send, sizeof(int));
    if (sendmsg(c, &msg, 0) < 0) { close(c); return -1; }
    close(c);
    return 1;
}
int ucc_sys_fd_recv(const char *name, int timeout_ms)
{
    struct sockaddr_un sa; socklen_t len; double t0 = ucc_get_time();
    struct msghdr msg; struct iovec iov; char data;
    union { struct cmsghdr h; char buf[CMSG_SPACE(sizeof(int))]; } u;
    struct cmsghdr *cm; int fd = -1, s;
    abstract_addr(name, &sa, &len);
    for (;;) {
        s = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
        if (s < 0) return -1;
        if (connect(s, (struct sockaddr *)&sa, len) == 0) break;
        close(s);
        if ((ucc_get_time() - t0) * 1e3 > timeout_ms) return -1;
        usleep(200);
    }
    memset(&msg, 0, sizeof(msg)); memset(&u, 0, sizeof(u));
    iov.iov_base = &data; iov.iov_len = 1; msg.msg_iov = &iov; msg.msg_iovlen = 1;
    msg.msg_control = u.buf; msg.msg_controllen = sizeof(u.buf);
    if (recvmsg(s, &msg, 0) <= 0) { close(s); return -1; }
    cm = CMSG_FIRSTHDR(&msg);
    if (cm && cm->cmsg_level == SOL_SOCKET && cm->cmsg_type == SCM_RIGHTS) memcpy(&fd, CMSG_DATA(cm), sizeof(int));
    close(s);
    return fd;
}
int ucc_sys_pidfd_getfd(int pid, int remote_fd)
{
#if defined(SYS_pidfd_open) && defined(SYS_pidfd_getfd)
    int pfd = (int)syscall(SYS_pidfd_open, pid, 0), fd;
    if (pfd < 0) return -1;
    fd = (int)syscall(SYS_pidfd_getfd, pfd, remote_fd, 0);
    close(pfd);
    return fd;
#else
    (void)pid; (void)remote_fd; return -1;
#endif
}
