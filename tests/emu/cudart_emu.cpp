// Minimal CUDA runtime stand-in for running the HOST side of tl/nvl (team creation, launch queue, zero-copy board, score
// selection, asymmetric staging in the core...) without a GPU: "device" memory is host memory with a registry behind
// cudaPointerGetAttributes, a stream is a worker thread draining a queue of closures, an event completes when the closure
// recorded behind the preceding work runs.  Kernels are the host-emulated ones (nvl_emu_launch.cpp) enqueued on the same queues.
// Only what libucc_tl_nvl / libucc_mc_cuda call is implemented.
#include <cuda_runtime_api.h>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
#include <sched.h>

namespace {
struct EmuStream {
    std::mutex m; std::condition_variable cv; std::deque<std::function<void()>> q; bool stop = false; bool busy = false; std::thread th;
    EmuStream() { th = std::thread([this] { run(); }); }
    void run()
    {
        for (;;) {
            std::function<void()> f;
            { std::unique_lock<std::mutex> l(m); cv.wait(l, [this] { return stop || !q.empty(); }); if (q.empty()) return; f = std::move(q.front()); q.pop_front(); busy = true; }
            f();
            { std::lock_guard<std::mutex> l(m); busy = false; }
            cv.notify_all();
        }
    }
    void push(std::function<void()> f) { { std::lock_guard<std::mutex> l(m); q.push_back(std::move(f)); } cv.notify_all(); }
    void sync() { std::unique_lock<std::mutex> l(m); cv.wait(l, [this] { return q.empty() && !busy; }); }
    ~EmuStream() { { std::lock_guard<std::mutex> l(m); stop = true; } cv.notify_all(); if (th.joinable()) th.join(); }
};
struct EmuEvent { std::atomic<uint64_t> recorded{0}, completed{0}; };

std::mutex g_lock;
std::map<uintptr_t, std::pair<size_t, int>> g_allocs;   // base -> (size, cudaMemoryType)
std::vector<EmuStream *> g_streams;
EmuStream *g_default;
thread_local cudaError_t t_last = cudaSuccess;

EmuStream *S(cudaStream_t s)
{
    if (s) return reinterpret_cast<EmuStream *>(s);
    std::lock_guard<std::mutex> l(g_lock);
    if (!g_default) { g_default = new EmuStream(); g_streams.push_back(g_default); }
    return g_default;
}
cudaError_t fail(cudaError_t e) { t_last = e; return e; }
void *reg_alloc(size_t n, int type)
{
    void *p = nullptr;
    if (posix_memalign(&p, 512, n ? n : 1)) return nullptr;
    std::lock_guard<std::mutex> l(g_lock);
    g_allocs[(uintptr_t)p] = {n, type};
    return p;
}
}  // namespace

// used by nvl_emu_launch.cpp: run `f` behind the work already queued on `s`
extern "C" void emu_stream_enqueue(cudaStream_t s, std::function<void()> *f) { S(s)->push(*f); }

extern "C" {
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { cudaError_t e = t_last; t_last = cudaSuccess; return e; }
cudaError_t cudaPeekAtLastError(void) { return t_last; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorNotReady ? "not ready" : "emulated CUDA error"); }
static int emu_stream_value_stub(void *, unsigned long long, unsigned, unsigned) { return 999; /* CUDA_ERROR_UNKNOWN: never called, tl/nvl only checks that the entry points exist before it enables the zero-copy board */ }
cudaError_t cudaGetDriverEntryPoint(const char *name, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *qr)
{
    if (!strcmp(name, "cuStreamWaitValue32") || !strcmp(name, "cuStreamWriteValue32")) { *fn = (void *)emu_stream_value_stub; if (qr) *qr = cudaDriverEntryPointSuccess; return cudaSuccess; }
    *fn = nullptr; if (qr) *qr = cudaDriverEntryPointSymbolNotFound; return fail(cudaErrorNotSupported);   /* no VMM / multicast: heaps are shared by pointer */
}
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp *p, int)
{ memset(p, 0, sizeof(*p)); snprintf(p->name, sizeof(p->name), "host-emulated device"); p->multiProcessorCount = 2; p->major = 10; p->minor = 0; p->totalGlobalMem = 1ull << 34; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 0; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 1; return cudaSuccess; }

cudaError_t cudaMalloc(void **p, size_t n) { *p = reg_alloc(n, cudaMemoryTypeDevice); return *p ? cudaSuccess : fail(cudaErrorMemoryAllocation); }
cudaError_t cudaMallocManaged(void **p, size_t n, unsigned) { *p = reg_alloc(n, cudaMemoryTypeManaged); return *p ? cudaSuccess : fail(cudaErrorMemoryAllocation); }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { *p = reg_alloc(n, cudaMemoryTypeHost); return *p ? cudaSuccess : fail(cudaErrorMemoryAllocation); }
cudaError_t cudaMallocHost(void **p, size_t n) { return cudaHostAlloc(p, n, 0); }
cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned) { *d = h; return cudaSuccess; }
cudaError_t cudaFree(void *p) { if (!p) return cudaSuccess; { std::lock_guard<std::mutex> l(g_lock); g_allocs.erase((uintptr_t)p); } free(p); return cudaSuccess; }
cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *ptr)
{
    memset(a, 0, sizeof(*a));
    a->type = cudaMemoryTypeUnregistered; a->hostPointer = const_cast<void *>(ptr);
    std::lock_guard<std::mutex> l(g_lock);
    auto it = g_allocs.upper_bound((uintptr_t)ptr);
    if (it != g_allocs.begin()) { --it; if ((uintptr_t)ptr < it->first + (it->second.first ? it->second.first : 1)) { a->type = (cudaMemoryType)it->second.second; a->device = 0; a->devicePointer = const_cast<void *>(ptr); } }
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t st) { S(st)->push([=] { memmove(d, s, n); }); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st) { S(st)->push([=] { memset(d, v, n); }); return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void)
{ std::vector<EmuStream *> all; { std::lock_guard<std::mutex> l(g_lock); all = g_streams; } for (auto *s : all) s->sync(); return cudaSuccess; }

/* leak checks of the tests: objects currently alive (what = 0 streams, 1 events, 2 device / pinned allocations) */
long emu_live_objects(int what);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { auto *e = new EmuStream(); { std::lock_guard<std::mutex> l(g_lock); g_streams.push_back(e); } *s = reinterpret_cast<cudaStream_t>(e); return cudaSuccess; }
cudaError_t cudaStreamCreate(cudaStream_t *s) { return cudaStreamCreateWithFlags(s, 0); }
cudaError_t cudaStreamDestroy(cudaStream_t s)
{ auto *e = reinterpret_cast<EmuStream *>(s); if (!e) return cudaSuccess; e->sync(); { std::lock_guard<std::mutex> l(g_lock); for (auto &x : g_streams) if (x == e) { x = g_streams.back(); g_streams.pop_back(); break; } } delete e; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t s) { S(s)->sync(); return cudaSuccess; }
cudaError_t cudaStreamQuery(cudaStream_t s) { auto *e = S(s); std::lock_guard<std::mutex> l(e->m); return e->q.empty() && !e->busy ? cudaSuccess : cudaErrorNotReady; }
cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus *st) { *st = cudaStreamCaptureStatusNone; return cudaSuccess; }

static std::atomic<long> g_live_events{0};
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = reinterpret_cast<cudaEvent_t>(new EmuEvent()); g_live_events++; return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { return cudaEventCreateWithFlags(e, 0); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete reinterpret_cast<EmuEvent *>(e); g_live_events--; return cudaSuccess; }   // (callers destroy only completed events)
long emu_live_objects(int what)
{ std::lock_guard<std::mutex> l(g_lock); return what == 0 ? (long)g_streams.size() : what == 1 ? g_live_events.load() : (long)g_allocs.size(); }
cudaError_t cudaEventRecord(cudaEvent_t ev, cudaStream_t s)
{ auto *e = reinterpret_cast<EmuEvent *>(ev); uint64_t seq = ++e->recorded; S(s)->push([e, seq] { uint64_t c = e->completed.load(); while (c < seq && !e->completed.compare_exchange_weak(c, seq)) {} }); return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t ev) { auto *e = reinterpret_cast<EmuEvent *>(ev); return e->completed.load() >= e->recorded.load() ? cudaSuccess : fail(cudaErrorNotReady); }
cudaError_t cudaEventSynchronize(cudaEvent_t ev) { auto *e = reinterpret_cast<EmuEvent *>(ev); while (e->completed.load() < e->recorded.load()) sched_yield(); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t ev, unsigned)
{ auto *e = reinterpret_cast<EmuEvent *>(ev); uint64_t seq = e->recorded.load(); S(s)->push([e, seq] { while (e->completed.load() < seq) sched_yield(); }); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return fail(cudaErrorNotSupported); }
cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return fail(cudaErrorNotSupported); }
cudaError_t cudaIpcCloseMemHandle(void *) { return fail(cudaErrorNotSupported); }
}
