/* ucc_perftest: collective micro-benchmark with the reference tool's CLI and metrics
 * (tools/perf/ucc_pt_config.cc:113-470, ucc_pt_benchmark.cc:190-283): element counts with k/M
 * suffixes, -F bus bandwidth columns with the reference formulas, persistent / in-place / triggered
 * modes, executor micro-benchmarks (memcpy, reducedt, reducedt_strided), traffic generators.
 * Differences: no MPI — ranks rendezvous over TCP using the torchrun / srun style environment
 * (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR, MASTER_PORT); for CUDA memory the table also shows the
 * device time of the operation measured with CUDA events (max over ranks). */
#include <ucc/api/ucc.h>
#include <arpa/inet.h>
#include <dlfcn.h>
#include <getopt.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

extern "C" {
#include "components/ec/ucc_ec.h"
#include "components/mc/ucc_mc.h"
#include "core/ucc_dt.h"
#include "utils/ucc_coll_utils.h"
}

#define CHECK(_c) do { ucc_status_t _s = (_c); if (_s < 0) { fprintf(stderr, "%s failed: %s\n", #_c, ucc_status_string(_s)); exit(1); } } while (0)

/* ------------------------------------------------------------------ CUDA runtime through dlopen (reference: ucc_pt_cuda.cc:36) */
struct Cuda {
    void *h = nullptr;
    int (*SetDevice)(int); int (*GetDeviceCount)(int *); int (*Malloc)(void **, size_t); int (*Free)(void *);
    int (*Memset)(void *, int, size_t); int (*StreamCreateWithFlags)(void **, unsigned); int (*StreamSynchronize)(void *);
    int (*EventCreate)(void **); int (*EventRecord)(void *, void *); int (*EventSynchronize)(void *); int (*EventElapsedTime)(float *, void *, void *);
    int (*DeviceSynchronize)(); int (*MallocManaged)(void **, size_t, unsigned);
    bool load() {
        for (const char *n : {"libcudart.so.12", "libcudart.so", "libcudart.so.13"}) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
        if (!h) return false;
#define S(_f) *(void **)&_f = dlsym(h, "cuda" #_f)
        S(SetDevice); S(GetDeviceCount); S(Malloc); S(Free); S(Memset); S(StreamCreateWithFlags); S(StreamSynchronize); S(EventCreate); S(EventRecord);
        S(EventSynchronize); S(EventElapsedTime); S(DeviceSynchronize); S(MallocManaged);
        int n = 0; return GetDeviceCount && GetDeviceCount(&n) == 0 && n > 0;
    }
} cuda;

/* ------------------------------------------------------------------ TCP bootstrap */
struct Boot {
    int rank = 0, size = 1, lrank = 0; std::vector<int> socks; int sock = -1;
    static void xsend(int s, const void *b, size_t n) { const char *p = (const char *)b; while (n) { ssize_t r = send(s, p, n, 0); if (r <= 0) { perror("send"); exit(1); } p += r; n -= (size_t)r; } }
    static void xrecv(int s, void *b, size_t n) { char *p = (char *)b; while (n) { ssize_t r = recv(s, p, n, 0); if (r <= 0) { perror("recv"); exit(1); } p += r; n -= (size_t)r; } }
    void init() {
        const char *r = getenv("RANK"), *w = getenv("WORLD_SIZE"), *l = getenv("LOCAL_RANK");
        if (!r) r = getenv("OMPI_COMM_WORLD_RANK"); if (!w) w = getenv("OMPI_COMM_WORLD_SIZE"); if (!r) r = getenv("SLURM_PROCID"); if (!w) w = getenv("SLURM_NTASKS");
        rank = r ? atoi(r) : 0; size = w ? atoi(w) : 1; lrank = l ? atoi(l) : rank;
        if (size == 1) return;
        const char *addr = getenv("MASTER_ADDR"); const char *port = getenv("UCC_PT_PORT"); int p = port ? atoi(port) : (getenv("MASTER_PORT") ? atoi(getenv("MASTER_PORT")) + 17 : 29517);
        sockaddr_in sa{}; sa.sin_family = AF_INET; sa.sin_port = htons((uint16_t)p); inet_pton(AF_INET, addr ? addr : "127.0.0.1", &sa.sin_addr);
        int one = 1;
        if (rank == 0) {
            int ls = socket(AF_INET, SOCK_STREAM, 0); setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one)); sa.sin_addr.s_addr = INADDR_ANY;
            if (bind(ls, (sockaddr *)&sa, sizeof(sa)) || listen(ls, size)) { perror("bootstrap bind/listen"); exit(1); }
            socks.assign(size, -1);
            for (int i = 1; i < size; i++) { int c = accept(ls, nullptr, nullptr); int pr; xrecv(c, &pr, sizeof(pr)); setsockopt(c, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one)); socks[pr] = c; }
            close(ls);
        } else {
            for (int tries = 0;; tries++) { sock = socket(AF_INET, SOCK_STREAM, 0); if (connect(sock, (sockaddr *)&sa, sizeof(sa)) == 0) break; close(sock); if (tries > 3000) { fprintf(stderr, "bootstrap connect timeout\n"); exit(1); } usleep(10000); }
            setsockopt(sock, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one)); xsend(sock, &rank, sizeof(rank));
        }
    }
    void allgather(const void *src, void *dst, size_t n) {
        if (size == 1) { memcpy(dst, src, n); return; }
        if (rank == 0) { memcpy(dst, src, n); for (int i = 1; i < size; i++) xrecv(socks[i], (char *)dst + (size_t)i * n, n); for (int i = 1; i < size; i++) xsend(socks[i], dst, n * size); }
        else { xsend(sock, src, n); xrecv(sock, dst, n * size); }
    }
    double reduce_d(double v, int op) { std::vector<double> all(size); allgather(&v, all.data(), sizeof(double)); double r = all[0]; for (double x : all) r = op == 0 ? std::min(r, x) : op == 1 ? std::max(r, x) : r; if (op == 2) { r = 0; for (double x : all) r += x; r /= size; } return r; }
    void barrier() { char c = 0; std::vector<char> all(size); allgather(&c, all.data(), 1); }
} boot;

static ucc_status_t oob_allgather(void *s, void *r, size_t n, void *, void **req) { boot.allgather(s, r, n); *req = (void *)1; return UCC_OK; }
static ucc_status_t oob_test(void *) { return UCC_OK; }
static ucc_status_t oob_free(void *) { return UCC_OK; }

/* ------------------------------------------------------------------ config */
struct Config {
    std::string coll = "allreduce", gen = "exp"; ucc_memory_type_t mt = UCC_MEMORY_TYPE_HOST; ucc_datatype_t dt = UCC_DT_FLOAT32; ucc_reduction_op_t op = UCC_OP_SUM;
    size_t min_count = 128, max_count = 128; int n_iter_small = 1000, n_warm_small = 100, n_iter_large = 200, n_warm_large = 20; size_t large_thresh = 64 * 1024;
    int mult = 2, n_bufs = 2; bool inplace = false, persistent = false, triggered = false, full = false; int root = 0, root_shift = 0; int thread_mode = 0;
    unsigned seed = 1; std::string gen_file; int matrix_kind = 0; size_t token_size = 0, num_tokens = 0; int tgt_group = 0; bool iters_set = false, warm_set = false;
} cfg;

static size_t parse_count(const char *s) { char *e; double v = strtod(s, &e); switch (*e) { case 'k': case 'K': v *= 1024; break; case 'm': case 'M': v *= 1024 * 1024; break; case 'g': case 'G': v *= 1024.0 * 1024 * 1024; break; } return (size_t)v; }
static void usage()
{
    printf("Usage: ucc_perftest [options]\n"
           "  -c <name>   collective: allgather allgatherv allreduce alltoall alltoallv barrier bcast gather gatherv reduce reduce_scatter\n"
           "              reduce_scatterv scatter scatterv | memcpy reducedt reducedt_strided (executor micro-benchmarks)\n"
           "  -b <count>  min element count (k/M suffix)     -e <count>  max element count\n"
           "  -d <dt>     datatype (float32, bfloat16, int32 ...)   -o <op>  sum prod min max avg land lor lxor band bor bxor\n"
           "  -m <mtype>  host | cuda | cuda-mng              -n <iter>  iterations   -w <iter>  warm-up iterations\n"
           "  -f <mult>   size multiplication factor (default 2)  -N <n> number of source buffers for reducedt\n"
           "  -i          in-place   -p persistent   -T triggered (stream-ordered) post   -F full print (bus bandwidth)\n"
           "  -M <mode>   thread mode single|multiple   -r <root>  root   -S <shift> root shift per iteration\n"
           "  --gen <g>   exp (default) | file:name=<f.ptini> | matrix:kind=<0-3>:token_size=<b>:num_tokens=<n>:tgt_group_size=<g>\n"
           "  --seed <n>\n");
}
static void parse_gen(const std::string &g)
{
    cfg.gen = g.substr(0, g.find(':'));
    std::stringstream ss(g); std::string tok; std::getline(ss, tok, ':');
    while (std::getline(ss, tok, ':')) {
        size_t eq = tok.find('='); if (eq == std::string::npos) continue; std::string k = tok.substr(0, eq), v = tok.substr(eq + 1);
        if (k == "name") cfg.gen_file = v; else if (k == "kind") cfg.matrix_kind = atoi(v.c_str()); else if (k == "token_size") cfg.token_size = parse_count(v.c_str());
        else if (k == "num_tokens") cfg.num_tokens = parse_count(v.c_str()); else if (k == "tgt_group_size") cfg.tgt_group = atoi(v.c_str());
    }
}
static void parse_args(int argc, char **argv)
{
    static option lo[] = {{"gen", required_argument, nullptr, 1000}, {"seed", required_argument, nullptr, 1001}, {"help", no_argument, nullptr, 'h'}, {nullptr, 0, nullptr, 0}};
    int c;
    while ((c = getopt_long(argc, argv, "c:b:e:d:o:m:n:w:f:N:ipTFM:r:S:h", lo, nullptr)) != -1) {
        switch (c) {
        case 'c': cfg.coll = optarg; break; case 'b': cfg.min_count = parse_count(optarg); break; case 'e': cfg.max_count = parse_count(optarg); break;
        case 'd': cfg.dt = ucc_datatype_from_str(optarg); if (cfg.dt == (ucc_datatype_t)-1) { fprintf(stderr, "bad datatype %s\n", optarg); exit(1); } break;
        case 'o': cfg.op = ucc_reduction_op_from_str(optarg); if (cfg.op == UCC_OP_LAST) { fprintf(stderr, "bad op %s\n", optarg); exit(1); } break;
        case 'm': cfg.mt = ucc_mem_type_from_str(optarg); if (cfg.mt == UCC_MEMORY_TYPE_LAST) { fprintf(stderr, "bad mem type %s\n", optarg); exit(1); } break;
        case 'n': cfg.n_iter_small = cfg.n_iter_large = atoi(optarg); cfg.iters_set = true; break; case 'w': cfg.n_warm_small = cfg.n_warm_large = atoi(optarg); cfg.warm_set = true; break;
        case 'f': cfg.mult = std::max(2, atoi(optarg)); break; case 'N': cfg.n_bufs = atoi(optarg); break; case 'i': cfg.inplace = true; break; case 'p': cfg.persistent = true; break;
        case 'T': cfg.triggered = true; break; case 'F': cfg.full = true; break; case 'M': cfg.thread_mode = !strcmp(optarg, "multiple") ? 2 : 0; break;
        case 'r': cfg.root = atoi(optarg); break; case 'S': cfg.root_shift = atoi(optarg); break; case 1000: parse_gen(optarg); break; case 1001: cfg.seed = (unsigned)atoi(optarg); break;
        default: usage(); exit(c == 'h' ? 0 : 1);
        }
    }
    if (cfg.max_count < cfg.min_count) cfg.max_count = cfg.min_count;
}

/* ------------------------------------------------------------------ buffers */
static void *alloc_buf(size_t bytes)
{
    void *p = nullptr; if (bytes == 0) bytes = 1;
    if (cfg.mt == UCC_MEMORY_TYPE_HOST) { if (posix_memalign(&p, 4096, bytes)) p = nullptr; if (p) memset(p, 0, bytes); }
    else if (cfg.mt == UCC_MEMORY_TYPE_CUDA) { if (cuda.Malloc(&p, bytes)) p = nullptr; else cuda.Memset(p, 0, bytes); }
    else if (cfg.mt == UCC_MEMORY_TYPE_CUDA_MANAGED) { if (cuda.MallocManaged(&p, bytes, 1)) p = nullptr; }
    if (!p) { fprintf(stderr, "failed to allocate %zu bytes\n", bytes); exit(1); }
    return p;
}
static void free_buf(void *p) { if (!p) return; if (cfg.mt == UCC_MEMORY_TYPE_HOST) free(p); else cuda.Free(p); }
static double now_us() { timeval tv; gettimeofday(&tv, nullptr); return tv.tv_sec * 1e6 + tv.tv_usec; }

/* ------------------------------------------------------------------ traffic generators for alltoallv */
struct Pattern { std::vector<uint64_t> scounts, sdispls, rcounts, rdispls; size_t src_total = 0, dst_total = 0; };
/* matrix generator: token routing model. kind 0 uniform, 1 normal around the diagonal group, 2 hot experts (zipf), 3 all to one group */
static std::vector<std::vector<uint64_t>> make_matrix(int N)
{
    std::vector<std::vector<uint64_t>> m(N, std::vector<uint64_t>(N, 0)); std::mt19937 rng(cfg.seed);
    size_t tokens = cfg.num_tokens ? cfg.num_tokens : 1024, ts = cfg.token_size ? cfg.token_size : 1024; int grp = cfg.tgt_group > 0 ? cfg.tgt_group : std::max(1, N / 4);
    for (int r = 0; r < N; r++) {
        std::vector<double> w(N, 1.0);
        if (cfg.matrix_kind == 1) for (int p = 0; p < N; p++) { double d = std::min(std::abs(p - r), N - std::abs(p - r)); w[p] = std::exp(-d * d / (2.0 * grp * grp)); }
        else if (cfg.matrix_kind == 2) for (int p = 0; p < N; p++) w[p] = 1.0 / std::pow(1.0 + p, 1.2);
        else if (cfg.matrix_kind == 3) for (int p = 0; p < N; p++) w[p] = p < grp ? 1.0 : 0.0;
        std::discrete_distribution<int> dist(w.begin(), w.end());
        for (size_t t = 0; t < tokens; t++) m[r][dist(rng)] += ts;
    }
    return m;
}
static std::vector<std::vector<uint64_t>> read_ptini(const std::string &f, int N)
{   /* rows of N byte counts, '#' comments; one matrix per file (reference generator/examples/*.ptini) */
    std::ifstream in(f); std::vector<std::vector<uint64_t>> m; std::string line;
    if (!in) { fprintf(stderr, "cannot open %s\n", f.c_str()); exit(1); }
    while (std::getline(in, line)) { if (line.empty() || line[0] == '#' || line[0] == '[') continue; std::replace(line.begin(), line.end(), ',', ' '); std::stringstream ss(line); std::vector<uint64_t> row; uint64_t v; while (ss >> v) row.push_back(v); if ((int)row.size() == N) m.push_back(row); }
    if ((int)m.size() != N) { fprintf(stderr, "%s: expected a %dx%d matrix\n", f.c_str(), N, N); exit(1); }
    return m;
}
static Pattern make_pattern(int N, int me, size_t dts)
{
    auto m = cfg.gen == "file" ? read_ptini(cfg.gen_file, N) : make_matrix(N);
    Pattern p; p.scounts.resize(N); p.sdispls.resize(N); p.rcounts.resize(N); p.rdispls.resize(N);
    for (int q = 0; q < N; q++) { p.scounts[q] = m[me][q] / dts; p.rcounts[q] = m[q][me] / dts; p.sdispls[q] = p.src_total; p.rdispls[q] = p.dst_total; p.src_total += p.scounts[q]; p.dst_total += p.rcounts[q]; }
    return p;
}

/* ------------------------------------------------------------------ main */
int main(int argc, char **argv)
{
    parse_args(argc, argv);
    boot.init();
    bool is_exec = cfg.coll == "memcpy" || cfg.coll == "reducedt" || cfg.coll == "reducedt_strided";
    bool need_cuda = cfg.mt != UCC_MEMORY_TYPE_HOST;
    if (need_cuda) { if (!cuda.load()) { fprintf(stderr, "CUDA runtime / device not available\n"); return 1; } int n = 0; cuda.GetDeviceCount(&n); cuda.SetDevice(boot.lrank % n); }
    ucc_lib_config_h lcfg; ucc_lib_params_t lp{}; ucc_lib_h lib; ucc_context_config_h ccfg; ucc_context_params_t cp{}; ucc_context_h ctx; ucc_team_params_t tp{}; ucc_team_h team;
    lp.mask = UCC_LIB_PARAM_FIELD_THREAD_MODE; lp.thread_mode = cfg.thread_mode == 2 ? UCC_THREAD_MULTIPLE : UCC_THREAD_SINGLE;
    CHECK(ucc_lib_config_read(nullptr, nullptr, &lcfg)); CHECK(ucc_init(&lp, lcfg, &lib)); ucc_lib_config_release(lcfg);
    CHECK(ucc_context_config_read(lib, nullptr, &ccfg));
    if (boot.size > 1) { cp.mask = UCC_CONTEXT_PARAM_FIELD_OOB; cp.oob.allgather = oob_allgather; cp.oob.req_test = oob_test; cp.oob.req_free = oob_free; cp.oob.n_oob_eps = (uint32_t)boot.size; cp.oob.oob_ep = (uint32_t)boot.rank; }
    CHECK(ucc_context_create(lib, &cp, ccfg, &ctx)); ucc_context_config_release(ccfg);
    tp.mask = UCC_TEAM_PARAM_FIELD_EP | UCC_TEAM_PARAM_FIELD_EP_RANGE | UCC_TEAM_PARAM_FIELD_OOB; tp.ep = (uint64_t)boot.rank; tp.ep_range = UCC_COLLECTIVE_EP_RANGE_CONTIG;
    tp.oob.allgather = oob_allgather; tp.oob.req_test = oob_test; tp.oob.req_free = oob_free; tp.oob.n_oob_eps = (uint32_t)boot.size; tp.oob.oob_ep = (uint32_t)boot.rank;
    CHECK(ucc_team_create_post(&ctx, 1, &tp, &team));
    { ucc_status_t st; while ((st = ucc_team_create_test(team)) == UCC_INPROGRESS) ucc_context_progress(ctx); CHECK(st); }

    /* iterations are separated by a UCC barrier on the team (reference ucc_pt_comm::barrier): much less arrival skew than a trip over the TCP bootstrap */
    auto team_barrier = [&]() {
        ucc_coll_args_t ba{}; ucc_coll_req_h br; ucc_status_t st;
        ba.coll_type = UCC_COLL_TYPE_BARRIER;
        CHECK(ucc_collective_init(&ba, &br, team)); CHECK(ucc_collective_post(br));
        while ((st = ucc_collective_test(br)) > 0) ucc_context_progress(ctx);
        CHECK(st); CHECK(ucc_collective_finalize(br));
    };
    const int N = boot.size, me = boot.rank; const size_t dts = ucc_dt_size(cfg.dt);
    ucc_coll_type_t ct = is_exec ? UCC_COLL_TYPE_LAST : ucc_coll_type_from_str(cfg.coll.c_str());
    if (!is_exec && ct == UCC_COLL_TYPE_LAST) { fprintf(stderr, "unknown collective %s\n", cfg.coll.c_str()); return 1; }
    void *stream = nullptr, *ev0 = nullptr, *ev1 = nullptr; ucc_ee_h ee = nullptr;
    if (need_cuda) { cuda.StreamCreateWithFlags(&stream, 1); cuda.EventCreate(&ev0); cuda.EventCreate(&ev1); }
    if (cfg.triggered) { if (!need_cuda) { fprintf(stderr, "-T needs -m cuda\n"); return 1; } ucc_ee_params_t ep{UCC_EE_CUDA_STREAM, stream, sizeof(void *)}; CHECK(ucc_ee_create(team, &ep, &ee)); }

    if (me == 0) {
        printf("\nCollective:\t\t%s\nMemory type:\t\t%s\nDatatype:\t\t%s\nReduction:\t\t%s\nInplace:\t\t%d\nPersistent:\t\t%d\nTriggered:\t\t%d\nRanks:\t\t\t%d\n\n", cfg.coll.c_str(), ucc_mem_type_str(cfg.mt), ucc_datatype_str(cfg.dt),
               ucc_reduction_op_str(cfg.op), cfg.inplace, cfg.persistent, cfg.triggered, N);
        printf("%12s %14s %12s %12s %12s", "Count", "Size", "avg(us)", "min(us)", "max(us)");
        if (need_cuda && !is_exec) printf(" %14s", "dev max(us)");
        if (cfg.full) printf(" %12s %12s %12s", "avgBW(GB/s)", "maxBW(GB/s)", "minBW(GB/s)");
        if (cfg.full && need_cuda && !is_exec) printf(" %12s", "devBW(GB/s)");
        printf("\n");
    }
    bool pattern_mode = (cfg.gen != "exp") && (ct == UCC_COLL_TYPE_ALLTOALLV);
    for (size_t count = cfg.min_count; count <= cfg.max_count; count *= (size_t)cfg.mult) {
        size_t bytes = count * dts; bool large = bytes >= cfg.large_thresh;
        int iters = cfg.iters_set ? cfg.n_iter_small : (large ? cfg.n_iter_large : cfg.n_iter_small), warm = cfg.warm_set ? cfg.n_warm_small : (large ? cfg.n_warm_large : cfg.n_warm_small);
        ucc_coll_args_t a{}; a.coll_type = ct; a.op = cfg.op; a.root = (uint64_t)cfg.root; a.mask = 0; a.flags = 0;
        if (cfg.inplace) { a.mask |= UCC_COLL_ARGS_FIELD_FLAGS; a.flags |= UCC_COLL_ARGS_FLAG_IN_PLACE; }
        if (cfg.persistent) { a.mask |= UCC_COLL_ARGS_FIELD_FLAGS; a.flags |= UCC_COLL_ARGS_FLAG_PERSISTENT; }
        void *src = nullptr, *dst = nullptr; Pattern pat; std::vector<uint64_t> cnts(N, count), dsp(N);
        for (int i = 0; i < N; i++) dsp[i] = (uint64_t)i * count;
        double bw_factor = 1.0; size_t S = bytes; /* S and factor follow the reference's per-collective get_bw() */
        if (is_exec) {
            src = alloc_buf(bytes * (size_t)std::max(2, cfg.n_bufs)); dst = alloc_buf(bytes);
        } else switch (ct) {
        case UCC_COLL_TYPE_ALLREDUCE: src = alloc_buf(bytes); dst = alloc_buf(bytes); a.src.info = {src, count, cfg.dt, cfg.mt}; a.dst.info = {dst, count, cfg.dt, cfg.mt}; bw_factor = 2.0 * (N - 1) / N; break;
        case UCC_COLL_TYPE_ALLGATHER: src = alloc_buf(bytes); dst = alloc_buf(bytes * N); a.src.info = {src, count, cfg.dt, cfg.mt}; a.dst.info = {dst, count * N, cfg.dt, cfg.mt}; S = bytes * N; bw_factor = (double)(N - 1) / N; break;
        case UCC_COLL_TYPE_ALLGATHERV: src = alloc_buf(bytes); dst = alloc_buf(bytes * N); a.src.info = {src, count, cfg.dt, cfg.mt}; a.dst.info_v = {dst, (ucc_count_t *)cnts.data(), (ucc_aint_t *)dsp.data(), cfg.dt, cfg.mt};
            a.mask |= UCC_COLL_ARGS_FIELD_FLAGS; a.flags |= UCC_COLL_ARGS_FLAG_COUNT_64BIT | UCC_COLL_ARGS_FLAG_DISPLACEMENTS_64BIT; S = bytes * N; bw_factor = (double)(N - 1) / N; break;
        case UCC_COLL_TYPE_ALLTOALL: src = alloc_buf(bytes * N); dst = alloc_buf(bytes * N); a.src.info = {src, count * N, cfg.dt, cfg.mt}; a.dst.info = {dst, count * N, cfg.dt, cfg.mt}; S = bytes * N; bw_factor = (double)(N - 1) / N; break;
        case UCC_COLL_TYPE_ALLTOALLV:
            if (pattern_mode) pat = make_pattern(N, me, dts); else { pat.scounts = cnts; pat.rcounts = cnts; pat.sdispls = dsp; pat.rdispls = dsp; pat.src_total = pat.dst_total = count * N; }
            src = alloc_buf(pat.src_total * dts); dst = alloc_buf(pat.dst_total * dts);
            a.src.info_v = {src, (ucc_count_t *)pat.scounts.data(), (ucc_aint_t *)pat.sdispls.data(), cfg.dt, cfg.mt}; a.dst.info_v = {dst, (ucc_count_t *)pat.rcounts.data(), (ucc_aint_t *)pat.rdispls.data(), cfg.dt, cfg.mt};
            a.mask |= UCC_COLL_ARGS_FIELD_FLAGS; a.flags |= UCC_COLL_ARGS_FLAG_COUNT_64BIT | UCC_COLL_ARGS_FLAG_DISPLACEMENTS_64BIT; S = pat.src_total * dts; bw_factor = (double)(N - 1) / N; break;
        case UCC_COLL_TYPE_BCAST: src = alloc_buf(bytes); a.src.info = {src, count, cfg.dt, cfg.mt}; break;
        case UCC_COLL_TYPE_REDUCE: src = alloc_buf(bytes); dst = alloc_buf(bytes); a.src.info = {src, count, cfg.dt, cfg.mt}; a.dst.info = {dst, count, cfg.dt, cfg.mt}; break;
        case UCC_COLL_TYPE_REDUCE_SCATTER: src = alloc_buf(bytes * N); dst = alloc_buf(cfg.inplace ? bytes * N : bytes); a.src.info = {src, count * N, cfg.dt, cfg.mt}; a.dst.info = {dst, cfg.inplace ? count * N : count, cfg.dt, cfg.mt}; S = bytes * N; bw_factor = (double)(N - 1) / N; break;
        case UCC_COLL_TYPE_REDUCE_SCATTERV: src = alloc_buf(bytes * N); dst = alloc_buf(cfg.inplace ? bytes * N : bytes); a.src.info = {src, count * N, cfg.dt, cfg.mt}; a.dst.info_v = {dst, (ucc_count_t *)cnts.data(), (ucc_aint_t *)dsp.data(), cfg.dt, cfg.mt};
            a.mask |= UCC_COLL_ARGS_FIELD_FLAGS; a.flags |= UCC_COLL_ARGS_FLAG_COUNT_64BIT | UCC_COLL_ARGS_FLAG_DISPLACEMENTS_64BIT; S = bytes * N; bw_factor = (double)(N - 1) / N; break;
        case UCC_COLL_TYPE_GATHER: src = alloc_buf(bytes); dst = alloc_buf(bytes * N); a.src.info = {src, count, cfg.dt, cfg.mt}; a.dst.info = {dst, count * N, cfg.dt, cfg.mt}; S = bytes * N; break;
        case UCC_COLL_TYPE_GATHERV: src = alloc_buf(bytes); dst = alloc_buf(bytes * N); a.src.info = {src, count, cfg.dt, cfg.mt}; a.dst.info_v = {dst, (ucc_count_t *)cnts.data(), (ucc_aint_t *)dsp.data(), cfg.dt, cfg.mt};
            a.mask |= UCC_COLL_ARGS_FIELD_FLAGS; a.flags |= UCC_COLL_ARGS_FLAG_COUNT_64BIT | UCC_COLL_ARGS_FLAG_DISPLACEMENTS_64BIT; S = bytes * N; break;
        case UCC_COLL_TYPE_SCATTER: src = alloc_buf(bytes * N); dst = alloc_buf(bytes); a.src.info = {src, count * N, cfg.dt, cfg.mt}; a.dst.info = {dst, count, cfg.dt, cfg.mt}; S = bytes * N; break;
        case UCC_COLL_TYPE_SCATTERV: src = alloc_buf(bytes * N); dst = alloc_buf(bytes); a.src.info_v = {src, (ucc_count_t *)cnts.data(), (ucc_aint_t *)dsp.data(), cfg.dt, cfg.mt}; a.dst.info = {dst, count, cfg.dt, cfg.mt};
            a.mask |= UCC_COLL_ARGS_FIELD_FLAGS; a.flags |= UCC_COLL_ARGS_FLAG_COUNT_64BIT | UCC_COLL_ARGS_FLAG_DISPLACEMENTS_64BIT; S = bytes * N; break;
        default: break; /* barrier fanin fanout */
        }
        ucc_coll_req_h req = nullptr; ucc_ee_executor_t *exec = nullptr;
        if (is_exec) { ucc_ee_executor_params_t ep{UCC_EE_EXECUTOR_PARAM_FIELD_TYPE, need_cuda ? UCC_EE_CUDA_STREAM : UCC_EE_CPU_THREAD, 0}; CHECK(ucc_ee_executor_init(&ep, &exec)); CHECK(ucc_ee_executor_start(exec, nullptr)); }
        if (cfg.persistent && !is_exec) CHECK(ucc_collective_init(&a, &req, team));
        double t_sum = 0, dev_sum = 0;
        boot.barrier(); if (!is_exec) team_barrier();
        for (int it = 0; it < warm + iters; it++) {
            if (cfg.root_shift && !is_exec) a.root = (uint64_t)((cfg.root + it * cfg.root_shift) % N);
            if (need_cuda) cuda.DeviceSynchronize();
            double t0 = now_us();
            if (is_exec) {
                ucc_ee_executor_task_args_t ta{}; ucc_ee_executor_task_t *tk;
                if (cfg.coll == "memcpy") { ta.task_type = UCC_EE_EXECUTOR_TASK_COPY; ta.copy.dst = dst; ta.copy.src = src; ta.copy.len = bytes; }
                else if (cfg.coll == "reducedt") { ta.task_type = UCC_EE_EXECUTOR_TASK_REDUCE; ta.reduce.dst = dst; ta.reduce.n_srcs = (uint16_t)std::min(9, std::max(2, cfg.n_bufs)); for (int k = 0; k < ta.reduce.n_srcs; k++) ta.reduce.srcs[k] = (char *)src + (size_t)k * bytes; ta.reduce.count = count; ta.reduce.dt = cfg.dt; ta.reduce.op = cfg.op == UCC_OP_AVG ? UCC_OP_SUM : cfg.op; }
                else { ta.task_type = UCC_EE_EXECUTOR_TASK_REDUCE_STRIDED; ta.reduce_strided.dst = dst; ta.reduce_strided.src1 = src; ta.reduce_strided.src2 = (char *)src + bytes; ta.reduce_strided.stride = bytes; ta.reduce_strided.n_src2 = (uint16_t)(std::max(2, cfg.n_bufs) - 1); ta.reduce_strided.count = count; ta.reduce_strided.dt = cfg.dt; ta.reduce_strided.op = cfg.op == UCC_OP_AVG ? UCC_OP_SUM : cfg.op; }
                CHECK(ucc_ee_executor_task_post(exec, &ta, &tk)); ucc_status_t st; while ((st = ucc_ee_executor_task_test(tk)) == UCC_INPROGRESS) {} CHECK(st); ucc_ee_executor_task_finalize(tk);
            } else {
                if (!cfg.persistent) CHECK(ucc_collective_init(&a, &req, team));
                if (need_cuda) cuda.EventRecord(ev0, stream);
                if (cfg.triggered) { ucc_ev_t ev{UCC_EVENT_COMPUTE_COMPLETE, nullptr, 0, req}, *post_ev; CHECK(ucc_collective_triggered_post(ee, &ev));
                    while (ucc_ee_get_event(ee, &post_ev) != UCC_OK) ucc_context_progress(ctx); ucc_ee_ack_event(ee, post_ev); }
                else CHECK(ucc_collective_post(req));
                ucc_status_t st; while ((st = ucc_collective_test(req)) > 0) ucc_context_progress(ctx); CHECK(st);
                if (need_cuda && cfg.triggered) { cuda.EventRecord(ev1, stream); cuda.EventSynchronize(ev1); float ms = 0; cuda.EventElapsedTime(&ms, ev0, ev1); if (it >= warm) dev_sum += ms * 1e3; }
                if (!cfg.persistent) CHECK(ucc_collective_finalize(req));
            }
            double t1 = now_us();
            if (it >= warm) t_sum += t1 - t0;
            if (!is_exec) team_barrier();
        }
        if (cfg.persistent && !is_exec) CHECK(ucc_collective_finalize(req));
        if (exec) { ucc_ee_executor_stop(exec); ucc_ee_executor_finalize(exec); }
        double t = t_sum / iters, tmin = boot.reduce_d(t, 0), tmax = boot.reduce_d(t, 1), tavg = boot.reduce_d(t, 2), dmax = boot.reduce_d(dev_sum / iters, 1);
        if (is_exec) { int nb = cfg.coll == "memcpy" ? 2 : std::max(2, cfg.n_bufs) + 1; S = bytes * (size_t)nb; bw_factor = 1.0; }
        if (me == 0) {
            printf("%12zu %14zu %12.2f %12.2f %12.2f", count, bytes, tavg, tmin, tmax);
            if (need_cuda && !is_exec) { if (cfg.triggered) printf(" %14.2f", dmax); else printf(" %14s", "-"); }
            if (cfg.full) printf(" %12.2f %12.2f %12.2f", S / tavg / 1e3 * bw_factor, S / tmin / 1e3 * bw_factor, S / tmax / 1e3 * bw_factor);
            if (cfg.full && need_cuda && !is_exec) { if (cfg.triggered && dmax > 0) printf(" %12.2f", S / dmax / 1e3 * bw_factor); else printf(" %12s", "-"); }
            printf("\n"); fflush(stdout);
        }
        free_buf(src); free_buf(dst);
        if (count == 0) break;
    }
    if (ee) ucc_ee_destroy(ee);
    { ucc_status_t st; while ((st = ucc_team_destroy(team)) == UCC_INPROGRESS) ucc_context_progress(ctx); }
    ucc_context_destroy(ctx); ucc_finalize(lib);
    return 0;
}
