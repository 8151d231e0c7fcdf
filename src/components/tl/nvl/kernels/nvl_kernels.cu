/* tl/nvl collective kernels for sm_100a.  See nvl_kernels.h for the contract.
 *
 * Reference parity (what each kernel replaces):
 *   allreduce_oneshot  - no reference equivalent below NVLS; tl_cuda has no P2P allreduce (tl/cuda/tl_cuda.h:40-44)
 *   reduce_staged      - tl_cuda reduce_scatterv_linear (copy to peer scratch via executor + separate
 *                        REDUCE_STRIDED kernel + CPU polled step counters, reduce_scatterv_linear.c:229-328)
 *                        and the NVLS kernels K1/K2/K4 with their two staging cudaMemcpyAsync
 *                        (kernels/allreduce_kernel.cu:21-141, allreduce_nvls.c:81-116) fused into one launch
 *   exchange           - allgatherv_linear / alltoallv_ce / bcast_linear state machines (host driven
 *                        copy-engine or executor copies) as one pull kernel with device-side flags
 */
#include "nvl_device.cuh"

/* ------------------------------------------------------------------ */
/* small helpers                                                       */
/* ------------------------------------------------------------------ */
static __device__ __forceinline__ size_t dmin(size_t a, size_t b) { return a < b ? a : b; }
static __device__ __forceinline__ char *data_of(const nvl_team_dev_t &t, int p) { return t.heap[p] + NVL_DATA_OFFSET; }

__global__ void nvl_ctrl_init_kernel(nvl_ctrl_t *c)
{
    uint32_t *w = reinterpret_cast<uint32_t *>(c);
    for (size_t i = threadIdx.x + (size_t)blockIdx.x * blockDim.x; i < sizeof(nvl_ctrl_t) / 4; i += (size_t)gridDim.x * blockDim.x) w[i] = 0;
}

__global__ void nvl_barrier_kernel(nvl_team_dev_t t)
{
    BlockSync bs; bs.init(t);
    bs.barrier(t, 1);
    bs.finish(1);
}

/* ------------------------------------------------------------------ */
/* one-shot push allreduce (latency path)                              */
/*   every rank stores its vector into a private slot of every peer's  */
/*   heap, raises a flag, then reduces the N vectors it received.      */
/*   Two slot sets alternate so no trailing barrier is needed.         */
/* ------------------------------------------------------------------ */
template <typename T, int OP>
static __device__ __forceinline__ void oneshot_body(const nvl_red_args_t &a, nvl_ctrl_t *mine, uint32_t seq)
{
    constexpr int E = 16 / sizeof(T);
    const nvl_team_dev_t &t = a.team;
    const int N = t.size, me = t.rank;
    const T *src = static_cast<const T *>(a.src);
    T *dst = static_cast<T *>(a.dst);
    const size_t n = a.count, nvec = (n + E - 1) / E;
    const size_t per = (nvec + gridDim.x - 1) / gridDim.x, v0 = dmin((size_t)blockIdx.x * per, nvec), v1 = dmin(v0 + per, nvec);
    const bool sal = ((uintptr_t)src & 15) == 0, dal = ((uintptr_t)dst & 15) == 0;
    const size_t slot = (size_t)((seq & 1) * NVL_MAX_PEERS) * NVL_LL_MAX;
    const float inv_n = 1.0f / (float)N;

    for (size_t v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        uint4 x = load_src_vec<T>(src, v * E, n, sal);
        for (int i = 1; i < N; i++) { /* start at my right neighbour so the N senders do not hit one peer at once */
            int p = (me + i) % N;
            st_v4(t.heap[p] + NVL_LL_OFFSET + slot + (size_t)me * NVL_LL_MAX + v * 16, x);
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < N && (int)threadIdx.x != me) {
        fence_sys();
        st_relaxed_sys_u32(&reinterpret_cast<nvl_ctrl_t *>(t.heap[threadIdx.x])->ll_flags[blockIdx.x][me], seq);
        BlockSync bs; bs.mine = mine;
        bs.spin(t, &mine->ll_flags[blockIdx.x][threadIdx.x], seq);
    }
    __syncthreads();
    const char *mybase = t.heap[me] + NVL_LL_OFFSET + slot;
    for (size_t v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        VecAcc<T, OP> acc;
        acc.set(load_src_vec<T>(src, v * E, n, sal));
#pragma unroll 4
        for (int i = 1; i < N; i++) { int p = (me + i) % N; acc.add(ld_peer_v4(mybase + (size_t)p * NVL_LL_MAX + v * 16)); }
        store_dst_vec<T>(dst, v * E, n, dal, acc.get(inv_n));
    }
}

template <typename T> __global__ void __launch_bounds__(1024) nvl_allreduce_oneshot_kernel(nvl_red_args_t a)
{
    nvl_ctrl_t *mine = reinterpret_cast<nvl_ctrl_t *>(a.team.heap[a.team.rank]);
    uint32_t seq = ld_volatile_u32(&mine->ll_seq[blockIdx.x]) + 1;
#define CALL_ONESHOT(_T, _OP) oneshot_body<_T, _OP>(a, mine, seq)
    NVL_DISPATCH_OP(T, a.op, CALL_ONESHOT);
    __syncthreads();
    if (threadIdx.x == 0) mine->ll_seq[blockIdx.x] = seq;
}

/* ------------------------------------------------------------------ */
/* staged two-shot reduction: allreduce / reduce_scatter(v) / reduce   */
/* ------------------------------------------------------------------ */
template <typename T> static __device__ __forceinline__ uint4 mc_ld_reduce(const void *mc) { return make_uint4(0, 0, 0, 0); }
template <> __device__ __forceinline__ uint4 mc_ld_reduce<float>(const void *mc) { return mc_ld_reduce_f32(mc); }
template <> __device__ __forceinline__ uint4 mc_ld_reduce<__half>(const void *mc) { return mc_ld_reduce_f16(mc); }
template <> __device__ __forceinline__ uint4 mc_ld_reduce<__nv_bfloat16>(const void *mc) { return mc_ld_reduce_bf16(mc); }

struct SlicePlan {
    size_t off[NVL_MAX_PEERS], cnt[NVL_MAX_PEERS]; /* elements of the user vector owned by slice s */
    size_t slice_max, cap_e;                       /* longest slice; elements of one slice per round */
    int    rounds;
};

template <typename T> static __device__ __forceinline__ void make_plan(const nvl_red_args_t &a, SlicePlan &pl)
{
    constexpr int E = 16 / sizeof(T);
    const int N = a.team.size;
    pl.slice_max = 0;
    if (a.kind == NVL_RED_REDUCE_SCATTER) {
        for (int s = 0; s < N; s++) { pl.off[s] = a.rs_offset[s]; pl.cnt[s] = a.rs_count[s]; if (pl.cnt[s] > pl.slice_max) pl.slice_max = pl.cnt[s]; }
    } else {
        size_t se = (a.count + N - 1) / N; se = (se + E - 1) / E * E;
        for (int s = 0; s < N; s++) { pl.off[s] = dmin((size_t)s * se, a.count); pl.cnt[s] = dmin(se, a.count - pl.off[s]); }
        pl.slice_max = se;
    }
    pl.cap_e = (a.team.data_size / N / 16) * E;
    pl.rounds = (int)((pl.slice_max + pl.cap_e - 1) / pl.cap_e);
}

template <typename T, int OP>
static __device__ __forceinline__ void staged_body(const nvl_red_args_t &a, BlockSync &bs, const SlicePlan &pl)
{
    constexpr int E = 16 / sizeof(T);
    const nvl_team_dev_t &t = a.team;
    const int N = t.size, me = t.rank, nb = gridDim.x, b = blockIdx.x;
    const T *src = static_cast<const T *>(a.src);
    T *dst = static_cast<T *>(a.dst);
    const size_t cap_bytes = pl.cap_e * sizeof(T);
    const float inv_n = 1.0f / (float)N;
    const bool nvls = a.use_nvls && t.mc_heap != nullptr;
    char *mydata = data_of(t, me);
    char *pd[NVL_MAX_PEERS]; /* pd[i] = data region of my i-th right neighbour (i = 0: myself) */
#pragma unroll
    for (int i = 0; i < NVL_MAX_PEERS; i++) pd[i] = i < N ? data_of(t, (me + i) % N) : nullptr;

    for (int k = 0; k < pl.rounds; k++) {
        const size_t e0 = (size_t)k * pl.cap_e;                       /* first element of every slice in this round */
        const size_t rmax = dmin(pl.cap_e, pl.slice_max - e0);        /* longest slice part in this round */
        const size_t nvmax = (rmax + E - 1) / E, per = (nvmax + nb - 1) / nb;
        const size_t j0 = dmin((size_t)b * per, nvmax), j1 = dmin(j0 + per, nvmax); /* my vector range inside every slice */

        /* phase A: stage vector range [j0,j1) of EVERY slice into my heap */
        for (int s = 0; s < N; s++) {
            const size_t rc = pl.cnt[s] > e0 ? dmin(pl.cap_e, pl.cnt[s] - e0) : 0;
            const size_t jend = dmin(j1, (rc + E - 1) / E);
            const T *sb = src + pl.off[s] + e0;
            const bool sal = ((uintptr_t)sb & 15) == 0;
            char *hb = mydata + (size_t)s * cap_bytes;
            for (size_t j = j0 + threadIdx.x; j < jend; j += blockDim.x) st_v4(hb + j * 16, load_src_vec<T>(sb, j * E, rc, sal));
        }
        bs.barrier(t, 2 * k + 1);

        /* phase B: reduce my slice from every member's staging area (over NVLink or in the switch) and publish it */
        {
            const size_t rc = pl.cnt[me] > e0 ? dmin(pl.cap_e, pl.cnt[me] - e0) : 0;
            const size_t nv = (rc + E - 1) / E, jend = dmin(j1, nv);
            T *db = dst + (a.kind == NVL_RED_REDUCE_SCATTER ? 0 : pl.off[me]) + e0;
            const bool dal = ((uintptr_t)db & 15) == 0;
            for (size_t j = j0 + threadIdx.x; j < jend; j += blockDim.x) {
                const size_t o = (size_t)me * cap_bytes + j * 16;
                uint4 r;
                if (nvls) {
                    r = mc_ld_reduce<T>(t.mc_heap + NVL_DATA_OFFSET + o);
                    if (OP == NVL_OP_AVG) { VecAcc<T, OP> acc; acc.set(r); r = acc.get(inv_n); }
                } else {
                    uint4 x[NVL_MAX_PEERS];
#pragma unroll
                    for (int i = 0; i < NVL_MAX_PEERS; i++) if (i < N) x[i] = ld_peer_v4(pd[i] + o);
                    VecAcc<T, OP> acc; acc.set(x[0]);
#pragma unroll
                    for (int i = 1; i < NVL_MAX_PEERS; i++) if (i < N) acc.add(x[i]);
                    r = acc.get(inv_n);
                }
                if (a.kind == NVL_RED_ALLREDUCE) {
                    if (nvls) mc_st_v4(t.mc_heap + NVL_DATA_OFFSET + o, r);
                    else {
#pragma unroll
                        for (int i = 0; i < NVL_MAX_PEERS; i++) if (i < N) st_v4(pd[i] + o, r);
                    }
                } else if (a.kind == NVL_RED_REDUCE) st_v4(data_of(t, a.root) + o, r);
                else store_dst_vec<T>(db, j * E, rc, dal, r);
            }
        }
        bs.barrier(t, 2 * k + 2);

        /* phase C: my heap now holds vector range [j0,j1) of every reduced slice */
        if (a.kind == NVL_RED_ALLREDUCE || (a.kind == NVL_RED_REDUCE && me == a.root)) {
            for (int s = 0; s < N; s++) {
                const size_t rc = pl.cnt[s] > e0 ? dmin(pl.cap_e, pl.cnt[s] - e0) : 0;
                const size_t jend = dmin(j1, (rc + E - 1) / E);
                T *db = dst + pl.off[s] + e0;
                const bool dal = ((uintptr_t)db & 15) == 0;
                const char *hb = mydata + (size_t)s * cap_bytes;
                for (size_t j = j0 + threadIdx.x; j < jend; j += blockDim.x) store_dst_vec<T>(db, j * E, rc, dal, ld_peer_v4(hb + j * 16));
            }
        }
    }
}

template <typename T> __global__ void __launch_bounds__(1024) nvl_reduce_staged_kernel(nvl_red_args_t a)
{
    BlockSync bs; bs.init(a.team);
    SlicePlan pl; make_plan<T>(a, pl);
#define CALL_STAGED(_T, _OP) staged_body<_T, _OP>(a, bs, pl)
    NVL_DISPATCH_OP(T, a.op, CALL_STAGED);
    bs.finish(2 * (uint32_t)pl.rounds);
}

/* ------------------------------------------------------------------ */
/* staged exchange (pull): allgather(v) alltoall(v) bcast gather(v) scatter(v) */
/* ------------------------------------------------------------------ */
static __device__ __forceinline__ void copy_bytes_grid(char *dst, const char *src, size_t n, bool src_is_peer)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    const uintptr_t mis = ((uintptr_t)dst | (uintptr_t)src);
    if ((mis & 15) == 0) {
        const size_t nv = n / 16;
        for (size_t v = tid; v < nv; v += nt) st_v4(dst + v * 16, src_is_peer ? ld_peer_v4(src + v * 16) : ld_src_v4(src + v * 16));
        for (size_t i = nv * 16 + tid; i < n; i += nt) dst[i] = *(volatile const char *)(src + i);
    } else if ((mis & 3) == 0) {
        const size_t nw = n / 4;
        for (size_t v = tid; v < nw; v += nt) ((uint32_t *)dst)[v] = ((volatile const uint32_t *)src)[v];
        for (size_t i = nw * 4 + tid; i < n; i += nt) dst[i] = *(volatile const char *)(src + i);
    } else {
        for (size_t i = tid; i < n; i += nt) dst[i] = *(volatile const char *)(src + i);
    }
}

__global__ void __launch_bounds__(1024) nvl_exchange_kernel(nvl_xchg_args_t a)
{
    const nvl_team_dev_t &t = a.team;
    const int N = t.size, me = t.rank;
    BlockSync bs; bs.init(t);
    char *mydata = data_of(t, me);
    /* phase A: publish my bytes (and the offset table peers use to find their block) */
    if (a.publish_table && blockIdx.x == 0 && (int)threadIdx.x < N) reinterpret_cast<uint64_t *>(mydata)[threadIdx.x] = (uint64_t)a.stage_off[threadIdx.x];
    if (a.src_bytes) copy_bytes_grid(mydata + NVL_XCHG_TABLE_BYTES, static_cast<const char *>(a.src), a.src_bytes, false);
    bs.signal(t, 1);
    bs.wait_all_blocks(t, 1);
    /* phase B: pull */
    for (int i = 0; i < N; i++) {
        const int p = (me + i) % N;
        const size_t n = a.pull_bytes[p];
        if (!n) continue;
        char *d = static_cast<char *>(a.dst) + a.dst_off[p];
        if (p == me) { if (d != static_cast<const char *>(a.src) + a.self_off) copy_bytes_grid(d, static_cast<const char *>(a.src) + a.self_off, n, false); continue; }
        size_t off = a.pull_off[p];
        if (off == NVL_XCHG_LOOKUP) off = (size_t) * reinterpret_cast<volatile const uint64_t *>(data_of(t, p) + (size_t)me * 8);
        copy_bytes_grid(d, data_of(t, p) + NVL_XCHG_TABLE_BYTES + off, n, true);
    }
    bs.signal(t, 2);
    bs.wait_all_blocks(t, 2);
    bs.finish(2);
}

/* ------------------------------------------------------------------ */
/* host-side launchers                                                 */
/* ------------------------------------------------------------------ */
extern "C" size_t nvl_dt_size(int dt)
{
    static const size_t s[NVL_DT_LAST] = {1, 2, 4, 8, 1, 2, 4, 8, 2, 4, 8, 2};
    return dt >= 0 && dt < NVL_DT_LAST ? s[dt] : 0;
}
extern "C" int nvl_dt_supports_op(int dt, int op)
{
    if (dt < 0 || dt >= NVL_DT_LAST || op < 0 || op >= NVL_OP_LAST) return 0;
    if (dt == NVL_DT_F16 || dt == NVL_DT_F32 || dt == NVL_DT_F64 || dt == NVL_DT_BF16)
        return op == NVL_OP_SUM || op == NVL_OP_PROD || op == NVL_OP_MAX || op == NVL_OP_MIN || op == NVL_OP_AVG;
    return 1;
}
extern "C" int nvl_nvls_supports(int dt, int op)
{ return (op == NVL_OP_SUM || op == NVL_OP_AVG) && (dt == NVL_DT_F32 || dt == NVL_DT_BF16 || dt == NVL_DT_F16); }

#define NVL_DISPATCH_DT(_dt, _KERNEL, _grid, _block, _s, _arg)                         \
    switch (_dt) {                                                                     \
    case NVL_DT_I8: _KERNEL<int8_t><<<_grid, _block, 0, _s>>>(_arg); break;            \
    case NVL_DT_I16: _KERNEL<int16_t><<<_grid, _block, 0, _s>>>(_arg); break;          \
    case NVL_DT_I32: _KERNEL<int32_t><<<_grid, _block, 0, _s>>>(_arg); break;          \
    case NVL_DT_I64: _KERNEL<int64_t><<<_grid, _block, 0, _s>>>(_arg); break;          \
    case NVL_DT_U8: _KERNEL<uint8_t><<<_grid, _block, 0, _s>>>(_arg); break;           \
    case NVL_DT_U16: _KERNEL<uint16_t><<<_grid, _block, 0, _s>>>(_arg); break;         \
    case NVL_DT_U32: _KERNEL<uint32_t><<<_grid, _block, 0, _s>>>(_arg); break;         \
    case NVL_DT_U64: _KERNEL<uint64_t><<<_grid, _block, 0, _s>>>(_arg); break;         \
    case NVL_DT_F16: _KERNEL<__half><<<_grid, _block, 0, _s>>>(_arg); break;           \
    case NVL_DT_F32: _KERNEL<float><<<_grid, _block, 0, _s>>>(_arg); break;            \
    case NVL_DT_F64: _KERNEL<double><<<_grid, _block, 0, _s>>>(_arg); break;           \
    case NVL_DT_BF16: _KERNEL<__nv_bfloat16><<<_grid, _block, 0, _s>>>(_arg); break;   \
    default: return cudaErrorInvalidValue;                                             \
    }

extern "C" cudaError_t nvl_launch_allreduce_oneshot(const nvl_red_args_t *a, int nblocks, int nthreads, cudaStream_t s)
{
    NVL_DISPATCH_DT(a->dt, nvl_allreduce_oneshot_kernel, nblocks, nthreads, s, *a);
    return cudaGetLastError();
}
extern "C" cudaError_t nvl_launch_reduce_staged(const nvl_red_args_t *a, int nblocks, int nthreads, cudaStream_t s)
{
    NVL_DISPATCH_DT(a->dt, nvl_reduce_staged_kernel, nblocks, nthreads, s, *a);
    return cudaGetLastError();
}
extern "C" cudaError_t nvl_launch_exchange(const nvl_xchg_args_t *a, int nblocks, int nthreads, cudaStream_t s)
{
    nvl_exchange_kernel<<<nblocks, nthreads, 0, s>>>(*a);
    return cudaGetLastError();
}
extern "C" cudaError_t nvl_launch_barrier(const nvl_team_dev_t *t, cudaStream_t s)
{
    nvl_barrier_kernel<<<1, 32, 0, s>>>(*t);
    return cudaGetLastError();
}
extern "C" cudaError_t nvl_launch_ctrl_init(void *heap_base, cudaStream_t s)
{
    nvl_ctrl_init_kernel<<<4, 256, 0, s>>>(static_cast<nvl_ctrl_t *>(heap_base));
    return cudaGetLastError();
}
