/* Reduction kernels of tl/nvl (templates).  Each datatype is instantiated in its own translation unit
 * (nvl_red_<type>.cu) so the build parallelises; see nvl_kernels.cu for the design notes. */
#ifndef UCC_TL_NVL_REDUCE_IMPL_CUH_
#define UCC_TL_NVL_REDUCE_IMPL_CUH_
#include "nvl_device.cuh"

#define NVL_THREADS_MAX 512

/* ------------------------------------------------------------------ */
/* small helpers                                                       */
/* ------------------------------------------------------------------ */
static __device__ __forceinline__ size_t dmin(size_t a, size_t b) { return a < b ? a : b; }
static __device__ __forceinline__ char *data_of(const nvl_team_dev_t &t, int p) { return t.heap[p] + NVL_DATA_OFFSET; }

/* block-strided copy of 16-byte vectors [v0,v1): U loads in flight per thread */
template <int U, bool PEER> static __device__ __forceinline__ void copy_vecs(char *dst, const char *src, size_t v0, size_t v1)
{
    const size_t nt = blockDim.x;
    size_t v = v0 + threadIdx.x;
    for (; v + (U - 1) * nt < v1; v += U * nt) {
        uint4 x[U];
#pragma unroll
        for (int u = 0; u < U; u++) x[u] = PEER ? ld_peer_v4(src + (v + u * nt) * 16) : ld_src_v4(src + (v + u * nt) * 16);
#pragma unroll
        for (int u = 0; u < U; u++) st_v4(dst + (v + u * nt) * 16, x[u]);
    }
    for (; v < v1; v += nt) st_v4(dst + v * 16, PEER ? ld_peer_v4(src + v * 16) : ld_src_v4(src + v * 16));
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
            int p = me + i; if (p >= N) p -= N;
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
        uint4 x[NVL_MAX_PEERS];
        x[0] = load_src_vec<T>(src, v * E, n, sal);
#pragma unroll
        for (int i = 1; i < NVL_MAX_PEERS; i++) if (i < N) { int p = me + i; if (p >= N) p -= N; x[i] = ld_peer_v4(mybase + (size_t)p * NVL_LL_MAX + v * 16); }
        VecAcc<T, OP> acc; acc.set(x[0]);
#pragma unroll
        for (int i = 1; i < NVL_MAX_PEERS; i++) if (i < N) acc.add(x[i]);
        store_dst_vec<T>(dst, v * E, n, dal, acc.get(inv_n));
    }
}

template <typename T> __global__ void __launch_bounds__(NVL_THREADS_MAX) nvl_allreduce_oneshot_kernel(nvl_red_args_t a)
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
/* (datatype, operator) pairs the NVSwitch can reduce; must mirror nvl_nvls_supports() in nvl_kernels.cu */
template <typename T, int OP> struct McRed { static __device__ __forceinline__ uint4 ld(const void *) { return make_uint4(0, 0, 0, 0); } };
#define NVL_MCRED(_T, _OP, _fn) template <> struct McRed<_T, _OP> { static __device__ __forceinline__ uint4 ld(const void *p) { return _fn(p); } };
NVL_MCRED(float, NVL_OP_SUM, mc_ld_reduce_f32) NVL_MCRED(float, NVL_OP_AVG, mc_ld_reduce_f32)
NVL_MCRED(__half, NVL_OP_SUM, mc_ld_reduce_f16) NVL_MCRED(__half, NVL_OP_AVG, mc_ld_reduce_f16)
NVL_MCRED(__nv_bfloat16, NVL_OP_SUM, mc_ld_reduce_bf16) NVL_MCRED(__nv_bfloat16, NVL_OP_AVG, mc_ld_reduce_bf16)
NVL_MCRED(int32_t, NVL_OP_SUM, mc_red_add_s32) NVL_MCRED(int32_t, NVL_OP_MIN, mc_red_min_s32) NVL_MCRED(int32_t, NVL_OP_MAX, mc_red_max_s32)
NVL_MCRED(int32_t, NVL_OP_BAND, mc_red_and_b32) NVL_MCRED(int32_t, NVL_OP_BOR, mc_red_or_b32) NVL_MCRED(int32_t, NVL_OP_BXOR, mc_red_xor_b32)
NVL_MCRED(uint32_t, NVL_OP_SUM, mc_red_add_u32) NVL_MCRED(uint32_t, NVL_OP_MIN, mc_red_min_u32) NVL_MCRED(uint32_t, NVL_OP_MAX, mc_red_max_u32)
NVL_MCRED(uint32_t, NVL_OP_BAND, mc_red_and_b32) NVL_MCRED(uint32_t, NVL_OP_BOR, mc_red_or_b32) NVL_MCRED(uint32_t, NVL_OP_BXOR, mc_red_xor_b32)
NVL_MCRED(int64_t, NVL_OP_SUM, mc_red_add_u64) NVL_MCRED(int64_t, NVL_OP_MIN, mc_red_min_s64) NVL_MCRED(int64_t, NVL_OP_MAX, mc_red_max_s64)
NVL_MCRED(int64_t, NVL_OP_BAND, mc_red_and_b64) NVL_MCRED(int64_t, NVL_OP_BOR, mc_red_or_b64) NVL_MCRED(int64_t, NVL_OP_BXOR, mc_red_xor_b64)
NVL_MCRED(uint64_t, NVL_OP_SUM, mc_red_add_u64) NVL_MCRED(uint64_t, NVL_OP_MIN, mc_red_min_u64) NVL_MCRED(uint64_t, NVL_OP_MAX, mc_red_max_u64)
NVL_MCRED(uint64_t, NVL_OP_BAND, mc_red_and_b64) NVL_MCRED(uint64_t, NVL_OP_BOR, mc_red_or_b64) NVL_MCRED(uint64_t, NVL_OP_BXOR, mc_red_xor_b64)

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

/* destination pointer table of the zero-copy kernel's cold tail path */
struct DOut { char *p[NVL_MAX_PEERS]; int n; };

/* phase B worker: reduce vectors [j0,jend) of my slice. U vectors x NP sources in flight per thread. */
template <typename T, int OP, int NP, int U>
static __device__ __forceinline__ void reduce_p2p(const nvl_red_args_t &a, char *const *pd, size_t base, size_t j0, size_t jend, T *db, size_t rc, bool dal, float inv_n)
{
    constexpr int E = 16 / sizeof(T);
    const size_t nt = blockDim.x;
    const int N = a.team.size;
    for (size_t j = j0 + threadIdx.x; j < jend; j += U * nt) {
        uint4 x[U][NP];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t jj = j + u * nt;
            if (jj < jend) {
#pragma unroll
                for (int i = 0; i < NP; i++) if (i < N) x[u][i] = ld_peer_v4(pd[i] + base + jj * 16);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t jj = j + u * nt;
            if (jj < jend) {
                VecAcc<T, OP> acc; acc.set(x[u][0]);
#pragma unroll
                for (int i = 1; i < NP; i++) if (i < N) acc.add(x[u][i]);
                const uint4 r = acc.get(inv_n);
                const size_t o = base + jj * 16;
                if (a.kind == NVL_RED_ALLREDUCE) {
#pragma unroll
                    for (int i = 0; i < NP; i++) if (i < N) st_v4(pd[i] + o, r);
                } else if (a.kind == NVL_RED_REDUCE) st_v4(data_of(a.team, a.root) + o, r);
                else store_dst_vec<T>(db, jj * E, rc, dal, r);
            }
        }
    }
}

template <typename T, int OP, int U>
static __device__ __forceinline__ void reduce_nvls(const nvl_red_args_t &a, size_t base, size_t j0, size_t jend, T *db, size_t rc, bool dal, float inv_n)
{
    constexpr int E = 16 / sizeof(T);
    const size_t nt = blockDim.x;
    char *mc = a.team.mc_heap + NVL_DATA_OFFSET + base;
    for (size_t j = j0 + threadIdx.x; j < jend; j += U * nt) {
        uint4 r[U];
#pragma unroll
        for (int u = 0; u < U; u++) if (j + u * nt < jend) r[u] = McRed<T, OP>::ld(mc + (j + u * nt) * 16);
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t jj = j + u * nt;
            if (jj < jend) {
                uint4 v = r[u];
                if (OP == NVL_OP_AVG) { VecAcc<T, OP> acc; acc.set(v); v = acc.get(inv_n); }
                if (a.kind == NVL_RED_ALLREDUCE) mc_st_v4(mc + jj * 16, v);
                else if (a.kind == NVL_RED_REDUCE) st_v4(data_of(a.team, a.root) + base + jj * 16, v);
                else store_dst_vec<T>(db, jj * E, rc, dal, v);
            }
        }
    }
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
    for (int i = 0; i < NVL_MAX_PEERS; i++) { int p = me + i; if (p >= N) p -= N; pd[i] = i < N ? data_of(t, p) : nullptr; }

    for (int k = 0; k < pl.rounds; k++) {
        const size_t e0 = (size_t)k * pl.cap_e;                       /* first element of every slice in this round */
        const size_t rmax = dmin(pl.cap_e, pl.slice_max - e0);        /* longest slice part in this round */
        /* Blocks only synchronise with the SAME block of the peers, so block b may be a round ahead of block b' of its own rank:
         * the vector range a block owns must not move between rounds.  With several rounds it is cut from the full round
         * capacity (a shorter last round leaves the upper blocks idle) - cutting it from the round's own length let block b's
         * staging of the last round overwrite heap vectors block b' had not copied out yet (found in the host emulation). */
        const size_t nvmax = (rmax + E - 1) / E, nvsplit = pl.rounds > 1 ? pl.cap_e / E : nvmax, per = (nvsplit + nb - 1) / nb;
        const size_t j0 = dmin((size_t)b * per, nvmax), j1 = dmin(j0 + per, nvmax); /* my vector range inside every slice */

        /* phase A: stage vector range [j0,j1) of EVERY slice into my heap */
        for (int s = 0; s < N; s++) {
            const size_t rc = pl.cnt[s] > e0 ? dmin(pl.cap_e, pl.cnt[s] - e0) : 0;
            const size_t nfull = rc / E, jend = dmin(j1, (rc + E - 1) / E);
            const T *sb = src + pl.off[s] + e0;
            char *hb = mydata + (size_t)s * cap_bytes;
            if (((uintptr_t)sb & 15) == 0) {
                copy_vecs<8, false>(hb, reinterpret_cast<const char *>(sb), j0, dmin(jend, nfull));
                /* ragged tail: by the thread that owns this vector index in copy_vecs (v = j0 + tid mod blockDim) - the same thread
                 * copied the vector out of the heap in the previous round's phase C, and there is no block barrier in between */
                if (nfull >= j0 && nfull < jend && threadIdx.x == (unsigned)((nfull - j0) % blockDim.x)) st_v4(hb + nfull * 16, load_src_vec<T>(sb, nfull * E, rc, false));
            } else {
                for (size_t j = j0 + threadIdx.x; j < jend; j += blockDim.x) st_v4(hb + j * 16, load_src_vec<T>(sb, j * E, rc, false));
            }
        }
        bs.barrier(t, 2 * k + 1);

        /* phase B: reduce my slice from every member's staging area (over NVLink or in the switch) and publish it */
        {
            const size_t rc = pl.cnt[me] > e0 ? dmin(pl.cap_e, pl.cnt[me] - e0) : 0;
            const size_t jend = dmin(j1, (rc + E - 1) / E), base = (size_t)me * cap_bytes;
            T *db = dst + (a.kind == NVL_RED_REDUCE_SCATTER ? 0 : pl.off[me]) + e0;
            const bool dal = ((uintptr_t)db & 15) == 0;
            if (nvls) reduce_nvls<T, OP, 8>(a, base, j0, jend, db, rc, dal, inv_n);
            else if (N <= 2) reduce_p2p<T, OP, 2, 4>(a, pd, base, j0, jend, db, rc, dal, inv_n);
            else if (N <= 4) reduce_p2p<T, OP, 4, 2>(a, pd, base, j0, jend, db, rc, dal, inv_n);
            else if (N <= 8) reduce_p2p<T, OP, 8, 1>(a, pd, base, j0, jend, db, rc, dal, inv_n);
            else reduce_p2p<T, OP, NVL_MAX_PEERS, 1>(a, pd, base, j0, jend, db, rc, dal, inv_n);
        }
        bs.barrier(t, 2 * k + 2);

        /* phase C: my heap now holds vector range [j0,j1) of every reduced slice */
        if (a.kind == NVL_RED_ALLREDUCE || (a.kind == NVL_RED_REDUCE && me == a.root)) {
            for (int s = 0; s < N; s++) {
                const size_t rc = pl.cnt[s] > e0 ? dmin(pl.cap_e, pl.cnt[s] - e0) : 0;
                const size_t nfull = rc / E, jend = dmin(j1, (rc + E - 1) / E);
                T *db = dst + pl.off[s] + e0;
                const char *hb = mydata + (size_t)s * cap_bytes;
                if (((uintptr_t)db & 15) == 0) {
                    copy_vecs<8, true>(reinterpret_cast<char *>(db), hb, j0, dmin(jend, nfull));
                    if (nfull >= j0 && nfull < jend && threadIdx.x == (unsigned)((nfull - j0) % blockDim.x)) store_dst_vec<T>(db, nfull * E, rc, false, ld_peer_v4(hb + nfull * 16));
                } else {
                    for (size_t j = j0 + threadIdx.x; j < jend; j += blockDim.x) store_dst_vec<T>(db, j * E, rc, false, ld_peer_v4(hb + j * 16));
                }
            }
        }
    }
}

template <typename T> __global__ void __launch_bounds__(NVL_THREADS_MAX, 2) nvl_reduce_staged_kernel(nvl_red_args_t a)
{
    BlockSync bs; bs.init(a.team);
    SlicePlan pl; make_plan<T>(a, pl);
#define CALL_STAGED(_T, _OP) staged_body<_T, _OP>(a, bs, pl)
    NVL_DISPATCH_OP(T, a.op, CALL_STAGED);
    bs.finish(2 * (uint32_t)pl.rounds);
}


/* ------------------------------------------------------------------ */
/* zero-copy two-shot: no staging, no copy-out                         */
/*   barrier (every member's kernel is running => its src is final and */
/*   its dst may be written) -> reduce my slice straight from the N    */
/*   src buffers -> store into every member's dst -> barrier.          */
/*   Each byte crosses NVLink exactly once per direction and HBM is    */
/*   touched only for the payload itself.                              */
/* ------------------------------------------------------------------ */
template <typename T, int OP, int NP, int U>
static __device__ __forceinline__ void direct_p2p(const nvl_red_args_t &a, size_t so, size_t j0, size_t jend, float inv_n)
{
    const size_t nt = blockDim.x;
    const int N = a.team.size, me = a.team.rank;
    /* pointer tables with compile-time indices only, so they live in registers (a runtime-indexed copy would sit in local memory) */
    const char *sp[NP]; char *dp[NP]; int nd = 0;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        int p = me + i; if (p >= N) p -= N;
        sp[i] = i < N ? a.d.src[p] + so : nullptr;
        dp[i] = (i < N && a.kind == NVL_RED_ALLREDUCE) ? a.d.dst[p] + so : nullptr;
    }
    if (a.kind == NVL_RED_ALLREDUCE) nd = N;
    else { dp[0] = a.kind == NVL_RED_REDUCE ? a.d.dst[a.root] + so : static_cast<char *>(a.dst); nd = 1; }
    for (size_t j = j0 + threadIdx.x; j < jend; j += U * nt) {
        uint4 x[U][NP];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t jj = j + u * nt;
            if (jj < jend) {
#pragma unroll
                for (int i = 0; i < NP; i++) if (i < N) x[u][i] = ld_peer_v4(sp[i] + jj * 16);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t jj = j + u * nt;
            if (jj < jend) {
                VecAcc<T, OP> acc; acc.set(x[u][0]);
#pragma unroll
                for (int i = 1; i < NP; i++) if (i < N) acc.add(x[u][i]);
                const uint4 r = acc.get(inv_n);
#pragma unroll
                for (int i = 0; i < NP; i++) if (i < nd) st_v4(dp[i] + jj * 16, r);
            }
        }
    }
}

template <typename T, int OP, int NP, int U>
static __device__ __forceinline__ void direct_body(const nvl_red_args_t &a, const SlicePlan &pl)
{
    constexpr int E = 16 / sizeof(T);
    const nvl_team_dev_t &t = a.team;
    const int N = t.size, me = t.rank, nb = gridDim.x, b = blockIdx.x;
    const size_t cnt = pl.cnt[me], nfull = cnt / E, per = (nfull + nb - 1) / nb;
    const size_t j0 = dmin((size_t)b * per, nfull), j1 = dmin(j0 + per, nfull);
    const size_t so = pl.off[me] * sizeof(T);
    const float inv_n = 1.0f / (float)N;
    /* in flight per thread: U x NP 16-byte loads; 2 CTAs x 512 threads per SM keep > 64 KB outstanding per SM, far above
     * the NVLink bandwidth-delay product, and U x NP x 4 registers stay within the 64-register budget */
    direct_p2p<T, OP, NP, U>(a, so, j0, j1, inv_n); /* NP (team size class) is a kernel template parameter: each class gets its own register allocation */
    /* ragged tail of the slice (fewer than E elements): one thread, element by element */
    if (b == nb - 1 && threadIdx.x == 0 && nfull * E < cnt) {
        const char *sp[NVL_MAX_PEERS]; DOut dout;
        for (int i = 0; i < N; i++) { int p = me + i; if (p >= N) p -= N; sp[i] = a.d.src[p] + so; }
        if (a.kind == NVL_RED_ALLREDUCE) { for (int i = 0; i < N; i++) { int p = me + i; if (p >= N) p -= N; dout.p[i] = a.d.dst[p] + so; } dout.n = N; }
        else if (a.kind == NVL_RED_REDUCE) { dout.p[0] = a.d.dst[a.root] + so; dout.n = 1; }
        else { dout.p[0] = static_cast<char *>(a.dst); dout.n = 1; } /* reduce_scatter(v): my block, already a local pointer */
        for (size_t e = nfull * E; e < cnt; e++) {
            typename AccOf<T>::type acc = to_acc<T>(reinterpret_cast<const T *>(sp[0])[e]);
            for (int i = 1; i < N; i++) acc = OpFn<OP, typename AccOf<T>::type>::f(acc, to_acc<T>(reinterpret_cast<const T *>(sp[i])[e]));
            if (OP == NVL_OP_AVG) acc = (typename AccOf<T>::type)(acc * (typename AccOf<T>::type)inv_n);
            for (int i = 0; i < dout.n; i++) reinterpret_cast<T *>(dout.p[i])[e] = from_acc<T>(acc);
        }
    }
}

template <typename T, int NP, int U> __global__ void __launch_bounds__(NVL_THREADS_MAX, 2) nvl_reduce_direct_kernel(nvl_red_args_t a)
{
    BlockSync bs; bs.init(a.team);
    SlicePlan pl; make_plan<T>(a, pl);
    bs.barrier(a.team, 1);
#define CALL_DIRECT(_T, _OP) direct_body<_T, _OP, NP, U>(a, pl)
    NVL_DISPATCH_OP(T, a.op, CALL_DIRECT);
    bs.barrier(a.team, 2);
    bs.finish(2);
}

/* ------------------------------------------------------------------ */
/* step-structured allreduce / reduce_scatter through the heaps:       */
/*   sched 1: ring (N-1 reduce-scatter steps + N-1 allgather steps,    */
/*            neighbour links only)                                    */
/*   sched 2: recursive halving + recursive doubling (log2 N + log2 N  */
/*            pairwise steps; "SRA" with radix 2, power-of-two teams)  */
/* Both keep every slice at the same heap slot on every rank and give  */
/* block b the same vector range of every slice, so one per-block flag */
/* wait per step is all the synchronisation needed.  Single round: the */
/* vector has to fit the heap (larger ones use the two-shot kernels).  */
/* ------------------------------------------------------------------ */
/* acc[j] = op(peer[j], mine[j]) for my block's vectors of one slice; optionally scaled (AVG) and mirrored into dst */
template <typename T, int OP>
static __device__ __forceinline__ void step_reduce(const char *peer, char *mine, size_t j0, size_t jend, bool last, float inv_n, T *db, size_t rc, bool dal)
{
    constexpr int E = 16 / sizeof(T);
    const size_t nt = blockDim.x;
    for (size_t j = j0 + threadIdx.x; j < jend; j += 4 * nt) {
        uint4 x[4], y[4];
#pragma unroll
        for (int u = 0; u < 4; u++) if (j + u * nt < jend) { x[u] = ld_peer_v4(peer + (j + u * nt) * 16); y[u] = ld_peer_v4(mine + (j + u * nt) * 16); }
#pragma unroll
        for (int u = 0; u < 4; u++) if (j + u * nt < jend) {
            VecAcc<T, OP> acc; acc.set(y[u]); acc.add(x[u]);
            const uint4 r = acc.get(last ? inv_n : 1.0f); /* get() scales for AVG only; partial sums must stay unscaled */
            st_v4(mine + (j + u * nt) * 16, r);
            if (db) store_dst_vec<T>(db, (j + u * nt) * E, rc, dal, r);
        }
    }
}
/* mine[j] = peer[j] (+ dst) for my block's vectors of one slice */
template <typename T>
static __device__ __forceinline__ void step_copy(const char *peer, char *mine, size_t j0, size_t jend, bool keep, T *db, size_t rc, bool dal)
{
    constexpr int E = 16 / sizeof(T);
    const size_t nt = blockDim.x;
    for (size_t j = j0 + threadIdx.x; j < jend; j += 4 * nt) {
        uint4 x[4];
#pragma unroll
        for (int u = 0; u < 4; u++) if (j + u * nt < jend) x[u] = ld_peer_v4(peer + (j + u * nt) * 16);
#pragma unroll
        for (int u = 0; u < 4; u++) if (j + u * nt < jend) {
            if (keep) st_v4(mine + (j + u * nt) * 16, x[u]);
            store_dst_vec<T>(db, (j + u * nt) * E, rc, dal, x[u]);
        }
    }
}

template <typename T, int OP>
static __device__ __forceinline__ void steps_body(const nvl_red_args_t &a, BlockSync &bs, const SlicePlan &pl, uint32_t &phase)
{
    constexpr int E = 16 / sizeof(T);
    const nvl_team_dev_t &t = a.team;
    const int N = t.size, me = t.rank, nb = gridDim.x, b = blockIdx.x;
    const T *src = static_cast<const T *>(a.src);
    T *dst = static_cast<T *>(a.dst);
    const size_t cap_bytes = pl.cap_e * sizeof(T);
    const float inv_n = 1.0f / (float)N;
    const size_t nvmax = (pl.slice_max + E - 1) / E, per = (nvmax + nb - 1) / nb;
    const size_t j0 = dmin((size_t)b * per, nvmax), j1 = dmin(j0 + per, nvmax);
    char *mydata = data_of(t, me);
    const bool allred = a.kind == NVL_RED_ALLREDUCE;
#define SLICE_JEND(_s) dmin(j1, (pl.cnt[_s] + E - 1) / E)
#define SLICE_DST(_s) (dst + (allred ? pl.off[_s] : 0))
    /* stage every slice of my vector (same loop as phase A of the two-shot kernel) */
    for (int s = 0; s < N; s++) {
        const size_t rc = pl.cnt[s], nfull = rc / E, jend = SLICE_JEND(s);
        const T *sb = src + pl.off[s];
        char *hb = mydata + (size_t)s * cap_bytes;
        if (((uintptr_t)sb & 15) == 0) {
            copy_vecs<8, false>(hb, reinterpret_cast<const char *>(sb), j0, dmin(jend, nfull));
            if (nfull >= j0 && nfull < jend && threadIdx.x == 0) st_v4(hb + nfull * 16, load_src_vec<T>(sb, nfull * E, rc, false));
        } else {
            for (size_t j = j0 + threadIdx.x; j < jend; j += blockDim.x) st_v4(hb + j * 16, load_src_vec<T>(sb, j * E, rc, false));
        }
    }
    bs.barrier(t, ++phase);
    if (a.sched == 1) {
        /* ---- ring ---- slice j starts at rank j+1 and ends, complete, at rank j */
        const int L = (me + N - 1) % N;
        char *left = data_of(t, L);
        for (int s = 0; s + 1 < N; s++) {
            const int c = (me - 2 - s + 3 * N) % N;
            const bool last = s + 2 == N;
            T *db = last ? SLICE_DST(c) : nullptr; /* c == me in the last step */
            step_reduce<T, OP>(left + (size_t)c * cap_bytes, mydata + (size_t)c * cap_bytes, j0, SLICE_JEND(c), last, inv_n, db, pl.cnt[c], db && ((uintptr_t)db & 15) == 0);
            bs.barrier(t, ++phase);
        }
        if (allred) for (int s = 0; s + 1 < N; s++) {
            const int c = (me - 1 - s + 2 * N) % N;
            T *db = SLICE_DST(c);
            step_copy<T>(left + (size_t)c * cap_bytes, mydata + (size_t)c * cap_bytes, j0, SLICE_JEND(c), s + 2 < N, db, pl.cnt[c], ((uintptr_t)db & 15) == 0);
            bs.barrier(t, ++phase);
        }
    } else {
        /* ---- recursive halving (reduce-scatter) ... ---- */
        int lo = 0, hi = N; /* slices [lo,hi) are still mine to reduce */
        for (int d = N >> 1; d >= 1; d >>= 1) {
            const int partner = me ^ d, mid = lo + (hi - lo) / 2;
            const int klo = (me & d) ? mid : lo, khi = (me & d) ? hi : mid; /* the half I keep */
            const bool last = d == 1;
            char *pd = data_of(t, partner);
            for (int c = klo; c < khi; c++) {
                T *db = last ? SLICE_DST(c) : nullptr;
                step_reduce<T, OP>(pd + (size_t)c * cap_bytes, mydata + (size_t)c * cap_bytes, j0, SLICE_JEND(c), last, inv_n, db, pl.cnt[c], db && ((uintptr_t)db & 15) == 0);
            }
            lo = klo; hi = khi;
            bs.barrier(t, ++phase);
        }
        /* ---- ... recursive doubling (allgather) ---- */
        if (allred) for (int d = 1; d < N; d <<= 1) {
            const int partner = me ^ d, span = hi - lo;
            const int plo = (me & d) ? lo - span : hi; /* the partner's complete range sits next to mine */
            char *pd = data_of(t, partner);
            for (int c = plo; c < plo + span; c++) {
                T *db = SLICE_DST(c);
                step_copy<T>(pd + (size_t)c * cap_bytes, mydata + (size_t)c * cap_bytes, j0, SLICE_JEND(c), (d << 1) < N, db, pl.cnt[c], ((uintptr_t)db & 15) == 0);
            }
            if (me & d) lo -= span; else hi += span;
            bs.barrier(t, ++phase);
        }
    }
#undef SLICE_JEND
#undef SLICE_DST
}

template <typename T> __global__ void __launch_bounds__(NVL_THREADS_MAX, 2) nvl_reduce_steps_kernel(nvl_red_args_t a)
{
    BlockSync bs; bs.init(a.team);
    SlicePlan pl; make_plan<T>(a, pl);
    uint32_t phase = 0;
#define CALL_STEPS(_T, _OP) steps_body<_T, _OP>(a, bs, pl, phase)
    NVL_DISPATCH_OP(T, a.op, CALL_STEPS);
    bs.finish(phase);
}

/* one translation unit per datatype defines these launchers */
#define NVL_INSTANTIATE_REDUCE(_T, _suffix)                                                                              \
    extern "C" cudaError_t nvl_launch_oneshot_##_suffix(const nvl_red_args_t *a, int nblocks, int nthreads, cudaStream_t s) \
    { nvl_allreduce_oneshot_kernel<_T><<<nblocks, nthreads, 0, s>>>(*a); return cudaGetLastError(); }                     \
    extern "C" cudaError_t nvl_launch_staged_##_suffix(const nvl_red_args_t *a, int nblocks, int nthreads, cudaStream_t s)  \
    { nvl_reduce_staged_kernel<_T><<<nblocks, nthreads, 0, s>>>(*a); return cudaGetLastError(); }                      \
    extern "C" cudaError_t nvl_launch_direct_##_suffix(const nvl_red_args_t *a, int nblocks, int nthreads, cudaStream_t s)  \
    { if (a->team.size <= 2) nvl_reduce_direct_kernel<_T, 2, 4><<<nblocks, nthreads, 0, s>>>(*a);                       \
      else if (a->team.size <= 4) nvl_reduce_direct_kernel<_T, 4, 2><<<nblocks, nthreads, 0, s>>>(*a);                  \
      else if (a->team.size <= 8) nvl_reduce_direct_kernel<_T, 8, 1><<<nblocks, nthreads, 0, s>>>(*a);                  \
      else nvl_reduce_direct_kernel<_T, NVL_MAX_PEERS, 1><<<nblocks, nthreads, 0, s>>>(*a);                             \
      return cudaGetLastError(); }                                                                                     \
    extern "C" cudaError_t nvl_launch_steps_##_suffix(const nvl_red_args_t *a, int nblocks, int nthreads, cudaStream_t s)   \
    { nvl_reduce_steps_kernel<_T><<<nblocks, nthreads, 0, s>>>(*a); return cudaGetLastError(); }
#endif
