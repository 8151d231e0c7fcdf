/* Point-to-point over the team heap: the two-member active-set broadcast (UCC's send / recv, what ProcessGroupUCC and every
 * pipeline- / context-parallel exchange use).  Reference: tl/cuda bcast_linear.c:8-60 + tl_cuda_coll.h:122-166 - a host
 * state machine that copies root -> scratch -> peer in steps, each acknowledged through a shared-memory barrier polled by the CPU.
 *
 * Here it is ONE kernel per side and the data crosses NVLink once, as stores:
 *     sender   : for every chunk  wait for ring space (ack counter in MY heap, advanced by the receiver)
 *                                 store the chunk into the RECEIVER's heap ring            (fire-and-forget NVLink writes)
 *                                 fence + bump the receiver's head counter
 *     receiver : for every chunk  wait for the head counter (in my heap), copy ring -> dst (local HBM), bump the sender's ack
 * Every ordered pair (src -> dst) of a team owns a channel in dst's heap: a ring of NVL_P2P_SLOTS chunks, cut into up to
 * NVL_P2P_MAX_CTAS independent lanes (CTA b of the sender talks to CTA b of the receiver; both derive the lane count from the
 * message size, which is a collective argument).  The produced / consumed counters live on the device and persist across
 * launches, so messages between a pair are matched in post order (as in NCCL), the sender of a message that fits the ring
 * never waits for the receiver's kernel, captured launches replay correctly, and nothing here touches the barrier epochs of the
 * team's collectives - a send / recv can be in flight next to an allreduce of the same team. */
#include "nvl_reduce_impl.cuh"

/* Waiting for the OTHER side of a message is not bounded by anything the library controls - a receive may be posted long before
 * its send (pre-posted receives of a pipeline stage) - so, like NCCL, the channel kernels wait without a deadline unless
 * P2P_TIMEOUT is set (args.spin_ns != 0); the TIMEOUT of the collectives, whose members all launch the same kernel, does not apply */
static __device__ __forceinline__ void p2p_spin(const nvl_p2p_args_t &a, nvl_ctrl_t *mine, const uint32_t *f, uint32_t target)
{
    uint64_t t0 = 0; uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys_u32(f) - target) < 0) {
        if (a.spin_ns && (++spins & 0x3ff) == 0) {
            uint64_t now = globaltimer_ns();
            if (!t0) t0 = now;
            else if (now - t0 > a.spin_ns) { if (a.team.host_err) *a.team.host_err = 1; mine->error = 1; break; }
        }
    }
}

/* CTA-wide copy of n bytes; `aligned` = both pointers 16-byte aligned.  PEER: the source is written by another GPU during the kernel */
template <bool PEER> static __device__ __forceinline__ void p2p_copy(char *dst, const char *src, size_t n, bool aligned)
{
    const size_t nt = blockDim.x, tid = threadIdx.x;
    if (aligned) {
        const size_t nv = n / 16;
        size_t v = tid;
        for (; v + 3 * nt < nv; v += 4 * nt) {
            uint4 x[4];
#pragma unroll
            for (int u = 0; u < 4; u++) x[u] = PEER ? ld_peer_v4(src + (v + u * nt) * 16) : ld_src_v4(src + (v + u * nt) * 16);
#pragma unroll
            for (int u = 0; u < 4; u++) st_v4(dst + (v + u * nt) * 16, x[u]);
        }
        for (; v < nv; v += nt) st_v4(dst + v * 16, PEER ? ld_peer_v4(src + v * 16) : ld_src_v4(src + v * 16));
        for (size_t i = nv * 16 + tid; i < n; i += nt) dst[i] = *(volatile const char *)(src + i);
    } else {
        for (size_t i = tid; i < n; i += nt) dst[i] = *(volatile const char *)(src + i);
    }
}

__global__ void __launch_bounds__(NVL_THREADS_MAX) nvl_p2p_kernel(const __grid_constant__ nvl_p2p_args_t a)
{
    const nvl_team_dev_t &t = a.team;
    const int me = t.rank, peer = a.peer, b = blockIdx.x, nb = gridDim.x;
    nvl_ctrl_t *mine = reinterpret_cast<nvl_ctrl_t *>(t.heap[me]);
    nvl_ctrl_t *theirs = reinterpret_cast<nvl_ctrl_t *>(t.heap[peer]);
    /* lane b of a chunk: bytes [b * lane, (b + 1) * lane) of the chunk.  The lane WIDTH is fixed (chunk / max lanes), only the
     * number of lanes a message uses follows from its size: every lane is then an independent ring with its own byte range and
     * its own counters.  (With a width of chunk / nb, messages of different lane counts that were in flight together overlapped in
     * the slots: lane 1 of a 4-lane message could land on unconsumed data of lane 0 of a 3-lane one.) */
    const size_t chunk = NVL_P2P_CHUNK, lane = NVL_P2P_LANE_BYTES;
    const size_t per_chunk = lane * nb;                              /* payload bytes one chunk carries with this lane count */
    const uint32_t nchunks = (uint32_t)((a.bytes + per_chunk - 1) / per_chunk);
    char *ubuf = static_cast<char *>(a.buf);
    if (a.send) {
        const uint32_t tx0 = ld_volatile_u32(&mine->p2p_tx[peer][b]);
        char *ring = t.heap[peer] + NVL_P2P_OFFSET + (size_t)me * NVL_P2P_CHAN_BYTES;   /* my channel in the receiver's heap */
        for (uint32_t c = 0; c < nchunks; c++) {
            const uint32_t seq = tx0 + c;
            const size_t off = (size_t)c * per_chunk + (size_t)b * lane;
            const size_t n = off < a.bytes ? dmin(lane, a.bytes - off) : 0;
            if (threadIdx.x == 0) p2p_spin(a, mine, &mine->p2p_ack[peer][b], seq + 1 - NVL_P2P_SLOTS);   /* ring slot free again */
            __syncthreads();
            char *slot = ring + (size_t)(seq % NVL_P2P_SLOTS) * chunk + (size_t)b * lane;
            if (n) p2p_copy<false>(slot, ubuf + off, n, (((uintptr_t)(ubuf + off)) & 15) == 0);
            __syncthreads();
            if (threadIdx.x == 0) { fence_sys(); st_relaxed_sys_u32(&theirs->p2p_head[me][b], seq + 1); }
        }
        if (threadIdx.x == 0) mine->p2p_tx[peer][b] = tx0 + nchunks;
    } else {
        const uint32_t rx0 = ld_volatile_u32(&mine->p2p_rx[peer][b]);
        const char *ring = t.heap[me] + NVL_P2P_OFFSET + (size_t)peer * NVL_P2P_CHAN_BYTES;
        for (uint32_t c = 0; c < nchunks; c++) {
            const uint32_t seq = rx0 + c;
            const size_t off = (size_t)c * per_chunk + (size_t)b * lane;
            const size_t n = off < a.bytes ? dmin(lane, a.bytes - off) : 0;
            if (threadIdx.x == 0) p2p_spin(a, mine, &mine->p2p_head[peer][b], seq + 1);
            __syncthreads();
            const char *slot = ring + (size_t)(seq % NVL_P2P_SLOTS) * chunk + (size_t)b * lane;
            if (n) p2p_copy<true>(ubuf + off, slot, n, (((uintptr_t)(ubuf + off)) & 15) == 0);
            __syncthreads();   /* every thread's loads of the slot have returned (their values were stored) */
            if (threadIdx.x == 0) st_relaxed_sys_u32(&theirs->p2p_ack[me][b], seq + 1);   /* (no fence: nothing the sender reads was written here) */
        }
        if (threadIdx.x == 0) mine->p2p_rx[peer][b] = rx0 + nchunks;
    }
}

/* Rendezvous sender: the receiver's buffer is mapped here (a.remote) and the receiver's stream has reached its recv (the host
 * published the buffer only then), so this kernel never waits: every CTA stores its contiguous part over NVLink, the last CTA
 * to finish makes the message visible - fence, then the message counter in the receiver's heap. */
__global__ void __launch_bounds__(NVL_THREADS_MAX) nvl_p2p_push_kernel(const __grid_constant__ nvl_p2p_args_t a)
{
    const nvl_team_dev_t &t = a.team;
    const int me = t.rank, peer = a.peer, b = blockIdx.x, nb = gridDim.x;
    nvl_ctrl_t *mine = reinterpret_cast<nvl_ctrl_t *>(t.heap[me]);
    nvl_ctrl_t *theirs = reinterpret_cast<nvl_ctrl_t *>(t.heap[peer]);
    const char *src = static_cast<const char *>(a.buf);
    const bool aligned = ((((uintptr_t)src) | ((uintptr_t)a.remote)) & 15) == 0;
    const size_t seg = ((a.bytes + nb - 1) / nb + 15) / 16 * 16;
    const size_t off = (size_t)b * seg;
    if (off < a.bytes) p2p_copy<false>(a.remote + off, src + off, dmin(seg, a.bytes - off), aligned);
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_sys();
        if (atomicAdd(&mine->p2p_rz_blocks[peer], 1u) == (uint32_t)nb - 1) {
            mine->p2p_rz_blocks[peer] = 0;
            fence_sys();
            st_relaxed_sys_u32(&theirs->p2p_rz_done[me], a.rz_seq + 1);
        }
    }
}

/* Rendezvous receiver: the data is written into a.buf by the peer's kernel; later work in this stream starts after it landed */
__global__ void __launch_bounds__(32) nvl_p2p_wait_kernel(const __grid_constant__ nvl_p2p_args_t a)
{
    nvl_ctrl_t *mine = reinterpret_cast<nvl_ctrl_t *>(a.team.heap[a.team.rank]);
    if (threadIdx.x == 0) p2p_spin(a, mine, &mine->p2p_rz_done[a.peer], a.rz_seq + 1);
}

#ifndef NVL_HOST_EMU
extern "C" int nvl_p2p_push_ctas(size_t bytes)
{
    size_t n = bytes / (64 * 1024);
    return n < 1 ? 1 : (n > NVL_P2P_PUSH_MAX_CTAS ? NVL_P2P_PUSH_MAX_CTAS : (int)n);
}
extern "C" int nvl_p2p_lanes(size_t bytes)
{
    size_t n = (bytes + NVL_P2P_LANE_BYTES - 1) / NVL_P2P_LANE_BYTES;
    return n < 1 ? 1 : (n > NVL_P2P_MAX_CTAS ? NVL_P2P_MAX_CTAS : (int)n);
}
extern "C" cudaError_t nvl_launch_p2p(const nvl_p2p_args_t *a, int nthreads, cudaStream_t s)
{
    if (nthreads > NVL_THREADS_MAX) nthreads = NVL_THREADS_MAX;
    if (a->peer < 0 || a->peer >= a->team.size || a->peer == a->team.rank) return cudaErrorInvalidValue;
    if (a->mode == NVL_P2P_PUSH) nvl_p2p_push_kernel<<<nvl_p2p_push_ctas(a->bytes), nthreads, 0, s>>>(*a);
    else if (a->mode == NVL_P2P_WAIT) nvl_p2p_wait_kernel<<<1, 32, 0, s>>>(*a);
    else nvl_p2p_kernel<<<nvl_p2p_lanes(a->bytes), nthreads, 0, s>>>(*a);
    return cudaGetLastError();
}
#endif
