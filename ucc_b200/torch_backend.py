"""`torch.distributed` backend "ucc_b200": the role PyTorch's ProcessGroupUCC plays for the reference
(torch/csrc/distributed/c10d/ProcessGroupUCC.cpp consumes ucc_collective_init / triggered_post / test).

    import ucc_b200.torch_backend            # registers the backend
    torch.distributed.init_process_group("ucc_b200", rank=..., world_size=...)
    torch.distributed.all_reduce(t)          # -> libucc: tl/nvl NVLink kernels for CUDA tensors, tl/shm for host tensors
    torch.nn.parallel.DistributedDataParallel(model)   # gradient buckets go through the same path

The process group is a Python subclass of `torch.distributed.ProcessGroup`; wire-up uses the rendezvous store as the
UCC out-of-band allgather.  CUDA collectives are posted stream-ordered on the caller's current stream, so `Work.wait()`
only has to make sure the kernel is in the stream (it does not block the host on the GPU)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from .dist import Communicator

_OPS = {dist.ReduceOp.SUM: "sum", dist.ReduceOp.PRODUCT: "prod", dist.ReduceOp.MIN: "min", dist.ReduceOp.MAX: "max", dist.ReduceOp.AVG: "avg",
        dist.ReduceOp.BAND: "band", dist.ReduceOp.BOR: "bor", dist.ReduceOp.BXOR: "bxor"}


def _op(o):
    o = getattr(o, "op", o)   # ReduceOp object or RedOpType
    for k, v in _OPS.items():
        if o == k:
            return v
    raise ValueError(f"unsupported reduce op {o}")


class _Work(dist._Work):
    def __init__(self, pg, reqs, result, cuda, side_stream=None):
        super().__init__()
        self._pg, self._reqs, self._result, self._cuda = pg, reqs, result, cuda
        self._side = side_stream      # CUDA p2p runs on an internal stream; wait() makes the caller's stream wait for it
        self._fut = torch.futures.Future()
        self._done = False

    def _finish(self):
        if self._done:
            return
        for r in self._reqs:
            if not self._cuda:
                r.wait()            # host buffers: completion == data valid
            else:
                r.wait_posted()     # the kernel is in the stream (idempotent; immediate for everything but p2p)
            r.finalize_later() if self._cuda else r.finalize()
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
        self._done = True
        self._fut.set_result(self._result)

    def is_completed(self):
        return all(r.test() == 0 for r in self._reqs)

    def is_success(self):
        return True

    def wait(self, timeout=None):
        self._finish()
        return True

    def get_future(self):
        self._finish()
        return self._fut

    def result(self):
        return self._result


class ProcessGroupUCCB200(dist.ProcessGroup):
    def __init__(self, store, rank, size, timeout=None):
        super().__init__(rank, size)
        self._comm = Communicator(store=store, rank=rank, size=size)
        self._pending = []   # CUDA requests whose kernels are in flight; finalized lazily
        self._p2p_streams = {}

    def getBackendName(self):
        return "ucc_b200"

    # ---- helpers
    def _run(self, reqs, tensors, result, wait_posted=True):
        cuda = bool(tensors) and tensors[0].is_cuda
        for r in reqs:
            if cuda:
                r.post_on_stream(wait_posted=wait_posted)   # returns once the kernel is in the current stream
                r.finalize_later = lambda r=r: self._pending.append(r)
            else:
                r.post()
        self._reap()
        return _Work(self, reqs, result, cuda)

    def _reap(self):
        keep = []
        for r in self._pending:
            if r.test() == 0:
                r.finalize()
            else:
                keep.append(r)
        self._pending = keep
        if len(keep) > 256:
            for r in keep:
                r.wait(); r.finalize()
            self._pending = []

    # ---- collectives (list-of-tensors signatures of c10d)
    def allreduce(self, tensors, opts=None):
        op = _op(opts.reduceOp) if opts is not None else "sum"
        reqs = [self._comm.allreduce_init(t, t, op=op) for t in tensors]
        return self._run(reqs, tensors, tensors)

    def allreduce_coalesced(self, tensors, opts=None):
        return self.allreduce(tensors, opts)

    def broadcast(self, tensors, opts=None):
        root = opts.rootRank if opts is not None else 0
        reqs = [self._comm.coll_init("bcast", t, None, root=root) for t in tensors]
        return self._run(reqs, tensors, tensors)

    def reduce(self, tensors, opts=None):
        root = opts.rootRank if opts is not None else 0
        op = _op(opts.reduceOp) if opts is not None else "sum"
        me = self.rank()
        reqs = [self._comm.coll_init("reduce", None if me == root else t, t if me == root else None, op=op, root=root, inplace=me == root) for t in tensors]
        return self._run(reqs, tensors, tensors)

    def allgather(self, output_lists, input_tensors, opts=None):
        reqs, flats = [], []
        for outs, inp in zip(output_lists, input_tensors):
            flat = torch.empty((len(outs),) + tuple(inp.shape), dtype=inp.dtype, device=inp.device)
            reqs.append(self._comm.coll_init("allgather", inp.contiguous(), flat))
            flats.append((flat, outs))
        w = self._run(reqs, input_tensors, output_lists)
        if input_tensors and input_tensors[0].is_cuda:
            for flat, outs in flats:      # stream-ordered behind the collective
                for i, o in enumerate(outs):
                    o.copy_(flat[i])
        else:
            w.wait()
            for flat, outs in flats:
                for i, o in enumerate(outs):
                    o.copy_(flat[i])
        return w

    def _allgather_base(self, output, input, opts=None):
        return self._run([self._comm.coll_init("allgather", input.contiguous(), output)], [input], output)

    def allgather_into_tensor_coalesced(self, outputs, inputs, opts=None):
        return self._run([self._comm.coll_init("allgather", i.contiguous(), o) for o, i in zip(outputs, inputs)], inputs, outputs)

    def _reduce_scatter_base(self, output, input, opts=None):
        op = _op(opts.reduceOp) if opts is not None else "sum"
        return self._run([self._comm.coll_init("reduce_scatter", input.contiguous(), output, op=op)], [input], output)

    def reduce_scatter(self, outputs, input_lists, opts=None):
        op = _op(opts.reduceOp) if opts is not None else "sum"
        reqs, keep = [], []
        for out, ins in zip(outputs, input_lists):
            flat = torch.cat([t.reshape(-1) for t in ins])
            keep.append(flat)
            reqs.append(self._comm.coll_init("reduce_scatter", flat, out, op=op))
        w = self._run(reqs, outputs, outputs)
        w._keep = keep
        return w

    def alltoall_base(self, output, input, output_split_sizes, input_split_sizes, opts=None):
        if not output_split_sizes and not input_split_sizes:
            return self._run([self._comm.coll_init("alltoall", input.contiguous(), output)], [input], output)
        row = input[0].numel() if input.dim() > 1 else 1
        sc = [int(s) * row for s in input_split_sizes]; rc = [int(s) * row for s in output_split_sizes]
        sd = [sum(sc[:i]) for i in range(len(sc))]; rd = [sum(rc[:i]) for i in range(len(rc))]
        req = self._comm.coll_init("alltoallv", input.contiguous(), output, src_counts=sc, src_displs=sd, dst_counts=rc, dst_displs=rd)
        return self._run([req], [input], output)

    # ---- point to point: a two-member active-set broadcast, exactly how ProcessGroupUCC maps send/recv onto UCC.
    # CUDA tensors travel over tl/nvl's heap channels: messages between an ordered pair of ranks are matched in POST ORDER (the NCCL
    # contract: the tag does not reorder them); host tensors go through tl/shm, which matches by tag.
    def _p2p(self, tensors, src, dst, tag):
        reqs = [self._comm.coll_init("bcast", t, None, root=src, active_set=(src, dst - src, 2), tag=int(tag) & 0x3fff) for t in tensors]
        # rendezvous: a large send enters the stream when the receiver has published its buffer, a receive is published when the
        # stream reaches it - both need progress, which Work.wait() provides (isend + irecv + wait, as with every backend)
        if not (tensors and tensors[0].is_cuda):
            return self._run(reqs, tensors, tensors, wait_posted=False)
        # CUDA: one internal stream per (peer, direction), as ProcessGroupNCCL keeps per-peer streams - a kernel of the eager ring
        # may wait for its peer (a receive posted before its send), and torch issues the ops of batch_isend_irecv one after the
        # other on the caller's stream: an irecv in front of an isend there would deadlock a ring of ranks
        peer, is_send = (dst, True) if src == self.rank() else (src, False)
        side = self._p2p_streams.get((peer, is_send))
        if side is None:
            side = self._p2p_streams[(peer, is_send)] = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        for r, t in zip(reqs, tensors):
            r.post_on_stream(side, wait_posted=False)
            r.finalize_later = lambda r=r: self._pending.append(r)
            t.record_stream(side)
        self._reap()
        return _Work(self, reqs, tensors, True, side_stream=side)

    def send(self, tensors, dstRank, tag=0):
        return self._p2p(tensors, self.rank(), dstRank, tag)

    def recv(self, tensors, srcRank, tag=0):
        return self._p2p(tensors, srcRank, self.rank(), tag)

    def barrier(self, opts=None):
        self._comm.barrier()
        w = _Work(self, [], None, False)
        return w

    def shutdown(self):
        for r in self._pending:
            r.wait(); r.finalize()
        self._pending = []
        self._comm.destroy()


def _create(store, rank, size, timeout):
    return ProcessGroupUCCB200(store, rank, size, timeout)


def register():
    if "UCC_B200" not in [b for b in dir(dist.Backend)]:
        dist.Backend.register_backend("ucc_b200", _create, devices=["cpu", "cuda"])


register()
