"""Data-parallel training wrapper: gradients live in flat per-bucket buffers; when the last gradient of a bucket
has been accumulated the bucket is averaged with ONE allreduce kernel (tl/nvl: one-shot / zero-copy two-shot /
NVLS picked by size) on a communication stream, overlapped with the rest of backward.

This is the consumer role that the reference serves through PyTorch's ProcessGroupUCC + DDP (BASELINE.json config
"torch-ucc ProcessGroupUCC DDP ResNet-50 bf16"): same algorithm (reverse-order buckets, allreduce(avg) per bucket,
stream-ordered via ucc_collective_triggered_post), without the c10d reducer in between."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class _Bucket:
    def __init__(self, params, dtype, device, alloc=None):
        self.params = params
        n = sum(p.numel() for p in params)
        # `alloc`: symmetric-memory allocator of the communicator (gradient buckets are then reduced in place in the switch)
        self.flat = alloc(n, dtype).zero_() if alloc is not None else torch.zeros(n, dtype=dtype, device=device)
        self.pending = 0
        self.req = None
        self.ready_event = torch.cuda.Event() if device.type == "cuda" else None
        off = 0
        for p in params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)   # autograd accumulates straight into the bucket
            off += p.numel()


class DistributedDataParallel(nn.Module):
    def __init__(self, module: nn.Module, comm=None, bucket_mb: float = 25.0, broadcast_params: bool = True, symmetric_buckets: bool = False):
        """`symmetric_buckets`: place the gradient buckets in the communicator's symmetric user region (Communicator(symm_size=...));
        every rank builds the same buckets in the same order, so they land at the same offsets."""
        super().__init__()
        self.module = module
        self.comm = comm or ops.default_comm()
        self._alloc = self.comm.symm_empty if symmetric_buckets and getattr(self.comm, "symm_region", lambda: None)() else None
        self.buckets = []
        self._of = {}
        params = [p for p in module.parameters() if p.requires_grad]
        dev = params[0].device
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        if broadcast_params and self.comm.size > 1:
            for p in module.state_dict().values():
                if torch.is_tensor(p) and p.numel():
                    ops.broadcast(p.data, 0, comm=self.comm)
        # reverse registration order ~ order in which gradients become ready
        cap = int(bucket_mb * (1 << 20))
        cur, cur_bytes, cur_dt = [], 0, None
        for p in reversed(params):
            nb = p.numel() * p.element_size()
            if cur and (cur_bytes + nb > cap or p.dtype != cur_dt):
                self._add_bucket(cur, cur_dt, dev)
                cur, cur_bytes = [], 0
            cur.append(p); cur_bytes += nb; cur_dt = p.dtype
        if cur:
            self._add_bucket(cur, cur_dt, dev)
        for p in params:
            p.register_post_accumulate_grad_hook(self._hook)
        self._arm()

    def _add_bucket(self, params, dtype, dev):
        b = _Bucket(params, dtype, dev, self._alloc if dev.type == "cuda" else None)
        for p in params:
            self._of[p] = b
        self.buckets.append(b)

    def _arm(self):
        for b in self.buckets:
            b.pending = len(b.params)

    def _hook(self, p):
        b = self._of[p]
        if p.grad.data_ptr() < b.flat.data_ptr() or p.grad.data_ptr() >= b.flat.data_ptr() + b.flat.numel() * b.flat.element_size():
            # somebody replaced .grad (e.g. zero_grad(set_to_none=True)): copy in and re-attach the view
            off = 0
            for q in b.params:
                if q is p:
                    view = b.flat[off:off + q.numel()].view_as(q)
                    view.copy_(p.grad)
                    p.grad = view
                off += q.numel()
        b.pending -= 1
        if b.pending == 0 and self.comm.size > 1:
            self._launch(b)

    def _launch(self, b):
        if self.comm_stream is None:
            b.req = self.comm.allreduce_init(b.flat, b.flat, op="avg"); b.req.post()
            return
        b.ready_event.record(torch.cuda.current_stream())
        self.comm_stream.wait_event(b.ready_event)
        b.req = self.comm.allreduce_init(b.flat, b.flat, op="avg")
        with torch.cuda.stream(self.comm_stream):
            b.req.post_on_stream(self.comm_stream)

    def forward(self, *a, **kw):
        return self.module(*a, **kw)

    def finish_gradient_sync(self):
        """Call after backward(): waits for the bucket collectives and orders the optimizer after them."""
        for b in self.buckets:
            if b.req is not None:
                b.req.wait(); b.req.finalize(); b.req = None
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._arm()

    def zero_grad(self, set_to_none: bool = False):
        for b in self.buckets:
            b.flat.zero_()
