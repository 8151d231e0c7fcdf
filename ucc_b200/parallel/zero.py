"""ZeRO-1 / FSDP-style sharded data parallelism on reduce_scatter + allgather ("ZeRO / FSDP" shape).

All parameters live in one flat buffer padded to a multiple of the world size; after backward the flat gradient is
reduce-scattered (average), each rank applies the optimizer to its shard only (optimizer state is 1/N per rank) and the
updated shards are allgathered back into the flat parameter buffer."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class ZeroRedundancyTrainer:
    def __init__(self, module: nn.Module, optimizer_cls=torch.optim.SGD, comm=None, symmetric: bool = False, **optim_kw):
        """`symmetric`: keep the flat parameter and gradient buffers in the communicator's symmetric user region
        (Communicator(symm_size=...)): the reduce_scatter then reduces in the switch straight out of the gradient buffer and
        the allgather multicasts the updated shards straight into every member's parameter buffer."""
        self.module = module
        self.comm = comm or ops.default_comm()
        n = self.comm.size
        params = [p for p in module.parameters() if p.requires_grad]
        self.params = params
        dev, dt = params[0].device, params[0].dtype
        total = sum(p.numel() for p in params)
        self.shard = (total + n - 1) // n
        if symmetric and dev.type == "cuda" and getattr(self.comm, "symm_region", lambda: None)():
            self.flat = self.comm.symm_empty(self.shard * n, dt).zero_()
            self.flat_grad = self.comm.symm_empty(self.shard * n, dt).zero_()
        else:
            self.flat = torch.zeros(self.shard * n, dtype=dt, device=dev)
            self.flat_grad = torch.zeros_like(self.flat)
        off = 0
        for p in params:                                  # parameters and gradients become views of the flat buffers
            self.flat[off:off + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + p.numel()].view_as(p)
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()
        if n > 1:
            ops.broadcast(self.flat, 0, comm=self.comm)
        r = self.comm.rank
        self.my_param = self.flat[r * self.shard:(r + 1) * self.shard]
        self.my_grad = torch.zeros_like(self.my_param)
        self._shadow = nn.Parameter(self.my_param.detach().clone())      # the optimizer owns only this shard
        self.optimizer = optimizer_cls([self._shadow], **optim_kw)

    def zero_grad(self):
        self.flat_grad.zero_()

    def step(self):
        """call after backward()"""
        if self.comm.size > 1:
            ops.reduce_scatter_tensor(self.my_grad, self.flat_grad, op="avg", comm=self.comm)
        else:
            self.my_grad.copy_(self.flat_grad)
        self._shadow.grad = self.my_grad
        self.optimizer.step()
        if self.comm.size > 1:
            ops.all_gather_into_tensor(self.flat, self._shadow.data.contiguous(), comm=self.comm)
        else:
            self.flat.copy_(self._shadow.data)
