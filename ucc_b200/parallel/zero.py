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


class FullyShardedModule(nn.Module):
    """FSDP / ZeRO-3 unit: between uses only 1/N of the wrapped module's parameters lives on a rank (`shard_param`, which is also
    what the optimizer sees).  The full parameters are all-gathered right before the unit's forward and again right before its
    backward and released afterwards (the flat buffer's storage is resized to zero - views and autograd's saved tensors keep
    pointing at it and find the re-gathered values when it is filled again); once every parameter of the unit has its gradient,
    the flat gradient is reduce-scattered (average) into `shard_param.grad` and the full gradients are dropped.
    Collectives used: allgather (twice per step and unit) and reduce_scatter - the SURVEY 2.9 "ZeRO / FSDP" shape."""

    def __init__(self, module: nn.Module, comm=None):
        super().__init__()
        self.module, self.comm = module, comm or ops.default_comm()
        n, r = self.comm.size, self.comm.rank
        self._params = [p for p in module.parameters() if p.requires_grad]
        dev, dt = self._params[0].device, self._params[0].dtype
        total = sum(p.numel() for p in self._params)
        self.shard = (total + n - 1) // n
        self._full = torch.zeros(self.shard * n, dtype=dt, device=dev)
        off = 0
        for p in self._params:
            self._full[off:off + p.numel()].copy_(p.data.reshape(-1))
            p.data = self._full[off:off + p.numel()].view_as(p)
            off += p.numel()
        if n > 1:
            ops.broadcast(self._full, 0, comm=self.comm)
        self.shard_param = nn.Parameter(self._full[r * self.shard:(r + 1) * self.shard].detach().clone())
        self._nbytes = self._full.untyped_storage().nbytes()
        self._resident = True
        self._pending = 0
        self._release()
        module.register_forward_pre_hook(lambda m, a: self._gather())
        module.register_forward_hook(lambda m, a, o: self._release())
        module.register_full_backward_pre_hook(lambda m, g: self._gather())
        for p in self._params:
            p.register_post_accumulate_grad_hook(self._grad_ready)

    def parameters_for_optimizer(self):
        return [self.shard_param]

    def _gather(self):
        if self._resident:
            return
        self._full.untyped_storage().resize_(self._nbytes)
        if self.comm.size > 1:
            ops.all_gather_into_tensor(self._full, self.shard_param.data.contiguous(), comm=self.comm)
        else:
            self._full.copy_(self.shard_param.data)
        self._resident = True

    def _release(self):
        if self._resident:
            self._full.untyped_storage().resize_(0)
            self._resident = False

    def _grad_ready(self, p):
        self._pending += 1
        if self._pending < len(self._params):
            return
        self._pending = 0
        flat = torch.zeros(self.shard * self.comm.size, dtype=self.shard_param.dtype, device=self.shard_param.device)
        off = 0
        for q in self._params:
            flat[off:off + q.numel()].copy_(q.grad.reshape(-1))
            q.grad = None
            off += q.numel()
        g = torch.empty_like(self.shard_param.data)
        if self.comm.size > 1:
            ops.reduce_scatter_tensor(g, flat, op="avg", comm=self.comm)
        else:
            g.copy_(flat)
        self.shard_param.grad = g if self.shard_param.grad is None else self.shard_param.grad + g
        self._release()

    def forward(self, *a, **kw):
        return self.module(*a, **kw)

    def full_parameters(self):
        """gathered copy of the unit's parameters (checkpointing / tests)"""
        out = torch.empty(self.shard * self.comm.size, dtype=self.shard_param.dtype, device=self.shard_param.device)
        if self.comm.size > 1:
            ops.all_gather_into_tensor(out, self.shard_param.data.contiguous(), comm=self.comm)
        else:
            out.copy_(self.shard_param.data)
        return out[:sum(p.numel() for p in self._params)]
