"""Megatron-style tensor parallel linear layers on ucc_b200 collectives (allgather / reduce_scatter / allreduce)."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class _AllReduce(torch.autograd.Function):
    """identity forward / allreduce backward (f) or allreduce forward / identity backward (g)"""

    @staticmethod
    def forward(ctx, x, comm, fwd):
        ctx.comm, ctx.fwd = comm, fwd
        if fwd:
            x = x.contiguous().clone()
            ops.all_reduce(x, comm=comm)
        return x

    @staticmethod
    def backward(ctx, g):
        if not ctx.fwd:
            g = g.contiguous().clone()
            ops.all_reduce(g, comm=ctx.comm)
        return g, None, None


class ColumnParallelLinear(nn.Module):
    """Y = X W^T with W split by output features; output stays sharded unless gather_output."""

    def __init__(self, in_features, out_features, comm=None, bias=True, gather_output=False, dtype=None, device=None):
        super().__init__()
        self.comm = comm or ops.default_comm()
        assert out_features % self.comm.size == 0
        self.local_out = out_features // self.comm.size
        self.gather_output = gather_output
        self.weight = nn.Parameter(torch.empty(self.local_out, in_features, dtype=dtype, device=device))
        self.bias = nn.Parameter(torch.zeros(self.local_out, dtype=dtype, device=device)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)

    def forward(self, x):
        x = _AllReduce.apply(x, self.comm, False)
        y = torch.nn.functional.linear(x, self.weight, self.bias)
        if not self.gather_output:
            return y
        return _AllGatherLastDim.apply(y, self.comm)


class _AllGatherLastDim(torch.autograd.Function):
    """forward: concatenate the members' output-feature shards along the last dimension (allgather); backward: every member
    keeps the slice of the gradient that belongs to its shard (the downstream computation is replicated, so the incoming
    gradient is already identical everywhere - Megatron's gather_from_tensor_model_parallel_region)."""

    @staticmethod
    def forward(ctx, y, comm):
        ctx.comm, ctx.width = comm, y.shape[-1]
        out = torch.empty((comm.size,) + tuple(y.shape), dtype=y.dtype, device=y.device)
        ops.all_gather_into_tensor(out, y.contiguous(), comm=comm)
        return out.movedim(0, -2).reshape(*y.shape[:-1], -1)

    @staticmethod
    def backward(ctx, grad):
        w, r = ctx.width, ctx.comm.rank
        return grad[..., r * w:(r + 1) * w].contiguous(), None


class RowParallelLinear(nn.Module):
    """Y = X W^T with W split by input features; partial products are summed with an allreduce."""

    def __init__(self, in_features, out_features, comm=None, bias=True, dtype=None, device=None):
        super().__init__()
        self.comm = comm or ops.default_comm()
        assert in_features % self.comm.size == 0
        self.local_in = in_features // self.comm.size
        self.weight = nn.Parameter(torch.empty(out_features, self.local_in, dtype=dtype, device=device))
        self.bias = nn.Parameter(torch.zeros(out_features, dtype=dtype, device=device)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)

    def forward(self, x_shard):
        y = torch.nn.functional.linear(x_shard, self.weight)
        y = _AllReduce.apply(y, self.comm, True)
        return y + self.bias if self.bias is not None else y
