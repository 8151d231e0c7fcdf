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


class _SeqGather(torch.autograd.Function):
    """Megatron sequence parallelism, entry of a tensor-parallel region: forward = allgather of the sequence shards along dim 0,
    backward = reduce_scatter (sum) of the gradient back to the shards"""

    @staticmethod
    def forward(ctx, x, comm):
        ctx.comm = comm
        out = torch.empty((comm.size * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        ops.all_gather_into_tensor(out, x.contiguous(), comm=comm)
        return out

    @staticmethod
    def backward(ctx, g):
        n = ctx.comm.size
        out = torch.empty((g.shape[0] // n,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        ops.reduce_scatter_tensor(out, g.contiguous(), op="sum", comm=ctx.comm)
        return out, None


class _SeqScatter(torch.autograd.Function):
    """exit of the region: forward = reduce_scatter (sum of the partial products, result sharded along the sequence),
    backward = allgather of the gradient"""

    @staticmethod
    def forward(ctx, y, comm):
        ctx.comm = comm
        out = torch.empty((y.shape[0] // comm.size,) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
        ops.reduce_scatter_tensor(out, y.contiguous(), op="sum", comm=comm)
        return out

    @staticmethod
    def backward(ctx, g):
        out = torch.empty((ctx.comm.size * g.shape[0],) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        ops.all_gather_into_tensor(out, g.contiguous(), comm=ctx.comm)
        return out, None


class ColumnParallelLinear(nn.Module):
    """Y = X W^T with W split by output features; output stays sharded unless gather_output.
    `sequence_parallel`: the input arrives sharded along dim 0 (the sequence) and is all-gathered here (backward: reduce_scatter) -
    the Megatron-SP pairing with RowParallelLinear(sequence_parallel=True), which replaces the allreduce by a reduce_scatter."""

    def __init__(self, in_features, out_features, comm=None, bias=True, gather_output=False, dtype=None, device=None, sequence_parallel=False):
        super().__init__()
        self.comm = comm or ops.default_comm()
        assert out_features % self.comm.size == 0
        self.local_out = out_features // self.comm.size
        self.gather_output = gather_output
        self.sequence_parallel = sequence_parallel
        self.weight = nn.Parameter(torch.empty(self.local_out, in_features, dtype=dtype, device=device))
        self.bias = nn.Parameter(torch.zeros(self.local_out, dtype=dtype, device=device)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)

    def forward(self, x):
        x = _SeqGather.apply(x, self.comm) if self.sequence_parallel else _AllReduce.apply(x, self.comm, False)
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

    def __init__(self, in_features, out_features, comm=None, bias=True, dtype=None, device=None, sequence_parallel=False):
        super().__init__()
        self.comm = comm or ops.default_comm()
        self.sequence_parallel = sequence_parallel    # output sharded along dim 0 by a reduce_scatter instead of an allreduce
        assert in_features % self.comm.size == 0
        self.local_in = in_features // self.comm.size
        self.weight = nn.Parameter(torch.empty(out_features, self.local_in, dtype=dtype, device=device))
        self.bias = nn.Parameter(torch.zeros(out_features, dtype=dtype, device=device)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)

    def forward(self, x_shard):
        y = torch.nn.functional.linear(x_shard, self.weight)
        y = _SeqScatter.apply(y, self.comm) if self.sequence_parallel else _AllReduce.apply(y, self.comm, True)
        return y + self.bias if self.bias is not None else y
