"""Sequence / context parallel helpers: DeepSpeed-Ulysses style all-to-all (swap the sharded dimension between sequence
and heads) and one ring-attention hop (pass the KV block to the next rank)."""
from __future__ import annotations

import torch

from .. import ops


def ulysses_all_to_all(x: torch.Tensor, scatter_dim: int, gather_dim: int, comm=None) -> torch.Tensor:
    """x is sharded along `gather_dim` across the ranks; the result is sharded along `scatter_dim` instead
    (e.g. [S/N, H, D] -> [S, H/N, D] with scatter_dim=1, gather_dim=0)."""
    comm = comm or ops.default_comm()
    n = comm.size
    if n == 1:
        return x
    assert x.shape[scatter_dim] % n == 0
    parts = torch.stack(x.chunk(n, dim=scatter_dim)).contiguous()        # [N, ...]: part p goes to rank p
    out = torch.empty_like(parts)
    ops.all_to_all_single(out, parts, comm=comm)                         # out[p] = what rank p had for me
    return torch.cat(list(out), dim=gather_dim)


def ring_pass(block: torch.Tensor, comm=None, tag=0) -> torch.Tensor:
    """one ring-attention hop: returns the block of the previous rank, hands mine to the next"""
    comm = comm or ops.default_comm()
    out = torch.empty_like(block)
    ops.ring_exchange(block.contiguous(), out, 1, tag, comm)
    return out
