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


def ring_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, comm=None, causal: bool = False, scale: float | None = None) -> torch.Tensor:
    """Context-parallel attention (ring attention): the sequence is sharded over the ranks, `q, k, v` are this rank's blocks
    [S/N, H, D] (rank r holds positions r*S/N ...).  The K/V blocks travel around the ring with `ops.ring_exchange` (neighbour
    send / recv: tl/nvl p2p kernels on CUDA) while every rank accumulates its queries' attention over the blocks it has seen with
    the online-softmax recurrence (running maximum m, normaliser l, unnormalised output o), so no rank ever holds more than two
    K/V blocks.  Returns this rank's rows of softmax(QK^T * scale [+ causal mask]) V."""
    comm = comm or ops.default_comm()
    n, r = comm.size, comm.rank
    sq, h, d = q.shape
    scale = scale if scale is not None else d ** -0.5
    qf = q.float().transpose(0, 1)                                        # [H, Sq, D]
    m = torch.full((h, sq, 1), float("-inf"), device=q.device)
    l_ = torch.zeros((h, sq, 1), device=q.device)
    o = torch.zeros((h, sq, d), device=q.device)
    kv = torch.stack([k, v]).contiguous()                                 # one message per hop
    for step in range(n):
        src = (r - step) % n                                              # whose block I hold in this step
        if not causal or src <= r:
            kb, vb = kv[0].float().transpose(0, 1), kv[1].float().transpose(0, 1)     # [H, Sk, D]
            s = torch.matmul(qf, kb.transpose(1, 2)) * scale              # [H, Sq, Sk]
            if causal and src == r:
                s = s.masked_fill(torch.ones(sq, kb.shape[1], dtype=torch.bool, device=q.device).triu(1), float("-inf"))
            m_new = torch.maximum(m, s.max(dim=-1, keepdim=True).values)
            p = torch.exp(s - m_new)
            corr = torch.exp(m - m_new)
            l_ = l_ * corr + p.sum(dim=-1, keepdim=True)
            o = o * corr + torch.matmul(p, vb)
            m = m_new
        if step + 1 < n:
            nxt = torch.empty_like(kv)
            ops.ring_exchange(kv, nxt, 1, tag=step, comm=comm)
            kv = nxt
    return (o / l_).transpose(0, 1).to(q.dtype)
