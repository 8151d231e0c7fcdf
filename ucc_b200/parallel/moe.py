"""Expert-parallel token exchange: dispatch / combine as alltoallv with per-rank skewed counts.  On CUDA the
tl/nvl exchange kernel publishes each rank's offset table on the device, so only the counts travel through the host."""
from __future__ import annotations

import torch

from .. import ops


def moe_dispatch(tokens: torch.Tensor, dest_rank: torch.Tensor, comm=None):
    """tokens [T, H] and their destination expert rank [T] -> (received tokens, send_counts, recv_counts, order)."""
    comm = comm or ops.default_comm()
    n = comm.size
    order = torch.argsort(dest_rank, stable=True)
    send = tokens[order].contiguous()
    send_counts = torch.bincount(dest_rank, minlength=n).to(torch.int64)
    recv_counts = torch.empty_like(send_counts)
    ops.all_to_all_single(recv_counts, send_counts, comm=comm)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    recv = torch.empty((sum(rc), tokens.shape[1]), dtype=tokens.dtype, device=tokens.device)
    ops.all_to_all_single(recv, send, rc, sc, comm=comm)
    return recv, sc, rc, order


def moe_combine(expert_out: torch.Tensor, send_counts, recv_counts, order, comm=None):
    """inverse of moe_dispatch: results travel back and are restored to the original token order."""
    comm = comm or ops.default_comm()
    back = torch.empty((sum(send_counts), expert_out.shape[1]), dtype=expert_out.dtype, device=expert_out.device)
    ops.all_to_all_single(back, expert_out.contiguous(), send_counts, recv_counts, comm=comm)
    out = torch.empty_like(back)
    out[order] = back
    return out
