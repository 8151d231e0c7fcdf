#!/usr/bin/env python3
"""Context parallelism + pipeline parallelism on the library's send / recv, in one small script.

    torchrun --nproc-per-node N examples/long_context_pipeline.py [--device cuda]

1. ring attention: a sequence of N * S tokens is sharded over the N ranks; K/V blocks travel around the ring
   (`ucc_b200.parallel.ring_attention`) and the result is checked against attention over the gathered sequence.
2. pipeline: a 2N-layer MLP is cut into N stages (`ucc_b200.parallel.PipelineStage`, 1F1B schedule); one optimisation step
   is checked against the unsplit model.
Runs on host tensors (tl/shm) as well as on GPUs (tl/nvl p2p kernels: eager ring below 1 MB, rendezvous above)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--seq-per-rank", type=int, default=128)
    ap.add_argument("--heads", type=int, default=4)
    ap.add_argument("--head-dim", type=int, default=32)
    a = ap.parse_args()
    from ucc_b200 import ops
    from ucc_b200.dist import init_distributed
    from ucc_b200.parallel import PipelineStage, ring_attention
    rank, world, _ = init_distributed("cpu:gloo,cuda:nccl" if a.device == "cuda" else "gloo")
    dev = torch.device("cuda", torch.cuda.current_device()) if a.device == "cuda" else torch.device("cpu")
    comm = ops.init()
    # ---- 1. ring attention
    g = torch.Generator().manual_seed(0)
    S, H, D = a.seq_per_rank, a.heads, a.head_dim
    Q, K, V = (torch.randn(S * world, H, D, generator=g).to(dev) for _ in range(3))
    sl = slice(rank * S, (rank + 1) * S)
    out = ring_attention(Q[sl].contiguous(), K[sl].contiguous(), V[sl].contiguous(), comm=comm, causal=True)
    att = torch.einsum("qhd,khd->hqk", Q, K) * D ** -0.5
    att = att.masked_fill(torch.ones(S * world, S * world, dtype=torch.bool, device=dev).triu(1), float("-inf"))
    ref = torch.einsum("hqk,khd->qhd", torch.softmax(att, dim=-1), V)[sl]
    err_att = (out - ref).abs().max().item()
    # ---- 2. pipeline (1F1B)
    torch.manual_seed(1)
    width, n_micro, mb = 64, 8, 16
    layers = [torch.nn.Sequential(torch.nn.Linear(width, width), torch.nn.GELU()) for _ in range(2 * world)]
    full = torch.nn.Sequential(*layers).to(dev)
    xs = [torch.randn(mb, width, generator=g).to(dev) for _ in range(n_micro)]
    ys = [torch.randn(mb, width, generator=g).to(dev) for _ in range(n_micro)]
    lf = torch.nn.functional.mse_loss
    (sum(lf(full(xs[i]), ys[i]) for i in range(n_micro)) / n_micro).backward()
    mine = torch.nn.Sequential(*layers[2 * rank:2 * rank + 2])
    ref_grads = [p.grad.clone() for p in mine.parameters()]
    for p in mine.parameters():
        p.grad = None
    loss = PipelineStage(mine, act_shape=(mb, width), comm=comm, device=dev).run(n_micro, inputs=xs, targets=ys, loss_fn=lf, schedule="1f1b")
    err_pp = max((p.grad - r).abs().max().item() for p, r in zip(mine.parameters(), ref_grads))
    errs = torch.tensor([err_att, err_pp], dtype=torch.float64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    if rank == world - 1:
        print(f"pipeline loss {loss.item():.6f}", flush=True)
    if rank == 0:
        print(f"ring attention max |err| {errs[0].item():.2e}, pipeline gradient max |err| {errs[1].item():.2e} on {world} ranks ({a.device})", flush=True)
        print("EXAMPLE_OK" if errs[0] < 1e-4 and errs[1] < 1e-4 else "EXAMPLE_FAIL", flush=True)
    ops.shutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
