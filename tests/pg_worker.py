"""Worker under torchrun: ucc_b200 as the torch.distributed backend (role of ProcessGroupUCC), incl. torch DDP on top."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ucc_b200.torch_backend  # noqa: F401,E402


def main():
    use_cuda = len(sys.argv) > 1 and sys.argv[1] == "cuda"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if use_cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)) % torch.cuda.device_count())
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    dist.init_process_group("ucc_b200", rank=rank, world_size=world)
    ok = True
    x = torch.full((1000,), float(rank + 1), device=dev)
    dist.all_reduce(x)
    ok &= bool((x == world * (world + 1) / 2).all())
    y = torch.full((5,), float(rank), device=dev)
    dist.all_reduce(y, op=dist.ReduceOp.MAX)
    ok &= bool((y == world - 1).all())
    b = torch.arange(10, dtype=torch.float32, device=dev) * (1 if rank == 1 % world else 0)
    dist.broadcast(b, src=1 % world)
    ok &= bool(torch.equal(b.cpu(), torch.arange(10, dtype=torch.float32)))
    outs = [torch.zeros(3, device=dev) for _ in range(world)]
    dist.all_gather(outs, torch.full((3,), float(rank), device=dev))
    ok &= all(bool((outs[r] == r).all()) for r in range(world))
    big = torch.zeros(world * 4, device=dev)
    dist.all_gather_into_tensor(big, torch.full((4,), float(rank), device=dev))
    ok &= bool(torch.equal(big.view(world, 4)[:, 0].cpu(), torch.arange(world, dtype=torch.float32)))
    rs = torch.zeros(4, device=dev)
    dist.reduce_scatter_tensor(rs, torch.ones(world * 4, device=dev) * (rank + 1))
    ok &= bool((rs == world * (world + 1) / 2).all())
    a2a = torch.zeros(world * 2, device=dev)
    dist.all_to_all_single(a2a, torch.arange(world * 2, dtype=torch.float32, device=dev) + 100 * rank)
    exp = torch.cat([torch.arange(rank * 2, rank * 2 + 2, dtype=torch.float32) + 100 * p for p in range(world)])
    ok &= bool(torch.equal(a2a.cpu(), exp))
    dist.barrier()
    # point to point (active-set bcast underneath): ring shift
    nxt, prv = (rank + 1) % world, (rank - 1 + world) % world
    got = torch.zeros(7, device=dev)
    if rank % 2 == 0:
        dist.send(torch.full((7,), float(rank), device=dev), nxt)
        dist.recv(got, prv)
    else:
        dist.recv(got, prv)
        dist.send(torch.full((7,), float(rank), device=dev), nxt)
    if world % 2 == 0 or rank not in (0, world - 1):     # (odd rings would need non-blocking p2p to avoid the wrap-around wait)
        ok &= bool((got == float(prv)).all())
    # batch_isend_irecv with the RECEIVE listed first on every rank (torch issues the ops one after the other): works on host tensors
    # (tag-matched, non-blocking transport) and on CUDA tensors (the backend keeps one internal stream per peer and direction)
    big, small = torch.zeros(70001, device=dev), torch.zeros(5, device=dev)
    msg_b, msg_s = torch.full((70001,), float(rank + 1), device=dev), torch.full((5,), float(rank + 10), device=dev)
    ops_ = [dist.P2POp(dist.irecv, big, prv), dist.P2POp(dist.irecv, small, prv), dist.P2POp(dist.isend, msg_b, nxt), dist.P2POp(dist.isend, msg_s, nxt)]
    for w in dist.batch_isend_irecv(ops_):
        w.wait()
    if use_cuda:
        torch.cuda.synchronize()
    ok &= bool((big == float(prv + 1)).all()) and bool((small == float(prv + 10)).all())
    # torch's own DDP on top of the backend
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4)).to(dev)
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index] if use_cuda else None)
    g = torch.Generator().manual_seed(1)
    xs, ys = torch.randn(world, 8, 16, generator=g).to(dev), torch.randn(world, 8, 4, generator=g).to(dev)
    torch.nn.functional.mse_loss(ddp(xs[rank]), ys[rank]).backward()
    ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4)).to(dev)
    ref.load_state_dict(model.state_dict())
    sum(torch.nn.functional.mse_loss(ref(xs[r]), ys[r]) for r in range(world)).div(world).backward()
    for p, q in zip(model.parameters(), ref.parameters()):
        ok &= bool(torch.allclose(p.grad, q.grad, rtol=1e-4, atol=1e-5))
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if use_cuda:
        torch.cuda.synchronize()
    if rank == 0:
        print("PG_WORKER_OK" if flag.item() == 1 else "PG_WORKER_FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
