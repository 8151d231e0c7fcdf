"""Worker run under torchrun by test_dist_*.py: multi-process collectives through ucc_b200.dist."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ucc_b200 import capi as U  # noqa: E402
from ucc_b200.dist import Communicator, init_distributed  # noqa: E402


def main():
    if os.environ.get("DW_FAULT"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["DW_FAULT"]), exit=False)
    use_cuda = len(sys.argv) > 1 and sys.argv[1] == "cuda"
    rank, world, lrank = init_distributed("cpu:gloo,cuda:nccl" if use_cuda else "gloo")
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    comm = Communicator()
    ok = True
    for count in (1, 1000, 100000, 3000001):
        for dt in (torch.float32, torch.bfloat16, torch.int32):
            g = torch.Generator().manual_seed(1234 + count)
            alls = [(torch.rand(count, generator=g) * 4 + r).to(dt) for r in range(world)]
            src = alls[rank].to(dev)
            dst = torch.zeros(count, dtype=dt, device=dev)
            req = comm.allreduce_init(src, dst)
            if use_cuda:
                req.post_on_stream()
            else:
                req.post()
            req.wait()
            req.finalize()
            if use_cuda:
                torch.cuda.synchronize()
            exp = sum(a.double() for a in alls)
            tol = 2e-2 * world if dt == torch.bfloat16 else 1e-5
            if not torch.allclose(dst.cpu().double(), exp, rtol=tol, atol=tol):
                print(f"rank {rank}: allreduce mismatch count {count} dt {dt}", flush=True)
                ok = False
    # allgather + alltoall + bcast + reduce_scatter: small (staged kernels) and large (zero-copy kernels on CUDA)
    for blk in (1000, 300000):
        ok &= other_colls(comm, rank, world, dev, use_cuda, blk)
    if os.environ.get("DW_MEMH") == "1":
        ok &= registered_alltoall(comm, rank, world, dev)
    if not use_cuda:   # (validated on host memory; on GPUs the WORLD team below is what the 2/4/8-GPU sessions exercised)
        ok &= team_kinds(rank, world, dev, use_cuda)
    if use_cuda and os.environ.get("DW_ASYM") == "1":
        ok &= asymmetric_root(comm, rank, world, dev)
    comm.barrier()
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    comm.destroy()
    if rank == 0:
        print("DIST_WORKER_OK" if flag.item() == 1 else "DIST_WORKER_FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


def asymmetric_root(comm, rank, world, dev):
    """reference test/gtest/asym_mem: at the ROOT src and dst live in different memory types (the core stages through a
    temporary of the other type around the TL collective).  reduce: CUDA contributions, the root wants the result on the HOST;
    scatter: the root's source is on the HOST, every destination is CUDA."""
    from ucc_b200.harness import coll_args
    ok, count, root = True, 5000, world - 1
    src = torch.arange(count, dtype=torch.float32, device=dev) + rank
    host_dst = torch.zeros(count, dtype=torch.float32)
    if rank == root:
        a = coll_args("reduce", dt="float32", root=root, src_ptr=src.data_ptr(), dst_ptr=host_dst.data_ptr(), count_src=count, count_dst=count,
                      src_mem_type=U.UCC_MEMORY_TYPE_CUDA, dst_mem_type=U.UCC_MEMORY_TYPE_HOST)
    else:
        a = coll_args("reduce", dt="float32", root=root, src_ptr=src.data_ptr(), count_src=count, count_dst=0, mem_type=U.UCC_MEMORY_TYPE_CUDA)
    r = comm.init(a, (src, host_dst))
    r.post(); r.wait(); r.finalize()
    torch.cuda.synchronize()
    if rank == root:
        exp = torch.arange(count, dtype=torch.float32) * world + sum(range(world))
        if not torch.allclose(host_dst, exp):
            print(f"rank {rank}: asymmetric reduce mismatch", flush=True)
            ok = False
    blk = 700
    host_src = torch.arange(blk * world, dtype=torch.float32) * 3
    dst = torch.zeros(blk, dtype=torch.float32, device=dev)
    if rank == 0:
        a = coll_args("scatter", dt="float32", root=0, src_ptr=host_src.data_ptr(), dst_ptr=dst.data_ptr(), count_src=blk * world, count_dst=blk,
                      src_mem_type=U.UCC_MEMORY_TYPE_HOST, dst_mem_type=U.UCC_MEMORY_TYPE_CUDA)
    else:
        a = coll_args("scatter", dt="float32", root=0, dst_ptr=dst.data_ptr(), count_dst=blk, count_src=0, mem_type=U.UCC_MEMORY_TYPE_CUDA)
    r = comm.init(a, (host_src, dst))
    r.post(); r.wait(); r.finalize()
    torch.cuda.synchronize()
    if not torch.equal(dst.cpu(), host_src[rank * blk:(rank + 1) * blk]):
        print(f"rank {rank}: asymmetric scatter mismatch", flush=True)
        ok = False
    return ok


def registered_alltoall(comm, rank, world, dev):
    """Communicator.register (ucc_mem_map export + OOB exchange + import) and an alltoall whose destination is part of the
    registered segment: with UCC_TL_SHM_TUNE=alltoall:@onesided the blocks are written straight into the peers' memory
    (process_vm_writev between processes) - no address exchange."""
    blk, pad = 5000, 128
    seg = torch.full((pad + blk * world + 7,), -1, dtype=torch.int32, device=dev)
    h = comm.register(seg)
    ok = True
    for it in range(3):
        g = torch.Generator().manual_seed(77 + it)
        alls = [torch.randint(0, 1 << 30, (blk * world,), generator=g, dtype=torch.int32) for _ in range(world)]
        seg.fill_(-1)
        comm.barrier()                                      # one-sided contract: every destination is ready before anybody writes
        dst = seg[pad:pad + blk * world]
        req = comm.coll_init("alltoall", alls[rank].to(dev), dst, dst_memh=h)
        req.post(); req.wait(); req.finalize()
        exp = torch.cat([alls[p][rank * blk:(rank + 1) * blk] for p in range(world)])
        if not torch.equal(dst.cpu(), exp) or not bool((seg[:pad] == -1).all()) or not bool((seg[pad + blk * world:] == -1).all()):
            print(f"rank {rank}: registered alltoall mismatch (iteration {it})", flush=True)
            ok = False
        comm.barrier()
    h.close()
    return ok


def team_kinds(rank, world, dev, use_cuda):
    """Teams other than WORLD (reference test/mpi team kinds: half, odd_even, reverse), each checked against an oracle."""
    ok = True
    if world < 3:
        return ok
    kinds = {"half": list(range((world + 1) // 2)), "odd_even": [r for r in range(world) if r % 2 == rank % 2]}
    groups = {"half": dist.new_group(kinds["half"], backend="gloo")}
    even, odd = dist.new_group([r for r in range(world) if r % 2 == 0], backend="gloo"), dist.new_group([r for r in range(world) if r % 2 == 1], backend="gloo")
    groups["odd_even"] = even if rank % 2 == 0 else odd
    for kind, members in kinds.items():
        if rank not in members or len(members) < 2:
            continue
        c = Communicator(groups[kind])
        x = torch.full((257,), float(rank + 1), device=dev)
        c.run(c.allreduce_init(x, x))
        if use_cuda:
            torch.cuda.synchronize()
        ok &= bool((x.cpu() == float(sum(m + 1 for m in members))).all())
        g = torch.zeros(len(members) * 3, device=dev)
        c.run(c.coll_init("allgather", torch.full((3,), float(rank), device=dev), g))
        if use_cuda:
            torch.cuda.synchronize()
        ok &= bool(torch.equal(g.view(len(members), 3)[:, 0].cpu(), torch.tensor([float(m) for m in members])))
        c.destroy()
    # reverse: every process takes part, UCC rank = world - 1 - torch rank
    perm = [world - 1 - g for g in range(world)]
    c = Communicator(perm=perm)
    g = torch.zeros(world * 2, device=dev)
    c.run(c.coll_init("allgather", torch.full((2,), float(rank), device=dev), g))
    b = torch.full((64,), 5.0 if c.rank == 0 else 0.0, device=dev)     # UCC root 0 is torch rank world-1
    c.run(c.coll_init("bcast", b, None, root=0))
    if use_cuda:
        torch.cuda.synchronize()
    ok &= bool(torch.equal(g.view(world, 2)[:, 0].cpu(), torch.tensor([float(world - 1 - u) for u in range(world)])))
    ok &= bool((b.cpu() == 5.0).all())
    c.destroy()
    if not ok:
        print(f"rank {rank}: team kinds mismatch", flush=True)
    return ok


def other_colls(comm, rank, world, dev, use_cuda, blk):
    ok = True
    src = torch.full((blk,), float(rank), device=dev)
    dst = torch.zeros(blk * world, device=dev)
    comm.run(comm.coll_init("allgather", src, dst))
    if use_cuda:
        torch.cuda.synchronize()
    ok &= bool((dst.view(world, blk).cpu() == torch.arange(world, dtype=torch.float32)[:, None]).all())
    src = torch.arange(blk * world, dtype=torch.float32, device=dev) + 1000 * rank
    dst = torch.zeros(blk * world, device=dev)
    comm.run(comm.coll_init("alltoall", src, dst))
    exp = torch.cat([torch.arange(rank * blk, (rank + 1) * blk, dtype=torch.float32) + 1000 * p for p in range(world)])
    ok &= bool(torch.equal(dst.cpu(), exp))
    b = torch.full((5 * blk,), 7.0 if rank == 0 else 0.0, device=dev)
    comm.run(comm.coll_init("bcast", b, None, root=0))
    ok &= bool((b.cpu() == 7.0).all())
    src = torch.ones(blk * world, device=dev) * (rank + 1)
    dst = torch.zeros(blk, device=dev)
    comm.run(comm.coll_init("reduce_scatter", src, dst))
    ok &= bool((dst.cpu() == world * (world + 1) / 2).all())
    # in-place allreduce + reduce to a non-zero root
    x = torch.full((blk * 3 + 5,), float(rank + 1), device=dev)
    comm.run(comm.allreduce_init(x, x))
    ok &= bool((x.cpu() == world * (world + 1) / 2).all())
    if not ok:
        print(f"rank {rank}: collective mismatch at blk {blk}", flush=True)
    return ok


if __name__ == "__main__":
    main()
