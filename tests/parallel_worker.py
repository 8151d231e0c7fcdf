"""Worker under torchrun: DDP / tensor-parallel / MoE helpers of ucc_b200.parallel against single-process references."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ucc_b200 import ops  # noqa: E402
from ucc_b200.dist import init_distributed  # noqa: E402
from ucc_b200.models import MLP, TPTransformerBlock  # noqa: E402
from ucc_b200.parallel import DistributedDataParallel, ZeroRedundancyTrainer, moe_combine, moe_dispatch, ring_pass, ulysses_all_to_all  # noqa: E402


def main():
    use_cuda = len(sys.argv) > 1 and sys.argv[1] == "cuda"
    rank, world, _ = init_distributed("cpu:gloo,cuda:nccl" if use_cuda else "gloo")
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    comm = ops.init()
    ok = True

    def stage(name):   # where a hang happened is the last line printed
        if rank == 0:
            print(f"[parallel_worker] {name}", flush=True)
    stage("ddp")

    # ---- DDP: averaged gradients == gradients of the mean loss over the global batch
    torch.manual_seed(0)
    model = MLP(32, 64, 2, 5).to(dev)
    ref = MLP(32, 64, 2, 5).to(dev)
    ref.load_state_dict(model.state_dict())
    ddp = DistributedDataParallel(model, comm=comm, bucket_mb=0.01)   # tiny buckets: several collectives in flight
    g = torch.Generator().manual_seed(1)
    xs = torch.randn(world, 8, 32, generator=g).to(dev)
    ys = torch.randn(world, 8, 5, generator=g).to(dev)
    for step in range(2):
        ddp.zero_grad()
        loss = torch.nn.functional.mse_loss(ddp(xs[rank]), ys[rank])
        loss.backward()
        ddp.finish_gradient_sync()
        ref.zero_grad()
        sum(torch.nn.functional.mse_loss(ref(xs[r]), ys[r]) for r in range(world)).div(world).backward()
        for (n1, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
            if not torch.allclose(p.grad, q.grad, rtol=1e-4, atol=1e-5):
                print(f"rank {rank}: DDP grad mismatch {n1} step {step}", flush=True)
                ok = False

    stage("tensor parallel")
    # ---- tensor parallel block == dense block assembled from the shards
    torch.manual_seed(100 + rank)
    d, h = 16, 8 * world
    blk = TPTransformerBlock(d, h, comm=comm, device=dev)
    with torch.no_grad():   # same LayerNorm / bias on every rank
        for p in (blk.norm.weight, blk.norm.bias, blk.down.bias):
            ops.broadcast(p.data, 0, comm=comm)
    x = torch.randn(4, d, generator=torch.Generator().manual_seed(7)).to(dev).requires_grad_()
    y = blk(x)
    y.sum().backward()
    w1 = torch.empty(world, h // world, d, device=dev); ops.all_gather_into_tensor(w1, blk.up.weight.data.contiguous(), comm=comm)
    b1 = torch.empty(world, h // world, device=dev); ops.all_gather_into_tensor(b1, blk.up.bias.data.contiguous(), comm=comm)
    w2 = torch.empty(world, d, h // world, device=dev); ops.all_gather_into_tensor(w2, blk.down.weight.data.contiguous(), comm=comm)
    xd = x.detach().clone().requires_grad_()
    hid = torch.nn.functional.gelu(torch.nn.functional.linear(blk.norm(xd), w1.reshape(h, d), b1.reshape(h)))
    yd = xd + torch.nn.functional.linear(hid, torch.cat(list(w2), dim=1), blk.down.bias)
    yd.sum().backward()
    if not torch.allclose(y, yd, rtol=1e-4, atol=1e-5) or not torch.allclose(x.grad, xd.grad, rtol=1e-4, atol=1e-5):
        print(f"rank {rank}: tensor-parallel mismatch", flush=True)
        ok = False

    stage("column parallel gather")
    # ---- ColumnParallelLinear(gather_output=True): the gathered output must carry gradients back to weight, bias and input
    from ucc_b200.parallel.tensor_parallel import ColumnParallelLinear
    torch.manual_seed(200 + rank)
    col = ColumnParallelLinear(d, 4 * world, comm=comm, gather_output=True, device=dev)
    xg = torch.randn(3, d, generator=torch.Generator().manual_seed(9)).to(dev).requires_grad_()
    coef = torch.arange(4 * world, dtype=torch.float32, device=dev) + 1
    (col(xg) * coef).sum().backward()
    wf = torch.empty(world, 4, d, device=dev); ops.all_gather_into_tensor(wf, col.weight.data.contiguous(), comm=comm)
    bf = torch.empty(world, 4, device=dev); ops.all_gather_into_tensor(bf, col.bias.data.contiguous(), comm=comm)
    wd, bd = wf.reshape(4 * world, d).clone().requires_grad_(), bf.reshape(4 * world).clone().requires_grad_()
    xr = xg.detach().clone().requires_grad_()
    (torch.nn.functional.linear(xr, wd, bd) * coef).sum().backward()
    sl = slice(4 * rank, 4 * rank + 4)
    if col.weight.grad is None or not torch.allclose(col.weight.grad, wd.grad[sl], rtol=1e-4, atol=1e-5) or not torch.allclose(col.bias.grad, bd.grad[sl], rtol=1e-4, atol=1e-5) \
            or not torch.allclose(xg.grad, xr.grad, rtol=1e-4, atol=1e-5):
        print(f"rank {rank}: ColumnParallelLinear(gather_output) gradient mismatch", flush=True)
        ok = False

    stage("sequence-parallel tensor parallelism (allgather in, reduce_scatter out)")
    # ---- Megatron-SP MLP: sequence-sharded input -> ColumnParallel(sequence_parallel) -> gelu -> RowParallel(sequence_parallel) ->
    #      sequence-sharded output; output and input gradient must equal the dense MLP's rows of this rank
    from ucc_b200.parallel.tensor_parallel import RowParallelLinear
    torch.manual_seed(300 + rank)
    sq, hd = 3, 4 * world
    up = ColumnParallelLinear(d, hd, comm=comm, device=dev, sequence_parallel=True)
    down = RowParallelLinear(hd, d, comm=comm, device=dev, sequence_parallel=True)
    with torch.no_grad():
        ops.broadcast(down.bias.data, 0, comm=comm)
    xfull = torch.randn(sq * world, d, generator=torch.Generator().manual_seed(13)).to(dev)
    xs_ = xfull[rank * sq:(rank + 1) * sq].clone().requires_grad_()
    ys_ = down(torch.nn.functional.gelu(up(xs_)))
    wgt = torch.arange(sq * world * d, dtype=torch.float32, device=dev).view(sq * world, d) / 100
    (ys_ * wgt[rank * sq:(rank + 1) * sq]).sum().backward()
    w1 = torch.empty(world, hd // world, d, device=dev); ops.all_gather_into_tensor(w1, up.weight.data.contiguous(), comm=comm)
    b1 = torch.empty(world, hd // world, device=dev); ops.all_gather_into_tensor(b1, up.bias.data.contiguous(), comm=comm)
    w2 = torch.empty(world, d, hd // world, device=dev); ops.all_gather_into_tensor(w2, down.weight.data.contiguous(), comm=comm)
    xd = xfull.clone().requires_grad_()
    yd = torch.nn.functional.linear(torch.nn.functional.gelu(torch.nn.functional.linear(xd, w1.reshape(hd, d), b1.reshape(hd))), torch.cat(list(w2), dim=1), down.bias)
    (yd * wgt).sum().backward()
    if not torch.allclose(ys_, yd[rank * sq:(rank + 1) * sq], rtol=1e-4, atol=1e-5) or not torch.allclose(xs_.grad, xd.grad[rank * sq:(rank + 1) * sq], rtol=1e-4, atol=1e-5):
        print(f"rank {rank}: sequence-parallel TP mismatch", flush=True)
        ok = False

    stage("moe")
    # ---- MoE dispatch / combine round trip with skewed routing
    T, H = 50 + 7 * rank, 12
    tok = torch.arange(T * H, dtype=torch.float32, device=dev).view(T, H) + 10000 * rank
    dest = (torch.arange(T, device=dev) * (rank + 2)) % world
    if world > 1:
        dest[: T // 2] = 0          # skew: rank 0's expert is hot
    recv, sc, rc, order = moe_dispatch(tok, dest, comm=comm)
    if recv.shape[0] != sum(rc):
        ok = False
    out = moe_combine(recv * 2.0, sc, rc, order, comm=comm)
    if not torch.equal(out, tok * 2.0):
        print(f"rank {rank}: MoE round trip mismatch", flush=True)
        ok = False

    stage("zero")
    # ---- ZeRO-1: sharded SGD step == plain SGD on the averaged gradient
    torch.manual_seed(3)
    zm, rm = MLP(16, 24, 2, 3).to(dev), MLP(16, 24, 2, 3).to(dev)
    rm.load_state_dict(zm.state_dict())
    zt = ZeroRedundancyTrainer(zm, torch.optim.SGD, comm=comm, lr=0.1, momentum=0.9)
    ropt = torch.optim.SGD(rm.parameters(), lr=0.1, momentum=0.9)
    for step in range(3):
        zt.zero_grad()
        torch.nn.functional.mse_loss(zm(xs[rank][:, :16]), ys[rank][:, :3]).backward()
        zt.step()
        ropt.zero_grad()
        sum(torch.nn.functional.mse_loss(rm(xs[r][:, :16]), ys[r][:, :3]) for r in range(world)).div(world).backward()
        ropt.step()
    for (n1, p), (_, q) in zip(zm.named_parameters(), rm.named_parameters()):
        if not torch.allclose(p, q, rtol=1e-4, atol=1e-5):
            print(f"rank {rank}: ZeRO parameter mismatch {n1}", flush=True)
            ok = False

    stage("fsdp (ZeRO-3 units: allgather before forward / backward, reduce_scatter of the gradients)")
    from ucc_b200.parallel import FullyShardedModule
    torch.manual_seed(21)
    blocks = [torch.nn.Sequential(torch.nn.Linear(20, 20), torch.nn.Tanh()) for _ in range(3)]
    ref_model = torch.nn.Sequential(*[torch.nn.Sequential(torch.nn.Linear(20, 20), torch.nn.Tanh()) for _ in range(3)]).to(dev)
    ref_model.load_state_dict(torch.nn.Sequential(*blocks).state_dict())
    units = [FullyShardedModule(b.to(dev), comm=comm) for b in blocks]
    fsdp = torch.nn.Sequential(*units)
    opt = torch.optim.SGD([u.shard_param for u in units], lr=0.1)
    ropt = torch.optim.SGD(ref_model.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(22)
    for step in range(3):
        xb = torch.randn(world, 4, 20, generator=g).to(dev)
        yb = torch.randn(world, 4, 20, generator=g).to(dev)
        opt.zero_grad()
        torch.nn.functional.mse_loss(fsdp(xb[rank]), yb[rank]).backward()
        opt.step()
        ropt.zero_grad()
        torch.nn.functional.mse_loss(ref_model(xb.reshape(-1, 20)), yb.reshape(-1, 20)).backward()   # mean over the global batch
        ropt.step()
    for u, rb in zip(units, ref_model):
        want = torch.cat([p.detach().reshape(-1) for p in rb.parameters()])
        if not torch.allclose(u.full_parameters(), want, rtol=1e-4, atol=1e-6):
            print(f"rank {rank}: FSDP parameters differ from the reference after 3 steps", flush=True)
            ok = False
        if u._full.untyped_storage().nbytes() != 0:
            print(f"rank {rank}: FSDP unit kept its full parameters resident", flush=True)
            ok = False

    stage("ulysses")
    # ---- Ulysses all-to-all: [S/N, H, D] -> [S, H/N, D], and one ring-attention hop
    S, H, D = 4 * world, 2 * world, 3
    full = torch.arange(S * H * D, dtype=torch.float32).view(S, H, D).to(dev)
    mine = full[rank * (S // world):(rank + 1) * (S // world)].contiguous()
    got = ulysses_all_to_all(mine, scatter_dim=1, gather_dim=0, comm=comm)
    if not torch.equal(got, full[:, rank * (H // world):(rank + 1) * (H // world)]):
        print(f"rank {rank}: ulysses mismatch", flush=True)
        ok = False
    stage("ring pass (send / recv)")
    kv = torch.full((5, 7), float(rank), device=dev)
    prev = ring_pass(kv, comm=comm)
    ok &= bool((prev == float((rank - 1) % world)).all())

    stage("ring attention (context parallel)")
    # ---- ring attention: every rank's rows of softmax(QK^T) V over the sharded sequence == full attention on the gathered sequence
    from ucc_b200.parallel import ring_attention
    Sb, Hh, Dd = 6, 3, 8
    g = torch.Generator().manual_seed(11)
    Q, K, V = (torch.randn(Sb * world, Hh, Dd, generator=g).to(dev) for _ in range(3))
    sl = slice(rank * Sb, (rank + 1) * Sb)
    for causal in (False, True):
        got = ring_attention(Q[sl].contiguous(), K[sl].contiguous(), V[sl].contiguous(), comm=comm, causal=causal)
        att = torch.einsum("qhd,khd->hqk", Q, K) * Dd ** -0.5
        if causal:
            att = att.masked_fill(torch.ones(Sb * world, Sb * world, dtype=torch.bool, device=dev).triu(1), float("-inf"))
        exp = torch.einsum("hqk,khd->qhd", torch.softmax(att, dim=-1), V)[sl]
        if not torch.allclose(got, exp, rtol=1e-4, atol=1e-5):
            print(f"rank {rank}: ring attention mismatch (causal={causal}) max err {(got - exp).abs().max().item()}", flush=True)
            ok = False

    stage("pipeline (GPipe / 1F1B over send / recv)")
    # ---- pipeline parallel: every rank owns two layers of a 2 * world layer MLP; loss and gradients must equal the unsplit model's
    from ucc_b200.parallel import PipelineStage
    torch.manual_seed(7)
    width, n_micro, mb = 24, 5, 6
    layers = [torch.nn.Sequential(torch.nn.Linear(width, width), torch.nn.Tanh()) for _ in range(2 * world)]
    full = torch.nn.Sequential(*layers).to(dev)
    g = torch.Generator().manual_seed(8)
    xs = [torch.randn(mb, width, generator=g).to(dev) for _ in range(n_micro)]
    ys = [torch.randn(mb, width, generator=g).to(dev) for _ in range(n_micro)]
    lf = torch.nn.functional.mse_loss
    ref_loss = sum(lf(full(xs[i]), ys[i]) for i in range(n_micro)) / n_micro
    ref_loss.backward()
    ref_grads = [p.grad.clone() for lyr in layers[2 * rank:2 * rank + 2] for p in lyr.parameters()]
    mine = torch.nn.Sequential(*layers[2 * rank:2 * rank + 2])
    for sched in ("gpipe", "1f1b"):
        for p in mine.parameters():
            p.grad = None
        pp = PipelineStage(mine, act_shape=(mb, width), comm=comm, device=dev)
        loss = pp.run(n_micro, inputs=xs, targets=ys, loss_fn=lf, schedule=sched)
        if rank == world - 1 and not torch.allclose(loss, ref_loss.detach(), rtol=1e-5, atol=1e-6):
            print(f"rank {rank}: pipeline {sched} loss {loss.item()} != {ref_loss.item()}", flush=True)
            ok = False
        for p, rg in zip(mine.parameters(), ref_grads):
            if p.grad is None or not torch.allclose(p.grad, rg, rtol=1e-4, atol=1e-6):
                print(f"rank {rank}: pipeline {sched} gradient mismatch", flush=True)
                ok = False

    stage("done")
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ops.shutdown()
    if rank == 0:
        print("PARALLEL_WORKER_OK" if flag.item() == 1 else "PARALLEL_WORKER_FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
