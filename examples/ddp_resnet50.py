#!/usr/bin/env python3
"""DDP ResNet-50 bf16 on synthetic ImageNet-shaped data (BASELINE.json configuration 5).

    torchrun --nproc-per-node N examples/ddp_resnet50.py --backend ucc    # gradients averaged by ucc_b200 (tl/nvl kernels)
    torchrun --nproc-per-node N examples/ddp_resnet50.py --backend nccl   # torch DDP over NCCL (what the reference's tl_nccl would issue)

Prints one JSON line from rank 0: images/s over all GPUs, device-timed (CUDA events), max over ranks."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="ucc", choices=["ucc", "nccl"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--res", type=int, default=224)
    ap.add_argument("--bucket-mb", type=float, default=25.0)
    a = ap.parse_args()
    from ucc_b200 import ops
    from ucc_b200.dist import init_distributed
    from ucc_b200.models import resnet50
    from ucc_b200.parallel import DistributedDataParallel
    rank, world, _ = init_distributed("cpu:gloo,cuda:nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(0)
    model = resnet50().to(dev).to(memory_format=torch.channels_last)
    if a.backend == "ucc":
        ops.init()
        ddp = DistributedDataParallel(model, bucket_mb=a.bucket_mb)
    else:
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], bucket_cap_mb=a.bucket_mb, gradient_as_bucket_view=True) if world > 1 else model
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
    x = torch.randn(a.batch, 3, a.res, a.res, device=dev).to(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (a.batch,), device=dev)

    def step():
        if a.backend == "ucc":
            ddp.zero_grad()
        else:
            opt.zero_grad(set_to_none=False)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(ddp(x), y)
        loss.backward()
        if a.backend == "ucc":
            ddp.finish_gradient_sync()
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    t = torch.tensor([ms], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "ddp_resnet50_images_per_s", "backend": a.backend, "value": round(a.batch * world / (t.item() / 1e3), 1), "unit": "img/s",
                          "n_gpus": world, "ms_per_step": round(t.item(), 3), "per_gpu_batch": a.batch, "dtype": "bf16 autocast", "data": "synthetic",
                          "loss": float(loss)}), flush=True)
    if a.backend == "ucc":
        ops.shutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
