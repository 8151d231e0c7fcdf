#!/bin/bash
# Performance session on N GPUs:  gpurun --gpus N -- 'bash tools/gpu_bench.sh N'
# allreduce sweep with the NCCL comparison, the other collectives, DDP ResNet-50 with both gradient-averaging paths
export PYTHONPATH=$PWD UCC_HANDLE_ERRORS=bt
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = 1 ]; then TR=python; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811"; fi
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --out gpurun_out/bench$N.json > gpurun_out/bench$N.log 2>&1
if [ "$N" != 1 ]; then
  timeout 300 $TR tools/coll_bench.py > gpurun_out/coll$N.log 2>&1
  timeout 200 $TR examples/ddp_resnet50.py --backend ucc --steps 15 --warmup 5 > gpurun_out/ddp${N}_ucc.log 2>&1
  timeout 200 $TR examples/ddp_resnet50.py --backend nccl --steps 15 --warmup 5 > gpurun_out/ddp${N}_nccl.log 2>&1
fi
tail -c 400 gpurun_out/bench$N.log; tail -n 1 gpurun_out/coll$N.log 2>/dev/null | cut -c1-400; tail -n 1 gpurun_out/ddp${N}_ucc.log gpurun_out/ddp${N}_nccl.log 2>/dev/null
