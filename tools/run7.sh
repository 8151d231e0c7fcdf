export PYTHONPATH=$PWD UCC_HANDLE_ERRORS=bt
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811"
timeout 200 $TR tests/dist_worker.py cuda > gpurun_out/dist$N.log 2>&1; echo "dist rc=$?" >> gpurun_out/dist$N.log
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --out gpurun_out/bench${N}_v2.json > gpurun_out/bench${N}_v2.log 2>&1
UCC_TL_NVL_TUNE=allreduce:cuda:1M-inf:@twoshot timeout 200 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-nccl > gpurun_out/bench${N}_p2pzc.log 2>&1
UCC_TL_NVL_ZCOPY=n timeout 200 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-nccl --no-sweep > gpurun_out/bench${N}_nvls_staged.log 2>&1
timeout 300 $TR tools/coll_bench.py > gpurun_out/coll$N.log 2>&1
timeout 200 $TR examples/ddp_resnet50.py --backend ucc --steps 15 --warmup 5 > gpurun_out/ddp${N}_ucc.log 2>&1
timeout 200 $TR examples/ddp_resnet50.py --backend nccl --steps 15 --warmup 5 > gpurun_out/ddp${N}_nccl.log 2>&1
tail -2 gpurun_out/dist$N.log; tail -c 300 gpurun_out/bench${N}_v2.log; tail -n 2 gpurun_out/coll$N.log | cut -c1-300; tail -n 1 gpurun_out/ddp${N}_ucc.log gpurun_out/ddp${N}_nccl.log
