export PYTHONPATH=$PWD UCC_HANDLE_ERRORS=bt
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811"
timeout 300 $TR tests/dist_worker.py cuda > gpurun_out/dist$N.log 2>&1; echo "dist rc=$?" >> gpurun_out/dist$N.log
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --out gpurun_out/bench$N.json > gpurun_out/bench$N.log 2>&1
UCC_TL_NVL_USE_NVLS=n timeout 200 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-nccl > gpurun_out/bench${N}_nonvls.log 2>&1
tail -c 400 gpurun_out/dist$N.log; tail -c 600 gpurun_out/bench$N.log
