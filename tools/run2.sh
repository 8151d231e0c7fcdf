export PYTHONPATH=$PWD UCC_HANDLE_ERRORS=bt
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811"
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --out gpurun_out/bench2.json > gpurun_out/bench2.log 2>&1
for nb in 16 64 128; do
  UCC_TL_NVL_MAX_BLOCKS=$nb timeout 200 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-sweep --no-e2e --no-nccl > gpurun_out/bench2_nb$nb.log 2>&1
done
UCC_TL_NVL_NTHREADS=1024 UCC_TL_NVL_MAX_BLOCKS=64 timeout 200 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-sweep --no-e2e --no-nccl > gpurun_out/bench2_nb64_t1024.log 2>&1
UCC_TL_NVL_USE_NVLS=n timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-nccl > gpurun_out/bench2_nonvls.log 2>&1
UCC_TL_NVL_USE_NVLS=n UCC_TL_NVL_MAX_BLOCKS=64 timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-nccl --no-sweep > gpurun_out/bench2_nonvls64.log 2>&1
timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench1.log 2>&1
tail -c 400 gpurun_out/bench2.log
