export PYTHONPATH=$PWD UCC_HANDLE_ERRORS=bt
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_nvl_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/nvl_test.log
timeout 300 python -m pytest tests/test_ec_mc.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/ec_test.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811"
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --out gpurun_out/bench2.json > gpurun_out/bench2.log 2>&1
for nb in 64 256; do
  UCC_TL_NVL_MAX_BLOCKS=$nb timeout 200 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-sweep --no-e2e --no-nccl > gpurun_out/bench2_nb$nb.log 2>&1
done
UCC_TL_NVL_USE_NVLS=n timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-nccl > gpurun_out/bench2_nonvls.log 2>&1
UCC_TL_NVL_USE_NVLS=n UCC_TL_NVL_MAX_BLOCKS=256 timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-nccl --no-sweep > gpurun_out/bench2_nonvls256.log 2>&1
tail -c 300 gpurun_out/bench2.log; tail -3 gpurun_out/nvl_test.log gpurun_out/ec_test.log
