export PYTHONPATH=$PWD UCC_HANDLE_ERRORS=bt
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nvl_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/nvl_test.log
timeout 900 python -m pytest tests/test_dist_gpu.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/dist_test.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811"
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --out gpurun_out/bench${N}_zc.json > gpurun_out/bench${N}_zc.log 2>&1
UCC_TL_NVL_ZCOPY_THRESH=0 timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-nccl > gpurun_out/bench${N}_zc0.log 2>&1
UCC_TL_NVL_ZCOPY=n timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-nccl > gpurun_out/bench${N}_nozc.log 2>&1
cat gpurun_out/nvl_test.log | tail -3; tail -5 gpurun_out/dist_test.log; tail -c 300 gpurun_out/bench${N}_zc.log
