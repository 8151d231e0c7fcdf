export PYTHONPATH=$PWD UCC_HANDLE_ERRORS=bt
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nvl_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/nvl_test.log
timeout 600 python -m pytest tests/test_dist_gpu.py -x -q -m gpu -k "all_gpus or zcopy_forced or nvls_everything or torch_backend" 2>&1 | tail -30 > gpurun_out/dist_test.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811"
timeout 300 $TR tools/coll_bench.py > gpurun_out/coll${N}_v3.log 2>&1
tail -3 gpurun_out/nvl_test.log; tail -4 gpurun_out/dist_test.log | cut -c1-300; tail -n 1 gpurun_out/coll${N}_v3.log | cut -c1-1200
