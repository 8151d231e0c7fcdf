export PYTHONPATH=$PWD UCC_HANDLE_ERRORS=bt
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811"
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --out gpurun_out/bench${N}_final.json > gpurun_out/bench${N}_final.log 2>&1
timeout 200 $TR tools/coll_bench.py > gpurun_out/coll${N}_final.log 2>&1
tail -c 300 gpurun_out/bench${N}_final.log; tail -n 1 gpurun_out/coll${N}_final.log | cut -c1-300
