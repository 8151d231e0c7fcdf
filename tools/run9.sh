export PYTHONPATH=$PWD UCC_HANDLE_ERRORS=bt
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811"
timeout 200 $TR tests/dist_worker.py cuda > gpurun_out/dist$N.log 2>&1; echo "dist rc=$?" >> gpurun_out/dist$N.log
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --out gpurun_out/bench${N}_v3.json > gpurun_out/bench${N}_v3.log 2>&1
timeout 300 $TR tools/coll_bench.py > gpurun_out/coll$N.log 2>&1
tail -2 gpurun_out/dist$N.log; tail -c 400 gpurun_out/bench${N}_v3.log; tail -n 2 gpurun_out/coll$N.log | cut -c1-1500
