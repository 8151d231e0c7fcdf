#!/bin/bash
# Second two-GPU session of the round: gpurun --gpus 2 --timeout 260 -- 'bash tools/gpu_r2_two_b.sh'
export PYTHONPATH=$PWD PYTHONFAULTHANDLER=1 UCC_HANDLE_ERRORS=bt
O=gpurun_out/r2two_b; mkdir -p $O
T0=$(date +%s); BUDGET=${BUDGET:-225}
left() { echo $(( BUDGET - ($(date +%s) - T0) )); }
cap() { local l=$(left); [ $l -lt $1 ] && echo $l || echo $1; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ARGS="-M cuda -t world,reverse -I 2 -P 2 -i 2 -m 64:4194304:32 -r all -d int32,float32,bfloat16 -o sum,max,avg --triggered 2"
timeout $(cap 100) $TR --master-port 29851 tools/ucc_test_dist.py $ARGS > $O/test_dist_cuda.log 2>&1; echo "test_dist rc=$?"
grep -A6 "TEST REPORT" $O/test_dist_cuda.log | cut -c1-200; grep -B2 -A12 "caught signal\|FAIL" $O/test_dist_cuda.log | head -40 | cut -c1-200
[ $(left) -gt 20 ] && timeout $(cap 40) $TR --master-port 29849 tools/p2p_bench.py > $O/p2p.log 2>&1; echo "p2p rc=$?"; grep '^{' $O/p2p.log
[ $(left) -gt 25 ] && timeout $(cap 45) $TR --master-port 29853 bench.py --gpus 2 --steps 10 --warmup 3 --no-sweep --out $O/bench_ours.json > $O/bench_ours.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$O/bench_ours.json"))
    print("ours busbw", d["busbw_per_gpu_GBps"], "us", d["latency_us"], "ok", d["correct"], "e2e us", d["e2e"]["us_per_step"], d["config"]["algorithm"][-40:])
except Exception as e:
    print("no json", e)
PY
[ $(left) -gt 20 ] && UCC_B200_EXPERIMENTAL_TESTS=1 timeout $(cap 70) python -m pytest tests/test_dist_gpu.py -m gpu -q -p no:cacheprovider -k "parallel_helpers or torch_backend or symm or nvls_pipe" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_sub.log | cut -c1-300
