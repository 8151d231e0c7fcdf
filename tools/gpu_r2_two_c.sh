#!/bin/bash
# Short two-GPU confirmation: gpurun --gpus 2 --timeout 110 -- 'bash tools/gpu_r2_two_c.sh'
export PYTHONPATH=$PWD PYTHONFAULTHANDLER=1 UCC_HANDLE_ERRORS=bt
O=gpurun_out/r2two_c; mkdir -p $O
T0=$(date +%s); BUDGET=${BUDGET:-90}
left() { echo $(( BUDGET - ($(date +%s) - T0) )); }
cap() { local l=$(left); [ $l -lt $1 ] && echo $l || echo $1; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ARGS="-M cuda -t world,reverse -I 2 -P 2 -i 2 -m 64:4194304:32 -r all -d int32,float32,bfloat16 -o sum,max,avg --triggered 2"
timeout $(cap 30) $TR --master-port 29851 tools/ucc_test_dist.py -c alltoall,barrier $ARGS > $O/test_dist_alltoall.log 2>&1; echo "test_dist rc=$?"
grep -A5 "TEST REPORT" $O/test_dist_alltoall.log | cut -c1-120; grep -A8 "caught signal" $O/test_dist_alltoall.log | head -12 | cut -c1-160
[ $(left) -gt 15 ] && timeout $(cap 30) $TR --master-port 29849 tools/p2p_bench.py > $O/p2p.log 2>&1; echo "p2p rc=$?"; grep '^{' $O/p2p.log
[ $(left) -gt 15 ] && UCC_B200_EXPERIMENTAL_TESTS=1 timeout $(cap 45) python -m pytest tests/test_dist_gpu.py -m gpu -q -p no:cacheprovider -k "parallel_helpers or torch_backend" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_sub.log | cut -c1-250
