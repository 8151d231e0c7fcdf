#!/bin/bash
# Two-GPU debug session for tools/ucc_test_dist.py -M cuda: full run with fatal-signal backtraces, then one run per collective.
export PYTHONPATH=$PWD PYTHONFAULTHANDLER=1 UCC_HANDLE_ERRORS=bt
O=gpurun_out/dbg2; mkdir -p $O
T0=$(date +%s); BUDGET=${BUDGET:-250}
left() { echo $(( BUDGET - ($(date +%s) - T0) )); }
cap() { local l=$(left); [ $l -lt $1 ] && echo $l || echo $1; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ARGS="-M cuda -t world,reverse -I 2 -P 2 -i 2 -m 64:4194304:32 -r all -d int32,float32,bfloat16 -o sum,max,avg --triggered 2 -v"
timeout $(cap 100) $TR --master-port 29851 tools/ucc_test_dist.py $ARGS > $O/full.log 2>&1; echo "full rc=$?"
grep -v -i "warn\|^\[OK\]\|^\[SKIP\]" $O/full.log | grep -B2 -A30 "caught signal\|Fatal Python\|FAIL\|TEST REPORT" | head -80 | cut -c1-220
grep "^\[OK\]\|^\[SKIP\]" $O/full.log | tail -3
timeout $(cap 60) $TR --master-port 29849 tools/p2p_bench.py > $O/p2p.log 2>&1; echo "p2p rc=$?"; grep '^{' $O/p2p.log; grep -i "error\|Traceback" -A5 $O/p2p.log | head -20
UCC_B200_EXPERIMENTAL_TESTS=1 timeout $(cap 90) python -m pytest tests/test_dist_gpu.py -m gpu -q -p no:cacheprovider -k "parallel_helpers or torch_backend or all_gpus" > $O/pytest_p2p.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_p2p.log | cut -c1-300
if ! grep -q UCC_TEST_DIST_OK $O/full.log; then
  port=29860
  for c in barrier allreduce allgather allgatherv bcast alltoall alltoallv reduce reduce_scatter reduce_scatterv gather gatherv scatter scatterv; do
    [ $(left) -lt 25 ] && break
    port=$((port+1))
    timeout $(cap 40) $TR --master-port $port tools/ucc_test_dist.py -c $c $ARGS > $O/$c.log 2>&1; rc=$?
    echo "== $c rc=$rc $(grep -c '^\[OK\]' $O/$c.log) ok; $(grep -A4 'TEST REPORT' $O/$c.log | tr '\n' ' ' | cut -c1-160)"
    [ $rc -ne 0 ] && { grep "^\[OK\]\|^\[SKIP\]" $O/$c.log | tail -2; grep -A25 "caught signal\|Fatal Python" $O/$c.log | head -50 | cut -c1-200; }
  done
fi
