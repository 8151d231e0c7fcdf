#!/bin/bash
# sweep every tl/shm algorithm through the multi-process matrix test
cd "$(dirname "$0")/.."
port=29700
declare -A ALGS=( [allgather]="knomial ring neighbor bruck sparbit linear batched" [allgatherv]="ring knomial linear" [allreduce]="knomial sra_knomial dbt ring sliding_window" [alltoall]="pairwise bruck onesided" [alltoallv]="pairwise hybrid onesided" [bcast]="knomial sag_knomial dbt" [gather]="knomial linear" [reduce]="knomial dbt srg" [reduce_scatter]="ring knomial" [scatter]="knomial linear" )
for n in 3 4; do
for coll in "${!ALGS[@]}"; do
  for alg in ${ALGS[$coll]}; do
    port=$((port+1))
    UCC_TL_SHM_TUNE="$coll:@$alg:inf" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port $port tools/ucc_test_dist.py -t world,odd_even,reverse -c $coll -I 2 -P 2 -i 2 -m 8:300000:40 -r all -d int32,float32,bfloat16 -o sum,max,avg -s 5 > /tmp/ucc_b200_sweep_${n}_${coll}_${alg}.log 2>&1
    rc=$?
    echo "n=$n $coll@$alg rc=$rc $(grep -A5 'TEST REPORT' /tmp/ucc_b200_sweep_${n}_${coll}_${alg}.log | grep -E 'passed|skipped|failed' | tr -s ' ' | tr '\n' ' ')"
  done
done
done
