#!/bin/bash
# Two-GPU session: gpurun --gpus 2 --timeout 700 -- 'bash tools/gpu_r2_two.sh'
export PYTHONPATH=$PWD
O=gpurun_out/r2two
mkdir -p $O
T0=$(date +%s); BUDGET=${BUDGET:-620}
left() { echo $(( BUDGET - ($(date +%s) - T0) )); }
cap() { local l=$(left); [ $l -lt $1 ] && echo $l || echo $1; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29831"
UCC_B200_EXPERIMENTAL_TESTS=1 timeout $(cap 330) python -m pytest tests/test_dist_gpu.py -m gpu -q --durations=6 -p no:cacheprovider > $O/pytest_dist.log 2>&1; echo "rc=$?" >> $O/pytest_dist.log
tail -12 $O/pytest_dist.log | cut -c1-300
[ $(left) -gt 60 ] && timeout $(cap 100) $TR bench.py --impl reference --gpus 2 --steps 10 --warmup 3 --no-sweep --out $O/bench_ref.json > $O/bench_ref.log 2>&1
[ $(left) -gt 60 ] && timeout $(cap 100) $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-sweep --out $O/bench_ours.json > $O/bench_ours.log 2>&1
[ $(left) -gt 40 ] && timeout $(cap 60) $TR tools/p2p_bench.py > $O/p2p.log 2>&1
[ $(left) -gt 40 ] && timeout $(cap 70) $TR tools/lanes_bench.py > $O/lanes.log 2>&1
[ $(left) -gt 40 ] && timeout $(cap 80) $TR tools/nvlink_traffic.py > $O/nvlink.log 2>&1
[ $(left) -gt 40 ] && COLL_VARIANTS=default,push,push_nobulk,push_bulk32,ce timeout $(cap 90) $TR tools/coll_bench.py > $O/colls.log 2>&1
[ $(left) -gt 60 ] && timeout $(cap 150) bash tools/gpu_perftest.sh 2 > $O/perftest.log 2>&1
[ $(left) -gt 50 ] && timeout $(cap 120) $TR tools/ucc_test_dist.py -M cuda -t world,reverse -I 2 -P 2 -i 2 -m 64:4194304:32 -r all -d int32,float32,bfloat16 -o sum,max,avg --triggered 2 > $O/test_dist_cuda.log 2>&1
grep -A8 "TEST REPORT" $O/test_dist_cuda.log | cut -c1-200
python - <<PY
import json
for f in ("ref", "ours"):
    try:
        d = json.load(open("$O/bench_%s.json" % f))
        print(f, "busbw", d["busbw_per_gpu_GBps"], "us", d["latency_us"], "ok", d["correct"], "e2e us", d["e2e"]["us_per_step"], "e2e ok", d["e2e"].get("correct"))
    except Exception as e:
        print(f, "no json", e)
PY
grep '^{' $O/colls.log | python -c "
import json,sys
for l in sys.stdin:
    for r in json.loads(l)['rows']:
        print({k:v for k,v in r.items() if k!='kernels'})
" | cut -c1-700
grep '^{' $O/p2p.log; grep '^{' $O/lanes.log; grep '^{' $O/nvlink.log | cut -c1-700; grep -v WARN $O/perftest.log | tail -n 60 | cut -c1-200
