#!/bin/bash
# Round-2 validation on N GPUs:  gpurun --gpus N --timeout 900 -- 'bash tools/gpu_r2_validate.sh N [sections]'
# sections: tests  = the whole multi-process GPU suite            smoke = only test_multiproc_all_gpus
#           bench  = both arms of bench.py                        matrix = allreduce algorithm / SM-budget matrix (one launch)
#           colls  = the other collectives, default / push / ce variants (one launch)
export PYTHONPATH=$PWD
N=${1:-2}
SECT=${2:-"tests bench matrix colls"}
O=gpurun_out/r2v$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29821"
T0=$(date +%s)
BUDGET=${BUDGET:-430}      # seconds: sections that would start after this are skipped (an N-GPU box is charged N x wall time)
left() { echo $(( BUDGET - ($(date +%s) - T0) )); }
has() { [[ " $SECT " == *" $1 "* ]] && [ $(left) -gt 20 ]; }
cap() { local want=$1; local l=$(left); [ $l -lt $want ] && echo $l || echo $want; }
if has tests; then
  UCC_B200_EXPERIMENTAL_TESTS=1 timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q --durations=20 -p no:cacheprovider > $O/pytest_dist.log 2>&1; echo "rc=$?" >> $O/pytest_dist.log
  tail -25 $O/pytest_dist.log | cut -c1-400
fi
if has smoke; then
  timeout $(cap 60) python -m pytest tests/test_dist_gpu.py -m gpu -q -k "all_gpus" -p no:cacheprovider > $O/pytest_smoke.log 2>&1; echo "rc=$?" >> $O/pytest_smoke.log
  tail -6 $O/pytest_smoke.log | cut -c1-400
fi
if has bench; then
  timeout $(cap 150) $TR bench.py --impl reference --gpus $N --steps 10 --warmup 3 --out $O/bench_ref.json > $O/bench_ref.log 2>&1
  timeout $(cap 150) $TR bench.py --gpus $N --steps 10 --warmup 3 --out $O/bench_ours.json > $O/bench_ours.log 2>&1
fi
if has matrix; then
  MATRIX_VARIANTS=${MATRIX_VARIANTS:-default,twoshot,nvls,nvls_pipe,nvls_pipe384,symm,symm_nb32,symm_nb64,default_nb64} timeout $(cap 120) $TR tools/alg_matrix.py > $O/matrix.log 2>&1
  grep '^{' $O/matrix.log | cut -c1-900
fi
if has colls; then
  COLL_VARIANTS=${COLL_VARIANTS:-default,push,ce} timeout $(cap 100) $TR tools/coll_bench.py > $O/colls.log 2>&1
  grep '^{' $O/colls.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    for r in d['rows']:
        print({k:v for k,v in r.items() if k!='kernels'}, '|', r.get('kernels'))
" | cut -c1-1200
fi
if has colls || has nvlink; then
  timeout $(cap 80) $TR tools/nvlink_traffic.py > $O/nvlink.log 2>&1
  grep '^{' $O/nvlink.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    r=(d.get('per_rank') or [d])[0]
    print(d.get('case'), '|', {k:r.get(k) for k in ('kernel','us','alg_tx_MB','nvlink_tx_MB','nvlink_rx_MB','tx_GBps','frac_of_900_tx','nvlink_counters','error') if r.get(k) is not None})
"
fi
python - <<PY
import glob, json, os
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d = json.load(open(f))
        print(os.path.basename(f), "busbw/gpu", d.get("busbw_per_gpu_GBps"), "us", d.get("latency_us"), "ok", d.get("correct"), d.get("config", {}).get("algorithm", "")[-60:], "| e2e", (d.get("e2e") or {}).get("us_per_step"), "| ref", d.get("reference_arms"), "| nccl", d.get("nccl_same_box"))
        for r in d.get("sweep", []):
            print("   ", {k: r[k] for k in r if k in ("bytes", "us", "busbw", "ok", "alg", "nccl_us", "nccl_busbw", "tl_cuda_us", "tl_cuda_busbw")})
    except Exception as e:
        print(f, "unreadable", e)
PY
for f in $O/bench_*.log; do [ -s ${f%.log}.json ] || { echo "--- $f (no json)"; tail -c 800 $f; }; done
