#!/bin/bash
# ucc_perftest (the tool BASELINE.json's metric is named after) on CUDA buffers, N ranks = N GPUs of this box, rendezvous over TCP:
#   gpurun --gpus N -- 'bash tools/gpu_perftest.sh N'
# reference CI configuration (.ci/scripts/run_tests_ucc_nvls.sh): allreduce -m cuda float32 / bfloat16 sweep; plus the executor
# micro-benchmarks (memcpy, reducedt) of ec/cuda on one rank.
export PYTHONPATH=$PWD
N=${1:-2}
O=gpurun_out/perftest$N
mkdir -p $O
PT=ucc_b200/bin/ucc_perftest
run() { # name, args...
  local name=$1; shift
  local pids=()
  for r in $(seq 0 $((N - 1))); do
    RANK=$r LOCAL_RANK=$r WORLD_SIZE=$N MASTER_ADDR=127.0.0.1 MASTER_PORT=29901 timeout 200 $PT "$@" > $O/${name}.r$r.log 2>&1 &
    pids+=($!)
  done
  for p in "${pids[@]}"; do wait $p; done
  cp $O/${name}.r0.log $O/${name}.log; echo "== $name: $*"; grep -v WARN $O/${name}.log | tail -n 25
}
run allreduce_f32 -c allreduce -m cuda -F -T -b 256 -e 256M -d float32 -f 4
run allreduce_bf16 -c allreduce -m cuda -F -T -b 256 -e 256M -d bfloat16 -f 4
run allreduce_f32_persistent -c allreduce -m cuda -F -T -p -b 256 -e 64K -d float32 -f 4
run alltoall_f32 -c alltoall -m cuda -F -T -b 1K -e 16M -d float32 -f 8
run allgather_f32 -c allgather -m cuda -F -T -b 1K -e 16M -d float32 -f 8
N=1 run memcpy -c memcpy -m cuda -b 4K -e 256M -f 8
N=1 run reducedt -c reducedt -m cuda -b 4K -e 64M -f 8 -d float32 -N 4
