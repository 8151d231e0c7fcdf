#!/bin/bash
# Round-2 validation on N GPUs:  gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_r2_validate.sh N'
# 1. the whole GPU suite incl. the kernels that had never run on hardware (UCC_B200_EXPERIMENTAL_TESTS=1)
# 2. reference arm + our arm of bench.py   3. allreduce algorithm matrix   4. the other collectives (default vs push)
export PYTHONPATH=$PWD UCC_HANDLE_ERRORS=bt
N=${1:-2}
O=gpurun_out/r2v$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29821"
nvidia-smi topo -m > $O/topo.txt 2>&1
UCC_B200_EXPERIMENTAL_TESTS=1 timeout 900 python -m pytest tests -m gpu -q --durations=15 -p no:cacheprovider > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 $TR bench.py --impl reference --gpus $N --steps 10 --warmup 3 --out $O/bench_ref.json > $O/bench_ref.log 2>&1
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --out $O/bench_ours.json > $O/bench_ours.log 2>&1
for alg in nvls nvls_pipe twoshot; do
  UCC_TL_NVL_TUNE="allreduce:cuda:inf:@$alg" UCC_TL_NVL_ALLREDUCE_ONESHOT_THRESH=0 timeout 200 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-sweep --no-e2e --out $O/bench_$alg.json > $O/bench_$alg.log 2>&1
done
UCC_TL_NVL_SYMMETRIC_SIZE=384M UCC_TL_NVL_TUNE="allreduce:cuda:inf:@nvls_pipe" UCC_TL_NVL_ALLREDUCE_ONESHOT_THRESH=0 timeout 200 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-sweep --no-e2e --out $O/bench_nvls_pipe_384.json > $O/bench_nvls_pipe_384.log 2>&1
timeout 200 $TR bench.py --gpus $N --steps 10 --warmup 3 --symm 3G --no-sweep --no-e2e --out $O/bench_symm.json > $O/bench_symm.log 2>&1
timeout 200 $TR tools/coll_bench.py > $O/coll_default.log 2>&1
UCC_TL_NVL_TUNE="allgather:cuda:inf:@push#alltoall:cuda:inf:@push#alltoallv:cuda:inf:@push" timeout 200 $TR tools/coll_bench.py > $O/coll_push.log 2>&1
tail -4 $O/pytest.log
for f in $O/bench_*.log; do echo "$f: $(tail -c 700 $f | tr '\n' ' ' | cut -c1-700)"; done
tail -n 2 $O/coll_default.log | cut -c1-600; tail -n 2 $O/coll_push.log | cut -c1-600
