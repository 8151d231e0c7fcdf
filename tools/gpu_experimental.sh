#!/bin/bash
# Round-2 opener: things written at the end of round 1 with no GPU time left.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_experimental.sh 2'      (then 4 / 8 for the numbers)
# 1. nvls_pipe correctness (2 heap sizes -> 1 chunk and many chunks), 2. its bandwidth vs the staged nvls and the default,
# 3. the multi-process matrix test on CUDA memory, 4. symmetric-memory allreduce (correctness + bandwidth), 5. the asymmetric-memory test that hung once.
export PYTHONPATH=$PWD UCC_HANDLE_ERRORS=bt UCC_B200_EXPERIMENTAL_TESTS=1
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29821"
timeout 600 python -m pytest tests/test_dist_gpu.py -x -q -m gpu -k nvls_pipe 2>&1 | tail -15 > gpurun_out/pipe_test.log
for alg in default nvls nvls_pipe; do
  if [ "$alg" = default ]; then T=""; else T="allreduce:cuda:inf:@$alg"; fi
  UCC_TL_NVL_TUNE="$T" UCC_TL_NVL_ALLREDUCE_ONESHOT_THRESH=0 timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${alg}_$N.log 2>&1
  UCC_TL_NVL_SYMMETRIC_SIZE=384M UCC_TL_NVL_TUNE="$T" UCC_TL_NVL_ALLREDUCE_ONESHOT_THRESH=0 timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${alg}_heap384_$N.log 2>&1
done
UCC_TL_NVL_TUNE="allgather:cuda:inf:@push#alltoall:cuda:inf:@push#alltoallv:cuda:inf:@push" timeout 300 $TR tests/dist_worker.py cuda > gpurun_out/push_worker_$N.log 2>&1
UCC_TL_NVL_TUNE="reduce_scatter:cuda:0-inf:@oneshot#reduce_scatterv:cuda:0-inf:@oneshot" timeout 300 $TR tests/dist_worker.py cuda > gpurun_out/oneshot_rs_worker_$N.log 2>&1
timeout 300 $TR tools/coll_bench.py > gpurun_out/coll_default_$N.log 2>&1
UCC_TL_NVL_TUNE="allgather:cuda:inf:@push#alltoall:cuda:inf:@push#alltoallv:cuda:inf:@push" timeout 300 $TR tools/coll_bench.py > gpurun_out/coll_push_$N.log 2>&1
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --symm 3G > gpurun_out/bench_symm_$N.log 2>&1
timeout 600 $TR tools/ucc_test_dist.py -M cuda -t world,reverse -I 2 -P 2 -i 2 -m 64:4194304:16 -r all -d int32,float32,bfloat16 -o sum,max,avg --triggered 2 > gpurun_out/test_dist_cuda_$N.log 2>&1
SYMM_SIZE=3G timeout 400 $TR tests/symm_worker.py > gpurun_out/symm_$N.log 2>&1
timeout 200 python -m pytest tests/test_nvl_gpu.py -x -q -m gpu -k asymmetric 2>&1 | tail -15 > gpurun_out/asym_test.log
tail -3 gpurun_out/pipe_test.log; for f in gpurun_out/bench_*_$N.log; do echo "$f: $(tail -c 300 $f | tr '\n' ' ')"; done; grep -A6 "TEST REPORT" gpurun_out/test_dist_cuda_$N.log; grep "SYMM_\|mismatch\|guard" gpurun_out/symm_$N.log | head; tail -3 gpurun_out/asym_test.log
