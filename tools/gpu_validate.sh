#!/bin/bash
# Functional validation on a GPU box:  gpurun --gpus 2 -- 'bash tools/gpu_validate.sh'
# (one GPU is enough for everything except tests/test_dist_gpu.py, which skips below two)
export PYTHONPATH=$PWD UCC_HANDLE_ERRORS=bt
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nvl_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/nvl_test.log
timeout 900 python -m pytest tests/test_dist_gpu.py -x -q -m gpu 2>&1 | tail -40 > gpurun_out/dist_test.log
timeout 300 python -m pytest tests/test_ec_mc.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/ec_test.log
tail -3 gpurun_out/nvl_test.log; tail -5 gpurun_out/dist_test.log | cut -c1-300; tail -2 gpurun_out/ec_test.log
