#!/bin/bash
# last sanity pass of the sharded path (2 GPUs): the three oracle-checked cases at world 2
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
rm -f gpurun_out/multi_gpu_check.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/final3_pytest.log
cp gpurun_out/multi_gpu_check.log gpurun_out/final3_multi_gpu_check.log 2>/dev/null
echo finished
