#!/bin/bash
# eighth GPU session (2 GPUs): flag-in-data exchange -- sharded parity (incl. the invariant-subspace case), cfg3 and cfg4 at N=2
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
nvidia-smi -L > gpurun_out/r02h_box.log 2>&1; free -g >> gpurun_out/r02h_box.log
rm -f gpurun_out/multi_gpu_check.log
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02h_pytest.log
cp gpurun_out/multi_gpu_check.log gpurun_out/r02h_multi_gpu_check.log 2>/dev/null
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $T bench.py --gpus 2 --workload cfg3 --steps 3 --warmup 2 > gpurun_out/r02h_cfg3_n2.json 2> gpurun_out/r02h_cfg3_n2.err
timeout 900 $T bench.py --gpus 2 --workload cfg4 --steps 2 --warmup 1 > gpurun_out/r02h_cfg4_n2.json 2> gpurun_out/r02h_cfg4_n2.err
env M3B_NO_MR_FUSED=1 timeout 600 $T bench.py --gpus 2 --workload cfg3 --steps 3 --warmup 2 --no-e2e > gpurun_out/r02h_cfg3_n2_nccl.json 2> gpurun_out/r02h_cfg3_n2_nccl.err
echo finished
