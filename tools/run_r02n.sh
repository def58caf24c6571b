#!/bin/bash
# fourteenth GPU session (8 GPUs), final code: sharded parity at world 8 against the oracle, cfg4 / cfg3 / cfg5 at N=8
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
rm -f gpurun_out/multi_gpu_check.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "50-8 or 6-8" 2>&1 | tail -8 > gpurun_out/r02n_pytest.log
cp gpurun_out/multi_gpu_check.log gpurun_out/r02n_multi_gpu_check_world8.log 2>/dev/null
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 600 $T8 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02n_cfg4_n8.json 2> gpurun_out/r02n_cfg4_n8.err
timeout 600 $T8 bench.py --gpus 8 --workload cfg3 --steps 3 --warmup 3 > gpurun_out/r02n_cfg3_n8.json 2> gpurun_out/r02n_cfg3_n8.err
timeout 600 $T8 bench.py --gpus 8 --workload cfg5 --steps 3 --warmup 2 > gpurun_out/r02n_cfg5_n8.json 2> gpurun_out/r02n_cfg5_n8.err
echo finished
