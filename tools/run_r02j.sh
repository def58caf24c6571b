#!/bin/bash
# tenth GPU session (8 GPUs): sharded parity at world 8, cfg4 at N=8 and N=4, cfg3 at N=8, cfg5 at N=8
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
nvidia-smi -L > gpurun_out/r02j_box.log 2>&1; free -g >> gpurun_out/r02j_box.log; nvidia-smi topo -m >> gpurun_out/r02j_box.log 2>&1
rm -f gpurun_out/multi_gpu_check.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "8-" 2>&1 | tail -8 > gpurun_out/r02j_pytest.log
cp gpurun_out/multi_gpu_check.log gpurun_out/r02j_multi_gpu_check.log 2>/dev/null
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542"
timeout 600 $T8 bench.py --gpus 8 --workload cfg4 --steps 3 --warmup 2 > gpurun_out/r02j_cfg4_n8.json 2> gpurun_out/r02j_cfg4_n8.err
timeout 600 $T8 bench.py --gpus 8 --workload cfg3 --steps 3 --warmup 2 > gpurun_out/r02j_cfg3_n8.json 2> gpurun_out/r02j_cfg3_n8.err
timeout 600 $T4 bench.py --gpus 4 --workload cfg4 --steps 2 --warmup 1 > gpurun_out/r02j_cfg4_n4.json 2> gpurun_out/r02j_cfg4_n4.err
timeout 600 $T8 bench.py --gpus 8 --workload cfg5 --steps 3 --warmup 2 > gpurun_out/r02j_cfg5_n8.json 2> gpurun_out/r02j_cfg5_n8.err
echo finished
