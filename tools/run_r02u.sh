#!/bin/bash
# cfg3 at N = 2 with the final code
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581"
timeout 400 $T bench.py --gpus 2 --workload cfg3 --steps 3 --warmup 3 > gpurun_out/r02u_cfg3_n2.json 2> gpurun_out/r02u_cfg3_n2.err
echo finished
