#!/bin/bash
# cfg4 at N = $1 with the final code (the table of DESIGN.md section 6)
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N"
timeout 900 $T bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r02o_cfg4_n$N.json 2> gpurun_out/r02o_cfg4_n$N.err
echo finished
