#!/bin/bash
# the driver's own command on eight GPUs with the final code
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571"
( time $T8 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/final_bench_n8.json 2> gpurun_out/final_bench_n8.err
( time $T8 bench.py --impl reference --gpus 8 --steps 1 --warmup 1 ) > gpurun_out/final_ref_n8.json 2> gpurun_out/final_ref_n8.err
echo finished
