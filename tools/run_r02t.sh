#!/bin/bash
# full ncu capture of the fused vector kernel at the per-rank cfg4 shape (250k cells, work = 107), late in a cycle
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_orth_fused -s 180 -c 2 -o gpurun_out/r02t_full_orth_fused -f python tools/profile_step.py --cells 250000 --genes 30000 --density 0.06 --num-dim 100 --maxit 3 > gpurun_out/r02t_ncu.log 2>&1
echo finished
