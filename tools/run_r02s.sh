#!/bin/bash
# product-kernel knobs at the cfg3 shape: batch depth (variants 1, 2) and piece sizes
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
S="python tools/spmv_sweep.py --cells 250000 --genes 30000 --density 0.06 --num-dim 50 --maxit 4"
timeout 300 $S --variants 1,2 > gpurun_out/r02s_v12.log 2>&1
for pg in 4096 16384; do timeout 300 $S --variants 1 --extra-env M3B_FLAT_PIECE_G=$pg > gpurun_out/r02s_pg$pg.log 2>&1; done
for pu in 16384 65536; do timeout 300 $S --variants 1 --extra-env M3B_FLAT_PIECE_U=$pu > gpurun_out/r02s_pu$pu.log 2>&1; done
echo finished
