#!/bin/bash
# ninth GPU session (1 GPU): restart GEMM v2 (A/B + ncu), balance-aware tiling, cfg4 on one GPU again
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02i_pytest.log
for v in v2 v1; do
  E=""; [ $v = v1 ] && E="M3B_GEMM_V1=1"
  env $E timeout 300 python bench.py --workload cfg3 --num-dim 100 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02i_cfg3nd100_$v.json 2> gpurun_out/r02i_cfg3nd100_$v.err
done
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_restart_gemm2 -s 3 -c 1 -o gpurun_out/r02i_full_restart_gemm2 -f python tools/profile_step.py --cells 250000 --genes 30000 --density 0.06 --num-dim 100 --maxit 3 > gpurun_out/r02i_ncu_gemm2.log 2>&1
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 2 > gpurun_out/r02i_cfg3_n1.json 2> gpurun_out/r02i_cfg3_n1.err
timeout 1200 python bench.py --workload cfg4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02i_cfg4_n1.json 2> gpurun_out/r02i_cfg4_n1.err
echo finished
