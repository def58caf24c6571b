#!/bin/bash
# bank-reorder size classes: tests, m3b_select_genes timing at the cfg3 shape, cfg3 line
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02r_pytest.log
env M3B_DEBUG_TIMING=1 timeout 300 python tools/profile_select.py --reps 3 > gpurun_out/r02r_select.log 2>&1
timeout 600 ncu --kernel-name regex:^k_ --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r02r_select_launches.csv python tools/profile_select.py --reps 1 > gpurun_out/r02r_ncu_select.log 2>&1
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_cfg3_n1.json 2> gpurun_out/r02r_cfg3_n1.err
echo finished
