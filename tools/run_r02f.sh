#!/bin/bash
# sixth GPU session: segment-major scatter (B3), one-pass A, packed count: tests + A/B + timings
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02f_pytest.log
timeout 300 python tools/ab_tables.py > gpurun_out/r02f_ab_tables.log 2>&1
env M3B_DEBUG_TIMING=1 timeout 300 python tools/profile_select.py --reps 3 > gpurun_out/r02f_select.log 2>&1
timeout 600 ncu --kernel-name regex:^k_ --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02f_select_launches.csv python tools/profile_select.py --reps 1 > gpurun_out/r02f_ncu_select.log 2>&1
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 2 > gpurun_out/r02f_cfg3.json 2> gpurun_out/r02f_cfg3.err
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_cfg2.json 2> gpurun_out/r02f_cfg2.err
echo finished
