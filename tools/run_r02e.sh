#!/bin/bash
# fifth GPU session: new bank-reorder kernel + chunks-per-tile, tiny branch / align tests, full ncu captures of the build kernels and K10
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02e_pytest.log
env M3B_DEBUG_TIMING=1 timeout 300 python tools/profile_select.py --reps 3 > gpurun_out/r02e_select.log 2>&1
timeout 300 python tools/ab_tables.py > gpurun_out/r02e_ab_tables.log 2>&1
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r02e_cfg3.json 2> gpurun_out/r02e_cfg3.err
timeout 600 ncu --kernel-name regex:^k_scatter_B2 --launch-count 1 --set full --clock-control none --import-source on -o gpurun_out/r02e_full_scatter_B2 -f python tools/profile_select.py --reps 1 > gpurun_out/r02e_ncu1.log 2>&1
timeout 600 ncu --kernel-name regex:^k_scatter_A2 --launch-count 1 --set full --clock-control none --import-source on -o gpurun_out/r02e_full_scatter_A2 -f python tools/profile_select.py --reps 1 > gpurun_out/r02e_ncu2.log 2>&1
timeout 600 ncu --kernel-name regex:^k_bank_reorder --launch-skip 1 --launch-count 1 --set full --clock-control none --import-source on -o gpurun_out/r02e_full_bank_reorder_unit -f python tools/profile_select.py --reps 1 > gpurun_out/r02e_ncu3.log 2>&1
timeout 900 ncu --kernel-name regex:^k_project_tiled --launch-count 1 --set full --clock-control none --import-source on -o gpurun_out/r02e_full_project_tiled -f python bench.py --workload cfg5 --cells 100000 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r02e_ncu4.log 2>&1
timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02e_cfg5_n1.json 2> gpurun_out/r02e_cfg5_n1.err
echo finished
