#!/bin/bash
# fourth GPU session: kernel-level view of m3b_select_genes, bank-reorder on/off, K10 v2, align tests
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_align.py -m gpu -q -x -k "transform or align or lsi or a549_model" 2>&1 | tail -8 > gpurun_out/r02d_pytest.log
env M3B_DEBUG_TIMING=1 timeout 300 python tools/profile_select.py --reps 3 > gpurun_out/r02d_select_default.log 2>&1
env M3B_NO_BANK_REORDER=1 M3B_DEBUG_TIMING=1 timeout 300 python tools/profile_select.py --reps 3 > gpurun_out/r02d_select_nobr.log 2>&1
env M3B_NO_BANK_REORDER=1 M3B_CPT=148 M3B_DEBUG_TIMING=1 timeout 300 python tools/profile_select.py --reps 3 > gpurun_out/r02d_select_nobr_cpt148.log 2>&1
timeout 600 ncu --kernel-name regex:^k_ --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02d_select_launches.csv python tools/profile_select.py --reps 2 > gpurun_out/r02d_ncu_select.log 2>&1
env M3B_NO_BANK_REORDER=1 timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r02d_cfg3_nobr.json 2> gpurun_out/r02d_cfg3_nobr.err
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r02d_cfg3_br.json 2> gpurun_out/r02d_cfg3_br.err
timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 2 > gpurun_out/r02d_cfg5_n1.json 2> gpurun_out/r02d_cfg5_n1.err
echo finished
