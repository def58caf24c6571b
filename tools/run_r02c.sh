#!/bin/bash
# third GPU session: where does m3b_select_genes spend its time (launch list, chunk sweep); tiled projection (K10) tests + cfg5 line
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "transform or lsi or errors or ragged or a549_model" 2>&1 | tail -8 > gpurun_out/r02c_pytest_transform.log
for cpt in 0 60 148; do
  if [ $cpt = 0 ]; then env M3B_DEBUG_TIMING=1 timeout 300 python tools/profile_select.py --reps 3 > gpurun_out/r02c_select_cpt_default.log 2>&1
  else env M3B_CPT=$cpt M3B_DEBUG_TIMING=1 timeout 300 python tools/profile_select.py --reps 3 > gpurun_out/r02c_select_cpt_$cpt.log 2>&1; fi
done
env M3B_TILE_W=10240 M3B_CPT=148 timeout 300 python tools/profile_select.py --reps 3 > gpurun_out/r02c_select_tw10240_cpt148.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_select_launches.csv python tools/profile_select.py --reps 1 > gpurun_out/r02c_ncu_select.log 2>&1
timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 2 > gpurun_out/r02c_cfg5_n1.json 2> gpurun_out/r02c_cfg5_n1.err
M3B_PROJECT_NO_TMA=1 timeout 900 python bench.py --workload cfg5 --cells 100000 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r02c_cfg5_100k_notma.json 2> gpurun_out/r02c_cfg5_100k_notma.err
timeout 900 python bench.py --workload cfg5 --cells 100000 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r02c_cfg5_100k.json 2> gpurun_out/r02c_cfg5_100k.err
echo finished
