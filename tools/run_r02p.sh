#!/bin/bash
# ncu evidence for the default bench command (cfg4 on one GPU): launch list (kernel shares of the step) and
# full captures of the projection kernel at the cfg5 shape
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-roofline-pass > gpurun_out/r02p_plain.json 2> gpurun_out/r02p_plain.err
timeout 1500 ncu --kernel-name regex:^k_ --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/r02p_cfg4_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-roofline-pass > gpurun_out/r02p_ncu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_project_tiled2 -c 1 -o gpurun_out/r02p_full_project_tiled2 -f python tools/bench_project.py --cells 120000 --genes 30000 --density 0.05 > gpurun_out/r02p_ncu_project.log 2>&1
echo finished
