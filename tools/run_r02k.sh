#!/bin/bash
# eleventh GPU session (1 GPU): projection kernel v2 (producer warp + register queue): tests, A/B at cfg5, ncu
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02k_pytest.log
timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 2 > gpurun_out/r02k_cfg5_v2.json 2> gpurun_out/r02k_cfg5_v2.err
env M3B_PROJECT_V1=1 timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02k_cfg5_v1.json 2> gpurun_out/r02k_cfg5_v1.err
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_project_tiled2 -c 1 -o gpurun_out/r02k_full_project_tiled2 -f python tools/bench_project.py > gpurun_out/r02k_ncu_project2.log 2>&1
echo finished
