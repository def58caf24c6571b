#!/bin/bash
# last sanity pass after the final edits: GPU tests + one quick bench line (cfg2)
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/final2_pytest.log
timeout 300 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final2_cfg2.json 2> gpurun_out/final2_cfg2.err
echo finished
