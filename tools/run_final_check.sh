#!/bin/bash
# last sanity pass after the final edits: smoke + GPU tests + one quick bench line (cfg2)
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final4_smoke.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/final4_pytest.log
timeout 300 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/final4_cfg2.json 2> gpurun_out/final4_cfg2.err
echo finished
