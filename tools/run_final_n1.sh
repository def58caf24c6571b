#!/bin/bash
# the driver's own commands on one GPU with the final code: build check, smoke, GPU tests, default bench, reference arm
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final_smoke.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/final_pytest.log
( time python bench.py ) > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
( time python bench.py --impl reference ) > gpurun_out/final_ref_n1.json 2> gpurun_out/final_ref_n1.err
echo finished
